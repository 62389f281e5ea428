"""One camera stream sharded by frame pair through the C ABI's own NCCL layer (rc_comm_* / rc_shard_* /
rc_allreduce_accumulators): per-frame thresholds, the all-reduced accumulator, the reporting-point
mask, the stream's cumulative histogram and the order-dependent window mean must equal the sequential pipeline BIT FOR BIT.

  - one rank (no communicator, then a 1-rank NCCL communicator): the whole code path on a single GPU;
  - two ranks on two GPUs (skipped with fewer than two devices): all-gather of counts, all-to-all hand-off of flow bands
    to the ranks that own the corresponding rows of the window mean, all-reduce of accumulators.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = (0.5, 2, 3, 2, 15, 1.2, 0)
W_IMG, H_IMG, NFRAMES, B, WIN = 320, 240, 26, 4, 5
FC0 = 25                       # loop counter of frame 0: the framecount > 30 gate falls inside the clip


def _clip():
    from ripcurrents_b200 import synth
    return np.stack(synth.clip(W_IMG, H_IMG, NFRAMES, seed=21))


def _sequential(ctx_cls, fr):
    c = ctx_cls(0)
    c.flow_configure_batch(W_IMG, H_IMG, *P, 8); c.hist_reset(); c.window_configure(W_IMG, H_IMG, WIN)
    ups, sums = [], []
    mask = np.zeros((8, H_IMG, W_IMG), np.uint8)
    last_mask = None
    for lo in range(0, NFRAMES, 8):
        nb = min(8, NFRAMES - lo)
        k, res = c.process_frames(fr[lo:lo + nb], FC0 + lo, mask[:nb])
        for i in range(nb):
            if res[i].produced:
                ups.append(res[i].UPPER); sums.append(res[i].histsum)
        last_mask = mask[nb - 1].copy()
    out = dict(ups=ups, sums=sums, acc=c.accumulator_get(W_IMG, H_IMG).copy(), hist=c.hist_get()[2].copy(),
               avg=c.window_get().copy(), mask=last_mask)
    c.close()
    return out


def _run_rank(ctx, fr, rank, world, owner):
    """Walks the clip in super-blocks of world * B pairs; returns this rank's {pair: (UPPER, histsum)}."""
    n_pairs = NFRAMES - 1
    got = {}
    for s0 in range(0, n_pairs, world * B):
        ppr = [max(0, min(B, n_pairs - (s0 + r * B))) for r in range(world)]
        lo = s0 + rank * B
        nb = ppr[rank]
        if nb:
            k, res = ctx.shard_step(fr[lo:lo + nb + 1], FC0 + lo + 1, ppr)
            assert k == nb
            for i in range(nb):
                got[lo + i] = (res[i].UPPER, res[i].histsum)
        else:
            ctx.shard_step(None, 0, ppr)
    return got


@pytest.mark.parametrize("with_nccl", [False, True])
def test_one_rank_equals_sequential(with_nccl):
    from ripcurrents_b200 import Context, capi
    fr = _clip()
    ref = _sequential(Context, fr)
    c = Context(0)
    if with_nccl:
        c.comm_init(capi.comm_unique_id(), 0, 1)
    c.flow_configure_batch(W_IMG, H_IMG, *P, B + 1)
    c.shard_configure(WIN, 0)
    got = _run_rank(c, fr, 0, 1, 0)
    assert [got[p][0] for p in range(NFRAMES - 1)] == ref["ups"]
    assert [got[p][1] for p in range(NFRAMES - 1)] == ref["sums"]
    mask, acc, hist = c.shard_report(FC0 + NFRAMES - 1)
    assert np.array_equal(acc, ref["acc"]) and acc.sum() > 0
    assert np.array_equal(hist, ref["hist"])
    assert np.array_equal(mask, ref["mask"])
    assert c.shard_window_get().tobytes() == ref["avg"].tobytes()
    if with_nccl:          # the stream-per-GPU collective on a 1-rank communicator: identity, in place
        c.allreduce_accumulators()
        c.synchronize()
        assert np.array_equal(c.accumulator_get(W_IMG, H_IMG), ref["acc"])
    c.close()


def _worker(rank, world, uid, q):
    sys.path.insert(0, ROOT)
    try:
        from ripcurrents_b200 import Context
        fr = _clip()
        c = Context(rank)
        c.comm_init(uid, rank, world)
        c.flow_configure_batch(W_IMG, H_IMG, *P, B + 1)
        owner = world - 1                                   # a non-zero owner: rank 0 only sends
        c.shard_configure(WIN, owner)
        got = _run_rank(c, fr, rank, world, owner)
        mask, acc, hist = c.shard_report(FC0 + NFRAMES - 1)
        avg = c.shard_window_get()                          # collective: every rank assembles the whole mean
        # stream-per-GPU collective on the same communicator: shared map = sum of the two ranks' accumulators
        c.allreduce_accumulators()
        c.synchronize()
        both = c.accumulator_get(W_IMG, H_IMG)
        q.put((rank, got, mask, acc, hist, avg, both))
        c.close()
    except Exception as e:       # noqa: BLE001
        q.put((rank, "ERROR: %r" % (e,)))


def test_two_gpus_equal_sequential():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from ripcurrents_b200 import Context, capi
    fr = _clip()
    ref = _sequential(Context, fr)
    uid = capi.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, uid, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert not (len(r) == 2 and isinstance(r[1], str)), r
    got = {}
    for r in res:
        got.update(r[1])
    assert [got[p][0] for p in range(NFRAMES - 1)] == ref["ups"]
    assert [got[p][1] for p in range(NFRAMES - 1)] == ref["sums"]
    for r in res:
        assert np.array_equal(r[3], ref["acc"]) and np.array_equal(r[4], ref["hist"]) and np.array_equal(r[2], ref["mask"])
        assert np.array_equal(r[6], ref["acc"])             # in-place all-reduce of the two ranks' own accumulators
    for r in res:                                            # the mean is sharded by pixel band and assembled on every rank
        assert r[5].tobytes() == ref["avg"].tobytes()
    assert res[0][3].sum() > 0


@pytest.mark.parametrize("ngpus", [1, 2])
def test_cpp_host_shards_one_stream(tmp_path, ngpus):
    """cpp/demo_multi_gpu.cpp: a C++ host (std::thread per GPU, the plain C ABI, no Python, no NCCL headers) shards one stream
    by frame pair; thresholds, accumulator, mask and window mean equal the sequential pipeline bit for bit."""
    import subprocess
    import torch
    if torch.cuda.device_count() < ngpus:
        pytest.skip("needs %d GPUs" % ngpus)
    from ripcurrents_b200 import Context, build
    build.build_cpp()
    fr = _clip()
    ref = _sequential(Context, fr)
    raw = tmp_path / "frames.raw"; out = tmp_path / "out.bin"
    fr.tofile(raw)
    r = subprocess.run([build.DEMO_MULTI, str(raw), str(W_IMG), str(H_IMG), str(NFRAMES), str(ngpus), str(B), str(out)],
                       capture_output=True, text=True, timeout=180)
    assert r.returncode == 0 and "demo_multi_gpu ok" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]
    blob = open(out, "rb").read()
    npairs = int(np.frombuffer(blob, np.int32, 1)[0]); off = 4
    assert npairs == NFRAMES - 1
    upper = np.frombuffer(blob, np.float32, npairs, off); off += 4 * npairs
    hsum = np.frombuffer(blob, np.int64, npairs, off); off += 8 * npairs
    n = W_IMG * H_IMG
    acc = np.frombuffer(blob, np.float32, n, off); off += 4 * n
    mask = np.frombuffer(blob, np.uint8, n, off); off += n
    avg = np.frombuffer(blob, np.float32, 2 * n, off)
    assert upper.tolist() == [np.float32(u) for u in ref["ups"]] and hsum.tolist() == ref["sums"]
    assert np.array_equal(acc.reshape(H_IMG, W_IMG), ref["acc"]) and np.array_equal(mask.reshape(H_IMG, W_IMG), ref["mask"])
    assert avg.tobytes() == ref["avg"].tobytes()
