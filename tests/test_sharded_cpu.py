"""world_size-2 gloo test (CPU) of the frame-pair sharding orchestration (ripcurrents_b200/sharded.py): per-frame
thresholds, total counts and the all-reduced accumulator must equal the sequential reference order
(ripcurrents.cpp:319-439) -- with the CPU oracle standing in for the GPU engine."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    def __init__(self, O, params, w, h):
        self.O, self.P, self.w, self.h = O, params, w, h
        self.acc = np.zeros(w * h, np.float32)
        self.flows = []

    def flows_and_counts(self, frames):
        self.flows = [self.O.farneback(frames[i], frames[i + 1], *self.P) for i in range(len(frames) - 1)]
        out = np.zeros((len(self.flows), 37, 50), np.int64)
        for i, f in enumerate(self.flows):
            st = self.O.HistState(); self.O.histogram(f, st); out[i] = st.hist2d
        return out

    def aggregate(self, prefix, framecounts):
        st = self.O.HistState()
        st.hist2d += prefix; st.hist += prefix.sum(0); st.histsum[0] = prefix.sum(); st.histsum2d += prefix.sum(1)
        ups = []
        for f, fc in zip(self.flows, framecounts):
            self.O.histogram(f, st)
            up, _, _ = self.O.thresholds(st)
            self.O.classify_accumulate(f, up, fc, self.acc)
            ups.append(up)
        return ups

    def accumulator(self):
        return self.acc


def _worker(rank, world, port, q, chunked=False):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle as O
    from ripcurrents_b200 import sharded, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, n = 96, 64, 8
    fr = np.stack(synth.clip(w, h, n, seed=3))
    P = (0.5, 1, 3, 2, 5, 1.1, 0)
    if chunked == "window":
        # the band-sharded window mean next to the aggregation: per super-block of world * 2 pairs every rank computes its
        # flows, the bands are exchanged, each rank updates its rows in stream order
        be = OracleBackend(O, P, w, h)
        win = sharded.BandWindow(w, h, 3, world, rank, O.window_update)
        n_pairs = n - 1
        for s0 in range(0, n_pairs, world * 2):
            lo = min(s0 + rank * 2, n_pairs); hi = min(lo + 2, n_pairs)
            flows = np.stack([O.farneback(fr[i], fr[i + 1], *P) for i in range(lo, hi)]) if hi > lo else np.zeros((0, h, w, 2), np.float32)
            win.update(sharded.exchange_bands(flows, h, dist=dist))
        q.put((rank, win.lo, win.hi, win.avg.copy(), win.count))
    elif chunked:      # super-blocks of world * 2 pairs: 7 pairs = two full super-blocks (one ragged) -> three calls
        res = sharded.run_stream(OracleBackend(O, P, w, h), fr, 2, lambda p: 28 + p, dist=dist)
        q.put((rank, sorted(res["upper"].items()), res["counts_total"], res["accumulator"]))
    else:
        lo, hi = sharded.block_range(n - 1, world, rank)
        res = sharded.run_block(OracleBackend(O, P, w, h), fr, lo, hi, lambda p: 28 + p, dist=dist)
        q.put((rank, lo, hi, res["upper"], res["counts_total"], res["accumulator"]))
    dist.barrier()
    dist.destroy_process_group()


def test_block_range():
    from ripcurrents_b200 import sharded
    assert [sharded.block_range(7, 2, r) for r in range(2)] == [(0, 4), (4, 7)]
    assert [sharded.block_range(5, 8, r) for r in range(8)] == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 5), (5, 5), (5, 5)]
    got = [sharded.block_range(300, 8, r) for r in range(8)]
    assert got[0][0] == 0 and got[-1][1] == 300 and all(a[1] == b[0] for a, b in zip(got, got[1:]))


def test_two_rank_sharding_matches_sequential(oracle):
    from ripcurrents_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # sequential reference
    w, h, n = 96, 64, 8
    fr = np.stack(synth.clip(w, h, n, seed=3))
    P = (0.5, 1, 3, 2, 5, 1.1, 0)
    st = oracle.HistState(); acc = np.zeros(w * h, np.float32); ups = []
    for i in range(n - 1):
        f = oracle.farneback(fr[i], fr[i + 1], *P)
        oracle.histogram(f, st)
        up, _, _ = oracle.thresholds(st)
        oracle.classify_accumulate(f, up, 28 + i, acc)
        ups.append(up)
    assert got[0][1:3] == (0, 4) and got[1][1:3] == (4, 7)
    assert got[0][3] + got[1][3] == ups
    for g in got:
        assert np.array_equal(g[4], st.hist2d)
        assert np.array_equal(g[5], acc)
    assert acc.sum() > 0


def test_two_rank_stream_in_super_blocks_matches_sequential(oracle):
    """Several run_block calls on the same backends (the cumulative counts must carry across calls): a 7-pair
    clip in super-blocks of 2 ranks x 2 pairs equals the sequential order frame by frame."""
    from ripcurrents_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w, h, n = 96, 64, 8
    fr = np.stack(synth.clip(w, h, n, seed=3))
    P = (0.5, 1, 3, 2, 5, 1.1, 0)
    st = oracle.HistState(); acc = np.zeros(w * h, np.float32); ups = {}
    for i in range(n - 1):
        f = oracle.farneback(fr[i], fr[i + 1], *P)
        oracle.histogram(f, st)
        ups[i], _, _ = oracle.thresholds(st)
        oracle.classify_accumulate(f, ups[i], 28 + i, acc)
    assert [p for p, _ in got[0][1]] == [0, 1, 4, 5] and [p for p, _ in got[1][1]] == [2, 3, 6]
    assert dict(got[0][1] + got[1][1]) == ups
    for g in got:
        assert np.array_equal(g[2], st.hist2d)
        assert np.array_equal(g[3], acc)


def test_two_rank_band_sharded_window_mean(oracle):
    """The order-dependent fp32 window mean sharded by row band (what csrc/comm.cu does over NCCL): two ranks, W = 3, a 7-pair
    clip in super-blocks of 2 x 2 pairs; the two bands put together equal the sequential mean bit for bit."""
    from ripcurrents_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + os.getpid() % 40
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, "window")) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w, h, n, W = 96, 64, 8, 3
    fr = np.stack(synth.clip(w, h, n, seed=3))
    P = (0.5, 1, 3, 2, 5, 1.1, 0)
    avg = np.zeros(h * w * 2, np.float32); ring = np.zeros((W, h * w * 2), np.float32)
    for i in range(n - 1):
        oracle.window_update(avg, ring[i % W], oracle.farneback(fr[i], fr[i + 1], *P), W)
    avg = avg.reshape(h, w, 2)
    assert (got[0][1], got[0][2], got[1][1], got[1][2]) == (0, 32, 32, 64)
    for _, lo, hi, band, count in got:
        assert count == n - 1
        assert band.tobytes() == avg[lo:hi].tobytes()
    assert np.abs(avg).max() > 0
