"""Round-2 parity cases (VERDICT r1 / ADVICE r1):
  - the BENCHED kernel configuration (1080p, batch >= 16 M pixel*pairs -> flow_strip_kernel on layers 0/1) against live cv2
    and the full aggregation against the oracle;
  - BASELINE configs[3] at its full length: 1 M seeds / the per-pixel field / 3 495 streaklines over a 300-frame flow
    sequence (vertex cap below the frame count, so `count == cap` saturation and long-horizon out-of-bounds stops happen);
  - odd image sizes through the batched pipeline (ring slots / mask rows off 16-byte alignment);
  - re-arming the sliding window / restarting a clip on a configured context;
  - several sharded super-blocks on the same backends;
  - averageVector's window update against the compiled reference (oracle/_ref) and the oracle.
"""
import numpy as np
import pytest

from util import EPE_MAX_TOL, EPE_MEAN_TOL, cv2_both, epe_vs_cv2

pytestmark = pytest.mark.gpu

P_DEFAULT = (0.5, 2, 3, 2, 15, 1.2, 0)                       # ripcurrents.cpp:215


def _sequential_oracle(oracle, flows, fc0, W, w, h, st=None, acc=None, avg=None, ring=None, pair0=0):
    st = st or oracle.HistState()
    acc = np.zeros(h * w, np.float32) if acc is None else acc
    avg = np.zeros(h * w * 2, np.float32) if avg is None else avg
    ring = np.zeros((W, h * w * 2), np.float32) if ring is None else ring
    ups, sums, masks = [], [], []
    for i, f in enumerate(flows):
        oracle.histogram(f, st)
        up, _, _ = oracle.thresholds(st)
        m, _, _ = oracle.classify_accumulate(f, up, fc0 + i, acc)
        oracle.window_update(avg, ring[(pair0 + i) % W], f, W)
        ups.append(up); sums.append(int(st.histsum[0])); masks.append(m)
    return st, acc, avg, ring, ups, sums, masks


def test_benched_configuration_1080p_strip_kernel(oracle):
    """bench.py's kernel selection: 1080p x 16 pairs = 33 M pixel*pairs per launch -> flow_strip_kernel with row segments on
    layers 0 and 1 (rc_launch_flows, default dispatch, no RC_FLOW_KERNEL override)."""
    cv2 = pytest.importorskip("cv2")
    import os
    assert "RC_FLOW_KERNEL" not in os.environ and "RC_STRIP_MINPX" not in os.environ
    from ripcurrents_b200 import Context, synth
    w, h, B, W = 1920, 1080, 16, 10
    fr = np.stack(synth.clip(w, h, B + 1, seed=2))
    c = Context(0)
    c.flow_configure_batch(w, h, *P_DEFAULT, B); c.hist_reset(); c.window_configure(w, h, W)
    masks = np.zeros((B, h, w), np.uint8)
    c.process_frames(fr[:1], 30, masks[:1])
    k, res = c.process_frames(fr[1:], 31, masks)
    assert k == B
    flows = [c.flow_host_at(B - 1 - i) for i in range(B)]
    for i in (0, 7, 15):                                        # live OpenCV CPU on three pairs of the batch
        r1, r2 = cv2_both(cv2, fr[i], fr[i + 1], P_DEFAULT)
        mean, mx, amb = epe_vs_cv2(flows[i], r1, r2)
        assert amb < 1e-4
        assert mean <= EPE_MEAN_TOL and mx <= EPE_MAX_TOL, (i, mean, mx)
    st, acc, avg, _, ups, sums, rmasks = _sequential_oracle(oracle, flows, 31, W, w, h)
    for i in range(B):
        assert res[i].UPPER == ups[i] and res[i].histsum == sums[i], i
        assert np.array_equal(masks[i], rmasks[i]), i
    assert np.array_equal(c.accumulator_get(w, h).ravel(), acc) and acc.sum() > 0
    assert np.array_equal(c.hist_get()[2], st.hist2d)
    assert np.array_equal(c.window_get().ravel(), avg)
    c.close()


def test_c4_full_length_300_frames(oracle):
    """BASELINE configs[3]: 1080p, 1 048 576 pathline seeds + the per-pixel particle field + 3 495 streakline emitters over
    a 300-frame flow sequence, bit-exact against the oracle after every 30-frame batch."""
    from ripcurrents_b200 import Context, synth
    w, h, nframes, B, cap = 1920, 1080, 300, 30, 256
    c = Context(0)
    c.flow_configure_batch(w, h, *P_DEFAULT, B)
    rng = np.random.default_rng(1)
    n = 1 << 20
    seeds = (rng.random((n, 2)) * [w - 3, h - 3] + 1).astype(np.float32)
    field = np.zeros((h * w, 2), np.float32); dist = np.zeros(h * w, np.float32)
    E = 3495
    em = (rng.random((E, 2)) * [w - 3, h - 3] + 1).astype(np.float32)
    verts = np.zeros((E, cap, 2), np.float32); cnt = np.ones(E, np.int32); verts[:, 0] = em
    r_seeds, r_field, r_dist, r_verts, r_cnt = seeds.copy(), field.copy(), dist.copy(), verts.copy(), cnt.copy()
    base = synth.base_texture(w, h, 4)
    frames_done = 0
    flows = np.empty((B, h, w, 2), np.float32)
    c.flow_push_batch(np.stack([synth.frame(base, w, h, 0)]))
    while frames_done < nframes - 1:
        nb = min(B, nframes - 1 - frames_done)
        fr = np.stack([synth.frame(base, w, h, frames_done + 1 + i) for i in range(nb)])
        assert c.flow_push_batch(fr, flows=flows[:nb]) == nb
        for i in range(nb):
            c.advect(flows[i], seeds, 1.0, 1, 0.0, 0)
            c.advect(flows[i], field, 2.0, 1, 2.0, 5, dist=dist)
            c.streakline_step(flows[i], em, verts, cnt)
            oracle.advect(flows[i], r_seeds, 1.0, 1, 0.0, oracle.ADV_PATHLINE)
            oracle.advect(flows[i], r_field, 2.0, 1, 2.0, oracle.ADV_FIELD, dist=r_dist)
            oracle.streakline_step(flows[i], em, r_verts, r_cnt)
        frames_done += nb
        assert np.array_equal(seeds.view(np.uint32), r_seeds.view(np.uint32)), frames_done
        assert np.array_equal(cnt, r_cnt), frames_done
    assert np.array_equal(field.view(np.uint32), r_field.view(np.uint32))
    assert np.array_equal(dist.view(np.uint32), r_dist.view(np.uint32))
    assert cnt.max() == cap and cnt.min() == cap                                       # saturated at the vertex cap
    assert np.array_equal(verts.view(np.uint32), r_verts.view(np.uint32))
    # long horizon: the texture drifts ~(1, 0.5) px/frame, so seeds that started near the right/bottom edge have stopped
    stopped = (np.floor(seeds[:, 0]) + 2 > w) | (np.floor(seeds[:, 1]) + 2 > h) | (seeds[:, 0] < 1) | (seeds[:, 1] < 1)
    assert stopped.sum() > 1000 and np.isfinite(seeds).all()
    c.close()


@pytest.mark.parametrize("w,h", [(225, 161), (226, 161), (642, 481)])
def test_odd_sizes_through_the_batched_pipeline(oracle, w, h):
    """w*h % 4 != 0: every other flow-ring slot and mask row is off 16-byte / 4-byte alignment."""
    from ripcurrents_b200 import Context, synth
    B, W, n = 5, 3, 11
    fr = np.stack(synth.clip(w, h, n, seed=9))
    P = (0.5, 2, 3, 2, 7, 1.5, 0)
    c = Context(0)
    c.flow_configure_batch(w, h, *P, B); c.hist_reset(); c.window_configure(w, h, W)
    st, acc, avg, ring, pair = None, None, None, None, 0
    masks = np.zeros((B, h, w), np.uint8)
    for lo in range(0, n, B):
        nb = min(B, n - lo)
        k, res = c.process_frames(fr[lo:lo + nb], 28 + lo, masks[:nb])
        first = 1 if lo == 0 else 0
        flows = [c.flow_host_at(nb - 1 - i) for i in range(first, nb)]
        st, acc, avg, ring, ups, sums, rmasks = _sequential_oracle(oracle, flows, 28 + lo + first, W, w, h, st, acc, avg, ring, pair)
        pair += len(flows)
        for j, i in enumerate(range(first, nb)):
            assert res[i].UPPER == ups[j] and res[i].histsum == sums[j]
            assert np.array_equal(masks[i], rmasks[j]), (lo, i)
        assert np.array_equal(c.window_get().ravel(), avg), lo
    assert np.array_equal(c.accumulator_get(w, h).ravel(), acc)
    # device-resident masks written straight into a caller buffer (the unchecked "direct" path of rc_submit_frames)
    c.close()


def test_window_rearm_and_clip_restart(oracle):
    """rc_window_configure twice with the same W, and a clip restart through rc_flow_configure_batch with unchanged
    parameters: the mean starts from zero buffers each time (main.cpp:1084-1092), no stale ring slot is subtracted."""
    from ripcurrents_b200 import Context, synth
    w, h, B, W = 160, 120, 4, 3
    fr = np.stack(synth.clip(w, h, 9, seed=12))
    P = (0.5, 1, 3, 2, 5, 1.1, 0)
    c = Context(0)
    c.flow_configure_batch(w, h, *P, B); c.hist_reset(); c.window_configure(w, h, W)
    c.process_frames(fr[:1], 0); c.process_frames(fr[1:5], 1); c.process_frames(fr[5:9], 5)      # 8 pairs: window is full
    assert np.abs(c.window_get()).max() > 0

    def check_fresh_run(tag):
        c.process_frames(fr[:1], 0)
        c.process_frames(fr[1:5], 1)
        flows = [c.flow_host_at(3 - i) for i in range(4)]
        avg = np.zeros(h * w * 2, np.float32); ring = np.zeros((W, h * w * 2), np.float32)
        for i, f in enumerate(flows):
            oracle.window_update(avg, ring[i % W], f, W)
        assert np.array_equal(c.window_get().ravel(), avg), tag
        c.process_frames(fr[5:9], 5)
        for i in range(4):
            oracle.window_update(avg, ring[(4 + i) % W], c.flow_host_at(3 - i), W)
        assert np.array_equal(c.window_get().ravel(), avg), tag

    c.window_configure(w, h, W)                     # re-arm: same W, ring kept
    c.flow_configure_batch(w, h, *P, B)             # and restart the clip (the next frame primes)
    assert not c.window_get().any()
    check_fresh_run("re-armed window")
    c.flow_configure_batch(w, h, *P, B)             # restart alone also starts the mean from zero
    assert not c.window_get().any()
    check_fresh_run("restarted clip")
    c.close()


@pytest.mark.parametrize("W,B,nframes", [(100, 12, 133), (1, 4, 14)])
def test_window_of_100_frames_wraps(oracle, W, B, nframes):
    """main.cpp:1505-1515: the W = 100 sliding window of the reference's second driver, through the batched pipeline, long
    enough for the ring to wrap (132 pairs): the mean equals the sequential oracle on the GPU's own flows bit for bit at
    every batch boundary.  W = 1 (the mean IS the latest flow, up to the rounding of avg - old + new) is the degenerate end."""
    from ripcurrents_b200 import Context, synth
    w, h = 96, 64
    fr = np.stack(synth.clip(w, h, nframes, seed=15))
    P = (0.5, 1, 3, 2, 5, 1.1, 0)
    c = Context(0)
    c.flow_configure_batch(w, h, *P, B); c.hist_reset(); c.window_configure(w, h, W)
    c.process_frames(fr[:1], 0)
    avg = np.zeros(h * w * 2, np.float32); ring = np.zeros((W, h * w * 2), np.float32)
    pair = 0
    for lo in range(1, len(fr), B):
        nb = min(B, len(fr) - lo)
        k, _ = c.process_frames(fr[lo:lo + nb], lo)
        assert k == nb
        for i in range(nb):
            oracle.window_update(avg, ring[pair % W], c.flow_host_at(nb - 1 - i), W)
            pair += 1
        assert np.array_equal(c.window_get().ravel(), avg), lo
    assert pair == nframes - 1
    c.close()


def test_sharded_super_blocks_on_gpu(oracle):
    """GpuBackend is re-entrant: a 13-pair clip in super-blocks of 2 'ranks' x 3 pairs (two contexts on this GPU, the
    all-gather / all-reduce done by hand) equals the sequential pipeline frame by frame."""
    from ripcurrents_b200 import Context, sharded, synth
    w, h, n, B = 224, 160, 14, 3
    fr = np.stack(synth.clip(w, h, n, seed=51))
    seq = Context(0); seq.flow_configure_batch(w, h, *P_DEFAULT, 16); seq.hist_reset()
    seq.process_frames(fr[:1], 27)
    _, res = seq.process_frames(fr[1:], 28)
    ups = [r.UPPER for r in res]
    acc_ref = seq.accumulator_get(w, h).ravel(); hist_ref = seq.hist_get()[2]
    backends = [sharded.GpuBackend(Context(0), w, h, P_DEFAULT, B) for _ in range(2)]
    base = np.zeros((37, 50), np.int64)
    got = {}
    for s0 in range(0, n - 1, 2 * B):
        blocks = [(min(s0 + r * B, n - 1), min(s0 + r * B + B, n - 1)) for r in range(2)]
        counts = [b.flows_and_counts(fr[lo:hi + 1]) if hi > lo else np.zeros((0, 37, 50), np.int64)
                  for b, (lo, hi) in zip(backends, blocks)]
        for r, (b, (lo, hi)) in enumerate(zip(backends, blocks)):
            if hi > lo:
                u = b.aggregate(base + sharded.exclusive_prefix_counts(counts, r), [28 + p for p in range(lo, hi)])
                got.update({lo + i: v for i, v in enumerate(u)})
        base = base + sum(c_.sum(0) for c_ in counts)
    assert [got[p] for p in range(n - 1)] == ups
    assert np.array_equal(base, hist_ref)
    assert np.array_equal(backends[0].accumulator() + backends[1].accumulator(), acc_ref)
    for b in backends:
        b.ctx.close()
    seq.close()


def test_average_vector_window_update(oracle):
    """rc_average_vector == averageVector's window update (module:386-400): against the oracle's composition at an odd size
    and, at the reference's compiled 640x480, against the reference's own compiled function (oracle/_ref)."""
    from oracle import ref as R
    from ripcurrents_b200 import Context
    c = Context(0)
    rng = np.random.default_rng(8)
    sizes = [(131, 77)] + ([R.dims()] if R.available() else [])
    for w, h in sizes:
        flow = (rng.normal(0, 0.7, (h, w, 2)) + [0.5, -0.2]).astype(np.float32)
        slot = rng.normal(0, 1, (h, w, 2)).astype(np.float32)
        avg0 = rng.normal(0, 1, (h, w, 2)).astype(np.float32)
        mine = avg0.copy()
        new = c.average_vector(slot, flow, mine, frames=300, dt=2.0, upper=1.1, want_new=True)
        rnew = np.zeros((h * w, 2), np.float32)
        oracle.advect(flow, rnew, 2.0, 1, 1.1, oracle.ADV_GET_DELTA)
        ravg = avg0.copy()
        oracle.window_update(ravg.reshape(-1), slot.copy().reshape(-1), rnew.reshape(-1), 300)
        assert new.tobytes() == rnew.tobytes() and mine.tobytes() == ravg.tobytes()
        assert (new != 0).any() and (new == 0).all(-1).any()
        if (w, h) == (640, 480) and R.available():
            ref_avg = avg0.copy()
            R.average_vector(slot, flow, ref_avg, 1.1)
            assert mine.tobytes() == ref_avg.tobytes()
    c.close()


def test_advection_against_the_compiled_reference():
    """The CUDA advection kernels against the reference's OWN compiled functions (oracle/_ref), bypassing the restatement."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    from ripcurrents_b200 import Context
    c = Context(0)
    rng = np.random.default_rng(77)
    h, w = 120, 200
    flow = rng.normal(0, 1.5, (h, w, 2)).astype(np.float32)
    seeds0 = (rng.random((20000, 2)) * [w + 4, h + 4] - 2).astype(np.float32)
    for variant, dt, it, upper in [(0, 2.0, 3, 0.0), (1, 2.0, 1, 2.5), (2, 0.1, 100, 45.0), (3, 0.3, 7, 0.0), (4, 0.0, 0, 0.0)]:
        a, b = seeds0.copy(), seeds0.copy()
        c.advect(flow, a, dt, it, upper, variant)
        R.advect(flow, b, dt, it, upper, variant)
        assert a.tobytes() == b.tobytes(), variant
    a = np.zeros((h * w, 2), np.float32); da = np.zeros(h * w, np.float32); b, db = a.copy(), da.copy()
    for _ in range(3):
        c.advect(flow, a, 2.0, 1, 1.4, 5, dist=da)
        R.advect(flow, b, 2.0, 1, 1.4, 5, dist=db)
    assert a.tobytes() == b.tobytes() and da.tobytes() == db.tobytes()
    c.close()


def test_failed_launch_is_named():
    """A launch that cannot run is reported by kernel class through rc_last_error (not at some later CUDA call)."""
    from ripcurrents_b200 import Context, capi
    c = Context(0)
    flow = np.zeros((8, 8, 2), np.float32)
    seeds = np.zeros((4, 2), np.float32)
    c.advect(flow, seeds, 1.0, 1, 0.0, 0)            # sanity: a good launch leaves no error behind
    assert c.lib.rc_last_error(c.h).decode() == ""
    with pytest.raises(capi.RcError):
        c.advect(flow, seeds, 1.0, 1, 0.0, 99)       # argument errors are caught before any launch
    c.close()


def test_packed_outmasks_equal_the_u8_masks():
    """rc_set_mask_format(RC_MASK_PACKED): bit p & 7 of byte p >> 3 == (u8 mask == 255), through the host-buffer pipeline."""
    from ripcurrents_b200 import Context, synth
    w, h, n, B = 320, 200, 9, 4
    fr = np.stack(synth.clip(w, h, n, seed=33))
    outs = []
    for packed in (False, True):
        c = Context(0)
        c.flow_configure_batch(w, h, *P_DEFAULT, B); c.hist_reset(); c.window_configure(w, h, 3)
        c.set_mask_format(packed)
        per = w * h // 8 if packed else w * h
        got = []
        masks = np.zeros((B, per), np.uint8)
        for lo in range(0, n, B):
            nb = min(B, n - lo)
            k, res = c.process_frames(fr[lo:lo + nb], 29 + lo, masks[:nb])
            for i in range(nb):
                if res[i].produced:
                    got.append(masks[i].copy())
        outs.append(got)
        c.close()
    assert len(outs[0]) == len(outs[1]) == n - 1
    for u8, pk in zip(*outs):
        assert np.array_equal(np.packbits(u8 == 255, bitorder="little"), pk)
    assert any((m == 0).any() for m in outs[0]) and any((m == 255).any() for m in outs[0])
    c = Context(0)
    c.flow_configure_batch(100, 60, *P_DEFAULT, 2); c.set_mask_format(True)      # 100 % 8 != 0
    from ripcurrents_b200 import capi
    with pytest.raises(capi.RcError):
        c.process_frames(np.zeros((2, 60, 100), np.uint8), 0, np.zeros((2, 750), np.uint8))
    c.close()
