"""GPU parity of A2-A6 through the C ABI: bit-exact / count-exact against the CPU oracle on identical flow
(ripcurrents.cpp:305-439, main.cpp:1143-1153, ripcurrents_module.cpp:810-863)."""
import os

import numpy as np
import pytest

from util import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctx():
    from ripcurrents_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _flow(rng, h, w, scale=0.8):
    f = (rng.normal(0, scale, (h, w, 2)) + [0.6, 0.2]).astype(np.float32)
    f[0, 0] = (1.0, -1e-7)     # angle == 360 -> direction row 36
    f[0, 1] = (40.0, 0.0)      # bin >= 50: not counted
    f[0, 2] = (0.0, 0.0)
    f[0, 3] = (-0.0, -3.0)
    return f


def test_cart_to_polar_bit_exact(ctx, oracle):
    z = np.load(os.path.join(GOLDEN, "cart_to_polar.npz"))
    mag, ang = ctx.cart_to_polar(np.stack([z["x"], z["y"]], -1))
    assert np.array_equal(mag.view(np.uint32), z["mag"].view(np.uint32))
    assert np.array_equal(ang.view(np.uint32), z["ang"].view(np.uint32))
    rng = np.random.default_rng(1)
    f = rng.normal(0, 3, (500000, 2)).astype(np.float32)
    m2, a2 = oracle.cart_to_polar(f[:, 0], f[:, 1])
    m1, a1 = ctx.cart_to_polar(f)
    assert np.array_equal(m1.view(np.uint32), m2.view(np.uint32))
    assert np.array_equal(a1.view(np.uint32), a2.view(np.uint32))


@pytest.mark.parametrize("shape", [(60, 80), (270, 482), (1080, 1920)])
def test_histogram_thresholds_exact(ctx, oracle, shape):
    rng = np.random.default_rng(shape[0])
    st = oracle.HistState()
    ctx.hist_reset()
    for f in range(3):
        flow = _flow(rng, *shape, scale=0.3 + 0.4 * f)
        oracle.histogram(flow, st)
        ctx.polar_hist(flow)
        hist, histsum, hist2d, histsum2d = ctx.hist_get()
        assert np.array_equal(hist, st.hist) and histsum == int(st.histsum[0])
        assert np.array_equal(hist2d, st.hist2d) and np.array_equal(histsum2d, st.histsum2d)
        assert hist2d[36].sum() >= f + 1      # the planted (1,-1e-7) pixel; random data adds a few more
        up, up2, prop = ctx.thresholds()
        rup, rup2, rprop = oracle.thresholds(st)
        assert up == rup and np.array_equal(up2, rup2) and np.array_equal(prop, rprop, equal_nan=True)


def test_histogram_uniform_motion_and_empty(ctx, oracle):
    # every pixel in ONE bin (worst case for atomics) and the empty-histogram thresholds (0/0 -> NaN as in C)
    up, up2, prop = ctx.thresholds()
    assert up == np.float32(2.45) and np.all(up2 == np.float32(2.45)) and np.all(np.isnan(prop))
    flow = np.zeros((300, 500, 2), np.float32); flow[..., 0] = 1.0; flow[..., 1] = 0.5
    st = oracle.HistState(); oracle.histogram(flow, st)
    ctx.polar_hist(flow)
    hist, histsum, hist2d, _ = ctx.hist_get()
    assert histsum == 150000 and np.array_equal(hist2d, st.hist2d) and (hist2d > 0).sum() == 1
    ctx.hist_add(st.hist2d)
    assert ctx.hist_get()[1] == 300000


def test_classify_accumulate_exact(ctx, oracle):
    rng = np.random.default_rng(5)
    h, w = 135, 241
    acc = np.zeros(h * w, np.float32)
    for fc in (1, 30, 31, 32, 33, 40, 41):
        flow = _flow(rng, h, w, 1.0)
        rmask, rwave, rwater = oracle.classify_accumulate(flow, 0.9, fc, acc)
        mask, wave, water = ctx.classify_accumulate(flow, 0.9, fc)
        assert np.array_equal(mask, rmask) and np.array_equal(wave, rwave) and np.array_equal(water, rwater)
        assert np.array_equal(ctx.accumulator_get(w, h).ravel(), acc)
    ctx.accumulator_reset()
    assert not ctx.accumulator_get(w, h).any()


def test_classify_uses_device_threshold(ctx, oracle):
    rng = np.random.default_rng(6)
    h, w = 90, 160
    flow = _flow(rng, h, w, 0.7)
    st = oracle.HistState(); oracle.histogram(flow, st)
    upper, _, _ = oracle.thresholds(st)
    ctx.hist_reset(); ctx.polar_hist(flow); ctx.thresholds()
    acc = np.zeros(h * w, np.float32)
    rmask, _, rwater = oracle.classify_accumulate(flow, upper, 35, acc)
    mask, _, water = ctx.classify_accumulate(flow, float("nan"), 35)
    assert np.array_equal(mask, rmask) and np.array_equal(water, rwater)
    assert np.array_equal(ctx.accumulator_get(w, h).ravel(), acc)


def test_window_mean_bit_exact(ctx, oracle):
    rng = np.random.default_rng(7)
    h, w, W = 67, 129, 10
    ctx.window_configure(w, h, W)
    avg = np.zeros(h * w * 2, np.float32); ring = np.zeros((W, h * w * 2), np.float32)
    for t in range(23):
        flow = rng.normal(0, 1, (h, w, 2)).astype(np.float32)
        oracle.window_update(avg, ring[t % W], flow, W)
        ctx.window_update(flow)
        assert np.array_equal(ctx.window_get().ravel(), avg), t


def test_subtract_mean(ctx, oracle):
    rng = np.random.default_rng(8)
    flow = (rng.normal(0, 1, (123, 211, 2)) + [0.7, -0.3]).astype(np.float32)
    ref = flow.copy(); rmean = oracle.subtract_mean(ref)
    mean = ctx.subtract_mean(flow)
    assert np.allclose(mean, rmean, rtol=0, atol=1e-12)
    assert np.abs(flow - ref).max() <= 6e-8      # fp64 sum order differs by ~1e-16 -> at most 1 ulp after rounding


def test_process_frame_matches_separate_calls(ctx, oracle):
    """Fused per-frame step == Farneback + histogram + thresholds + classify (+ window) on the same frames."""
    from ripcurrents_b200 import Context, synth
    w, h, W = 320, 240, 4
    fr = synth.clip(w, h, 8, seed=12)
    P = (0.5, 2, 3, 2, 15, 1.2, 0)
    ctx.flow_configure(w, h, *P); ctx.hist_reset(); ctx.window_configure(w, h, W)
    other = Context(0)
    st = oracle.HistState(); acc = np.zeros(h * w, np.float32)
    avg = np.zeros(h * w * 2, np.float32); ring = np.zeros((W, h * w * 2), np.float32)
    mask = np.empty((h, w), np.uint8)
    rc, res = ctx.process_frame(fr[0], 0, mask)
    assert rc == 0 and res.produced == 0
    for i in range(1, 8):
        fc = 28 + i                                   # crosses the framecount > 30 gate
        rc, res = ctx.process_frame(fr[i], fc, mask)
        assert rc == 1 and res.produced == 1
        flow = other.farneback(fr[i - 1], fr[i], *P)
        assert np.array_equal(ctx.flow_host(), flow)
        oracle.histogram(flow, st)
        up, up2, prop = oracle.thresholds(st)
        assert res.UPPER == up and np.array_equal(np.array(res.UPPER2d[:], np.float32), up2)
        assert np.array_equal(np.array(res.prop_above_upper[:], np.float32), prop, equal_nan=True)
        assert res.histsum == int(st.histsum[0])
        rmask, _, _ = oracle.classify_accumulate(flow, up, fc, acc)
        assert np.array_equal(mask, rmask)
        oracle.window_update(avg, ring[(i - 1) % W], flow, W)
        assert np.array_equal(ctx.window_get().ravel(), avg)
    assert np.array_equal(ctx.accumulator_get(w, h).ravel(), acc)
    other.close()


def test_process_frames_batched(ctx, oracle):
    """rc_process_frames over batches == the frame-by-frame reference order: per-frame cumulative histograms and
    thresholds, accumulator gate at framecount > 30, masks, and the W-frame window mean (W < batch and W > batch)."""
    from ripcurrents_b200 import Context, synth
    w, h = 256, 160
    fr = np.stack(synth.clip(w, h, 14, seed=31))
    P = (0.5, 2, 3, 2, 15, 1.2, 0)
    other = Context(0)
    flows = [other.farneback(fr[i], fr[i + 1], *P).copy() for i in range(13)]
    other.close()
    for W, B in [(3, 5), (6, 4)]:
        c = Context(0)
        c.flow_configure_batch(w, h, *P, B)
        c.hist_reset(); c.window_configure(w, h, W)
        st = oracle.HistState(); acc = np.zeros(h * w, np.float32)
        avg = np.zeros(h * w * 2, np.float32); ring = np.zeros((W, h * w * 2), np.float32)
        masks = np.zeros((B, h, w), np.uint8)
        fc0, pair = 25, 0
        for lo in range(0, 14, B):
            hi = min(14, lo + B)
            n, res = c.process_frames(np.ascontiguousarray(fr[lo:hi]), fc0 + lo, masks)
            first = 1 if lo == 0 else 0
            assert n == hi - lo - first
            for i in range(hi - lo):
                if i < first:
                    assert res[i].produced == 0
                    continue
                fc = fc0 + lo + i
                flow = flows[pair]
                oracle.histogram(flow, st)
                up, up2, prop = oracle.thresholds(st)
                assert res[i].produced == 1 and res[i].UPPER == up and res[i].histsum == int(st.histsum[0])
                assert np.array_equal(np.array(res[i].UPPER2d[:], np.float32), up2)
                rmask, _, _ = oracle.classify_accumulate(flow, up, fc, acc)
                assert np.array_equal(masks[i], rmask), (W, B, lo, i)
                oracle.window_update(avg, ring[pair % W], flow, W)
                pair += 1
            assert np.array_equal(c.window_get().ravel(), avg), (W, B, lo)
        assert np.array_equal(c.accumulator_get(w, h).ravel(), acc)
        assert np.array_equal(c.hist_get()[2], st.hist2d)
        c.close()


def test_submit_async_matches_sync(ctx):
    """rc_submit_frames / rc_wait (double-buffered copies on side streams) gives the same masks and thresholds."""
    import ctypes as C
    from ripcurrents_b200 import Context, capi, synth
    w, h, B = 192, 128, 3
    fr = np.stack(synth.clip(w, h, 13, seed=41))
    P = (0.5, 2, 3, 2, 15, 1.2, 0)
    ref = Context(0); ref.flow_configure_batch(w, h, *P, B); ref.hist_reset()
    rmasks = np.zeros((13, h, w), np.uint8); rres = []
    for lo in range(0, 12, B):
        n, res = ref.process_frames(np.ascontiguousarray(fr[lo:lo + B]), 40 + lo, rmasks[lo:lo + B])
        rres += [(r.produced, r.UPPER, r.histsum) for r in res]
    ref.close()
    c = Context(0); c.flow_configure_batch(w, h, *P, B); c.hist_reset()
    masks = np.zeros((13, h, w), np.uint8)
    results = [(capi.FrameResult * B)() for _ in range(4)]
    for k, lo in enumerate(range(0, 12, B)):
        c.process_frames(np.ascontiguousarray(fr[lo:lo + B]), 40 + lo, masks[lo:lo + B], submit_only=True,
                         results=results[k])
    c.wait()
    got = [(r.produced, r.UPPER, r.histsum) for rs in results for r in rs]
    assert got == rres
    assert np.array_equal(masks[1:12], rmasks[1:12])
    c.close()


def test_sharded_two_phase_matches_sequential(oracle):
    """Frame-pair sharding on the C ABI (rc_batch_hist / rc_hist_add / rc_aggregate_last / rc_accumulator_mask): two
    'ranks' (two contexts on this GPU, run one after the other) reproduce the sequential per-frame thresholds, the
    total counts and -- after summing their accumulators, which is what the NCCL all-reduce does -- the final mask."""
    from ripcurrents_b200 import Context, sharded, synth
    w, h, n = 224, 160, 9
    fr = np.stack(synth.clip(w, h, n, seed=51))
    P = (0.5, 2, 3, 2, 15, 1.2, 0)
    seq = Context(0); seq.flow_configure_batch(w, h, *P, 8); seq.hist_reset()
    masks = np.zeros((n, h, w), np.uint8)
    _, res = seq.process_frames(fr[:1], 27, masks[:1]); _, res = seq.process_frames(fr[1:], 28, masks[1:])
    ups = [r.UPPER for r in res]
    acc_ref = seq.accumulator_get(w, h); hist_ref = seq.hist_get()[2]
    blocks = [sharded.block_range(n - 1, 2, r) for r in range(2)]
    backends = [sharded.GpuBackend(Context(0), w, h, P, 8) for _ in range(2)]
    counts = [b.flows_and_counts(fr[lo:hi + 1]) for b, (lo, hi) in zip(backends, blocks)]
    got_ups, acc = [], np.zeros(h * w, np.float32)
    for r, (b, (lo, hi)) in enumerate(zip(backends, blocks)):
        prefix = sharded.exclusive_prefix_counts(counts, r)
        got_ups += b.aggregate(prefix, [28 + p for p in range(lo, hi)])
        acc += b.accumulator()
    assert got_ups == ups
    assert np.array_equal(counts[0].sum(0) + counts[1].sum(0), hist_ref)
    assert np.array_equal(acc.reshape(h, w), acc_ref)
    # reporting-point mask from the combined accumulator == the sequential mask of the last frame
    import ctypes as C
    c = backends[1].ctx
    p, _, _ = c.accumulator_device()
    c.accumulator_reset(); c.synchronize()
    # upload the all-reduced accumulator (what NCCL leaves in place) and ask for the mask
    from ripcurrents_b200 import capi
    capi._memcpy_h2d(p, acc)
    assert np.array_equal(c.accumulator_mask(28 + n - 2), masks[n - 1])
    for b in backends:
        b.ctx.close()
    seq.close()
