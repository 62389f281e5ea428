"""Pins the reference-OWNED half of the oracle (aggregation A3-A5, window mean A6, advection / streaklines A7) to the
reference itself: oracle/_ref/librc_ref.so is compiled by oracle/ref_build.py from the reference's own function
bodies (pathlines.cpp and Streakline.cpp whole; ripcurrents_module.cpp, ripcurrents.cpp and main.cpp sliced by
function name at build time) and every result of oracle/*.c must equal it BIT FOR BIT on random and edge inputs.

The reference compiles XDIM x YDIM = 640 x 480 into several of these loops, so those cases run at that size."""
import numpy as np
import pytest

from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref is not built and /root/reference is absent")


def _flows(rng, n, h, w, scale=1.0):
    """Smooth-ish random flows whose magnitudes spread over the 50 speed bins and all 36 directions."""
    out = []
    for i in range(n):
        f = rng.normal(0, 0.35 * scale, (h, w, 2)).astype(np.float32)
        f += np.float32(0.6 * scale) * np.array([np.cos(0.7 * i), np.sin(0.7 * i)], np.float32)
        f[rng.random((h, w)) < 0.01] *= np.float32(6)       # a tail beyond the last bin
        out.append(f)
    return out


def _polar(oracle, flow):
    mag, ang = oracle.cart_to_polar(flow[..., 0], flow[..., 1])
    h, w, _ = flow.shape
    return np.stack([ang.reshape(h, w), mag.reshape(h, w), mag.reshape(h, w)], -1)


def test_build_reports_reference_geometry():
    assert R.dims() == (640, 480)


@pytest.mark.parametrize("w", [640, 37])
def test_create_histogram_and_thresholds(oracle, w):
    """module:89-144: cumulative counts + the three threshold scans, several frames on the same counters."""
    rng = np.random.default_rng(11 + w)
    h = 480
    st, rst = oracle.HistState(), R.HistState()
    for fi, flow in enumerate(_flows(rng, 4, h, w)):
        if fi == 1:        # the reference's out-of-bounds direction index: angle == 360.0f exactly
            flow[0, 0] = (1.0, -1e-7); flow[0, 1] = (2.0, -1e-7)
        oracle.histogram(flow, st)
        up, up2, prop = oracle.thresholds(st)
        rup, rup2, rprop = R.create_histogram(_polar(oracle, flow), rst)
        assert np.array_equal(st.hist, rst.hist) and int(st.histsum[0]) == int(rst.histsum[0])
        assert np.array_equal(st.hist2d, rst.hist2d) and np.array_equal(st.histsum2d, rst.histsum2d)
        assert np.float32(up).tobytes() == np.float32(rup).tobytes()
        assert up2.tobytes() == rup2.tobytes() and prop.tobytes() == rprop.tobytes()
    assert rst.hist2d[36].sum() == 2          # both edge vectors landed in the extra row


def test_thresholds_on_empty_histogram(oracle):
    """all magnitudes beyond the last bin: histsum == 0, the scans do not move, prop = 0/0"""
    flow = np.full((480, 16, 2), 9.0, np.float32)
    st, rst = oracle.HistState(), R.HistState()
    oracle.histogram(flow, st)
    up, up2, prop = oracle.thresholds(st)
    rup, rup2, rprop = R.create_histogram(_polar(oracle, flow), rst)
    assert int(rst.histsum[0]) == 0 and up == rup
    assert up2.tobytes() == rup2.tobytes() and prop.tobytes() == rprop.tobytes()


def test_classify_accumulate_across_the_frame_gate(oracle):
    """module:153-212 over frames 28..36 (framecount > 30 gate, .1 / .2 * framecount classes)"""
    rng = np.random.default_rng(5)
    h, w = 60, 83
    acc = np.zeros(h * w, np.float32); racc = np.zeros(h * w, np.float32)
    up2 = np.full(36, 0.5, np.float32)
    for fc, flow in zip(range(28, 37), _flows(rng, 9, h, w)):
        upper = 0.55 + 0.05 * (fc % 3)
        m, wv, wt = oracle.classify_accumulate(flow, upper, fc, acc)
        rm, rwv, rwt = R.classify_accumulate(_polar(oracle, flow), upper, fc, racc, up2)
        assert np.array_equal(m, rm) and np.array_equal(wv, rwv) and np.array_equal(wt, rwt)
        assert acc.tobytes() == racc.tobytes()
    assert acc.max() >= 4 and (m == 0).any() and (m == 255).any()


def test_legacy_frame_loop_end_to_end(oracle):
    """ripcurrents.cpp:305-439 run as written in main() (polar -> histogram -> thresholds -> classify -> accumulate ->
    mask) with its cumulative state, against the oracle's composition of the same steps, 34 frames at 640x480."""
    rng = np.random.default_rng(3)
    w, h = R.dims()
    loop = R.LegacyLoop()
    st = oracle.HistState()
    acc = np.zeros(h * w, np.float32)
    base = _flows(rng, 3, h, w)
    for fc in range(1, 35):
        flow = base[fc % 3] * np.float32(1.0 + 0.02 * fc)
        oracle.histogram(flow, st)
        up, up2, prop = oracle.thresholds(st)
        mask, _, _ = oracle.classify_accumulate(flow, up, fc, acc)
        r = loop.frame(flow, fc)
        assert np.float32(up).tobytes() == r["UPPER"].tobytes(), fc
        assert up2.tobytes() == r["UPPER2d"].tobytes() and prop.tobytes() == r["prop"].tobytes(), fc
        assert np.array_equal(st.hist, r["hist"]) and np.array_equal(st.hist2d[:36], r["hist2d"]), fc
        assert int(st.histsum[0]) == int(r["histsum"][0]) and st.hist2d[36].sum() == 0
        assert np.array_equal(mask, r["mask"]), fc
        assert acc.tobytes() == r["acc"].tobytes(), fc
    assert acc.max() == 4.0          # frames 31..34 accumulated
    loop.close()


@pytest.mark.parametrize("W", [10, 3, 100, 1])
def test_window_mean(oracle, W):
    """main.cpp:1143-1153 (W = 10), :1505-1515 (W = 100): avg -= buf[i]/W; buf[i] = flow.clone(); avg += buf[i]/W;
    2.5 windows of updates (W = 1: the mean follows the latest flow)"""
    rng = np.random.default_rng(W)
    h, w = 33, 47
    win = R.Window(w, h, W)
    avg = np.zeros((h, w, 2), np.float32); ring = np.zeros((W, h, w, 2), np.float32)
    for i, flow in enumerate(_flows(rng, 2 * W + W // 2, h, w, 2.0)):
        oracle.window_update(avg.reshape(-1), ring[i % W].reshape(-1), flow, W)
        win.update(flow)
        assert avg.tobytes() == win.avg.tobytes(), i
        assert ring.tobytes() == win.ring.tobytes(), i
    assert int(win.cur[0]) == (2 * W + W // 2) % W


def test_average_vector_window_update(oracle):
    """module:386-400 (BUFFER_FRAME = 300): average -= buffer[i]/300; new = get_delta field (dt 2); average += new/300.
    `buffer` is a by-value vector in the reference: the caller's slot is never rewritten."""
    rng = np.random.default_rng(8)
    w, h = R.dims()
    flow = _flows(rng, 1, h, w, 2.0)[0]
    slot = rng.normal(0, 1, (h, w, 2)).astype(np.float32)
    slot0 = slot.copy()
    avg = rng.normal(0, 1, (h, w, 2)).astype(np.float32)
    ravg = avg.copy()
    R.average_vector(slot, flow, ravg, 1.1)
    assert slot.tobytes() == slot0.tobytes()
    new = np.zeros((h * w, 2), np.float32)
    oracle.advect(flow, new, 2.0, 1, 1.1, oracle.ADV_GET_DELTA)
    scratch = slot0.copy().reshape(-1)
    oracle.window_update(avg.reshape(-1), scratch, new.reshape(-1), 300)
    assert avg.tobytes() == ravg.tobytes()


@pytest.mark.parametrize("variant,dt,it,upper", [(0, 2.0, 3, 0.0), (0, 1.0, 1, 0.0), (1, 2.0, 1, 2.5), (1, 1.5, 4, 1.0),
                                                 (2, 0.1, 100, 45.0), (2, 0.1, 20, 1.5), (3, 0.3, 7, 0.0), (4, 9.0, 1, 0.0)])
def test_streamline_variants(oracle, variant, dt, it, upper):
    """pathlines.cpp:9-46, ripcurrents.cpp:656-698, module:486-606 on seeds inside, on and outside the border"""
    rng = np.random.default_rng(20 + variant)
    h, w = 48, 64
    flow = rng.normal(0, 1.5, (h, w, 2)).astype(np.float32)
    flow[rng.random((h, w)) < 0.02] *= np.float32(5)                      # some steps beyond the r > 5 cut-off
    seeds = (rng.random((4000, 2)) * [w + 4, h + 4] - 2).astype(np.float32)
    seeds[:8] = [(1, 1), (0.999, 5), (w - 2, h - 2), (w - 1.0001, 3), (5, h - 1), (1.5, 1.5), (-0.5, 3), (w, h)]
    a, b = seeds.copy(), seeds.copy()
    oracle.advect(flow, a, dt, it, upper, variant)
    R.advect(flow, b, dt, it, upper, variant)
    assert a.tobytes() == b.tobytes()
    assert (a != seeds).any()


@pytest.mark.parametrize("rv", [5, 7])
def test_streamline_field_and_get_delta(oracle, rv):
    """module:608-679 and the legacy copy ripcurrents.cpp:611-651 (rv = 7): per-pixel particle state + path length"""
    rng = np.random.default_rng(30)
    h, w = 40, 56
    flows = [rng.normal(0, 1.0, (h, w, 2)).astype(np.float32) for _ in range(4)]
    a = np.zeros((h * w, 2), np.float32); da = np.zeros(h * w, np.float32)
    b = a.copy(); db = da.copy()
    for f in flows:
        oracle.advect(f, a, 2.0, 1, 1.4, oracle.ADV_FIELD, dist=da)
        R.advect(f, b, 2.0, 1, 1.4, rv, dist=db)
        assert a.tobytes() == b.tobytes() and da.tobytes() == db.tobytes()
    home = np.stack([rng.integers(0, w, 500), rng.integers(0, h, 500)], -1).astype(np.int32)
    a = rng.normal(0, 2, (500, 2)).astype(np.float32); b = a.copy()
    oracle.advect(flows[0], a, 3.0, 2, 1.4, oracle.ADV_FIELD, dist=None, home=home)
    R.advect(flows[0], b, 3.0, 2, 1.4, rv, home=home)
    assert a.tobytes() == b.tobytes()
    a = rng.normal(0, 2, (500, 2)).astype(np.float32); b = a.copy()
    oracle.advect(flows[1], a, 2.0, 1, 1.4, oracle.ADV_GET_DELTA, home=home)
    R.advect(flows[1], b, 2.0, 1, 1.4, 6, home=home)
    assert a.tobytes() == b.tobytes()


def test_streakline_life_cycle(oracle):
    """Streakline.cpp:11-71 (compiled whole): advect every vertex, reject moves beyond 10 % of the frame, insert the
    generation point at the front -- 40 frames, including emitters in fast regions (rejections) and at the border."""
    rng = np.random.default_rng(40)
    w, h = R.dims()
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    E, cap, nframes = 24, 64, 40
    emit = (rng.random((E, 2)) * [w - 40, h - 40] + 20).astype(np.float32)
    emit[0] = (2.0, 2.0); emit[1] = (w - 3.0, h - 3.0)
    va = np.zeros((E, cap, 2), np.float32); ca = np.ones(E, np.int32)
    va[:, 0] = emit
    vb, cb = va.copy(), ca.copy()
    rejected = 0
    for t in range(nframes):
        flow = np.stack([3.0 * np.sin(yy / 37 + 0.2 * t) + 1.0, 2.0 * np.cos(xx / 53 - 0.1 * t)], -1).astype(np.float32)
        flow[200:260, 300:380] = (70.0, -50.0)                          # |dx| > 64 and |dy| > 48: rejected moves
        before = va.copy()
        oracle.streakline_step(flow, emit, va, ca, 1.0)
        R.streakline_step(flow, emit, vb, cb, 1.0)
        assert np.array_equal(ca, cb) and va.tobytes() == vb.tobytes(), t
        rejected += int(((before[:, :cap - 1] == va[:, 1:]).all(-1) & (np.arange(cap - 1) < ca[:, None] - 1)).sum())
    assert ca.max() == nframes + 1 and rejected > 0
