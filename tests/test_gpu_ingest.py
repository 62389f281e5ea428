"""GPU parity of the frame ingest: cv::resize(INTER_LINEAR) + cvtColor(BGR2GRAY)
(ripcurrents.cpp:209-210) -- bit-exact against the oracle, the committed cv2 fixtures and live cv2; and the BGR-fed
pipeline equals the gray-fed one."""
import os

import numpy as np
import pytest

from util import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ripcurrents_b200 import Context
    c = Context(0)
    yield c
    c.close()


def test_ingest_golden_and_oracle(ctx, oracle):
    z = np.load(os.path.join(GOLDEN, "ingest.npz"))
    for i in range(4):
        g = z["gray%d" % i]
        assert np.array_equal(ctx.ingest_bgr(z["bgr%d" % i], g.shape[1], g.shape[0]), g), i
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    for dw, dh in [(640, 480), (1920, 1080), (2000, 1200), (333, 977)]:
        assert np.array_equal(ctx.ingest_bgr(img, dw, dh), oracle.ingest_bgr(img, dw, dh)), (dw, dh)
    assert np.array_equal(ctx.ingest_bgr(img, 640, 480, flags=1), oracle.ingest_bgr(img, 640, 480, legacy14=True))


def test_ingest_live_cv2(ctx):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    for dw, dh in [(640, 480), (1280, 720), (1500, 900)]:
        ref = cv2.cvtColor(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
        assert np.array_equal(ctx.ingest_bgr(img, dw, dh), ref), (dw, dh)


def test_bgr_fed_pipeline_equals_gray_fed(oracle):
    from ripcurrents_b200 import Context, capi, synth
    sw, sh, w, h, n, B = 480, 360, 320, 240, 7, 4
    gray_src = np.stack(synth.clip(sw, sh, n, seed=9))
    rng = np.random.default_rng(0)
    bgr = np.stack([gray_src, np.roll(gray_src, 3, 2), 255 - gray_src], -1).astype(np.uint8)    # (n, sh, sw, 3)
    P = (0.5, 2, 3, 2, 15, 1.2, 0)
    gray = np.stack([oracle.ingest_bgr(bgr[i], w, h) for i in range(n)])
    a = Context(0); a.flow_configure_batch(w, h, *P, B); a.hist_reset()
    b = Context(0); b.flow_configure_batch(w, h, *P, B); b.hist_reset()
    ma = np.zeros((n, h, w), np.uint8); mb = np.zeros((n, h, w), np.uint8)
    ra, rb = [], []
    for lo in range(0, n, B):
        hi = min(n, lo + B)
        _, res = a.process_frames(np.ascontiguousarray(gray[lo:hi]), 29 + lo, ma[lo:hi])
        ra += [(r.produced, r.UPPER, r.histsum) for r in res]
        res2 = (capi.FrameResult * (hi - lo))()
        b.submit_frames_bgr(np.ascontiguousarray(bgr[lo:hi]), 29 + lo, mb[lo:hi], res2)
        b.wait()
        rb += [(r.produced, r.UPPER, r.histsum) for r in res2]
    assert ra == rb and np.array_equal(ma[1:], mb[1:])
    assert np.array_equal(a.flow_host(), b.flow_host())
    a.close(); b.close()


def test_mask_edges_bit_exact(ctx, oracle):
    """create_edges (ripcurrents_module.cpp:216-220): batch of masks, ragged size, border pixels set."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(6)
    for shape, thr in [((5, 1080, 1920), 0.97), ((3, 97, 131), 0.8), ((1, 32, 32), 0.5)]:
        m = (rng.random(shape) > thr).astype(np.uint8) * 255
        m[:, 0, :7] = 255; m[:, -1, -2:] = 255; m[:, :, 0] |= m[:, :, 1]
        got = ctx.mask_edges(m)
        k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5))
        for i in range(shape[0]):
            assert np.array_equal(got[i], oracle.edges(m[i])), (shape, i)
            assert np.array_equal(got[i], cv2.morphologyEx(cv2.dilate(m[i], k), cv2.MORPH_GRADIENT, k)), (shape, i)


def test_ingest_area_priming_frame(oracle):
    """RC_INGEST_AREA: resize(INTER_AREA) + cvtColor of the priming frame (ripcurrents.cpp:186-187), bit-exact against the
    cv2 fixtures, live cv2 at the reference's 1080p -> 640x480, and the oracle; then a pipeline whose first frame is ingested
    with INTER_AREA and the rest with INTER_LINEAR equals the pipeline fed with the same gray frames."""
    import os
    from util import GOLDEN
    from ripcurrents_b200 import Context, capi
    c = Context(0)
    z = np.load(os.path.join(GOLDEN, "ingest.npz"))
    for i in range(5):
        g = z["area_gray%d" % i]
        assert np.array_equal(c.ingest_bgr(z["area_bgr%d" % i], g.shape[1], g.shape[0], flags=2), g), i
    rng = np.random.default_rng(4)
    cases = [(1920, 1080, 640, 480), (1280, 960, 640, 480), (960, 720, 320, 240), (1280, 720, 640, 240), (701, 503, 640, 480)]
    try:
        import cv2
    except Exception:
        cv2 = None
    for sw, sh, dw, dh in cases:
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        got = c.ingest_bgr(img, dw, dh, flags=2)
        assert np.array_equal(got, oracle.ingest_bgr_area(img, dw, dh)), (sw, sh, dw, dh)
        if cv2 is not None:
            ref = cv2.cvtColor(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA), cv2.COLOR_BGR2GRAY)
            assert np.array_equal(got, ref), (sw, sh, dw, dh)
    with pytest.raises(capi.RcError):
        c.ingest_bgr(np.zeros((100, 100, 3), np.uint8), 160, 120, flags=2)
    # priming frame AREA + following frames LINEAR through the batched BGR pipeline == gray-fed pipeline
    from ripcurrents_b200 import synth
    P = (0.5, 2, 3, 2, 15, 1.2, 0)
    w, h, sw, sh, n = 160, 120, 400, 270, 5
    big = np.stack(synth.clip(sw, sh, n, seed=6))
    bgr = np.stack([big, big[:, ::-1], big[:, :, ::-1]], -1).copy()
    gray = [oracle.ingest_bgr_area(bgr[0], w, h)] + [oracle.ingest_bgr(bgr[i], w, h) for i in range(1, n)]
    c.flow_configure_batch(w, h, *P, 4); c.hist_reset()
    res1 = (capi.FrameResult * 1)(); res = (capi.FrameResult * 4)()
    masks = np.zeros((4, h, w), np.uint8)
    c.submit_frames_bgr(bgr[:1], 30, None, res1, ingest_flags=2); c.wait()
    c.submit_frames_bgr(bgr[1:], 31, masks, res, ingest_flags=0); c.wait()
    a_hist = c.hist_get()[2].copy(); a_up = [r.UPPER for r in res]
    d = Context(0)
    d.flow_configure_batch(w, h, *P, 4); d.hist_reset()
    masks2 = np.zeros((4, h, w), np.uint8)
    d.process_frames(np.stack(gray[:1]), 30)
    _, res2 = d.process_frames(np.stack(gray[1:]), 31, masks2)
    assert a_up == [r.UPPER for r in res2] and np.array_equal(a_hist, d.hist_get()[2]) and np.array_equal(masks, masks2)
    c.close(); d.close()
