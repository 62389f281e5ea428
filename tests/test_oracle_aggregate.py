"""Pins oracle/aggregate_oracle.c: cartToPolar bit-exact against cv2 fixtures (and live), histogram /
threshold / accumulate logic against a literal numpy re-reading of ripcurrents.cpp:319-439."""
import os

import numpy as np
import pytest

from util import GOLDEN


def test_cart_to_polar_bit_exact_golden(oracle):
    z = np.load(os.path.join(GOLDEN, "cart_to_polar.npz"))
    mag, ang = oracle.cart_to_polar(z["x"], z["y"])
    assert np.array_equal(mag.view(np.uint32), z["mag"].view(np.uint32))
    assert np.array_equal(ang.view(np.uint32), z["ang"].view(np.uint32))
    # the edge the survey found: tiny negative y -> exactly 360.0 -> direction index 36
    m, a = oracle.cart_to_polar(np.float32([1, 100]), np.float32([-1e-7, -1e-6]))
    assert a[0] == np.float32(360.0) and a[1] == np.float32(360.0)


def test_cart_to_polar_live(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    x = rng.normal(0, 3, 200000).astype(np.float32)
    y = rng.normal(0, 3, 200000).astype(np.float32)
    mag, ang = cv2.cartToPolar(x, y, angleInDegrees=True)
    m2, a2 = oracle.cart_to_polar(x, y)
    assert np.array_equal(mag.ravel().view(np.uint32), m2.view(np.uint32))
    assert np.array_equal(ang.ravel().view(np.uint32), a2.view(np.uint32))


def _np_hist(mag, ang):
    hist = np.zeros(50, np.int64); hist2d = np.zeros((37, 50), np.int64)
    b = (mag * np.float32(20)).astype(np.int32)
    d = ((ang * np.float32(36)) / np.float32(360)).astype(np.int32)
    ok = (b < 50) & (b >= 0)
    np.add.at(hist, b[ok], 1)
    np.add.at(hist2d, (d[ok], b[ok]), 1)
    return hist, hist2d


def test_histogram_and_thresholds(oracle):
    rng = np.random.default_rng(3)
    st = oracle.HistState()
    H = np.zeros(50, np.int64); H2 = np.zeros((37, 50), np.int64)
    for f in range(3):
        flow = (rng.normal(0, 0.8, (60, 80, 2)) + [0.6, 0.2]).astype(np.float32)
        flow[0, 0] = (1.0, -1e-7)          # angle == 360 -> overflow row 36
        flow[0, 1] = (40.0, 0.0)           # bin >= 50 -> not counted
        oracle.histogram(flow, st)
        mag, ang = oracle.cart_to_polar(flow[..., 0], flow[..., 1])
        h, h2 = _np_hist(mag, ang)
        H += h; H2 += h2
        assert np.array_equal(st.hist, H) and np.array_equal(st.hist2d, H2)
        assert st.histsum[0] == H.sum() and np.array_equal(st.histsum2d, H2.sum(1))
        assert st.hist2d[36].sum() == f + 1
        # thresholds: literal re-reading of ripcurrents.cpp:333-366
        upper, upper2d, prop = oracle.thresholds(st)
        ts, b = 0, 49
        while ts < H.sum() * .05:
            ts += H[b]; b -= 1
        assert upper == np.float32(b) / np.float32(20)
        for a in range(36):
            t2, bb = 0, 49
            while t2 < H2[a].sum() * .05:
                t2 += H2[a][bb]; bb -= 1
            assert upper2d[a] == max(np.float32(bb) / np.float32(20), np.float32(0.01))
            assert prop[a] == np.float32(H2[a][b + 1:].sum()) / np.float32(ts)


def test_thresholds_empty(oracle):
    st = oracle.HistState()
    upper, upper2d, prop = oracle.thresholds(st)
    assert upper == np.float32(49) / np.float32(20)
    assert np.all(upper2d == np.float32(2.45)) and np.all(np.isnan(prop))   # 0/0 in the reference too


def test_classify_accumulate(oracle):
    rng = np.random.default_rng(5)
    acc = np.zeros(40 * 50, np.float32)
    ref = np.zeros(40 * 50, np.float32)
    for fc in (1, 30, 31, 32, 33, 40):
        flow = rng.normal(0, 1, (40, 50, 2)).astype(np.float32)
        mask, wave, water = oracle.classify_accumulate(flow, 0.9, fc, acc)
        mag, _ = oracle.cart_to_polar(flow[..., 0], flow[..., 1])
        if fc > 30:
            ref += (mag > np.float32(0.9))
        assert np.array_equal(acc, ref)
        val = ref.astype(np.int32)
        m = np.where(val > .1 * fc, 0, 255).astype(np.uint8)
        assert np.array_equal(mask.ravel(), m)
        wc = np.where(val > .1 * fc, np.where(val < .2 * fc, 1, 2), 0)
        assert np.array_equal(wave.ravel(), wc)
        assert np.array_equal(water.ravel() == 3, mag > np.float32(0.9))


def test_window_update(oracle):
    rng = np.random.default_rng(9)
    W, n = 10, 1000
    avg = np.zeros(n, np.float32); ring = np.zeros((W, n), np.float32)
    avg2 = np.zeros(n, np.float32); ring2 = np.zeros((W, n), np.float32)
    inv = np.float32(1.0 / W)
    for t in range(25):
        flow = rng.normal(0, 1, n).astype(np.float32)
        oracle.window_update(avg, ring[t % W], flow, W)
        avg2 = avg2 - ring2[t % W] * inv
        ring2[t % W] = flow
        avg2 = avg2 + ring2[t % W] * inv
        assert np.array_equal(avg, avg2) and np.array_equal(ring, ring2)
    assert np.allclose(avg, ring.mean(0), atol=1e-5)
