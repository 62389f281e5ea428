"""Pins oracle/ingest_oracle.c bit-exactly to OpenCV (cv2 4.13.0 fixtures + live): resize(INTER_LINEAR) followed by
cvtColor(BGR2GRAY), the two calls that precede the flow in every frame loop of the reference (ripcurrents.cpp:209-210)."""
import os

import numpy as np
import pytest

from util import GOLDEN


def test_ingest_golden(oracle):
    z = np.load(os.path.join(GOLDEN, "ingest.npz"))
    for i in range(4):
        g = z["gray%d" % i]
        assert np.array_equal(oracle.ingest_bgr(z["bgr%d" % i], g.shape[1], g.shape[0]), g), i


def test_ingest_live(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (270, 480, 3), dtype=np.uint8)
    for dw, dh in [(160, 120), (480, 270), (333, 211), (600, 400), (480, 300)]:
        ref = cv2.cvtColor(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
        assert np.array_equal(oracle.ingest_bgr(img, dw, dh), ref)


def test_ingest_area_golden_and_live(oracle):
    """the PRIMING frame: resize(INTER_AREA) + cvtColor (ripcurrents.cpp:186-187) -- fixtures, then live cv2 at the
    reference's own 1080p -> 640x480 (3 x 2.25), a 2x2, a 3x3 and a mixed integer ratio"""
    z = np.load(os.path.join(GOLDEN, "ingest.npz"))
    for i in range(5):
        g = z["area_gray%d" % i]
        assert np.array_equal(oracle.ingest_bgr_area(z["area_bgr%d" % i], g.shape[1], g.shape[0]), g), i
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for sw, sh, dw, dh in [(1920, 1080, 640, 480), (1280, 960, 640, 480), (960, 720, 320, 240), (1280, 720, 640, 240), (701, 503, 640, 480)]:
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        ref = cv2.cvtColor(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA), cv2.COLOR_BGR2GRAY)
        assert np.array_equal(oracle.ingest_bgr_area(img, dw, dh), ref), (sw, sh, dw, dh)
    with pytest.raises(ValueError):
        oracle.ingest_bgr_area(np.zeros((100, 100, 3), np.uint8), 160, 120)


def test_edges_vs_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for shape, thr in [((120, 160), 0.93), ((33, 47), 0.8), ((64, 64), 0.995)]:
        m = (rng.random(shape) > thr).astype(np.uint8) * 255
        m[0, :5] = 255; m[-1, -3:] = 255
        k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5))
        ref = cv2.morphologyEx(cv2.dilate(m, k), cv2.MORPH_GRADIENT, k)
        assert np.array_equal(oracle.edges(m), ref)
    assert not oracle.edges(np.zeros((20, 30), np.uint8)).any()
    assert not oracle.edges(np.full((20, 30), 255, np.uint8)).any()      # all-wave mask has no edges
