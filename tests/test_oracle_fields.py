"""Pins oracle/fields_oracle.c to OpenCV (cv2 4.13.0 fixture + live): the split/magnitude/minMaxLoc/convertTo/
applyColorMap/divide chain and the position scatter of ripcurrents.cpp:231-279 == ripcurrents_module.cpp:13-59."""
import os

import numpy as np
import pytest

from util import GOLDEN


def reference_positions(field):
    """ripcurrents_module.cpp:44-59 restated literally in numpy (no cv2 primitive involved)."""
    h, w, _ = field.shape
    ys, xs = np.mgrid[0:h, 0:w]
    with np.errstate(all="ignore"):
        fx = np.floor(field[..., 0] + xs.astype(np.float32)); fy = np.floor(field[..., 1] + ys.astype(np.float32))
        ok = (fx >= 1) & (fy >= 1) & (fx + 2 <= w) & (fy + 2 <= h)
    d = np.zeros((h, w, 3), np.float32)
    d[fy[ok].astype(int), fx[ok].astype(int)] = 1.0
    return d


def test_fields_golden(oracle):
    z = np.load(os.path.join(GOLDEN, "fields.npz"))
    assert np.array_equal(oracle.jet_lut(), z["jet_lut"])
    sf = oracle.field_magnitude(z["field"])
    assert np.array_equal(sf, z["streamfield"])
    ratio = oracle.divide(sf, z["dist"])
    assert np.array_equal(ratio, z["ratio"])
    for name, src in (("disp", sf), ("motion", z["dist"]), ("ratio", ratio)):
        mx, gray, bgr = oracle.normalize_jet(src)
        assert mx == float(z[name + "_max"])
        assert np.array_equal(gray, z[name + "_gray"]) and np.array_equal(bgr, z[name + "_bgr"]), name


def test_fields_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    f = (rng.standard_normal((120, 161, 2)) * 7).astype(np.float32)
    mine = oracle.field_magnitude(f)
    cv2.setUseOptimized(False)
    try:
        assert np.array_equal(cv2.magnitude(f[..., 0].copy(), f[..., 1].copy()), mine)
    finally:
        cv2.setUseOptimized(True)
    ipp = cv2.magnitude(f[..., 0].copy(), f[..., 1].copy())                 # IPP path: close, not identical
    assert (np.abs(ipp - mine) <= 2 * np.spacing(mine)).all()
    d = (mine + np.abs(rng.standard_normal(mine.shape))).astype(np.float32)
    d.ravel()[::101] = 0
    a = mine.copy(); a.ravel()[::202] = 0                                       # 0/0 and x/0 both present
    with np.errstate(all="ignore"):
        assert np.array_equal(cv2.divide(a, d), oracle.divide(a, d), equal_nan=True)
        z = oracle.divide(a, d, div0_zero=True)
    assert np.isfinite(z).all() and (z[d == 0] == 0).all() and np.array_equal(z[d != 0], (a / np.where(d == 0, 1, d))[d != 0])
    # convertTo edge values: halves round to even, saturation, NaN / inf / >= 2^31 -> 0 (cvtss2si)
    s = np.array([0, 0.5, 1.5, 2.5, 254.5, 255.5, 300, 1e9, 3e9, 1e20, np.inf, np.nan] * 4, np.float32).reshape(1, -1)
    assert np.array_equal(cv2.convertScaleAbs(s, alpha=1.0), oracle.normalize_jet(s, 255.0)[1])
    # maxima: exact without NaNs; NaNs ignored; all-NaN -> NaN
    assert oracle.fmax(mine) == cv2.minMaxLoc(mine)[1]
    t = mine.copy(); t[0, 0] = np.nan; t[5, 7] = np.nan
    assert oracle.fmax(t) == np.nanmax(t)
    assert np.isnan(oracle.fmax(np.full(8, np.nan, np.float32)))
    assert oracle.fmax(np.array([-3, -1, -2], np.float32)) == -1.0


def test_positions(oracle):
    rng = np.random.default_rng(1)
    f = (rng.standard_normal((50, 70, 2)) * 9).astype(np.float32)
    f[3, 3] = (np.nan, 0); f[4, 4] = (np.inf, 1); f[5, 5] = (-np.inf, 1); f[6, 6] = (1e20, 0)
    assert np.array_equal(oracle.positions(f), reference_positions(f))
    keep = np.full((50, 70, 3), 0.25, np.float32)
    out = oracle.positions(f, keep)
    ref = reference_positions(f)
    assert np.array_equal(out, np.where(ref == 1, 1, 0.25).astype(np.float32))
