"""Pins oracle/diag_oracle.c (ripcurrents_module.cpp:900-1138): 8-bit HSV->BGR against cv2 (fixture + exhaustive live),
the three functions against an independent numpy restatement of the reference's loops."""
import os

import numpy as np
import pytest

from util import GOLDEN

f32 = np.float32


def to_uchar(a):
    """`uchar = float` as GCC/x86-64 compiles it: cvttss2si + low byte."""
    a = np.asarray(a, np.float32)
    with np.errstate(all="ignore"):
        ok = (a >= f32(-2147483648.0)) & (a < f32(2147483648.0))
        i = np.where(ok, np.trunc(np.where(ok, a, 0)), -2147483648.0).astype(np.int64)
    return (i & 0xff).astype(np.uint8)


def np_vector_hsv(flow, maxd):
    x, y = flow[..., 0], flow[..., 1]
    with np.errstate(all="ignore"):
        mag = np.sqrt(x * x + y * y)
        theta = (np.arctan2(y, x).astype(np.float64) * 180 / np.pi).astype(np.float32)
        theta = np.where(theta < 0, theta + f32(360), theta)
        hsv = np.stack([to_uchar(theta / f32(2)), np.full(x.shape, 255, np.uint8), to_uchar(mag * f32(255) / f32(maxd))], -1)
    return hsv, float(np.nanmax(np.where(mag > 0, mag, 0)))


def np_shear_hsv(flow, img, maxf, off=10):
    h, w, _ = flow.shape
    c = flow[off:h - off, off:w - off]
    above = flow[0:h - 2 * off, off:w - off]; below = flow[2 * off:h, off:w - off]
    left = flow[off:h - off, 0:w - 2 * off]; right = flow[off:h - off, 2 * off:w]
    j00 = right[..., 0] - left[..., 0]; j01 = above[..., 0] - below[..., 0]
    j10 = right[..., 1] - left[..., 1]; j11 = above[..., 1] - below[..., 1]
    with np.errstate(all="ignore"):
        frob = np.sqrt(((j00 * j00 + j01 * j01) + j10 * j10) + j11 * j11)
        hue = to_uchar(f32(128) - frob * f32(128) / f32(maxf))
    out = img.copy()
    out[off:h - off, off:w - off, 0] = hue; out[off:h - off, off:w - off, 1:] = 255
    return out, float(frob.max()) if frob.size else 0.0


def test_hsv2bgr_golden(oracle):
    z = np.load(os.path.join(GOLDEN, "hsv2bgr.npz"))
    for k in ("plane", "rnd"):
        assert np.array_equal(oracle.hsv2bgr(z[k + "_hsv"]), z[k + "_bgr"]), k
    assert np.array_equal(oracle.hsv2bgr(z["rnd_hsv"], fma=False), z["rnd_bgr_noopt"])


def test_hsv2bgr_exhaustive_live(oracle):
    cv2 = pytest.importorskip("cv2")
    h, s, v = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
    hsv = np.stack([h, s, v], -1).astype(np.uint8).reshape(256, 65536, 3)           # all 2^24 inputs
    assert np.array_equal(oracle.hsv2bgr(hsv), cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR))
    # the row tail cv2 converts with rounding instead of truncation: documented, not reproduced
    px = np.zeros((1, 33, 3), np.uint8); px[:] = (7, 128, 99)
    ref = cv2.cvtColor(px, cv2.COLOR_HSV2BGR)[0]
    assert tuple(ref[0]) == (49, 60, 99) and tuple(oracle.hsv2bgr(px)[0, 32]) == (49, 60, 99)


def flow_field(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    fl = np.stack([np.sin(xx / 17) * 2 + rng.standard_normal((h, w)) * 0.3, np.cos(yy / 11) * 1.5 + rng.standard_normal((h, w)) * 0.3], -1)
    return fl.astype(np.float32)


def test_vector_to_color(oracle):
    cv2 = pytest.importorskip("cv2")
    fl = flow_field(96, 128, 1)
    fl[0, :6] = [(0, 0), (1, 0), (-1, 0), (-1, -0.0), (0, 1), (0, -1)]                 # angles 0, 0, 180, 360-, 90, 270
    fl[1, 0] = (np.nan, 1); fl[1, 1] = (np.inf, 1)
    maxd = 0.0                                                                       # the reference's static starts at 0
    for it in range(3):
        hsv, bgr, newmax = oracle.vector_to_color(fl * f32(1 + 0.4 * it), maxd)
        ref_hsv, ref_max = np_vector_hsv(fl * f32(1 + 0.4 * it), maxd)
        assert np.array_equal(hsv, ref_hsv), it
        assert np.array_equal(bgr, cv2.cvtColor(ref_hsv, cv2.COLOR_HSV2BGR)), it       # width 128: all block pixels
        assert newmax == ref_max or (np.isinf(newmax) and np.isinf(ref_max))
        maxd = 3.0 if it == 0 else newmax                                             # finite from the second frame on
    # values above the previous maximum wrap modulo 256 (the reference's unchecked uchar store)
    hsv, _, _ = oracle.vector_to_color(fl, 1.0)
    big = np.hypot(fl[2:, :, 0], fl[2:, :, 1]) * 255 >= 256
    assert big.any() and (hsv[2:, :, 2][big] < 255).any()


def test_shear_to_color(oracle):
    cv2 = pytest.importorskip("cv2")
    fl = flow_field(96, 128, 2)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    maxf = 0.0
    for it in range(3):
        work = img.copy()
        newmax = oracle.shear_to_color(fl, work, maxf)
        ref_hsv, ref_max = np_shear_hsv(fl, img, maxf)
        assert np.array_equal(work, cv2.cvtColor(ref_hsv, cv2.COLOR_HSV2BGR)), it
        assert newmax == ref_max
        img = work; maxf = newmax if it else 2.5
    small = np.zeros((12, 40, 3), np.uint8) + 77                                      # no interior at all: only the conversion
    assert oracle.shear_to_color(flow_field(12, 40, 0), small, 1.0) == 0.0


def test_subtract_mean_magnitude(oracle):
    fl = flow_field(64, 96, 5); fl[0, 0] = 0
    ref = fl.copy()
    acc = f32(0)
    mags = np.sqrt(ref[..., 0] * ref[..., 0] + ref[..., 1] * ref[..., 1]).ravel()
    for m in mags:
        acc = f32(acc + m)                                                            # sequential fp32, as the loop
    mv = f32(acc / f32(mags.size))
    work = fl.copy()
    assert oracle.subtract_mean_magnitude(work) == mv
    mag = mags.reshape(64, 96)
    with np.errstate(all="ignore"):
        ux = np.where(mag != 0, ref[..., 0] / mag, 0).astype(np.float32); uy = np.where(mag != 0, ref[..., 1] / mag, 0).astype(np.float32)
    assert np.array_equal(work, np.stack([ux * (mag - mv), uy * (mag - mv)], -1))
