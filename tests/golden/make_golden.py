"""Generates tests/golden/*.npz from the oracle of record: cv2 (OpenCV) -- the library whose
calcOpticalFlowFarneback / cartToPolar the reference calls (ripcurrents.cpp:215,308).

Run here (container with cv2 4.13.0):   python tests/golden/make_golden.py
The fixtures are small (<= 160x128) so that they can be committed; the frames are stored too, so the
tests never depend on the generator being bit-reproducible on another machine.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402

from ripcurrents_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (name, w, h, params) -- params = (pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags)
CASES = [
    ("default_box", 160, 128, (0.5, 2, 3, 2, 15, 1.2, 0)),           # ripcurrents.cpp:215
    ("gauss_win10", 160, 128, (0.5, 2, 10, 3, 15, 1.2, 256)),        # main.cpp:1119,1481
    ("gauss_win20", 160, 128, (0.5, 2, 20, 3, 15, 1.2, 256)),        # main.cpp:609,961
    ("android_box", 160, 128, (0.5, 3, 5, 3, 15, 1.2, 0)),           # RipCurrents_android ripcurrents.cpp:167
    ("odd_size_box", 131, 97, (0.5, 2, 3, 2, 15, 1.2, 0)),           # ragged size, not a multiple of anything
    ("poly5_box", 160, 128, (0.5, 3, 15, 3, 5, 1.1, 0)),
    ("poly7_gauss", 160, 128, (0.5, 4, 21, 3, 7, 1.5, 256)),
    ("scale08_box", 160, 128, (0.8, 3, 7, 2, 7, 1.5, 0)),            # non power-of-two pyramid
]


def main():
    cv2.setNumThreads(1)
    for name, w, h, P in CASES:
        fr = synth.clip(w, h, 3, seed=3, vx=1.3, vy=-0.7, omega=0.004)
        flows = [cv2.calcOpticalFlowFarneback(fr[i], fr[i + 1], None, *P) for i in range(2)]
        np.savez_compressed(os.path.join(HERE, "farneback_%s.npz" % name), frames=np.stack(fr),
                            flows=np.stack(flows), params=np.array(P, np.float64),
                            cv2_version=np.array(cv2.__version__))
    # cartToPolar known answers (degrees), incl. the angle==360 edge and zeros
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.normal(0, 2, 4096), [0, 1, 100, -1, 0, 0, 1e-30, 3, -3],
                        rng.normal(0, 1e-3, 512)]).astype(np.float32)
    y = np.concatenate([rng.normal(0, 2, 4096), [0, -1e-7, -1e-6, 0, 1, -1, 1e-30, 3, -3],
                        rng.normal(0, 1e-3, 512)]).astype(np.float32)
    mag, ang = cv2.cartToPolar(x, y, angleInDegrees=True)
    np.savez_compressed(os.path.join(HERE, "cart_to_polar.npz"), x=x, y=y, mag=mag.ravel(), ang=ang.ravel(),
                        cv2_version=np.array(cv2.__version__))


def ingest():
    # frame ingest: cv2.resize INTER_LINEAR + cvtColor BGR2GRAY, several size ratios
    rng = np.random.default_rng(5)
    out = {}
    for i, (sw, sh, dw, dh) in enumerate([(320, 180, 160, 120), (200, 150, 200, 150), (131, 97, 64, 48), (100, 80, 160, 120)]):
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        g = cv2.cvtColor(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
        out["bgr%d" % i] = img; out["gray%d" % i] = g
    # the PRIMING frame of every loop is resized with INTER_AREA (ripcurrents.cpp:186): fractional, integer and 2x2 ratios
    for i, (sw, sh, dw, dh) in enumerate([(480, 270, 160, 120), (300, 200, 128, 96), (256, 192, 128, 96), (384, 288, 128, 96),
                                          (200, 150, 200, 150)]):
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        g = cv2.cvtColor(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA), cv2.COLOR_BGR2GRAY)
        out["area_bgr%d" % i] = img; out["area_gray%d" % i] = g
    np.savez_compressed(os.path.join(HERE, "ingest.npz"), cv2_version=np.array(cv2.__version__), **out)


def fields():
    # derived particle fields: cv2 with OpenCV's own kernels (setUseOptimized(False): the IPP
    # magnitude differs by <= 2 ulp and cannot be restated) on a synthetic displacement / path-length pair
    cv2.setUseOptimized(False)
    rng = np.random.default_rng(11)
    h, w = 48, 67
    field = (rng.standard_normal((h, w, 2)) * 6).astype(np.float32)
    field[:4] = 0                                                   # particles that never moved
    sf = cv2.magnitude(field[..., 0].copy(), field[..., 1].copy())
    dist = (sf + np.abs(rng.standard_normal((h, w)) * 5)).astype(np.float32)
    dist[:4] = 1.0                                                  # keep the divisor non-zero: NaN-free fixture
    out = {"field": field, "dist": dist, "streamfield": sf}
    ratio = cv2.divide(sf, dist)
    for name, src in (("disp", sf), ("motion", dist), ("ratio", ratio)):
        mx = cv2.minMaxLoc(src)[1]
        gray = cv2.convertScaleAbs(src, alpha=255 / mx)             # == Mat::convertTo(CV_8UC1, 255/max) for src >= 0
        out[name + "_max"] = np.float64(mx); out[name + "_gray"] = gray
        out[name + "_bgr"] = cv2.applyColorMap(gray, cv2.COLORMAP_JET)
    out["ratio"] = ratio
    out["jet_lut"] = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, -1), cv2.COLORMAP_JET).reshape(256, 3)
    cv2.setUseOptimized(True)
    np.savez_compressed(os.path.join(HERE, "fields.npz"), cv2_version=np.array(cv2.__version__), **out)


def diag():
    # 8-bit HSV -> BGR (cvtColor, used by vectorToColor / shearRateToColor): the S = 255 plane both functions produce,
    # plus random triples; rows are multiples of 32 pixels (cv2 rounds the last width % 32 pixels of a row differently)
    hh, vv = np.meshgrid(np.arange(192), np.arange(0, 256, 2), indexing="ij")
    plane = np.stack([hh, np.full_like(hh, 255), vv], -1).astype(np.uint8)
    rng = np.random.default_rng(13)
    rnd = rng.integers(0, 256, (32, 256, 3), dtype=np.uint8)
    out = {"plane_hsv": plane, "plane_bgr": cv2.cvtColor(plane, cv2.COLOR_HSV2BGR),
           "rnd_hsv": rnd, "rnd_bgr": cv2.cvtColor(rnd, cv2.COLOR_HSV2BGR)}
    cv2.setUseOptimized(False)
    out["rnd_bgr_noopt"] = cv2.cvtColor(rnd, cv2.COLOR_HSV2BGR)
    cv2.setUseOptimized(True)
    np.savez_compressed(os.path.join(HERE, "hsv2bgr.npz"), cv2_version=np.array(cv2.__version__), **out)


if __name__ == "__main__":
    main()
    ingest()
    fields()
    diag()
