"""Deterministic synthetic "moving texture" clips.

base  = uniform noise, Gaussian-blurred (sigma 2.5 px), min-max normalised to 0..255
frame = base warped by a known smooth motion: translation (1.0, 0.5) px/frame plus a slow
        rotation about the centre (so that every direction bin fills), bicubic, re-quantised to u8.

Used by tests/ and bench.py to feed the SAME frames to the CUDA path and to the CPU oracle.
"""
import numpy as np


def _blur(img, sigma):
    import cv2
    return cv2.GaussianBlur(img, (0, 0), sigma, borderType=cv2.BORDER_REFLECT_101)


def base_texture(w, h, seed=0, pad=64, sigma=2.5):
    rng = np.random.default_rng(seed)
    b = rng.random((h + 2 * pad, w + 2 * pad), dtype=np.float32)
    b = _blur(b, sigma)
    b -= b.min()
    b *= 255.0 / max(float(b.max()), 1e-12)
    return b


def frame(base, w, h, t, pad=64, vx=1.0, vy=0.5, omega=0.002):
    """Frame t of the clip: content moves by (vx,vy) px/frame and rotates by omega rad/frame."""
    import cv2
    cx, cy = pad + w * 0.5, pad + h * 0.5
    a = omega * t
    ca, sa = np.cos(a), np.sin(a)
    # dst(x,y) = base(R(-a) * ((x,y)+pad - c - v t) + c)
    tx, ty = vx * t, vy * t
    Minv = np.array([[ca, sa, cx - ca * (cx + tx) - sa * (cy + ty)],
                     [-sa, ca, cy + sa * (cx + tx) - ca * (cy + ty)]], np.float64)
    Minv[:, 2] += Minv[:, 0] * pad + Minv[:, 1] * pad
    out = cv2.warpAffine(base, Minv, (w, h), flags=cv2.INTER_CUBIC | cv2.WARP_INVERSE_MAP,
                         borderMode=cv2.BORDER_REFLECT_101)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def clip(w, h, nframes, seed=0, **kw):
    base = base_texture(w, h, seed)
    return [frame(base, w, h, t, **kw) for t in range(nframes)]
