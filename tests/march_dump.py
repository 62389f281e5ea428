"""Writes flows of the winsize > 3 configurations (marching kernel, half-widths 2 / 5 / 10, box and Gaussian) to an .npz
(argv[1]); tests/test_gpu_variants.py runs it with and without RC_MARCH_TMA=0 and requires identical bits (TMA staging of
the structure matrices against per-element cp.async)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ripcurrents_b200 import Context, synth  # noqa: E402

c = Context(0)
out = {}
for w, h, P in [(640, 480, (0.5, 3, 5, 3, 15, 1.2, 0)), (640, 480, (0.5, 2, 10, 3, 15, 1.2, 256)), (960, 540, (0.5, 2, 20, 3, 15, 1.2, 256)),
                (644, 484, (0.5, 2, 21, 3, 15, 1.2, 0)), (1920, 1080, (0.5, 2, 10, 2, 15, 1.2, 256)), (150, 97, (0.5, 1, 10, 3, 7, 1.5, 0))]:
    fr = synth.clip(w, h, 2, seed=w + h)
    out["%dx%d_w%d_%d" % (w, h, P[2], P[6])] = c.farneback(fr[0], fr[1], *P).copy()
np.savez(sys.argv[1], **out)
print("dumped")
