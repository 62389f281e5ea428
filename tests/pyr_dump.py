"""Writes the flows of a few frame pairs to an .npz (argv[1]); tests/test_gpu_variants.py runs it with and without
RC_PYR=separate and requires identical bits (fused three-layer pyramid kernel against the per-layer kernels)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ripcurrents_b200 import Context, synth  # noqa: E402

c = Context(0)
out = {}
for w, h, P in [(1920, 1080, (0.5, 2, 3, 2, 15, 1.2, 0)), (320, 240, (0.5, 2, 3, 2, 15, 1.2, 0)), (644, 484, (0.5, 4, 5, 1, 7, 1.5, 0)),
                (132, 36, (0.5, 2, 3, 1, 5, 1.1, 0)), (640, 480, (0.5, 2, 3, 2, 15, 1.2, 0x10000))]:
    fr = synth.clip(w, h, 2, seed=w + h)
    out["%dx%d" % (w, h)] = c.farneback(fr[0], fr[1], *P).copy()
np.savez(sys.argv[1], **out)
print("dumped")
