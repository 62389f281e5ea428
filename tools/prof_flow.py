#!/usr/bin/env python
"""One batched Farneback push of a named configuration, for ncu captures:
    ncu --set full --import-source on -k regex:flow_ -c 12 -o gpurun_out/x python tools/prof_flow.py gauss10 [B]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ripcurrents_b200 import Context, synth  # noqa: E402

CFG = {"default": (1920, 1080, (0.5, 2, 3, 2, 15, 1.2, 0)),
       "gauss10": (1920, 1080, (0.5, 2, 10, 3, 15, 1.2, 256)),
       "gauss20": (1920, 1080, (0.5, 2, 20, 3, 15, 1.2, 256)),
       "box21_4k": (3840, 2160, (0.5, 4, 21, 3, 15, 1.2, 0)),
       "gauss21_4k": (3840, 2160, (0.5, 4, 21, 3, 15, 1.2, 256))}

name = sys.argv[1] if len(sys.argv) > 1 else "gauss10"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
w, h, P = CFG[name]
fr = np.stack(synth.clip(w, h, B + 1, seed=1))
c = Context(0)
c.flow_configure_batch(w, h, *P, B)
c.hist_reset()
c.process_frames(fr[:1], 30, None, want_results=False)
c.process_frames(fr[1:], 31, None, want_results=False)
c.synchronize()
c.close()
print("ok", name, B)
