#!/usr/bin/env python
"""Per-kernel device time (rc_profile_*) of one batched step of a named configuration:
    python tools/config_profile.py 4k_box|4k_gauss|gauss10|gauss20|default [B]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ripcurrents_b200 import Context, synth  # noqa: E402

CFG = {"4k_box": (3840, 2160, (0.5, 4, 21, 3, 15, 1.2, 0), 16), "4k_gauss": (3840, 2160, (0.5, 4, 21, 3, 15, 1.2, 256), 16),
       "gauss10": (1920, 1080, (0.5, 2, 10, 3, 15, 1.2, 256), 64), "gauss20": (1920, 1080, (0.5, 2, 20, 3, 15, 1.2, 256), 64),
       "default": (1920, 1080, (0.5, 2, 3, 2, 15, 1.2, 0), 64)}
name = sys.argv[1] if len(sys.argv) > 1 else "4k_box"
W, H, P, B = CFG[name]
if len(sys.argv) > 2:
    B = int(sys.argv[2])
dev = torch.device("cuda", 0)
base = synth.clip(W, H, 5, seed=0)
d = torch.from_numpy(np.stack([base[i % 5] for i in range(2 * B + 1)])).to(dev)
c = Context(0)
c.flow_configure_batch(W, H, *P, B); c.hist_reset(); c.window_configure(W, H, 10)
masks = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
c.process_frames(d.data_ptr(), 30, None, want_results=False, count=1)
for i in range(3):
    c.process_frames(d.data_ptr() + (1 + (i % 2) * B) * W * H, 31, masks.data_ptr(), want_results=False, count=B)
c.synchronize()
c.profile_reset(); c.profile_enable(True)
N = 10
for i in range(N):
    c.process_frames(d.data_ptr() + (1 + (i % 2) * B) * W * H, 31, masks.data_ptr(), want_results=False, count=B)
c.synchronize()
prof = c.profile_read()
tot = sum(v["ms"] for v in prof.values())
print(json.dumps({"config": name, "B": B, "us_per_pair": round(tot * 1e3 / N / B, 1),
                  "kernels_us_per_pair": {k: [round(v["ms"] * 1e3 / N / B, 1), v["launches"] // N, round(v["ms"] / tot, 3)]
                                          for k, v in prof.items()}}))
