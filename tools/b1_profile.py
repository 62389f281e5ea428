#!/usr/bin/env python
"""Per-kernel device time of the default 1080p pipeline at batch size B (default 1): python tools/b1_profile.py [B]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ripcurrents_b200 import Context, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
W, H, P = 1920, 1080, (0.5, 2, 3, 2, 15, 1.2, 0)
dev = torch.device("cuda", 0)
fr = synth.clip(W, H, 2 * B + 1, seed=0)
d = torch.from_numpy(np.stack(fr)).to(dev)
c = Context(0)
c.flow_configure_batch(W, H, *P, B); c.hist_reset(); c.window_configure(W, H, 10)
masks = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
c.process_frames(d.data_ptr(), 30, None, want_results=False, count=1)
for i in range(10):
    c.process_frames(d.data_ptr() + (1 + (i % 2) * B) * W * H, 31, masks.data_ptr(), want_results=False, count=B)
c.synchronize()
c.profile_reset(); c.profile_enable(True)
N = 50
for i in range(N):
    c.process_frames(d.data_ptr() + (1 + (i % 2) * B) * W * H, 31, masks.data_ptr(), want_results=False, count=B)
c.synchronize()
prof = c.profile_read()
print(json.dumps({"B": B, "us_per_step_sum_of_kernels": round(sum(v["ms"] for v in prof.values()) * 1e3 / N, 1),
                  "kernels_us_per_step": {k: [round(v["ms"] * 1e3 / N, 1), v["launches"] // N] for k, v in prof.items()}}))
