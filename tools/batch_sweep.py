#!/usr/bin/env python
"""Device-resident pairs/s of the default 1080p pipeline (flow + aggregation + window mean) for small batches:
    python tools/batch_sweep.py            (RC_OVERLAP_MAXPX=0 disables the two-stream expansion overlap)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ripcurrents_b200 import Context, synth  # noqa: E402

W, H, P = 1920, 1080, (0.5, 2, 3, 2, 15, 1.2, 0)
dev = torch.device("cuda", 0)
fr = synth.clip(W, H, 65, seed=0)
out = {}
for B in (1, 2, 4, 8, 16, 64):
    order = list(range(B + 1)) + list(range(B - 1, 0, -1)) if B > 1 else [0, 1]
    d_seq = torch.from_numpy(np.stack([fr[i] for i in order])).to(dev)
    c = Context(0)
    stream = torch.cuda.Stream(dev)
    c.set_stream(stream.cuda_stream)
    c.flow_configure_batch(W, H, *P, B); c.hist_reset(); c.window_configure(W, H, 10)
    masks = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    n_seq = d_seq.shape[0] // B
    st = {"s": 0}

    def step():
        s = st["s"]
        c.process_frames(d_seq.data_ptr() + (s % n_seq) * B * W * H, 31 + s * B, masks.data_ptr(), want_results=False, count=B)
        st["s"] = s + 1

    steps = max(200 // B, 10)
    for _ in range(5):
        step()
    torch.cuda.synchronize(dev)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    out[B] = {"pairs_per_s_device": round(B * steps / (e0.elapsed_time(e1) * 1e-3), 1), "pairs_per_s_wall": round(B * steps / wall, 1)}
    c.close()
print(json.dumps({"overlap_maxpx": os.environ.get("RC_OVERLAP_MAXPX", "default"), "batch": out}))
