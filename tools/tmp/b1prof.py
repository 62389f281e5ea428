import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from ripcurrents_b200 import Context, synth
dev = torch.device('cuda', 0)
W, H = 1920, 1080
fr = torch.from_numpy(np.stack(synth.clip(W, H, 9, seed=0))).to(dev)
for B in (1, 4):
    c = Context(0)
    c.flow_configure_batch(W, H, 0.5, 2, 3, 2, 15, 1.2, 0, B); c.hist_reset(); c.window_configure(W, H, 10)
    masks = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    def step(i):
        c.process_frames(fr.data_ptr() + (i % (8 // B)) * B * W * H, 31 + i * B, masks.data_ptr(), want_results=False, count=B, submit_only=True)
    for i in range(8): step(i)
    c.wait(); c.profile_reset(); c.profile_enable(True)
    n = 16
    for i in range(n): step(i)
    c.wait()
    prof = c.profile_read(); c.profile_enable(False)
    tot = sum(v["ms"] for v in prof.values())
    print("B=%d: kernel time per step %.1f us:" % (B, tot / n * 1e3), {k: round(v["ms"] / n * 1e3, 1) for k, v in prof.items()}, flush=True)
    c.close()
