"""Timing + parity of the non-headline Farneback parameter sets (device timing through rc_profile_*)."""
import sys, time, numpy as np
sys.path.insert(0, ".")
from ripcurrents_b200 import Context, synth
import cv2
CASES = [(1920, 1080, (0.5, 2, 10, 3, 15, 1.2, 256), 8), (1920, 1080, (0.5, 2, 20, 3, 15, 1.2, 256), 8),
         (3840, 2160, (0.5, 4, 21, 3, 15, 1.2, 0), 4), (3840, 2160, (0.5, 4, 21, 3, 15, 1.2, 256), 4),
         (640, 480, (0.5, 3, 5, 3, 15, 1.2, 0), 16), (640, 480, (0.5, 2, 3, 2, 15, 1.2, 0), 16)]
for (w, h, P, B) in CASES:
    fr = np.stack(synth.clip(w, h, B + 1, seed=1))
    c = Context(0); c.flow_configure_batch(w, h, *P, B)
    flows = np.empty((B, h, w, 2), np.float32)
    c.flow_push_batch(fr[:1]); c.flow_push_batch(fr[1:], flows=flows)
    ref = cv2.calcOpticalFlowFarneback(fr[0], fr[1], None, *P)
    d = np.sqrt(((flows[0] - ref) ** 2).sum(-1))
    c.flow_configure_batch(w, h, *P, B); c.flow_push_batch(fr[:1])
    c.flow_push_batch(fr[1:]); c.synchronize()
    c.profile_reset(); c.profile_enable(True)
    for _ in range(3): c.flow_push_batch(fr[1:])
    c.synchronize()
    pr = c.profile_read(); c.profile_enable(False)
    tot = sum(v["ms"] for v in pr.values()) / 3
    print((w, h), P, "EPE mean %.2e max %.2e" % (d.mean(), d.max()), "device pairs/s %.0f" % (B / (tot * 1e-3)),
          {k: round(v["ms"] / 3, 2) for k, v in pr.items()})
    c.close()
