#!/usr/bin/env python
"""Flow error against OpenCV for the current RC_POLY_TRUNC setting (tap truncation of the fast expansion kernel):
golden fixtures (cv2 4.13.0) + live cv2 on 1080p pairs.  Prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ripcurrents_b200 import Context, synth  # noqa: E402
from util import golden_cases, load_golden, cv2_both, epe_vs_cv2  # noqa: E402
import cv2  # noqa: E402

c = Context(0)
out = {"trunc": os.environ.get("RC_POLY_TRUNC", "1e-9"), "golden": {}, "live_1080p": []}
for name in golden_cases():
    frames, flows, P = load_golden(name)
    worst = [0.0, 0.0]
    for i in range(len(flows)):
        f = c.farneback(frames[i], frames[i + 1], *P)
        d = np.sqrt(((f - flows[i]) ** 2).sum(-1))
        worst = [max(worst[0], float(d.mean())), max(worst[1], float(d.max()))]
    out["golden"][name] = {"mean": worst[0], "max": worst[1]}
P = (0.5, 2, 3, 2, 15, 1.2, 0)
for seed in (0, 1, 2):
    fr = synth.clip(1920, 1080, 2, seed=seed)
    f = c.farneback(fr[0], fr[1], *P)
    r1, r2 = cv2_both(cv2, fr[0], fr[1], P)
    m, mx, amb = epe_vs_cv2(f, r1, r2)
    out["live_1080p"].append({"mean": m, "max_to_nearer": mx})
for PP, nm in (((0.5, 2, 10, 3, 15, 1.2, 256), "gauss10"), ((0.5, 2, 20, 3, 15, 1.2, 256), "gauss20")):
    fr = synth.clip(1920, 1080, 2, seed=3)
    f = c.farneback(fr[0], fr[1], *PP)
    r1, r2 = cv2_both(cv2, fr[0], fr[1], PP)
    m, mx, amb = epe_vs_cv2(f, r1, r2)
    out["live_1080p_" + nm] = {"mean": m, "max_to_nearer": mx}
c.close()
print(json.dumps(out))
