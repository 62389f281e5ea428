"""Run under a chosen fused-layer kernel (env RC_FLOW_KERNEL, RC_STRIP_MINPX=0 forces it on every launch): checks the
golden fixtures, ragged sizes against the oracle, batched == streaming, and the fused histogram.  Used by
tests/test_gpu_variants.py through a subprocess (the kernel choice is read once per process)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import epe, golden_cases, load_golden  # noqa: E402
from oracle import oracle as O  # noqa: E402
from ripcurrents_b200 import Context, synth  # noqa: E402

c = Context(0)
for name in golden_cases():
    frames, flows, P = load_golden(name)
    for i in range(flows.shape[0]):
        mean, mx = epe(c.farneback(frames[i], frames[i + 1], *P), flows[i])
        assert mean <= 2e-5 and mx <= 3e-3, (name, i, mean, mx)
for (w, h), P in [((131, 97), (0.5, 2, 3, 2, 15, 1.2, 0)), ((64, 48), (0.5, 4, 3, 3, 7, 1.5, 0)), ((257, 129), (0.5, 1, 2, 1, 15, 1.2, 0)),
                  ((190, 130), (0.5, 2, 3, 2, 15, 1.2, 256)), ((29, 300), (0.5, 0, 3, 2, 5, 1.1, 0))]:
    fr = synth.clip(w, h, 2, seed=w)
    mean, mx = epe(c.farneback(fr[0], fr[1], *P), O.farneback(fr[0], fr[1], *P))
    if P[6] & 256 and P[2] == 3:
        # GAUSSIAN winsize 3 (main.cpp:264,742) is ill-conditioned in OpenCV itself: mean only
        assert mean <= 1e-3, ((w, h), P, mean, mx)
    else:
        assert mean <= 1e-5 and mx <= 2e-3, ((w, h), P, mean, mx)
w, h = 300, 200
fr = np.stack(synth.clip(w, h, 10, seed=21))
P = (0.5, 2, 3, 2, 15, 1.2, 0)
single = [c.farneback(fr[i], fr[i + 1], *P).copy() for i in range(9)]
c2 = Context(0); c2.flow_configure_batch(w, h, *P, 4); c2.hist_reset()
st = O.HistState()
masks = np.zeros((4, h, w), np.uint8)
k = 0
for lo in range(0, 10, 4):
    n, res = c2.process_frames(np.ascontiguousarray(fr[lo:lo + 4]), 40 + lo, masks)
    for i in range(len(fr[lo:lo + 4])):
        if lo == 0 and i == 0:
            continue
        O.histogram(single[k], st)
        assert res[i].UPPER == O.thresholds(st)[0] and res[i].histsum == int(st.histsum[0]), (lo, i)
        k += 1
assert np.array_equal(c2.flow_host(), single[8]) and np.array_equal(c2.hist_get()[2], st.hist2d)
print("variant ok", os.environ.get("RC_FLOW_KERNEL", "default"))
