import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerance of the task statement (BASELINE.json north_star): flow within
# mean endpoint error <= 1e-3 px and max <= 1e-2 px of the OpenCV reference.
EPE_MEAN_TOL = 1e-3
EPE_MAX_TOL = 1e-2


def epe(a, b):
    d = np.sqrt(((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2).sum(-1))
    return float(d.mean()), float(d.max())


def golden_cases():
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "farneback_*.npz"))):
        out.append(os.path.basename(f)[len("farneback_"):-4])
    return out


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, "farneback_%s.npz" % name))
    P = z["params"]
    params = (float(P[0]), int(P[1]), int(P[2]), int(P[3]), int(P[4]), float(P[5]), int(P[6]))
    return z["frames"], z["flows"], params
