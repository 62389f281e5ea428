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


def cv2_both(cv2, a, b, P):
    """OpenCV's two code paths (setUseOptimized on / off).  They disagree with each other on a handful of
    ill-conditioned pixels per frame, so the max-EPE gate uses the nearer of the two."""
    cv2.setUseOptimized(True)
    r1 = cv2.calcOpticalFlowFarneback(a, b, None, *P)
    cv2.setUseOptimized(False)
    r2 = cv2.calcOpticalFlowFarneback(a, b, None, *P)
    cv2.setUseOptimized(True)
    return r1, r2


def epe_vs_cv2(mine, r1, r2):
    """-> (mean EPE vs the default path, max over pixels of the distance to the nearer OpenCV answer,
           fraction of pixels where OpenCV disagrees with itself by more than 1e-3 px)"""
    d1 = np.sqrt(((mine - r1) ** 2).sum(-1)); d2 = np.sqrt(((mine - r2) ** 2).sum(-1))
    self_d = np.sqrt(((r1 - r2) ** 2).sum(-1))
    return float(d1.mean()), float(np.minimum(d1, d2).max()), float((self_d > 1e-3).mean())
