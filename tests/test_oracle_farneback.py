"""Pins the CPU oracle (oracle/farneback_oracle.c) to OpenCV, the library the reference calls at
RipCurrents_main/ripcurrents.cpp:215 -- (a) against committed cv2 4.13.0 fixtures, (b) live against cv2
when it is importable.  The reference itself holds no golden vectors for this path."""
import numpy as np
import pytest

from util import EPE_MAX_TOL, EPE_MEAN_TOL, epe, golden_cases, load_golden


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_golden(oracle, name):
    frames, flows, P = load_golden(name)
    for i in range(flows.shape[0]):
        mine = oracle.farneback(frames[i], frames[i + 1], *P)
        mean, mx = epe(mine, flows[i])
        assert mean <= 2e-5 and mx <= 2e-3, (name, i, mean, mx)   # far inside the task tolerance
        assert mean <= EPE_MEAN_TOL and mx <= EPE_MAX_TOL


def test_layer_selection(oracle):
    # levels = L gives L+1 layers; cropped when a layer would fall under 32 px
    assert oracle.layers(1920, 1080, 0.5, 2) == [(1920, 1080), (960, 540), (480, 270)]
    assert oracle.layers(640, 480, 0.5, 4) == [(640, 480), (320, 240), (160, 120), (80, 60)]
    assert oracle.layers(3840, 2160, 0.5, 4)[-1] == (240, 135)
    assert oracle.layers(40, 40, 0.5, 3) == [(40, 40)]


def test_oracle_live_vs_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    from ripcurrents_b200 import synth
    fr = synth.clip(322, 242, 2, seed=5)
    for P in [(0.5, 2, 3, 2, 15, 1.2, 0), (0.5, 2, 10, 3, 15, 1.2, 256)]:
        ref = cv2.calcOpticalFlowFarneback(fr[0], fr[1], None, *P)
        mean, mx = epe(oracle.farneback(fr[0], fr[1], *P), ref)
        assert mean <= 2e-5 and mx <= 2e-3, (P, mean, mx)


def test_polyexp_constant_and_ramp(oracle):
    # known answers: a constant image has zero derivatives; a ramp I = 2x + 3y has (dy,dx) = (3,2)
    # away from the replicated borders and zero second-order terms.
    h, w, n = 80, 96, 15
    R = oracle.polyexp(np.full((h, w), 7.0, np.float32), n, 1.2)
    assert np.abs(R).max() < 1e-5
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    R = oracle.polyexp(2 * xx + 3 * yy, n, 1.2)[n:-n, n:-n]
    assert np.allclose(R[..., 0], 3, atol=1e-4) and np.allclose(R[..., 1], 2, atol=1e-4)
    assert np.abs(R[..., 2:]).max() < 1e-4
