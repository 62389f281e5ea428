"""Parity on the BASELINE.json configurations that are not the bench line:
  configs[0]  640x480, 60-frame clip, reference default Farneback parameters + temporal aggregation (vs OpenCV CPU)
  configs[3]  1080p pathline / streakline advection, 1M seeds over a flow sequence (bit-exact vs the CPU oracle)
"""
import numpy as np
import pytest

from util import EPE_MAX_TOL, EPE_MEAN_TOL, cv2_both, epe_vs_cv2

pytestmark = pytest.mark.gpu


def test_c1_640x480_60_frames(oracle):
    cv2 = pytest.importorskip("cv2")
    from ripcurrents_b200 import Context, synth
    w, h, n, B, W = 640, 480, 60, 12, 10
    fr = np.stack(synth.clip(w, h, n, seed=0))
    P = (0.5, 2, 3, 2, 15, 1.2, 0)                       # ripcurrents.cpp:215
    c = Context(0)
    c.flow_configure_batch(w, h, *P, B); c.hist_reset(); c.window_configure(w, h, W)
    st = oracle.HistState(); acc = np.zeros(h * w, np.float32)
    avg = np.zeros(h * w * 2, np.float32); ring = np.zeros((W, h * w * 2), np.float32)
    masks = np.zeros((B, h, w), np.uint8)
    pair, worst_mean, worst_max = 0, 0.0, 0.0
    for lo in range(0, n, B):
        k, res = c.process_frames(fr[lo:lo + B], lo, masks)          # framecount = frame index, as in the reference loop
        first = 1 if lo == 0 else 0
        for i in range(first, B):
            back = (B - 1) - i                                        # flow of frame lo+i is `back` pairs before the newest
            flow = c.flow_host_at(back)
            if pair % 7 == 0:                                         # OpenCV CPU on a sample of pairs (1 in 7)
                r1, r2 = cv2_both(cv2, fr[lo + i - 1], fr[lo + i], P)
                mean, mx, amb = epe_vs_cv2(flow, r1, r2)
                assert amb < 1e-4
                worst_mean = max(worst_mean, mean); worst_max = max(worst_max, mx)
            oracle.histogram(flow, st)
            up, up2, prop = oracle.thresholds(st)
            assert res[i].UPPER == up and res[i].histsum == int(st.histsum[0])
            rmask, _, _ = oracle.classify_accumulate(flow, up, lo + i, acc)
            assert np.array_equal(masks[i], rmask), (lo, i)
            oracle.window_update(avg, ring[pair % W], flow, W)
            pair += 1
        assert np.array_equal(c.window_get().ravel(), avg), lo
    assert pair == n - 1
    assert np.array_equal(c.accumulator_get(w, h).ravel(), acc) and acc.sum() > 0
    assert np.array_equal(c.hist_get()[2], st.hist2d)
    assert worst_mean <= EPE_MEAN_TOL and worst_max <= EPE_MAX_TOL, (worst_mean, worst_max)
    c.close()


def test_c4_1m_seeds_over_flow_sequence(oracle):
    from ripcurrents_b200 import Context, synth
    w, h, nframes = 1920, 1080, 7
    fr = np.stack(synth.clip(w, h, nframes, seed=4))
    P = (0.5, 2, 3, 2, 15, 1.2, 0)
    c = Context(0)
    c.flow_configure_batch(w, h, *P, 1)
    rng = np.random.default_rng(1)
    n = 1 << 20
    seeds = (rng.random((n, 2)) * [w - 3, h - 3] + 1).astype(np.float32)           # uniform in [1,W-2) x [1,H-2)
    field = np.zeros((h * w, 2), np.float32); dist = np.zeros(h * w, np.float32)
    E, cap = 3495, nframes + 1
    em = (rng.random((E, 2)) * [w - 3, h - 3] + 1).astype(np.float32)
    verts = np.zeros((E, cap, 2), np.float32); cnt = np.ones(E, np.int32); verts[:, 0] = em
    r_seeds, r_field, r_dist, r_verts, r_cnt = seeds.copy(), field.copy(), dist.copy(), verts.copy(), cnt.copy()
    c.flow_push(fr[0])
    for t in range(1, nframes):
        assert c.flow_push(fr[t]) == 1
        flow = c.flow_host()
        # device-resident flow (flow=None): pathline step dt=1 it=1, per-pixel particle field dt=2 it=1, streaklines
        c.advect(None, seeds, 1.0, 1, 0.0, 0)
        c.advect(None, field, 2.0, 1, 2.0, 5, dist=dist)
        c.streakline_step(None, em, verts, cnt)
        oracle.advect(flow, r_seeds, 1.0, 1, 0.0, oracle.ADV_PATHLINE)
        oracle.advect(flow, r_field, 2.0, 1, 2.0, oracle.ADV_FIELD, dist=r_dist)
        oracle.streakline_step(flow, em, r_verts, r_cnt)
    assert np.array_equal(seeds.view(np.uint32), r_seeds.view(np.uint32))
    assert np.array_equal(field.view(np.uint32), r_field.view(np.uint32))
    assert np.array_equal(dist.view(np.uint32), r_dist.view(np.uint32))
    assert np.array_equal(cnt, r_cnt) and cnt.max() == nframes
    valid = np.arange(cap)[None, :] < cnt[:, None]                                    # slots beyond count are scratch
    assert np.array_equal(verts[valid].view(np.uint32), r_verts[valid].view(np.uint32))
    assert np.isfinite(seeds).all()
    c.close()
