"""The drop-in boundary from the reference's side: a C++ frame loop written against the reference's own entry points
(ripcurrents.hpp / Streakline.hpp signatures, ripcurrents_b200/cpp/demo_main.cpp mirrors ripcurrents.cpp:184-439)
runs on the GPU library and must reproduce the CPU oracle's aggregation on the same flows."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cpp_frame_loop_matches_oracle(tmp_path, oracle):
    from ripcurrents_b200 import Context, build, synth
    _, demo = build.build_cpp()
    w, h, n = 320, 240, 7
    fr = np.stack(synth.clip(w, h, n, seed=17))
    raw = tmp_path / "frames.raw"; out = tmp_path / "out.bin"
    fr.tofile(raw)
    subprocess.check_call([demo, str(raw), str(w), str(h), str(n), str(out)], timeout=300)
    rec = np.fromfile(out, np.float32).reshape(n - 1, 13)

    ctx = Context(0)
    P = (0.5, 2, 3, 2, 15, 1.2, 0)
    st = oracle.HistState(); acc = np.zeros(h * w, np.float32)
    disp = np.zeros((h * w, 2), np.float32); dist = np.zeros(h * w, np.float32)
    em = np.array([[w * 0.3, h * 0.4], [w * 0.6, h * 0.5]], np.float32).astype(np.float32)
    em = np.array([[np.float32(w * np.float32(0.3)), np.float32(h * np.float32(0.4))],
                   [np.float32(w * np.float32(0.6)), np.float32(h * np.float32(0.5))]], np.float32)
    cap = n + 1
    verts = np.zeros((2, cap, 2), np.float32); cnt = np.ones(2, np.int32); verts[:, 0] = em
    upper = np.float32(100.0)
    vmax = smax = 0.0; simg = np.zeros((h, w, 3), np.uint8)
    for i in range(1, n):
        flow = ctx.farneback(fr[i - 1], fr[i], *P).copy()
        oracle.advect(flow, disp, 2.0, 1, float(upper), oracle.ADV_FIELD, dist=dist)
        oracle.streakline_step(flow, em, verts, cnt)
        sf = oracle.field_magnitude(disp.reshape(h, w, 2)); d2 = dist.reshape(h, w)
        with np.errstate(all="ignore"):
            fields = [oracle.normalize_jet(sf)[2], oracle.normalize_jet(d2)[2], oracle.normalize_jet(oracle.divide(sf, d2))[2]]
        dens = oracle.positions(disp.reshape(h, w, 2))
        oracle.histogram(flow, st)
        upper, _, _ = oracle.thresholds(st)
        mask, _, _ = oracle.classify_accumulate(flow, upper, i + 28, acc)
        # the reference's int counters exclude nothing: direction-36 pixels are in hist/histsum too
        assert rec[i - 1, 0] == np.float32(upper), i
        assert rec[i - 1, 1] == np.float32(int(st.histsum[0])), i
        assert rec[i - 1, 2] == np.float32(acc.astype(np.float64).sum()), i
        assert rec[i - 1, 3] == np.float32((mask == 255).sum()), i
        assert np.isclose(rec[i - 1, 4], dist.astype(np.float64).sum(), rtol=1e-6), i
        assert rec[i - 1, 5] == np.float32(verts[0, cnt[0] - 1, 0] + verts[1, cnt[1] - 1, 1]), i
        # derived particle fields (ripcurrents.cpp:231-279) from the state before this frame's histogram update
        for k, img in enumerate(fields):
            assert rec[i - 1, 6 + k] == np.float32((img.astype(np.float64) * [1, 2, 3]).sum()), (i, k)
        assert rec[i - 1, 9] == np.float32(dens[..., 2].astype(np.float64).sum()), i
        # flow diagnostics (module:900-1138); the statics start at 0 and carry the previous frame's maxima
        _, vimg, vmax = oracle.vector_to_color(flow, vmax)
        smax = oracle.shear_to_color(flow, simg, smax)
        centred = flow.copy(); oracle.subtract_mean_magnitude(centred)
        w123 = np.array([1, 2, 3])
        assert rec[i - 1, 10] == np.float32((vimg.astype(np.float64) * w123).sum()), i
        assert rec[i - 1, 11] == np.float32((simg.astype(np.float64) * w123).sum()), i
        assert rec[i - 1, 12] == np.float32((centred[..., 0].astype(np.float64) + 2.0 * centred[..., 1]).sum()), i
    ctx.close()
