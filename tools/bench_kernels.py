#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): the hot-path rows that bench.py's headline step does not
time -- particle advection / streaklines (BASELINE.json configs[3]), the large-window Farneback parameter sets of the
reference's other call sites and the 4K configuration (configs[2]), frame ingest and mask clean-up.

Device time comes from the library's per-launch CUDA events (rc_profile_*); "frac" = algorithmic bytes / time / measured
HBM peak (MEASURED_PEAKS.json).  Writes one JSON object per line to stdout.

    python tools/bench_kernels.py > profiles/r01_aux_kernels.jsonl
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ripcurrents_b200 import Context, synth  # noqa: E402

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0


def emit(name, prof, units, unit_name, extra=None):
    ms = sum(v["ms"] for v in prof.values())
    by = sum(v["bytes"] for v in prof.values())
    line = {"bench": name, "ms": round(ms, 4), unit_name + "_per_s": units / (ms * 1e-3), "alg_GBps": round(by / ms / 1e6, 1),
            "frac_of_measured_hbm": round(by / ms / 1e6 / PEAK, 4),
            "kernels": {k: {"ms": round(v["ms"], 4), "launches": v["launches"]} for k, v in prof.items()}}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def timed(c, fn, reps=5):
    fn(); c.synchronize()
    c.profile_reset(); c.profile_enable(True)
    for _ in range(reps):
        fn()
    c.synchronize()
    prof = c.profile_read(); c.profile_enable(False)
    return {k: dict(ms=v["ms"] / reps, launches=v["launches"] // reps, bytes=v["bytes"] / reps) for k, v in prof.items()}


def main():
    import torch
    dev = torch.device("cuda", 0)
    c = Context(0)
    w, h = 1920, 1080
    fr = np.stack(synth.clip(w, h, 3, seed=0))
    P = (0.5, 2, 3, 2, 15, 1.2, 0)
    c.flow_configure_batch(w, h, *P, 2)
    c.flow_push_batch(fr[:1]); c.flow_push_batch(fr[1:])            # a device-resident 1080p flow

    # ---- configs[3]: 1M seeds, pathline step (dt=1, 1 iteration) / module streamline (100 x 0.1) / per-pixel field
    rng = np.random.default_rng(1)
    n = 1 << 20
    seeds = torch.from_numpy((rng.random((n, 2)) * [w - 3, h - 3] + 1).astype(np.float32)).to(dev)
    field = torch.zeros((h * w, 2), device=dev); dist = torch.zeros(h * w, device=dev)
    emit("advect_pathline_1M_seeds_1080p", timed(c, lambda: c.advect(None, seeds.data_ptr(), 1.0, 1, 0.0, 0, n=n)), n, "seed_steps")
    emit("advect_streamline_1M_seeds_100_iterations", timed(c, lambda: c.advect(None, seeds.data_ptr(), 0.1, 100, 45.0, 2, n=n), 3),
         n * 100, "seed_steps")
    emit("advect_field_per_pixel_1080p", timed(c, lambda: c.advect(None, field.data_ptr(), 2.0, 1, 2.0, 5, dist=dist.data_ptr(), n=w * h)),
         w * h, "seed_steps")
    E, cap = 3495, 300
    em = torch.from_numpy((rng.random((E, 2)) * [w - 3, h - 3] + 1).astype(np.float32)).to(dev)
    verts = torch.zeros((E, cap, 2), device=dev); verts[:, 0] = em
    cnt = torch.full((E,), cap - 1, dtype=torch.int32, device=dev)   # steady state: ~1.05M vertices per frame
    emit("streakline_3495_emitters_x_300_vertices", timed(c, lambda: c.streakline_step(None, em.data_ptr(), verts.data_ptr(),
                                                                                      cnt.data_ptr(), E=E, cap=cap)),
         E * (cap - 1), "vertex_steps")

    # ---- derived particle fields of the 1080p state (3 JET images + magnitude + position scatter)
    pf = torch.randn((h * w, 2), device=dev) * 6
    pd = pf.norm(dim=1) + torch.rand(h * w, device=dev) * 5
    sf = torch.empty(h * w, device=dev); dens = torch.empty(h * w * 3, device=dev)
    imgs = [torch.empty(h * w * 3, dtype=torch.uint8, device=dev) for _ in range(3)]
    vp = lambda t: C.c_void_p(t.data_ptr())
    emit("particle_fields_1080p", timed(c, lambda: c._chk(c.lib.rc_particle_fields(
        c.h, vp(pf), vp(pd), C.c_int(w), C.c_int(h), C.c_int(0), vp(sf), vp(imgs[0]), vp(imgs[1]), vp(imgs[2]), vp(dens), None))),
        1, "frames")

    # ---- diagnostics on the resident 1080p flow (device pointers, state kept in the context)
    dflow = torch.randn((h, w, 2), device=dev)
    dimg = torch.zeros((h, w, 3), dtype=torch.uint8, device=dev)
    emit("vector_to_color_1080p", timed(c, lambda: c._chk(c.lib.rc_vector_to_color(
        c.h, vp(dflow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h), vp(dimg), C.c_size_t(w * 3), None, C.c_int(0)))), 1, "frames")
    emit("shear_rate_to_color_1080p", timed(c, lambda: c._chk(c.lib.rc_shear_rate_to_color(
        c.h, vp(dflow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h), vp(dimg), C.c_size_t(w * 3), None, C.c_int(0)))), 1, "frames")
    emit("subtract_mean_magnitude_1080p", timed(c, lambda: c._chk(c.lib.rc_subtract_mean_magnitude(
        c.h, vp(dflow), C.c_size_t(w * 8), C.c_int(w), C.c_int(h), C.c_int(0), None))), 1, "frames")

    # ---- ingest and mask clean-up
    bgr = torch.randint(0, 256, (8, 1080, 1920, 3), dtype=torch.uint8, device=dev)
    gray = torch.empty((8, 480, 640), dtype=torch.uint8, device=dev)

    def ingest():
        for i in range(8):
            c._chk(c.lib.rc_ingest_bgr(c.h, C.c_void_p(bgr[i].data_ptr()), C.c_size_t(1920 * 3), C.c_int(1920), C.c_int(1080),
                                       C.c_void_p(gray[i].data_ptr()), C.c_size_t(640), C.c_int(640), C.c_int(480), C.c_int(0)))
    emit("ingest_bgr_1080p_to_640x480", timed(c, ingest), 8, "frames")
    masks = (torch.rand((16, h, w), device=dev) > 0.95).to(torch.uint8) * 255
    edges = torch.empty_like(masks)
    emit("mask_edges_1080p_x16", timed(c, lambda: c._chk(c.lib.rc_mask_edges(
        c.h, C.c_void_p(masks.data_ptr()), C.c_size_t(w), C.c_size_t(w * h), C.c_int(w), C.c_int(h), C.c_int(16),
        C.c_void_p(edges.data_ptr()), C.c_size_t(w), C.c_size_t(w * h)))), 16, "frames")
    c.close()

    # ---- other Farneback parameter sets of the reference + the 4K configuration (configs[2]); device-resident frames
    for name, (ww, hh, PP, B) in {
            "farneback_1080p_gauss_win10_it3 (main.cpp:1119)": (1920, 1080, (0.5, 2, 10, 3, 15, 1.2, 256), 8),
            "farneback_1080p_gauss_win20_it3 (main.cpp:609)": (1920, 1080, (0.5, 2, 20, 3, 15, 1.2, 256), 8),
            "farneback_4k_5layers_win21_it3_box (configs[2])": (3840, 2160, (0.5, 4, 21, 3, 15, 1.2, 0), 4),
            "farneback_4k_5layers_win21_it3_gauss (configs[2])": (3840, 2160, (0.5, 4, 21, 3, 15, 1.2, 256), 4),
            "farneback_640x480_default (configs[0])": (640, 480, (0.5, 2, 3, 2, 15, 1.2, 0), 16)}.items():
        cc = Context(0)
        frames = torch.from_numpy(np.stack(synth.clip(ww, hh, B + 1, seed=1))).to(dev)
        cc.flow_configure_batch(ww, hh, *PP, B)
        cc.flow_push_batch(frames.data_ptr(), count=1)
        f = lambda: cc.flow_push_batch(frames.data_ptr() + ww * hh, count=B)
        emit(name, timed(cc, f, 3), B, "pairs")
        cc.close()


if __name__ == "__main__":
    main()
