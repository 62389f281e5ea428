#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: 1080p Farneback flow + aggregation frame pairs / s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload "C2" (BASELINE.json configs[1]): one 1920x1080 camera stream per GPU, reference default Farneback
parameters (ripcurrents.cpp:215: pyr_scale 0.5, levels 2 -> 3 pyramid layers, winsize 3, 2 iterations, poly_n 15,
poly_sigma 1.2, box window), followed per frame by polar conversion + cumulative speed/direction histograms +
tail thresholds + classify/accumulate/mask (ripcurrents.cpp:305-439) + the 10-frame sliding-window flow mean
(main.cpp:1143-1153).  A "step" is FRAMES_PER_STEP consecutive new frames of the stream (= that many frame pairs).

  value : pairs/s summed over all ranks, frames already resident in HBM when the timed region starts
  e2e   : same metric through the C-ABI call with HOST (pinned) frames: per frame one H2D copy of the u8 frame and
          one D2H read of the outmask + thresholds, inside the timed region
  roofline : dominant kernel (largest share of device time), algorithmic bytes / CUDA-event duration
  cpu_baseline / --impl reference : OpenCV's CPU calcOpticalFlowFarneback (cv2, the library the reference calls)
          + the C port of the reference's aggregation (oracle/), on the box's host cores

Multi-GPU: streams are independent: one stream per rank, no data-path collective in the
flow; once per step the per-rank accumulators and histograms are all-reduced (NCCL) into a shared wave-activity map.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 1920, 1080
PARAMS = (0.5, 2, 3, 2, 15, 1.2, 0)      # ripcurrents.cpp:215
WINDOW = 10                               # main.cpp:1084
# frames per step = the context's max_batch (64 = RC_MAX_BATCH): one batched launch sequence per step.  Measured on one
# B200 (pairs/s, device-resident): 8 -> 8.0 k, 16 -> 8.9 k, 32 -> 9.4 k, 64 -> 9.7 k (fewer partial waves per launch); the
# reference reads recorded video, so the two seconds of buffering at 30 fps cost nothing there -- a live stream would
# configure a smaller max_batch.  RC_BENCH_BATCH overrides it (<= 64).
FRAMES_PER_STEP = int(os.environ.get("RC_BENCH_BATCH", "64"))
CLIP_FRAMES = FRAMES_PER_STEP + 1         # distinct synthetic frames per rank, played 0..B,B-1..1 (ping-pong, period 2B =
                                          # two steps) so that consecutive frames always differ by one motion step
METRIC = "1080p Farneback flow+aggregation frame pairs/s"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs per worker for the CPU legs (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configurations (C2-Gaussian, C3, C4)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the frame-pair-sharded single-stream measurement")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle check of the benched configuration")
    return ap.parse_args()


def config_dict(n_gpus):
    return {"workload": "C2: 1920x1080 single stream per GPU, Farneback(0.5,2,3,2,15,1.2,box) -> 3 pyramid layers, "
                        "+ polar/histogram/thresholds/classify+accumulate + window mean W=10",
            "frames_per_step": FRAMES_PER_STEP, "streams": n_gpus, "parallelism": "stream-per-GPU x%d" % n_gpus,
            "l2": "per-step working set (expansion/matrix planes, flow ring) ~0.5 GB per GPU > 126 MB L2; "
                  "frames cycle through a %d-frame clip" % CLIP_FRAMES}


# ------------------------------------------------------------------------------------------------ CPU legs
def _cpu_worker(args):
    """One host process: its own clip, `pairs` frame pairs of flow + aggregation.  Returns seconds of work."""
    seed, pairs, w, h = args
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        cv2 = None
    from oracle import oracle as O
    from ripcurrents_b200 import synth
    fr = synth.clip(w, h, pairs + 1, seed=seed)
    st = O.HistState()
    acc = np.zeros(w * h, np.float32)
    avg = np.zeros(w * h * 2, np.float32)
    ring = np.zeros((WINDOW, w * h * 2), np.float32)
    t0 = time.perf_counter()
    for i in range(pairs):
        if cv2 is not None:
            flow = cv2.calcOpticalFlowFarneback(fr[i], fr[i + 1], None, *PARAMS)
        else:
            flow = O.farneback(fr[i], fr[i + 1], *PARAMS)
        O.histogram(flow, st)
        up, _, _ = O.thresholds(st)
        O.classify_accumulate(flow, up, 31 + i, acc)
        O.window_update(avg, ring[i % WINDOW], flow, WINDOW)
    return time.perf_counter() - t0, cv2 is not None


def cpu_leg(pairs_per_worker, workers, pool=None):
    """Throughput of the CPU path with `workers` independent processes (OpenCV's Farneback is single-threaded,
    so frame-pair parallelism is how a host uses its cores).  Returns (pairs/s, cores, kind, sample)."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(workers)
    t0 = time.perf_counter()
    try:
        res = pool.map(_cpu_worker, [(100 + i, pairs_per_worker, W, H) for i in range(workers)])
    finally:
        if own:
            pool.close(); pool.join()
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    have_cv2 = all(r[1] for r in res)
    total = pairs_per_worker * workers
    kind = "reference" if have_cv2 else "port"
    what = ("cv2 %s calcOpticalFlowFarneback (OpenCV CPU: the routine ripcurrents.cpp:215 calls)" % _cv2_version()
            if have_cv2 else "C port oracle/farneback_oracle.c")
    sample = ("%d pairs of 1080p (%d worker processes x %d pairs), %s + C port of ripcurrents.cpp:305-439 / "
              "main.cpp:1143-1153 aggregation; timed region %.1f s (clip synthesis excluded, wall %.1f s)"
              % (total, workers, pairs_per_worker, what, busy, wall))
    return total / busy, workers, kind, sample


def _cv2_version():
    try:
        import cv2
        return cv2.__version__
    except Exception:
        return "n/a"


def run_reference(args, rank, world):
    if rank != 0:
        return
    import multiprocessing as mp
    workers = min(os.cpu_count() or 1, 32)
    per = args.cpu_pairs or 3
    # W warm-up + K steps: each "step" of this arm is one bounded sample (`per` pairs on each of `workers` processes);
    # the whole run is capped at ~4 minutes
    vals = []
    t_all = time.perf_counter()
    with mp.get_context("spawn").Pool(workers) as pool:
        for s in range(args.warmup + args.steps):
            v, cores, kind, sample = cpu_leg(per, workers, pool)
            if s >= args.warmup:
                vals.append(v)
            if time.perf_counter() - t_all > 200 and len(vals) >= 1:
                break
    value = float(np.mean(vals))
    # a step of this arm is a bounded SAMPLE of the workload (per * workers pairs), not the GPU arm's 64-frame batch:
    # ms_per_step and config.frames_per_step both describe that sample
    cfg = config_dict(args.gpus)
    cfg["frames_per_step"] = per * workers
    cfg["sample_of"] = "%d-frame steps of the GPU arm" % FRAMES_PER_STEP
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * per * workers / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (+f64 accumulators)",
            "data": "synthetic moving texture", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=OUT, flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """nvidia-smi sampling DURING the timed regions (one persistent `-lms 20` process, parsed afterwards)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.3)          # let the first samples arrive before the region starts
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()

    def summary(self, t0=None, t1=None):
        rows = []
        for t, line in self.lines:
            if t0 is not None and not (t0 <= t <= t1):
                continue
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) >= 7:
                rows.append(f)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------ secondary configurations
def survey_bytes_per_pair(w, h, P, layers, aggregation=True):
    """n0 + sum_k (118 + 80 (T - 1)) n_k for the flow, + (25 + 40) n0 for aggregation + window mean."""
    T = P[3]
    n0 = w * h
    nk = sum(lw * lh for lw, lh in layers)
    return n0 + (118 + 80 * (T - 1)) * nk + (65 * n0 if aggregation else 0)


def pyramid_layers(w, h, pyr_scale, levels):
    out, scale = [], 1.0
    k = 0
    while k < levels:
        scale *= pyr_scale
        if w * scale < 32 or h * scale < 32:
            break
        k += 1
    scale = 1.0
    for _ in range(k + 1):
        out.append((int(np.rint(w * scale)), int(np.rint(h * scale))))
        scale *= pyr_scale
    return out


def measure_flow_config(torch, dev, stream, name, ref, w, h, P, B, steps, peak, sampler):
    """One secondary line: flow + aggregation + window mean of a device-resident clip through rc_process_frames."""
    from ripcurrents_b200 import Context, synth
    frames = synth.clip(w, h, B + 1, seed=5)
    order = list(range(B + 1)) + list(range(B - 1, 0, -1))
    d_seq = torch.from_numpy(np.stack([frames[i] for i in order])).to(dev)
    ctx = Context(dev.index)
    ctx.set_stream(stream.cuda_stream)
    ctx.flow_configure_batch(w, h, *P, B)
    ctx.hist_reset()
    ctx.window_configure(w, h, WINDOW)
    NBY = w * h * B
    st = {"s": 0}

    def step():
        s = st["s"]
        ctx.process_frames(d_seq.data_ptr() + (s & 1) * NBY, 31 + s * B, None, want_results=False, count=B)
        st["s"] = s + 1

    step(); step(); step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    ctx.profile_reset(); ctx.profile_enable(True)
    step(); step()
    torch.cuda.synchronize(dev)
    prof = ctx.profile_read(); ctx.profile_enable(False)
    ctx.close()
    del d_seq
    torch.cuda.empty_cache()
    pairs = B * steps
    value = pairs / (ms * 1e-3)
    by = survey_bytes_per_pair(w, h, P, pyramid_layers(w, h, P[0], P[1]))
    by_flow = survey_bytes_per_pair(w, h, P, pyramid_layers(w, h, P[0], P[1]), aggregation=False)
    tot = sum(v["ms"] for v in prof.values()) or 1.0
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
    dgb = dom[1]["bytes"] / (dom[1]["ms"] * 1e-3) / 1e9
    return {"name": name, "reference_call_site": ref, "workload": "%dx%d Farneback%s + aggregation + window mean W=%d" % (w, h, str(P), WINDOW),
            "value": value, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "frames_per_step": B,
            "roofline": {"bound": "hbm", "scope": "whole step on SURVEY 8(d) algorithmic bytes", "bytes_per_pair": by,
                         "achieved": round(by * value / 1e9, 1), "peak": peak, "unit": "GB/s", "frac": round(by * value / 1e9 / peak, 4),
                         "pairs_per_s_at_100pct": round(peak * 1e9 / by, 1),
                         "flow_bytes_only": {"bytes_per_pair": by_flow, "frac": round(by_flow * value / 1e9 / peak, 4),
                                             "pairs_per_s_at_100pct": round(peak * 1e9 / by_flow, 1),
                                             "note": "the same measured rate against the Farneback bytes alone (n0 + sum_k (118 + 80 (T-1)) n_k), "
                                                     "although the timed step also does aggregation and the window mean"}},
            "dominant_kernel": {"name": dom[0], "share": round(dom[1]["ms"] / tot, 4), "avg_us": round(1e3 * dom[1]["ms"] / dom[1]["launches"], 1),
                                "alg_GBps": round(dgb, 1), "frac": round(dgb / peak, 4)},
            "kernel_shares": {k: round(v["ms"] / tot, 4) for k, v in prof.items()},
            "clocks": sampler.summary(t0, t1 + 0.05)}


def measure_advection(torch, dev, stream, steps, peak, sampler):
    """BASELINE configs[3]: per 1080p frame one pathline step of 1 M seeds, the per-pixel particle field and one streakline
    frame of 3 495 emitters x 299 vertices, on a device-resident flow."""
    from ripcurrents_b200 import Context, synth
    fr = np.stack(synth.clip(W, H, 3, seed=0))
    c = Context(dev.index)
    c.set_stream(stream.cuda_stream)
    c.flow_configure_batch(W, H, *PARAMS, 2)
    c.flow_push_batch(fr[:1]); c.flow_push_batch(fr[1:])
    rng = np.random.default_rng(1)
    n = 1 << 20
    seeds = torch.from_numpy((rng.random((n, 2)) * [W - 3, H - 3] + 1).astype(np.float32)).to(dev)
    field = torch.zeros((H * W, 2), device=dev); dist = torch.zeros(H * W, device=dev)
    E, cap = 3495, 300
    em = torch.from_numpy((rng.random((E, 2)) * [W - 3, H - 3] + 1).astype(np.float32)).to(dev)
    verts = torch.zeros((E, cap, 2), device=dev); verts[:, 0] = em
    cnt = torch.full((E,), cap - 1, dtype=torch.int32, device=dev)

    def step():
        c.advect(None, seeds.data_ptr(), 1.0, 1, 0.0, 0, n=n)
        c.advect(None, field.data_ptr(), 2.0, 1, 2.0, 5, dist=dist.data_ptr(), n=W * H)
        c.streakline_step(None, em.data_ptr(), verts.data_ptr(), cnt.data_ptr(), E=E, cap=cap)
        cnt.fill_(cap - 1)

    for _ in range(5):
        step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    c.close()
    # 16 B per seed-step (+8 path length) + the flow field once per frame per kernel
    by = 16 * n + 8 * W * H + (16 + 8) * W * H + 8 * W * H + 16 * E * (cap - 1) + 8 * W * H
    fps = steps / (ms * 1e-3)
    return {"name": "C4 advection", "reference_call_site": "pathlines.cpp:9-46, ripcurrents.cpp:229-231,611-651, Streakline.cpp:22-48",
            "workload": "per 1080p flow: 1 048 576 pathline seeds (dt 1, 1 it.) + per-pixel particle field (2 073 600 px) + "
                        "3 495 streaklines x 299 vertices", "value": fps, "unit": "frames/s", "ms_per_step": ms / steps, "steps": steps,
            "seed_steps_per_s": fps * (n + W * H + E * (cap - 1)),
            "roofline": {"bound": "hbm", "bytes_per_frame": by, "achieved": round(by * fps / 1e9, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(by * fps / 1e9 / peak, 4)},
            "clocks": sampler.summary(t0, t1 + 0.05)}


def measure_sharded_stream(torch, dist, dev, stream, rank, world, steps, frames):
    """ONE 1080p stream split by frame pair over the ranks (BASELINE configs[4] second half): per super-block
    every rank computes 32 pairs (33 device-resident frames, one duplicated frame per block edge); per-frame counts are
    all-gathered for exact thresholds, the window mean is sharded by pixel band (all-to-all of flow bands), accumulators are
    all-reduced at the end.  Strong scaling: the stream is the same at every N, value = pairs of the stream per second."""
    from ripcurrents_b200 import Context, capi
    BS = 32        # pairs per rank and super-block (63 was measured too: 9.29 k at one GPU but only 1.74x at two -- the larger
                   # NCCL send/recv group shares the SMs with the flow kernels for longer; 32: 9.04 k and 1.92x)
    period = 2 * (CLIP_FRAMES - 1)
    order = (list(range(CLIP_FRAMES)) + list(range(CLIP_FRAMES - 2, 0, -1)))
    order = order + order[:BS + 2]
    d_seq = torch.from_numpy(np.stack([frames[i] for i in order])).to(dev)
    ctx = Context(dev.index)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        box = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], rank, world)
    ctx.flow_configure_batch(W, H, *PARAMS, BS + 1)
    ctx.shard_configure(WINDOW, 0)
    ppr = [BS] * world
    st = {"s": 0}

    def step():
        g = (st["s"] * world + rank) * BS                      # first pair of this rank's run
        ctx.shard_step(d_seq.data_ptr() + (g % period) * W * H, 31 + g, ppr, count=BS + 1, want_results=False)
        st["s"] += 1

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    mask, acc, hist = ctx.shard_report(31 + st["s"] * world * BS, want_mask=True, want_acc=True, want_hist=True)
    pairs_total = int(hist.sum()) // (W * H) if hist is not None else 0
    ctx.close()
    del d_seq
    torch.cuda.empty_cache()
    return {"value": steps * world * BS / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "scaling": "strong",
            "ms_per_super_block": ms / steps, "pairs_per_rank_per_super_block": BS, "super_blocks": steps,
            "workload": "ONE 1920x1080 stream, reference default Farneback parameters + histograms/thresholds/classify + window "
                        "mean W=%d, sharded by frame pair" % WINDOW,
            "exchange": "per super-block and rank: ncclAllGather of %d x 1850 u32 per-frame counts; all-to-all (ncclSend/Recv) of the "
                        "row bands of %d flows (16.6 MB each, (N-1)/N of them leave the rank) to the ranks that own those rows of the "
                        "window mean; ncclAllReduce of the accumulators at the reporting point" % (BS, BS),
            "check": {"accumulator_max": float(acc.max()), "mask_calm_fraction": float((mask == 255).mean()),
                      "pairs_counted_lower_bound": pairs_total}}


def copy_roofline(torch, dist, dev, world, h_frames, h_masks, nbytes, steps):
    """What the box allows for the e2e leg's transfers alone: every rank copies the SAME bytes per step as step_e2e
    (nbytes of pinned frames host->device, nbytes of masks device->host) with no kernels -- H2D only, D2H only, and both at
    once on two streams -- all ranks concurrently.  pairs/s at the copy limit = frames per step / time of the 'both' leg."""
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    hf = h_frames.view(-1)[:nbytes]; hm = h_masks.view(-1)[:nbytes]
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    out = {}
    for mode in ("h2d", "d2h", "both"):
        def once():
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s_in):
                    d_in.copy_(hf, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s_out):
                    hm.copy_(d_out, non_blocking=True)
        once(); once()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(main)
        s_in.wait_event(e0); s_out.wait_event(e0)
        for _ in range(steps):
            once()
        main.wait_stream(s_in); main.wait_stream(s_out)
        e1.record(main)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        out[mode] = ms / steps
    gb = nbytes / 1e9
    return {"bytes_each_way_per_step_per_gpu": nbytes, "ranks_copying_concurrently": world,
            "h2d_only_GBps_per_gpu": round(gb / (out["h2d"] * 1e-3), 2), "d2h_only_GBps_per_gpu": round(gb / (out["d2h"] * 1e-3), 2),
            "both_GBps_per_gpu_each_way": round(gb / (out["both"] * 1e-3), 2),
            "aggregate_GBps_each_way": round(world * gb / (out["both"] * 1e-3), 2), "ms_per_step_both": round(out["both"], 3),
            "pairs_per_s_at_copy_limit": round(world * FRAMES_PER_STEP / (out["both"] * 1e-3), 1)}


def measure_cpp_dropin(frames):
    """The SOURCE-COMPATIBLE route: ripcurrents_b200/cpp/demo_main (the reference's legacy frame loop, ripcurrents.cpp:184-439,
    written against the reference's own entry points) -- one frame per call, every intermediate through a host cv::Mat."""
    import tempfile
    from ripcurrents_b200 import build
    demo = os.path.join(os.path.dirname(build.SO), "demo_main")
    if not os.path.exists(demo):
        return {"name": "C++ drop-in loop", "unavailable": "demo_main not built"}
    n = 14
    vals = {}
    with tempfile.TemporaryDirectory() as tmp:
        raw = os.path.join(tmp, "frames.raw")
        np.stack(frames[:n]).tofile(raw)
        for mode in ("--time", "--time-fused"):
            r = subprocess.run([demo, raw, str(W), str(H), str(n), os.path.join(tmp, "out.bin"), mode], capture_output=True,
                               text=True, timeout=300)
            if r.returncode != 0:
                return {"name": "C++ drop-in loop", "unavailable": "demo_main %s failed: %s" % (mode, r.stderr[-200:])}
            vals[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    return {"name": "C++ drop-in loop, frame by frame (cpp/demo_main.cpp = ripcurrents.cpp:184-439)", "unit": UNIT,
            "value": vals["--time"]["pairs_per_s"],
            "source_compatible_route": {"value": vals["--time"]["pairs_per_s"], "pairs": vals["--time"]["pairs"],
                                        "what": "the reference's own call sequence with the reference's signatures: rc::calcOpticalFlowFarneback "
                                                "+ streamline_field_all + flowToPolar + create_histogram + create_flow + "
                                                "create_accumulationbuffer; every intermediate (CV_32FC2 flow, three CV_32FC3 images) crosses the "
                                                "host in a pageable cv::Mat on every call, as in the reference: ~0.4 GB of host copies per "
                                                "1080p frame"},
            "one_call_per_frame_route": {"value": vals["--time-fused"]["pairs_per_s"], "pairs": vals["--time-fused"]["pairs"],
                                         "what": "the same loop with that block replaced by rc_process_frame(frame, framecount, outmask, "
                                                 "&result): pageable host gray frame in, outmask + thresholds out (INTEGRATION.md section 3)"}}

def verify_against_oracle(ctx, frames, order, fps, w, h):
    """Outside every timed region: ONE step of the benched configuration (same context, same batch size -> same kernel
    selection) from a clean temporal state, through the host-buffer API, checked frame by frame against the CPU oracle on
    the flows the GPU produced, plus live OpenCV on two pairs.  Raises on any mismatch."""
    from oracle import oracle as O
    ctx.wait()
    ctx.flow_configure_batch(w, h, *PARAMS, fps)         # restart the clip (same parameters: buffers are kept)
    ctx.hist_reset(); ctx.accumulator_reset(); ctx.window_configure(w, h, WINDOW)
    seq = np.stack([frames[i] for i in order[:fps + 1]])
    masks = np.zeros((fps, h, w), np.uint8)
    ctx.process_frames(seq[:1], 30, masks[:1])
    k, res = ctx.process_frames(seq[1:], 31, masks)
    assert k == fps
    st = O.HistState(); acc = np.zeros(h * w, np.float32)
    avg = np.zeros(h * w * 2, np.float32); ring = np.zeros((WINDOW, h * w * 2), np.float32)
    epe = {}
    try:
        import cv2
    except Exception:
        cv2 = None
    for i in range(fps):
        flow = ctx.flow_host_at(fps - 1 - i)
        if cv2 is not None and i in (0, fps - 1):
            ref = cv2.calcOpticalFlowFarneback(seq[i], seq[i + 1], None, *PARAMS)
            d = np.sqrt(((flow - ref) ** 2).sum(-1))
            cv2.setUseOptimized(False)
            ref2 = cv2.calcOpticalFlowFarneback(seq[i], seq[i + 1], None, *PARAMS)
            cv2.setUseOptimized(True)
            d2 = np.minimum(d, np.sqrt(((flow - ref2) ** 2).sum(-1)))
            epe["pair_%d" % i] = {"mean": float(d.mean()), "max_to_nearer_cv2_path": float(d2.max())}
            if not (d.mean() <= 1e-3 and d2.max() <= 1e-2):
                raise SystemExit("bench.py check FAILED: flow of pair %d differs from OpenCV: %r" % (i, epe))
        O.histogram(flow, st)
        up, _, _ = O.thresholds(st)
        rmask, _, _ = O.classify_accumulate(flow, up, 31 + i, acc)
        O.window_update(avg, ring[i % WINDOW], flow, WINDOW)
        if not (res[i].UPPER == up and res[i].histsum == int(st.histsum[0]) and np.array_equal(masks[i], rmask)):
            raise SystemExit("bench.py check FAILED: aggregation of frame %d differs from the oracle" % i)
    if not (np.array_equal(ctx.window_get().ravel(), avg) and np.array_equal(ctx.accumulator_get(w, h).ravel(), acc)
            and np.array_equal(ctx.hist_get()[2], st.hist2d)):
        raise SystemExit("bench.py check FAILED: window mean / accumulator / histogram differ from the oracle")
    return {"oracle_match": True, "frames_checked": fps, "last_UPPER": float(res[fps - 1].UPPER),
            "histsum": int(res[fps - 1].histsum), "epe_vs_cv2": epe or None,
            "what": "one %d-frame step of the benched configuration re-run from a clean state outside the timed regions: per-frame "
                    "UPPER, histsum and outmask, final window mean, accumulator and histogram bit-equal to oracle/ on the GPU's "
                    "flows; flow of the first and last pair within 1e-3 mean / 1e-2 max px of cv2" % fps}


class DevArr:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": shape, "typestr": typestr, "version": 2}


def bind_to_gpu_numa_node(index):
    """Pins this rank to the CPU cores NVML reports as local to its GPU, so that the pinned host buffers it allocates
    next (first touch) and its copy submissions stay on the GPU's NUMA node.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_ours(args, rank, world, local_rank):
    ncpu_local = bind_to_gpu_numa_node(local_rank)
    import torch
    import torch.distributed as dist
    from ripcurrents_b200 import Context, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # synthetic clip for this rank's camera stream
    frames = synth.clip(W, H, CLIP_FRAMES, seed=rank)
    order = list(range(CLIP_FRAMES)) + list(range(CLIP_FRAMES - 2, 0, -1))          # 2B frames = 2 steps
    assert len(order) == 2 * FRAMES_PER_STEP
    seq = np.stack([frames[i] for i in order])
    h_seq = torch.from_numpy(seq).pin_memory()                                       # [32, H, W] u8, pinned host
    d_seq = h_seq.to(dev)                                                            # same, resident in HBM
    h_masks = [torch.empty((FRAMES_PER_STEP, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
    from ripcurrents_b200 import capi
    h_results = [(capi.FrameResult * FRAMES_PER_STEP)() for _ in range(2)]
    NB = W * H * FRAMES_PER_STEP

    ctx = Context(local_rank)
    # all work of this rank (kernels, NCCL, timing events) goes to ONE explicit stream
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    ctx.flow_configure_batch(W, H, *PARAMS, FRAMES_PER_STEP)
    ctx.hist_reset()
    ctx.window_configure(W, H, WINDOW)

    shared = {}
    if world > 1:
        # the library's own NCCL communicator (rc_comm_init): the unique id travels over torch.distributed, the collective
        # itself is the C ABI's rc_allreduce_accumulators -- what a C++ host would call
        from ripcurrents_b200 import capi as _capi
        ok = torch.ones(1, device=dev)
        try:
            box = [_capi.comm_unique_id() if rank == 0 else None]
        except Exception:       # noqa: BLE001 -- libnccl not loadable by the library: fall back to torch's communicator
            box = [None]
        dist.broadcast_object_list(box, src=0)
        try:
            if box[0] is None:
                raise RuntimeError("no unique id")
            ctx.comm_init(box[0], rank, world)
        except Exception:       # noqa: BLE001
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        shared["capi_nccl"] = bool(ok.item() > 0)

    def allreduce_shared():
        # shared (all-camera) wave-activity map: all-reduce(SUM) of every rank's accumulator and histogram into separate
        # buffers (each camera keeps its own state), enqueued behind the step's kernels and overlapped with the next step
        if "acc_all" not in shared:
            p, aw, ah = ctx.accumulator_device()
            shared["acc"] = torch.as_tensor(DevArr(p, (ah, aw), "<f4"), device=dev)
            shared["hist"] = torch.as_tensor(DevArr(ctx.hist_device(), (37 * 50,), "<i8"), device=dev)
            shared["acc_all"] = torch.empty((ah, aw), dtype=torch.float32, device=dev)
            shared["hist_all"] = torch.empty((37 * 50,), dtype=torch.int64, device=dev)
        if shared["capi_nccl"]:
            ctx.allreduce_accumulators(shared["acc_all"].data_ptr(), shared["hist_all"].data_ptr())
        else:
            shared["acc_all"].copy_(shared["acc"]); shared["hist_all"].copy_(shared["hist"])
            dist.all_reduce(shared["acc_all"]); dist.all_reduce(shared["hist_all"])

    def step_device():
        s = state["step"]
        ctx.process_frames(d_seq.data_ptr() + (s & 1) * NB, 31 + s * FRAMES_PER_STEP, None, want_results=False,
                           count=FRAMES_PER_STEP)
        state["step"] = s + 1
        if world > 1:
            allreduce_shared()

    last = {}

    def step_e2e():
        # public API with HOST buffers: pinned frames in, per-frame masks + threshold records out; the submit is
        # asynchronous (copies of step s+1 / s-1 overlap the kernels of step s), results are complete after wait()
        s = state["step"]
        ctx.process_frames(h_seq.data_ptr() + (s & 1) * NB, 31 + s * FRAMES_PER_STEP, h_masks[s & 1].data_ptr(),
                           count=FRAMES_PER_STEP, submit_only=True, results=h_results[s & 1])
        state["step"] = s + 1
        if world > 1:
            allreduce_shared()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
        ctx.wait()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = ctx.kernel_launches
        e0.record(stream)
        for _ in range(steps):
            step_fn()
        ctx.wait()                 # all device->host copies of the region have landed (no-op for the device leg)
        e1.record(stream)          # behind the last all-reduce, which is enqueued on the same stream
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.kernel_launches - l0

    # prime (the very first frame produces no flow) so that every timed frame completes a pair
    state = {"step": 0}
    step_device(); step_device()

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_region0 = time.perf_counter()
    ms_dev, launches = timed(step_device, args.steps, max(args.warmup, 3))
    ms_e2e, _ = timed(step_e2e, args.steps, max(args.warmup, 3))
    t_region1 = time.perf_counter()
    sampler.stop()

    # the same end-to-end leg with bit-packed outmasks (rc_set_mask_format: the reference's mask holds only 0 / 255), and what
    # the box's host<->device links allow for the u8 leg's transfers alone
    h_masks_packed = [torch.empty((FRAMES_PER_STEP, H * W // 8), dtype=torch.uint8).pin_memory() for _ in range(2)]

    def step_e2e_packed():
        s = state["step"]
        ctx.process_frames(h_seq.data_ptr() + (s & 1) * NB, 31 + s * FRAMES_PER_STEP, h_masks_packed[s & 1].data_ptr(),
                           count=FRAMES_PER_STEP, submit_only=True, results=h_results[s & 1])
        state["step"] = s + 1
        if world > 1:
            allreduce_shared()

    ctx.wait()
    ctx.set_mask_format(True)
    ms_e2e_packed, _ = timed(step_e2e_packed, args.steps, max(args.warmup, 3))
    ctx.set_mask_format(False)
    copy_rf = copy_roofline(torch, dist, dev, world, h_seq, h_masks[0], NB, max(args.steps // 2, 5))

    pairs = FRAMES_PER_STEP * args.steps * world
    value = pairs / (ms_dev * 1e-3)
    e2e_value = pairs / (ms_e2e * 1e-3)

    # per-kernel device time (CUDA events around every launch, same steps, separate pass so that the event
    # overhead does not perturb `value`)
    ctx.profile_reset(); ctx.profile_enable(True)
    for _ in range(min(args.steps, 5)):
        step_device()
    torch.cuda.synchronize(dev)
    prof = ctx.profile_read()
    ctx.profile_enable(False)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kernels = {}
    total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    for name, v in prof.items():
        gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0
        kernels[name] = {"share": round(v["ms"] / total_ms, 4), "avg_us": round(1e3 * v["ms"] / v["launches"], 2),
                         "launches": v["launches"], "alg_GBps": round(gbs, 1), "frac": round(gbs / peak, 4)}
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])[0] if prof else None
    roofline = None
    if dom:
        v = prof[dom]
        achieved = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        roofline = {"kernel": dom, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                    "alg_bytes_per_launch": v["bytes"] / v["launches"], "avg_launch_us": 1e3 * v["ms"] / v["launches"],
                    "share_of_step": round(v["ms"] / total_ms, 4),
                    "note": "algorithmic bytes are those of the kernel AS BUILT (the fused layer: 50 B per pixel and layer); "
                            "SURVEY 8(d) budgets 170 B per pixel and layer for the same work done as one updateMatrices + T "
                            "updateFlow passes (see survey_bytes).  The fused kernel is bound by instruction issue (ncu: "
                            "~70 % of issue slots, ~700 thread-instructions per pixel, DRAM 13 %), not by HBM: DESIGN.md "
                            "section 7"}
        if dom == "flow_layer_fused":
            # the same launches measured against SURVEY.md 8(d)'s per-pixel figure for this step of the path
            # (62 + 80 (T - 1) + 28 bytes per pixel and layer at T iterations, T = 2 here)
            sv = v["bytes"] / 50.0 * (62.0 + 80.0 * (PARAMS[3] - 1) + 28.0)
            roofline["survey_bytes"] = {"per_launch": sv / v["launches"],
                                        "achieved": round(sv / (v["ms"] * 1e-3) / 1e9, 1),
                                        "frac": round(sv / (v["ms"] * 1e-3) / 1e9 / peak, 4)}
        tfile = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu capture
        try:
            roofline["traffic"] = json.load(open(tfile)).get(dom)
        except Exception:
            pass

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic moving texture (ripcurrents_b200/synth.py), %d-frame clip per stream" % CLIP_FRAMES,
                "config": config_dict(world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": FRAMES_PER_STEP * W * H,
                        "d2h_bytes_per_step": FRAMES_PER_STEP * (W * H + 320), "ms_per_step": ms_e2e / args.steps,
                        "api": "rc_submit_frames(%d pinned host frames) -> %d outmasks + threshold records on the host, rc_wait" % (FRAMES_PER_STEP, FRAMES_PER_STEP),
                        "copy_roofline": copy_rf,
                        "frac_of_copy_limit": round(e2e_value / copy_rf["pairs_per_s_at_copy_limit"], 4),
                        "packed_masks": {"value": pairs / (ms_e2e_packed * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_packed / args.steps,
                                         "h2d_bytes_per_step": FRAMES_PER_STEP * W * H,
                                         "d2h_bytes_per_step": FRAMES_PER_STEP * (W * H // 8 + 320),
                                         "api": "rc_set_mask_format(RC_MASK_PACKED): the same call, outmasks as 1 bit per pixel"}},
                "host_binding": "rank pinned to %d GPU-local cores (NVML affinity)" % ncpu_local if ncpu_local else "none",
                "collective": ("rc_allreduce_accumulators (NCCL inside the C ABI, overlapped on a communication stream)"
                               if shared.get("capi_nccl") else ("torch.distributed all_reduce" if world > 1 else None)),
                "gpu_launches": int(launches), "clocks": sampler.summary(t_region0, t_region1 + 0.05), "roofline": roofline, "kernels": kernels,
                "check": None}
        if not args.no_check:
            line["check"] = verify_against_oracle(ctx, frames, order, FRAMES_PER_STEP, W, H)
        if world == 1 and not args.no_secondary:
            sampler2 = ClockSampler(local_rank); sampler2.start()
            sec = []
            def guarded(fn, name, *a):
                try:
                    return fn(*a)
                except Exception as e:       # noqa: BLE001 -- an optional line must not cost the headline
                    return {"name": name, "unavailable": "%s: %s" % (type(e).__name__, e)}

            for name, ref, (ww, hh, PP, BB, st_) in [
                    ("C2 Gaussian winsize 10 (the reference's live driver)", "main.cpp:1119,1481", (W, H, (0.5, 2, 10, 3, 15, 1.2, 256), int(os.environ.get("RC_BENCH_SEC_B", "64")), 8)),
                    ("C2 Gaussian winsize 20", "main.cpp:609,961", (W, H, (0.5, 2, 20, 3, 15, 1.2, 256), int(os.environ.get("RC_BENCH_SEC_B", "64")), 8)),
                    ("C3 4K 5 layers winsize 21 box", "BASELINE configs[2]", (3840, 2160, (0.5, 4, 21, 3, 15, 1.2, 0), int(os.environ.get("RC_BENCH_SEC_B4K", "32")), 8)),
                    ("C3 4K 5 layers winsize 21 Gaussian", "BASELINE configs[2]", (3840, 2160, (0.5, 4, 21, 3, 15, 1.2, 256), int(os.environ.get("RC_BENCH_SEC_B4K", "32")), 8))]:
                sec.append(guarded(measure_flow_config, name, torch, dev, stream, name, ref, ww, hh, PP, BB, st_, peak, sampler2))
            sec.append(guarded(measure_advection, "C4 advection", torch, dev, stream, 50, peak, sampler2))
            sec.append(guarded(measure_cpp_dropin, "C++ drop-in loop", frames))
            sampler2.stop()
            line["secondary"] = sec
        if world == 1 and not args.no_cpu_baseline:
            workers = min(os.cpu_count() or 1, 32)
            v, cores, kind, sample = cpu_leg(args.cpu_pairs or 2, workers)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    ctx.close()
    if not args.no_sharded:
        # optional leg: a failure here (e.g. no usable libnccl) must not cost the headline line
        try:
            sh = measure_sharded_stream(torch, dist, dev, stream, rank, world, 6, frames)
        except Exception as e:       # noqa: BLE001
            sh = {"unavailable": "%s: %s" % (type(e).__name__, e)}
        if rank == 0:
            line["sharded_stream"] = sh
    if rank == 0:
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


OUT = sys.stdout


def main():
    global OUT
    args = parse()
    # the driver reads ONE JSON line from stdout: keep native libraries (NCCL's version banner, ...) off it by pointing fd 1
    # at stderr for the duration of the run and printing the line to a private duplicate of the real stdout
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)]
        cmd += sys.argv[1:]
        raise SystemExit(subprocess.call(cmd, stdout=OUT.fileno()))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
