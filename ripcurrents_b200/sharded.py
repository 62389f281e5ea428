"""Host-side logic for ONE camera stream sharded by frame pair across ranks.

Flow of pair (t, t+1) depends only on those two frames, so contiguous blocks of pairs go to different ranks with one
duplicated frame at each block edge and NO collective in the flow path.  The temporal aggregation is sequential in
three places, each solved with a tiny exchange:

  cumulative histograms -> thresholds at frame t need the counts of every frame <= t:
        all_gather of per-frame counts (37*50 int64 per frame) + an exclusive prefix over ranks;
  accumulator (plain sum of 0/1 per pixel, frames > 30): all_reduce(SUM) -- integer-valued, exact in any order;
  outmask: computed from the all-reduced accumulator at the reporting point (block end);
  sliding-window mean (main.cpp:1143-1153): fp32 and order-dependent, but per pixel -- its state is sharded by ROW BAND:
        rank r owns rows [r h / N, (r+1) h / N) and receives that band of every flow of the super-block (BandWindow below;
        the C ABI does this exchange with ncclSend/ncclRecv, csrc/comm.cu).

`backend` abstracts the compute engine so that the same orchestration is exercised on CPU (tests: the oracle, gloo)
and on GPUs (GpuBackend: the C ABI, NCCL).  Only torch.distributed is used for communication.
"""
import numpy as np


def block_range(n_pairs, world, rank):
    """Contiguous block [lo, hi) of pair indices owned by `rank`; the first n_pairs % world ranks get one more."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def exclusive_prefix_counts(all_counts, rank):
    """all_counts: list over ranks of int64 arrays [n_r, 37, 50] (per-frame counts).  Sum of all frames of earlier ranks."""
    out = np.zeros(all_counts[0].shape[1:], np.int64)
    for r in range(rank):
        if all_counts[r].shape[0]:
            out += all_counts[r].sum(0)
    return out


def run_block(backend, frames, lo, hi, framecount_of_pair, dist=None, device=None, base_counts=None):
    """Processes pairs [lo, hi) of `frames` (frame i and i+1 form pair i) on this rank and combines across ranks.

    Every pair that precedes ANY rank's [lo, hi) of this call must already be inside `base_counts` (int64 [37, 50], the
    global cumulative counts of earlier calls; None = nothing earlier): a long stream is processed in super-blocks of
    world * B pairs, rank r taking the r-th run of B (run_stream below).

    backend.flows_and_counts(frames[lo:hi+1])        -> int64 [hi-lo, 37, 50] per-frame counts (flows stay inside)
    backend.aggregate(prefix_counts, framecounts)    -> per-frame UPPER thresholds; prefix_counts = the cumulative counts of
                                                        every frame before this rank's first one (the backend SETS its
                                                        counters to it); adds into the backend's own accumulator
    backend.accumulator()                            -> float32 [h*w] (numpy): this rank's contributions so far
    Returns dict(upper=[...this rank's frames], counts_total=int64[37,50] (all ranks, THIS call),
                 accumulator=float32[h*w] (all ranks, all calls so far)).
    """
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n_local = hi - lo
    counts = backend.flows_and_counts(frames[lo:hi + 1]) if n_local > 0 else np.zeros((0, 37, 50), np.int64)
    if world > 1:
        # ragged all_gather: pad every rank's block to the largest block
        n_all = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(n_all, torch.tensor([n_local], dtype=torch.int64, device=device))
        n_all = [int(t.item()) for t in n_all]
        nmax = max(max(n_all), 1)
        pad = np.zeros((nmax, 37, 50), np.int64)
        pad[:n_local] = counts
        bufs = [torch.zeros((nmax, 37, 50), dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(bufs, torch.from_numpy(pad).to(device) if device is not None else torch.from_numpy(pad))
        all_counts = [b.cpu().numpy()[:n_all[r]] for r, b in enumerate(bufs)]
    else:
        all_counts = [counts]
    prefix = exclusive_prefix_counts(all_counts, rank)
    if base_counts is not None:
        prefix = prefix + np.asarray(base_counts, np.int64)
    uppers = backend.aggregate(prefix, [framecount_of_pair(p) for p in range(lo, hi)]) if n_local > 0 else []
    acc = backend.accumulator()
    if world > 1:
        t = torch.from_numpy(acc.copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t)          # integer-valued sums: exact in any order; the backend keeps its own contributions
        acc = t.cpu().numpy()
    total = sum((c.sum(0) for c in all_counts if c.shape[0]), np.zeros((37, 50), np.int64))
    return {"upper": uppers, "counts_total": total, "accumulator": acc}


def run_stream(backend, frames, B, framecount_of_pair, dist=None, device=None):
    """A whole clip in super-blocks of world * B pairs (rank r takes the r-th run of B pairs of each super-block), carrying
    the global cumulative counts from one super-block to the next.  Returns dict(upper={pair: UPPER} for this rank's
    pairs, counts_total, accumulator (global, after the last super-block))."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n_pairs = len(frames) - 1
    base = np.zeros((37, 50), np.int64)
    uppers, res = {}, None
    for s0 in range(0, n_pairs, world * B):
        lo = min(s0 + rank * B, n_pairs)
        hi = min(lo + B, n_pairs)
        res = run_block(backend, frames, lo, hi, framecount_of_pair, dist, device, base)
        uppers.update({lo + i: u for i, u in enumerate(res["upper"])})
        base = base + res["counts_total"]
    return {"upper": uppers, "counts_total": base, "accumulator": res["accumulator"] if res else backend.accumulator()}


def band_rows(h, world, rank):
    """Rows [lo, hi) of the window mean owned by `rank` (same split as csrc/comm.cu: r * h // world)."""
    return rank * h // world, (rank + 1) * h // world


class BandWindow:
    """This rank's row band of the sliding-window flow mean.  update() is called once per super-block with the band of
    EVERY flow of the super-block, in stream order; `step(avg, slot, flow, W)` is the per-flow arithmetic of
    main.cpp:1143-1153 (tests pass the oracle's window_update)."""

    def __init__(self, w, h, W, world, rank, step):
        self.lo, self.hi = band_rows(h, world, rank)
        self.w, self.W, self.step = w, W, step
        n = (self.hi - self.lo) * w * 2
        self.avg = np.zeros(n, np.float32)
        self.ring = np.zeros((W, n), np.float32)
        self.count = 0

    def update(self, flow_bands):
        for f in flow_bands:
            self.step(self.avg, self.ring[self.count % self.W], np.ascontiguousarray(f, np.float32).reshape(-1), self.W)
            self.count += 1


def exchange_bands(flows, h, dist=None, device=None):
    """flows: this rank's [n_local, h, w, 2] float32.  Returns the list, in stream order, of THIS rank's row band of every
    flow of the super-block (all ranks).  torch.distributed has no ragged all-to-all on gloo, so the host-side statement
    gathers whole flows and slices; the C ABI sends only the bands."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    lo, hi = band_rows(h, world, rank)
    if world == 1:
        return [f[lo:hi] for f in flows]
    n_all = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(n_all, torch.tensor([len(flows)], dtype=torch.int64, device=device))
    n_all = [int(t.item()) for t in n_all]
    nmax = max(max(n_all), 1)
    shape = (nmax,) + tuple(flows.shape[1:]) if len(flows) else None
    shp = [torch.zeros(3, dtype=torch.int64, device=device) for _ in range(world)]
    mine = torch.tensor(list(flows.shape[1:]) if len(flows) else [0, 0, 0], dtype=torch.int64, device=device)
    dist.all_gather(shp, mine)
    hw2 = max((tuple(int(v) for v in t.tolist()) for t in shp), key=lambda t: t[0])
    pad = np.zeros((nmax,) + hw2, np.float32)
    if len(flows):
        pad[:len(flows)] = flows
    bufs = [torch.zeros((nmax,) + hw2, dtype=torch.float32, device=device) for _ in range(world)]
    t = torch.from_numpy(pad)
    dist.all_gather(bufs, t.to(device) if device is not None else t)
    out = []
    for r in range(world):
        b = bufs[r].cpu().numpy()
        out += [b[j, lo:hi] for j in range(n_all[r])]
    return out


class GpuBackend:
    """The C-ABI engine: rc_flow_push_batch -> rc_batch_hist -> counters := prefix -> rc_aggregate_last.
    At most max_batch pairs per call; re-entrant (run_block / run_stream may call it once per super-block)."""

    def __init__(self, ctx, w, h, params, max_batch):
        self.ctx, self.w, self.h, self.P, self.B = ctx, w, h, tuple(params), max_batch
        ctx.flow_configure_batch(w, h, *self.P, max_batch)
        ctx.hist_reset()
        ctx.accumulator_reset()
        self._nb = 0

    def flows_and_counts(self, frames):
        frames = np.ascontiguousarray(frames, np.uint8)
        n_pairs = frames.shape[0] - 1
        if n_pairs > self.B:
            raise ValueError("block of %d pairs is larger than max_batch = %d: use run_stream, which walks a clip in "
                             "super-blocks of world * max_batch pairs" % (n_pairs, self.B))
        self.ctx.flow_configure_batch(self.w, self.h, *self.P, self.B)   # restart: the block's first frame primes
        assert self.ctx.flow_push_batch(frames[:1]) == 0
        assert self.ctx.flow_push_batch(frames[1:]) == n_pairs
        self._nb = n_pairs
        return self.ctx.batch_hist(n_pairs)

    def aggregate(self, prefix_counts, framecounts):
        # the device counters become exactly "every frame before this rank's first one" (earlier super-blocks of all ranks
        # + lower ranks of this one); whatever an earlier call left there is discarded
        self.ctx.hist_reset()
        self.ctx.hist_add(prefix_counts)
        res = self.ctx.aggregate_last(self._nb, framecounts[0])
        return [float(r.UPPER) for r in res]

    def accumulator(self):
        p, w, h = self.ctx.accumulator_device()
        return self.ctx.accumulator_get(w, h).ravel()
