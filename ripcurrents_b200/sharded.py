"""Host-side logic for ONE camera stream sharded by frame pair across ranks.

Flow of pair (t, t+1) depends only on those two frames, so contiguous blocks of pairs go to different ranks with one
duplicated frame at each block edge and NO collective in the flow path.  The temporal aggregation is sequential in
three places, each solved with a tiny exchange:

  cumulative histograms -> thresholds at frame t need the counts of every frame <= t:
        all_gather of per-frame counts (37*50 int64 per frame) + an exclusive prefix over ranks;
  accumulator (plain sum of 0/1 per pixel, frames > 30): all_reduce(SUM) -- integer-valued, exact in any order;
  outmask: computed from the all-reduced accumulator at the reporting point (block end).

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


def run_block(backend, frames, lo, hi, framecount_of_pair, dist=None, device=None):
    """Processes pairs [lo, hi) of `frames` (frame i and i+1 form pair i) on this rank and combines across ranks.

    backend.flows_and_counts(frames[lo:hi+1])        -> int64 [hi-lo, 37, 50] per-frame counts (flows stay inside)
    backend.aggregate(prefix_counts, framecounts)    -> list of per-frame UPPER thresholds; adds into backend's accumulator
    backend.accumulator() / set_accumulator(a)       -> float32 [h*w] (numpy)
    Returns dict(upper=[...this rank's frames], counts_total=int64[37,50], accumulator=float32[h*w] (global)).
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
    uppers = backend.aggregate(prefix, [framecount_of_pair(p) for p in range(lo, hi)]) if n_local > 0 else []
    acc = backend.accumulator()
    if world > 1:
        t = torch.from_numpy(acc.copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t)
        acc = t.cpu().numpy()
        backend.set_accumulator(acc)
    total = sum((c.sum(0) for c in all_counts if c.shape[0]), np.zeros((37, 50), np.int64))
    return {"upper": uppers, "counts_total": total, "accumulator": acc}


class GpuBackend:
    """The C-ABI engine: rc_flow_push_batch -> rc_batch_hist -> rc_hist_add(prefix) -> rc_aggregate_last.
    One block of at most max_batch pairs per run_block call."""

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
            raise ValueError("block larger than max_batch: call run_block once per sub-block")
        self.ctx.flow_configure_batch(self.w, self.h, *self.P, self.B)   # restart: the block's first frame primes
        assert self.ctx.flow_push_batch(frames[:1]) == 0
        assert self.ctx.flow_push_batch(frames[1:]) == n_pairs
        self._nb = n_pairs
        return self.ctx.batch_hist(n_pairs)

    def aggregate(self, prefix_counts, framecounts):
        self.ctx.hist_add(prefix_counts)
        res = self.ctx.aggregate_last(self._nb, framecounts[0])
        return [float(r.UPPER) for r in res]

    def accumulator(self):
        p, w, h = self.ctx.accumulator_device()
        return self.ctx.accumulator_get(w, h).ravel()

    def set_accumulator(self, acc):
        self.global_accumulator = acc     # reporting copy; the device keeps this rank's own contribution
