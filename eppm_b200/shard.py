"""Host-side sharding of independent frame pairs across ranks (one process per GPU, no data-path collective).

The dense-correspondence path has no exchange step for batched pairs: every pair is an independent unit and
the RNG stream depends on the level geometry only, so results do not depend on how a batch is split.  The only
cross-rank traffic is the timing reduction of the benchmark (max over ranks)."""


def shard_range(n_total, rank, world):
    """Contiguous shard [lo, hi) of `n_total` units for `rank`; sizes differ by at most one, earlier ranks get the extra unit."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def stream_shard(n_frames, rank, world):
    """Video stream (BASELINE config 5): frames [f_lo, f_hi) for `rank` such that its consecutive pairs (t, t+1) are the contiguous
    pair shard [p_lo, p_hi) of the n_frames - 1 pairs -- neighbouring ranks share exactly one frame.
    Returns (f_lo, f_hi, p_lo, p_hi); a rank without pairs gets an empty range."""
    if n_frames < 1:
        raise ValueError("a stream needs at least one frame")
    p_lo, p_hi = shard_range(n_frames - 1, rank, world)
    if p_hi == p_lo:
        return p_lo, p_lo, p_lo, p_hi
    return p_lo, p_hi + 1, p_lo, p_hi


def max_over_ranks(value, world, device=None):
    """MAX all-reduce of a python float over the default process group (gloo on CPU, nccl on GPU); identity when world == 1."""
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, world, device=None):
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
