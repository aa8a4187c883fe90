"""Stream sharding across GPUs: streams are independent, so rank r of W owns a contiguous
block of streams for the whole run; there is no collective on the data path.  torch.distributed is used
only to agree on the timing (max over ranks) and to gather a few scalars of statistics."""


def shard_range(n_streams, rank, world):
    """[lo, hi) of the streams owned by `rank`: contiguous blocks, sizes differing by at most one."""
    base, rem = divmod(n_streams, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_stats(dist, device, elapsed_ms, audio_seconds, checksum=0):
    """all_reduce the per-rank numbers: returns (max elapsed ms, total audio seconds, xor/sum checksum)."""
    import torch
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    a = torch.tensor([audio_seconds, float(checksum)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
    return float(t[0]), float(a[0]), float(a[1])
