"""Multi-GPU plumbing: channels (or bursts) shard over ranks as contiguous blocks with no exchange on
the data path; the only collective is a sum of small statistics vectors."""


def partition(nunits, world, rank):
    """Contiguous block partition: rank r of `world` owns units [start, start + count)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank %d of %d" % (rank, world))
    base, extra = divmod(nunits, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def reduce_stats(values, device=None):
    """Sum a flat list of numbers over all ranks (NCCL on GPUs, gloo on CPU); returns Python floats.
    With no process group this is the identity, so single-GPU callers need no special case."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
