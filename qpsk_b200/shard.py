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


def bind_host_to_gpu(local_index):
    """Pin this process to the CPU cores NVML reports as local to GPU `local_index` (its NUMA node), so that
    pinned host buffers allocated afterwards sit behind the same PCIe root as the GPU that reads them.  With
    one process per GPU and 55 GB/s of host reads per process this decides whether N GPUs share one socket's
    memory controllers.  Returns the core list, or None when NVML or the affinity call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
