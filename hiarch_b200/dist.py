"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL on GPUs, gloo in CPU tests).

The hot path shards with NO data-path collective: samples, and inside a
sample chromosome blocks, are independent. The only collectives are the timing ones of
bench.py (barrier, max over ranks). The row-sharded genome-wide chain (config C5) is the single
place that needs an exchange (all-gather of the operand planes / Lanczos panels, small
all-reduces): hiarch_b200/sharded.py.
"""
import os

import numpy as np


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend=None):
    """Initialise the default process group when WORLD_SIZE > 1. Returns (rank, local, world)."""
    import torch
    import torch.distributed as dist
    rank, local, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kwargs)
    return rank, local, world


def barrier():
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value):
    """Max of a python float over all ranks (device time of the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def yield_mp_idxes(size, thread):
    """The reference's contiguous sharding (publish/run_hic_dataset.py:259-269): shards of
    int(size / thread) items; the remainder makes extra shards."""
    cur = 0
    n_sub = int(size / thread) if size > thread else 1
    while cur < size:
        yield np.arange(cur, min(cur + n_sub, size)).astype(int)
        cur += n_sub


def shard_for_rank(weights, world, rank):
    """Indices of the work items rank `rank` owns: largest-first greedy balance by weight
    (sum n_c^2 for chromosome maps), deterministic, disjoint, complete."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    load = [0.0] * world
    owner = [0] * len(weights)
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        owner[i] = r
        load[r] += weights[i]
    return [i for i in range(len(weights)) if owner[i] == rank]
