"""Multi-GPU plumbing of the inference path: one process per GPU (torch.distributed), clips sharded over ranks.

Frames / clips are independent on the try-on inference path, so there is NO data-path collective:
each rank owns `shard_clips(...)` of the clip list.  The process group is only used for the barrier and for the
max-over-ranks reduction of device-timed intervals (NCCL on GPUs, gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_process_group(backend=None, device=None):
    """Initialise from the torchrun environment (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).  No-op for world 1."""
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def shard_clips(n_clips, rank, world):
    """Round-robin assignment of clip indices to ranks (the reference's DistributedSampler without shuffling,
    models/base_model.py:113-115, minus its padding: every clip is processed exactly once)."""
    return list(range(rank, n_clips, world))


def barrier():
    if dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (device-timed milliseconds -> slowest rank)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
