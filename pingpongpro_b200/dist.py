"""Multi-GPU plumbing: one process per GPU, each holding one position-range shard of the genome
(SURVEY.md §8(e)).  The data path has no collective; the only exchange is the all-reduce of three small
tables (maximum stack height, height -> frequency histogram, grouped overlap counts), which the CUDA layer
requests through its collective hook (ppp_set_allreduce).  On GPUs the hook is the library's own NCCL binding
(attach_nccl: torch.distributed only carries the 128-byte communicator id to the other ranks); a
torch.distributed-backed hook is kept for comparison, and gloo stands in for the CPU tests of the sharding logic.

PyTorch is plumbing only: it never touches the stack arrays or the kernels.
"""
from __future__ import annotations

import numpy as np


class _DevView:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, count: int, dtype: int):
    """torch tensor aliasing `count` uint32 (dtype 0) / uint64 (dtype 1) values at device address `ptr`.
    Viewed as signed integers of the same width (sum and max of values < 2^31 / 2^63 are identical)."""
    import torch

    return torch.as_tensor(_DevView(ptr, count, "<i4" if dtype == 0 else "<i8"), device="cuda")


def shard_ranges(total_positions: int, world_size: int):
    """Equal position ranges of the concatenated genome; rank r owns [cuts[r], cuts[r+1])."""
    return [total_positions * k // world_size for k in range(world_size + 1)]


def attach_nccl(ctx, group=None):
    """Binds the context's all-reduce hook to the library's NCCL communicator.  Collective over `group`."""
    import torch.distributed as dist

    from . import api

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [api.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.attach_nccl(box[0], rank, world)


def torch_allreduce_hook(group=None):
    """Collective hook for Context.set_allreduce backed by torch.distributed (NCCL).  The reduction is enqueued
    behind the context's kernels by running it with the context's stream as torch's current stream."""
    import torch
    import torch.distributed as dist

    def hook(ptr, count, dtype, op, stream):
        t = device_tensor(ptr, count, dtype)
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM, group=group)

    return hook


def reduce_tables_cpu(tables, op: str, group=None):
    """Host-side twin of the hook for the gloo tests: all-reduces a numpy integer table in place."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(tables).view(np.int64 if tables.dtype.itemsize == 8 else np.int32))
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=group)
    return t.numpy().view(tables.dtype)
