"""Data-parallel plumbing: views shard across ranks (one process per GPU), every rank holds a full replica of the
Gaussians, and the gradient of the trainable tensor(s) is summed with one all-reduce per step.
The reference has no distributed code; semantics = running the reference on all views of the step and adding
the gradients (its loss is a sum, utils/contrastive_utils.py:71)."""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend: str | None = None) -> tuple[int, int, int]:
    """Initialises torch.distributed from the torchrun environment (no-op for WORLD_SIZE=1)."""
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """View i -> rank i mod world (200 views -> 25 per rank at 8 GPUs)."""
    return list(range(rank, n_views, world))


def step_views(step: int, n_views: int, rank: int, world: int) -> int:
    """The view rank `rank` renders at global step `step` (one view per rank per step)."""
    return (step * world + rank) % n_views


def allreduce_grads(tensors: Iterable[torch.Tensor], world: int) -> None:
    """Sum gradients over ranks in place (one flat all-reduce per tensor; these are few and large: [P,F])."""
    if world <= 1:
        return
    for t in tensors:
        if t is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)


def max_over_ranks(value: float, world: int, device) -> float:
    if world <= 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world: int) -> None:
    if world > 1:
        dist.barrier()
