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


class ShardedAdam:
    """Data-parallel optimizer step for ONE [P,F] parameter (the `_seg_feature` of semantic training):

        reduce-scatter of the gradient  ->  Adam on this rank's rows  ->  all-gather of the updated rows

    instead of all-reduce + the full Adam step on every rank: same bytes on the wire, but the optimizer pass (and, with a
    deferred row-normalisation gradient, its chain rule) touches 1/N of the rows on every rank and the two moments are
    sharded (1/N of their memory).  The rows are cut into `chunks` contiguous pieces, each reduce-scattered / updated /
    all-gathered on its own, so that the collective of one piece overlaps the update of the previous one; within piece
    c (rows [c*Pc, (c+1)*Pc)) rank r owns rows [c*Pc + r*Pc/N, c*Pc + (r+1)*Pc/N).  Synchronous-SGD semantics: the result
    equals Adam on the sum of the ranks' gradients.

    update_fn(param_rows, grad_rows, exp_avg_rows, exp_avg_sq_rows, step, deferred_cfg) performs the update of one piece
    in place; the default is the fused CUDA kernel (isr_adam_step / isr_adam_rownorm_step); the CPU tests inject a torch
    implementation.  Backends without reduce_scatter_tensor (gloo) fall back to all_reduce + slicing (same result)."""

    def __init__(self, param: torch.Tensor, world: int, rank: int, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, chunks: int = 4,
                 update_fn=None):
        if param.dim() != 2:
            raise ValueError("ShardedAdam handles one [P,F] parameter")
        P = int(param.shape[0])
        chunks = max(1, int(chunks))
        while chunks > 1 and P % (chunks * world) != 0:
            chunks -= 1
        if P % (chunks * world) != 0:
            raise ValueError(f"P = {P} is not divisible by the world size {world}")
        self.param, self.world, self.rank, self.chunks = param, world, rank, chunks
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.rows = P // (chunks * world)            # rows per (piece, rank)
        F = int(param.shape[1])
        mk = lambda: torch.zeros((chunks, self.rows, F), dtype=param.dtype, device=param.device)
        self.exp_avg, self.exp_avg_sq = mk(), mk()
        self.grad_shard = torch.empty((chunks, self.rows, F), dtype=param.dtype, device=param.device)
        self.step_count = 0
        self.update_fn = update_fn if update_fn is not None else self._fused_update
        self._rs_ok = world > 1 and dist.is_initialized() and dist.get_backend() == "nccl"

    # -- default update: the fused CUDA kernels
    def _fused_update(self, p_rows, g_rows, m_rows, v_rows, step, deferred_cfg):
        from . import _lib
        from .rasterizer import _require_cuda_lib, _stream
        L = _require_cuda_lib()
        hyper = (self.lr, self.betas[0], self.betas[1], self.eps, int(step), None, _stream())
        if deferred_cfg is None:
            _lib.check(L.isr_adam_step(p_rows.numel(), p_rows.data_ptr(), g_rows.data_ptr(), m_rows.data_ptr(), v_rows.data_ptr(),
                                       *hyper), "isr_adam_step")
        else:
            e1, e2, stages = deferred_cfg
            _lib.check(L.isr_adam_rownorm_step(int(p_rows.shape[0]), int(p_rows.shape[1]), p_rows.data_ptr(), g_rows.data_ptr(),
                                               None, m_rows.data_ptr(), v_rows.data_ptr(), float(e1), float(e2), int(stages),
                                               *hyper), "isr_adam_rownorm_step")

    def _piece(self, t, c):
        n = self.rows * self.world
        return t[c * n:(c + 1) * n]

    def _mine(self, t, c):
        n = self.rows * self.world
        return t[c * n + self.rank * self.rows: c * n + (self.rank + 1) * self.rows]

    @torch.no_grad()
    def step(self, grad: torch.Tensor, deferred_cfg=None):
        """grad: this rank's [P,F] gradient (or dL/d(normalised rows) with deferred_cfg = (eps1, eps2, stages))."""
        self.step_count += 1
        p = self.param
        if self.world <= 1:
            for c in range(self.chunks):
                self.update_fn(self._mine(p, c), self._mine(grad, c), self.exp_avg[c], self.exp_avg_sq[c], self.step_count, deferred_cfg)
            return
        works = [None] * self.chunks

        def start_rs(c):
            if self._rs_ok:
                works[c] = dist.reduce_scatter_tensor(self.grad_shard[c], self._piece(grad, c), op=dist.ReduceOp.SUM, async_op=True)
            else:  # gloo: all-reduce the piece, keep my rows
                works[c] = dist.all_reduce(self._piece(grad, c), op=dist.ReduceOp.SUM, async_op=True)

        ahead = 2
        for c in range(min(ahead, self.chunks)):
            start_rs(c)
        gathers = []
        for c in range(self.chunks):
            works[c].wait()
            g_rows = self.grad_shard[c] if self._rs_ok else self._mine(grad, c)
            mine = self._mine(p, c)
            self.update_fn(mine, g_rows, self.exp_avg[c], self.exp_avg_sq[c], self.step_count, deferred_cfg)
            if self._rs_ok:
                gathers.append(dist.all_gather_into_tensor(self._piece(p, c), mine, async_op=True))
            else:
                parts = list(self._piece(p, c).split(self.rows))
                gathers.append(dist.all_gather(parts, mine.clone(), async_op=True))
            if c + ahead < self.chunks:
                start_rs(c + ahead)
        for w in gathers:
            w.wait()
