"""Fused single-kernel Adam for the trainable Gaussian tensors.  Same update rule and state
names as torch.optim.Adam (exp_avg, exp_avg_sq, step) without weight decay / amsgrad / maximize -- the configuration
the reference uses for `_seg_feature` (scene/gaussian_model.py:217-249: Adam(lr=0.025, eps=1e-15))."""
from __future__ import annotations

import torch

from . import _lib
from .rasterizer import _require_cuda_lib, _stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        L = _require_cuda_lib()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdam handles contiguous fp32 CUDA parameters only")
                g = p.grad
                if g.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                g = g.contiguous()
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                _lib.check(L.isr_adam_step(p.numel(), p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                           st["exp_avg_sq"].data_ptr(), float(group["lr"]), float(b1), float(b2),
                                           float(group["eps"]), int(st["step"]), _stream()), "isr_adam_step")
        return loss
