"""Fused single-kernel Adam for the trainable Gaussian tensors.  Same update rule and state
names as torch.optim.Adam (exp_avg, exp_avg_sq, step) without weight decay / amsgrad / maximize -- the configuration
the reference uses for `_seg_feature` (scene/gaussian_model.py:217-249: Adam(lr=0.025, eps=1e-15)).

Two extras of this implementation:
  * deferred row-normalisation gradient: when the parameter was fed to the rasterizer through
    `normalize_rows(..., defer_to=param)` (render() with `pipe.defer_seg_feature_grad = True`), its gradient arrives as
    dL/d(normalised rows) in `param._isr_deferred_dy` and the chain rule of the normalisation is applied INSIDE the Adam
    kernel (isr_adam_rownorm_step): parameter, gradient and both moments are each read once and the chained gradient is
    never written.  An ordinary `param.grad` from other uses of the parameter is added after the chain rule.
  * `capturable=True`: the step count lives on the device (incremented on the stream), so a captured CUDA graph replays
    with the right bias correction."""
from __future__ import annotations

import torch

from . import _lib
from .rasterizer import _ptr, _require_cuda_lib, _stream


def deferred_grad(p):
    """(dy, (eps1, eps2, stages)) of a parameter whose row-normalisation backward was deferred, else None."""
    dy = getattr(p, "_isr_deferred_dy", None)
    return None if dy is None else (dy, p._isr_deferred_cfg)


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.capturable = bool(capturable)

    def _state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            if self.capturable:
                st["step_dev"] = torch.zeros(1, dtype=torch.int32, device=p.device)
        return st

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)
        for group in self.param_groups:
            for p in group["params"]:
                if getattr(p, "_isr_deferred_dy", None) is not None:
                    p._isr_deferred_dy = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        L = _require_cuda_lib()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                deferred = deferred_grad(p)
                if p.grad is None and deferred is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdam handles contiguous fp32 CUDA parameters only")
                g = p.grad
                if g is not None:
                    if g.is_sparse:
                        raise RuntimeError("FusedAdam does not support sparse gradients")
                    g = g.contiguous()
                st = self._state(p)
                st["step"] += 1
                step, step_dev = int(st["step"]), None
                if self.capturable:
                    st["step_dev"].add_(1)
                    step, step_dev = 0, st["step_dev"].data_ptr()
                hyper = (float(group["lr"]), float(b1), float(b2), float(group["eps"]), step, step_dev, _stream())
                if deferred is not None:
                    dy, (e1, e2, stages) = deferred
                    dy = dy.contiguous()
                    if dy.shape != p.shape or p.dim() != 2:
                        raise RuntimeError("deferred row-normalisation gradient must have the parameter's [P,F] shape")
                    _lib.check(L.isr_adam_rownorm_step(int(p.shape[0]), int(p.shape[1]), p.data_ptr(), dy.data_ptr(), _ptr(g),
                                                       st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), float(e1),
                                                       float(e2), int(stages), *hyper), "isr_adam_rownorm_step")
                else:
                    _lib.check(L.isr_adam_step(p.numel(), p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                               st["exp_avg_sq"].data_ptr(), *hyper), "isr_adam_step")
        return loss
