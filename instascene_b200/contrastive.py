"""Drop-in mirror of ``utils.contrastive_utils.contrastive_loss`` (utils/contrastive_utils.py:18-73) backed by the
fused CUDA kernels of libisr.so (isr_contrastive_forward / _backward): no torch.unique, no host sync."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .rasterizer import _ptr, _require_cuda_lib, _stream


class _ContrastiveLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, labels, predef_u, K, temp_lambda, min_pixnum=0):
        L = _require_cuda_lib()
        feats = features.detach().float().contiguous()
        N, F = int(feats.shape[0]), int(feats.shape[1])
        labels = labels.to(torch.int32).contiguous()
        pu = None if predef_u is None else predef_u.detach().float().contiguous()
        ws_bytes = L.isr_contrastive_workspace_bytes(N, F, K)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=feats.device)
        loss = torch.empty((), dtype=torch.float32, device=feats.device)
        _lib.check(L.isr_contrastive_forward(N, F, K, _ptr(feats), _ptr(labels), _ptr(pu), float(temp_lambda),
                                             int(min_pixnum), ws.data_ptr(), ws_bytes, loss.data_ptr(), _stream()),
                   "isr_contrastive_forward")
        ctx.save_for_backward(feats, labels, ws)
        ctx.pu = pu
        ctx.K = K
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        L = _require_cuda_lib()
        feats, labels, ws = ctx.saved_tensors
        N, F = int(feats.shape[0]), int(feats.shape[1])
        g = grad_loss.detach().float().contiguous().reshape(1)
        dfeat = torch.empty_like(feats)
        _lib.check(L.isr_contrastive_backward(N, F, ctx.K, _ptr(feats), _ptr(labels), _ptr(ctx.pu), ws.data_ptr(),
                                              g.data_ptr(), dfeat.data_ptr(), _stream()), "isr_contrastive_backward")
        return dfeat, None, None, None, None, None


def contrastive_loss(features: torch.Tensor, masks: torch.Tensor, predef_u_list: Optional[torch.Tensor] = None,
                     min_pixnum: int = 0, temp_lambda: float = 1000, consider_negative: bool = False,
                     num_labels: Optional[int] = None) -> torch.Tensor:
    """ProtoNCE loss over sampled (feature, label) pairs; same semantics as the reference.

    `num_labels` (optional) bounds the label ids (K = num_labels); when omitted it is taken from predef_u_list or
    from `masks.max()` (one host sync, as the reference's `mask_ids.max() + 1`)."""
    labels = masks.to(torch.int32)
    if not consider_negative:
        labels = labels - 1  # valid ids start at 0 (:39-40); label 0 (unlabelled) becomes -1 = ignored
    if num_labels is not None:
        K = int(num_labels)
    elif predef_u_list is not None:
        K = int(predef_u_list.shape[0])
    else:
        K = int(labels.max().item()) + 1
    K = max(K, 1)
    # (:33-35 clusters with <= min_pixnum samples are dropped inside the kernels)
    return _ContrastiveLoss.apply(features, labels, predef_u_list, K, float(temp_lambda), int(min_pixnum))
