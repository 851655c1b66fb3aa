"""Drop-in mirror of ``simple_knn._C.distCUDA2`` (submodules/simple-knn/spatial.cu:15-25)."""
from __future__ import annotations

import torch

from . import _lib
from .rasterizer import _require_cuda_lib, _stream


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    L = _require_cuda_lib()
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor")
    pts = points.detach().float().contiguous()
    P = int(pts.shape[0])
    out = torch.zeros(P, dtype=torch.float32, device=pts.device)
    if P == 0:
        return out
    ws_bytes = L.isr_knn_workspace_bytes(P)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pts.device)
    _lib.check(L.isr_knn_mean_dist2(P, pts.data_ptr(), out.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
               "isr_knn_mean_dist2")
    return out
