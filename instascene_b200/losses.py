"""Mirror of the reference's photometric losses (utils/loss_utils.py: l1_loss :18-19, ssim :39-83) plus the fused
combination the RGB training step uses (train.py:76-77), SURVEY.md §8 row f-4.  One CUDA kernel each way
(csrc/isr_photometric.cu) instead of 5 depthwise convolutions + ~20 elementwise kernels and their autograd mirror.
No CPU fallback."""
from __future__ import annotations

import torch

from . import _lib
from .rasterizer import _f32c, _require_cuda_lib, _stream


class _Photometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        L = _require_cuda_lib()
        img = _f32c(image.detach(), "image")
        g = _f32c(gt.detach(), "gt")
        if img.shape != g.shape or img.dim() not in (3, 4):
            raise RuntimeError("image and gt must share a [C,H,W] (or [B,C,H,W]) shape")
        H, W = int(img.shape[-2]), int(img.shape[-1])
        C = int(img.numel() // (H * W))  # a batch is just more channels: every plane is filtered independently
        ws_bytes = L.isr_photometric_workspace_bytes(C, H, W)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=img.device)
        out = torch.empty(3, dtype=torch.float32, device=img.device)
        _lib.check(L.isr_photometric_forward(C, H, W, img.data_ptr(), g.data_ptr(), float(lambda_dssim), ws.data_ptr(),
                                             ws_bytes, out.data_ptr(), _stream()), "isr_photometric_forward")
        ctx.save_for_backward(img, g, ws)
        ctx.cfg = (C, H, W, float(lambda_dssim), image.shape)
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, grad_loss, _grad_parts):
        L = _require_cuda_lib()
        img, g, ws = ctx.saved_tensors
        C, H, W, lam, shape = ctx.cfg
        scale = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        dimg = torch.empty_like(img)
        _lib.check(L.isr_photometric_backward(C, H, W, img.data_ptr(), g.data_ptr(), lam, ws.data_ptr(), scale.data_ptr(),
                                              dimg.data_ptr(), _stream()), "isr_photometric_backward")
        return dimg.reshape(shape), None, None


def photometric_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float = 0.2, return_parts: bool = False):
    """(1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))   (train.py:76-77), fused.
    `return_parts`: also returns the device tensor [loss, L1, SSIM] (no gradient)."""
    loss, parts = _Photometric.apply(image, gt, lambda_dssim)
    return (loss, parts) if return_parts else loss


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """utils/loss_utils.py:18-19."""
    return photometric_loss(network_output, gt, 0.0)


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """utils/loss_utils.py:46-55 (window 11, mean over everything)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("only the configuration the training loop uses: window_size=11, size_average=True")
    return 1.0 - photometric_loss(img1, img2, 1.0)


def add_densification_stats(max_radii2D: torch.Tensor, xyz_gradient_accum: torch.Tensor, denom: torch.Tensor,
                            radii: torch.Tensor, viewspace_grad: torch.Tensor) -> None:
    """In place, for the Gaussians with radii > 0 (the `visibility_filter` of render()):
        max_radii2D[vis] = max(max_radii2D[vis], radii[vis])                                  train.py:140-141
        xyz_gradient_accum[vis] += norm(viewspace_points.grad[vis], dim=-1, keepdim=True)     gaussian_model.py:602-604
        denom[vis] += 1                                                                       gaussian_model.py:605
    one kernel, no boolean-mask indexing (each masked read/write of the reference syncs the host for the mask count)."""
    L = _require_cuda_lib()
    P = int(radii.shape[0])
    for t, n in ((max_radii2D, "max_radii2D"), (xyz_gradient_accum, "xyz_gradient_accum"), (denom, "denom")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == P):
            raise RuntimeError(f"{n} must be a contiguous CUDA float32 tensor with {P} elements")
    r = radii.to(torch.int32).contiguous()
    g = _f32c(viewspace_grad, "viewspace_grad")
    if g.numel() != 3 * P:
        raise RuntimeError("viewspace_grad must be [P,3]")
    _lib.check(L.isr_densify_stats(P, r.data_ptr() if P else None, g.data_ptr() if P else None, max_radii2D.data_ptr() if P else None,
                                   xyz_gradient_accum.data_ptr() if P else None, denom.data_ptr() if P else None, _stream()),
               "isr_densify_stats")
