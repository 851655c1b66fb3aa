"""instascene_b200 -- B200 (sm_100a) surfel rasterizer, sampled-pixel contrastive loss and k-NN initialiser behind
the reference InstaScene interfaces (diff_surfel_rasterization, gaussian_renderer.render, contrastive_loss,
distCUDA2).  Host side = Python mirror of the reference API; device side = libisr.so (C ABI, include/isr.h)."""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, sample_pixels, normalize_rows,  # noqa: F401
                         sample_labelled_pixels, set_arithmetic, arithmetic,
                         _C)
from .renderer import render, render_sampled, depth_to_normal, prefetch_geometry  # noqa: F401
from .contrastive import contrastive_loss  # noqa: F401
from .knn import distCUDA2  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .tracker import get_segmap_gaussians, segmap_gaussians  # noqa: F401
from .losses import photometric_loss, l1_loss, ssim, add_densification_stats  # noqa: F401
from . import io, optim  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "sample_pixels", "sample_labelled_pixels", "normalize_rows", "render",
           "render_sampled", "depth_to_normal", "prefetch_geometry", "contrastive_loss", "distCUDA2", "FusedAdam", "get_segmap_gaussians", "segmap_gaussians",
           "photometric_loss", "l1_loss", "ssim", "add_densification_stats", "io", "set_arithmetic", "arithmetic"]
