"""Device-side Gaussian-tracker extraction: mirror of ``get_segmap_gaussians``
(spatial_track/modules/init_tracker.py:16-47), SURVEY.md §8 row f-1.

The reference moves the whole (gaussian, pixel) pair list of a view to Python (`.tolist()` of millions of entries) and
builds one Python set per mask.  Here the pair list stays in HBM (libisr.so: isr_tracker_mark / isr_tracker_fill, one
P-bit set per mask) and only the distinct ids per mask come back -- already sorted and unique.

    segmap_gaussians(pairs, segmap, P)          -> TrackerSets (device CSR: mask ids, offsets, Gaussian ids, frame ids)
    get_segmap_gaussians(gaussian, view)        -> (mask_info, frame_gaussian_ids)   the reference's return value
"""
from __future__ import annotations

from typing import Callable, NamedTuple, Optional

import numpy as np
import torch

from . import _lib
from .rasterizer import _require_cuda_lib, _stream

MIN_GAUSSIANS_PER_MASK = 50  # init_tracker.py:41


class TrackerSets(NamedTuple):
    mask_ids: np.ndarray      # [M] kept mask ids, ascending (host)
    offsets: np.ndarray       # [M+1] int64 offsets into gaussian_ids (host)
    gaussian_ids: torch.Tensor  # [nnz] int32, ascending within each mask (device)
    frame_ids: torch.Tensor     # [n_frame] int32 ascending: every Gaussian in the pair list (device)
    counts: np.ndarray        # [K-1] distinct Gaussians of EVERY non-zero mask id of the view (before the threshold)
    all_mask_ids: np.ndarray  # [K-1] those mask ids, ascending


def segmap_gaussians(pairs: torch.Tensor, segmap: torch.Tensor, P: int,
                     min_gaussians: int = MIN_GAUSSIANS_PER_MASK) -> TrackerSets:
    """pairs: int32 [G,2] (gaussian id, pixel id = W*y+x) as returned by the rasterizer; segmap: integer mask ids per
    pixel (any shape, flattened row-major), 0 = background."""
    L = _require_cuda_lib()
    if not pairs.is_cuda:
        raise RuntimeError("pairs must be a CUDA tensor")
    dev = pairs.device
    pairs = pairs.to(torch.int32).contiguous()
    G = int(pairs.shape[0])
    seg = segmap.to(dev).reshape(-1)
    HW = int(seg.numel())
    ids = torch.unique(seg)  # sorted (init_tracker.py:31-32)
    ids_nz = ids[ids != 0]
    K = int(ids_nz.numel()) + 1  # row 0 = frame set / background
    # dense row per pixel: 0 for background, 1 + rank of the mask id otherwise
    rows = torch.where(seg != 0, torch.searchsorted(ids_nz, seg) + 1, torch.zeros_like(seg)).to(torch.int32).contiguous()
    ws_bytes = L.isr_tracker_workspace_bytes(P, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    counts = torch.empty(K, dtype=torch.int32, device=dev)
    _lib.check(L.isr_tracker_mark(pairs.data_ptr() if G else None, G, rows.data_ptr() if HW else None, HW, int(P), K,
                                  ws.data_ptr(), ws_bytes, counts.data_ptr(), _stream()), "isr_tracker_mark")
    counts_h = counts.cpu().numpy().astype(np.int64)  # K integers: the only host read before the fill
    all_ids = ids_nz.cpu().numpy()
    keep = np.zeros(K, dtype=bool)
    keep[0] = True
    keep[1:] = counts_h[1:] >= min_gaussians
    sizes = np.where(keep, counts_h, 0)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    row_off = np.where(keep, starts[:-1], -1).astype(np.int64)
    total = int(starts[-1])
    out = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    if P > 0:
        row_off_d = torch.from_numpy(row_off).to(dev)
        _lib.check(L.isr_tracker_fill(int(P), K, ws.data_ptr(), row_off_d.data_ptr(), out.data_ptr(), _stream()),
                   "isr_tracker_fill")
    out = out[:total]
    n_frame = int(counts_h[0])
    kept_rows = np.nonzero(keep[1:])[0]
    offsets = np.concatenate([[0], np.cumsum(counts_h[1:][kept_rows])]).astype(np.int64)
    return TrackerSets(all_ids[kept_rows], offsets, out[n_frame:], out[:n_frame], counts_h[1:], all_ids)


def get_segmap_gaussians(gaussian, view, render_fn: Optional[Callable] = None, as_sets: bool = True):
    """Same call and return value as the reference: ({mask_id: set of Gaussian ids}, list of frame Gaussian ids).
    `as_sets=False` returns ascending numpy arrays instead of Python sets / a list (same contents, no per-element
    Python objects); the reference's consumers (`construct_mask2gs_tracker`, init_tracker.py:108-120) accept both."""
    if render_fn is None:
        from .renderer import render as render_fn
    background = torch.tensor([0, 0, 0], dtype=torch.float32, device="cuda")
    pairs = render_fn(view, gaussian, gaussian.pipelineparams, background)["gau_related_pixels"]
    ts = segmap_gaussians(pairs, view.segmap, len(gaussian.get_xyz))
    ids_h = ts.gaussian_ids.cpu().numpy()
    frame_h = ts.frame_ids.cpu().numpy()
    mask_info = {}
    for j, mask_id in enumerate(ts.mask_ids):
        arr = ids_h[ts.offsets[j]:ts.offsets[j + 1]]
        mask_info[mask_id] = set(arr.tolist()) if as_sets else arr
    return mask_info, (frame_h.tolist() if as_sets else frame_h)
