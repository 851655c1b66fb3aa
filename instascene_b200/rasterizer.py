"""Drop-in mirror of the reference package ``diff_surfel_rasterization``
(submodules/diff-surfel-rasterization/diff_surfel_rasterization/__init__.py), backed by libisr.so.

Same public names, argument order, return tuples and error behaviour:
  GaussianRasterizationSettings  (NamedTuple, 12 fields)            reference __init__.py:179-191
  GaussianRasterizer(nn.Module).forward / .markVisible              reference __init__.py:194-248
  rasterize_gaussians(...) / _RasterizeGaussians                     reference __init__.py:22-176
  _C.rasterize_gaussians / rasterize_gaussians_backward / mark_visible   DSR/ext.cpp:15-19

Differences that do not change results:
  * kernels run on torch's CURRENT stream (the reference uses the legacy default stream);
  * gau_related_pixels is allocated as [9*H*W, 2] and not pre-filled (reference: [100*H*W, 2] filled with -1,
    then sliced to the same [:count] view);
  * backward honours ctx.needs_input_grad and skips all-zero (None) cotangents -- exact, adds of 0;
  * an additional hidden sparse-cotangent handle lets `sample_pixels` route the <=32768 sampled-pixel
    gradients of train_semantic.py to a sparse backward kernel.
There is no CPU / eager fallback: everything raises if libisr.so or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib

_pinned_cache = {}

# Arithmetic of the MUFU-based functions (include/isr.h ISR_FLAG_SPEC_ARITH): "reference" (default) evaluates expf /
# rsqrtf exactly like the reference CUDA build, "spec" uses the CPU-reproducible stand-ins of oracle/isr_oracle.c.
_arith_flag = 0


def set_arithmetic(mode: str) -> None:
    """'reference' (default): bit-identical to the unmodified reference CUDA rasterizer; 'spec': bit-identical to the
    CPU oracle (tests).  Applies to forwards launched afterwards; a backward always uses its forward's mode."""
    global _arith_flag
    if mode not in ("reference", "spec"):
        raise ValueError("arithmetic mode must be 'reference' or 'spec'")
    _arith_flag = _lib.FLAG_SPEC_ARITH if mode == "spec" else 0


class arithmetic:
    """Context manager: `with arithmetic("spec"): ...`"""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        global _arith_flag
        self.prev = _arith_flag
        set_arithmetic(self.mode)
        return self

    def __exit__(self, *a):
        global _arith_flag
        _arith_flag = self.prev


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _pinned_i64(device: torch.device) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    t = _pinned_cache.get(key)
    if t is None:
        t = torch.zeros(2, dtype=torch.int64).pin_memory()  # [reference num_rendered, emitted instances]
        _pinned_cache[key] = t
    return t


# High-water mark of the emitted instance count per (P, W, H): sizes the binning workspace of views whose binning is
# queued before their count is known on the host (prefetch_geometry).
_instance_high_water = {}


def note_instances(P: int, W: int, H: int, n_inst: int) -> None:
    key = (int(P), int(W), int(H))
    if n_inst > _instance_high_water.get(key, 0):
        _instance_high_water[key] = int(n_inst)


def binning_capacity_hint(P: int, W: int, H: int) -> int:
    """0 until a view of this shape has been rendered; then 1.25x the largest instance count seen, in whole millions."""
    hw = _instance_high_water.get((int(P), int(W), int(H)), 0)
    return 0 if hw == 0 else ((int(hw * 1.25) + (1 << 20)) >> 20) << 20


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")  # CHECK_INPUT, DSR/rasterize_points.cu:27-28
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _require_cuda_lib():
    L = _lib.lib()
    if not torch.cuda.is_available():
        raise _lib.IsrError("instascene_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return L


# --------------------------------------------------------------------------------------------------------------
# _C-level functions: same signatures / returns as the reference pybind module (DSR/rasterize_points.h:17-73)
# --------------------------------------------------------------------------------------------------------------
class _ForwardState:
    """Everything between the two phases of the forward (isr_forward_geometry -> isr_forward_render)."""
    __slots__ = ("args", "keep", "P", "H", "W", "dev", "want_pairs", "out_color", "out_others", "radii", "geom", "img",
                 "pairs", "pair_count", "nr_host", "ready_event", "binning", "bin_capacity")


def launch_geometry(background, means3D, colors, opacity, scales, rotations, scale_modifier, transMat_precomp, viewmatrix,
                    projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                    want_pairs: bool = True, pinned_counts: Optional[torch.Tensor] = None,
                    bin_capacity: int = 0) -> Optional[_ForwardState]:
    """Phase A of the forward: K1 projection + depth order + offsets, enqueued asynchronously on the current stream.
    Nothing here depends on extra_attrs (the semantic features), so a caller may start it before the features of the
    step are final (e.g. while the previous step's gradient all-reduce / optimizer step runs on another stream).
    bin_capacity > 0: the binning (instance partition + tile ranges, which read no features either) is queued right
    behind it into a workspace of that many instances -- the kernels take the actual count from device memory;
    finish_render repeats the binning inline in the (rare) case that the count exceeds the capacity."""
    L = _require_cuda_lib()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    P, H, W = int(means3D.shape[0]), int(image_height), int(image_width)
    if P == 0:
        return None
    dev = means3D.device
    f32 = dict(dtype=torch.float32, device=dev)
    background = _f32c(background, "background")
    means3D = _f32c(means3D, "means3D")
    opacity = _f32c(opacity, "opacity")
    viewmatrix = _f32c(viewmatrix, "viewmatrix")
    projmatrix = _f32c(projmatrix, "projmatrix")
    campos = _f32c(campos, "campos")
    colors = _f32c(colors, "colors") if colors.numel() else colors
    scales = _f32c(scales, "scales") if scales.numel() else scales
    rotations = _f32c(rotations, "rotations") if rotations.numel() else rotations
    transMat_precomp = _f32c(transMat_precomp, "transMat_precomp") if transMat_precomp.numel() else transMat_precomp
    sh = _f32c(sh, "sh") if sh.numel() else sh
    M = int(sh.shape[1]) if sh.numel() else 0
    st = _ForwardState()
    st.ready_event = None  # set when phase A was launched on another stream (renderer.prefetch_geometry)
    st.binning, st.bin_capacity = None, 0
    st.P, st.H, st.W, st.dev, st.want_pairs = P, H, W, dev, want_pairs
    st.out_color = torch.empty((3, H, W), **f32)
    st.out_others = torch.empty((7, H, W), **f32)
    st.radii = torch.empty(P, dtype=torch.int32, device=dev)
    geom_bytes, img_bytes = L.isr_geom_bytes(P), L.isr_image_bytes(W, H)
    st.geom = torch.empty(geom_bytes, dtype=torch.uint8, device=dev)
    st.img = torch.empty(img_bytes, dtype=torch.uint8, device=dev)
    pair_cap = 9 * H * W if want_pairs else 0
    st.pairs = torch.empty((pair_cap, 2), dtype=torch.int32, device=dev)
    st.pair_count = torch.zeros(1, dtype=torch.int32, device=dev)
    st.nr_host = pinned_counts if pinned_counts is not None else _pinned_i64(dev)
    a = _lib.IsrForwardArgs()
    a.P, a.sh_degree, a.sh_coeffs, a.F, a.W, a.H = P, int(degree), M, 0, W, H
    a.flags = (0 if want_pairs else _lib.FLAG_NO_PAIRS) | _arith_flag
    a.tan_fovx, a.tan_fovy, a.scale_modifier = float(tan_fovx), float(tan_fovy), float(scale_modifier)
    a.background, a.viewmatrix, a.projmatrix, a.campos = _ptr(background), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos)
    a.means3D, a.opacities = _ptr(means3D), _ptr(opacity)
    a.scales, a.rotations, a.transMat_precomp = _ptr(scales), _ptr(rotations), _ptr(transMat_precomp)
    a.shs, a.colors_precomp, a.extra_attrs = _ptr(sh), _ptr(colors), None
    a.geom, a.geom_bytes, a.image, a.image_bytes = st.geom.data_ptr(), geom_bytes, st.img.data_ptr(), img_bytes
    a.radii, a.out_color, a.out_others, a.out_extra = st.radii.data_ptr(), st.out_color.data_ptr(), st.out_others.data_ptr(), None
    a.pairs, a.pair_capacity, a.pair_count = _ptr(st.pairs), pair_cap, st.pair_count.data_ptr()
    a.num_rendered_host = st.nr_host.data_ptr()
    st.args = a
    st.keep = (background, means3D, colors, opacity, scales, rotations, transMat_precomp, sh, viewmatrix, projmatrix, campos)
    _lib.check(L.isr_forward_geometry(C.byref(a), _stream()), "isr_forward_geometry")
    if bin_capacity > 0:
        cap = int(bin_capacity)
        bin_bytes = L.isr_binning_bytes(P, cap, W, H)
        st.binning = torch.empty(bin_bytes, dtype=torch.uint8, device=dev)
        st.bin_capacity = cap
        a.binning, a.binning_bytes = st.binning.data_ptr(), bin_bytes
        flags = a.flags
        a.flags = flags | _lib.FLAG_SKIP_BLEND
        _lib.check(L.isr_forward_render(C.byref(a), cap, _stream()), "isr_forward_render(binning only)")
        a.flags = flags
    return st


def finish_render(st: _ForwardState, extra_attrs, F: int, debug: bool = False, return_args: bool = False):
    """Phase B: wait for the instance count, size the binning workspace, emit + tile-sort + blend."""
    L = _require_cuda_lib()
    a, dev = st.args, st.dev
    if F > 0:
        extra_attrs = _f32c(extra_attrs, "extra_attrs")
        out_extra = torch.empty((F, st.H, st.W), dtype=torch.float32, device=dev)
        a.F, a.extra_attrs, a.out_extra = F, extra_attrs.data_ptr(), out_extra.data_ptr()
    else:
        out_extra = torch.empty(0, dtype=torch.float32, device=dev)
    if st.ready_event is not None:
        # phase A ran on another stream (usually a whole step earlier): order this stream behind it and wait on the host
        # for THAT event only -- nothing queued on the current stream is waited for
        torch.cuda.current_stream().wait_event(st.ready_event)
        st.ready_event.synchronize()
    else:
        torch.cuda.current_stream().synchronize()  # the reference blocks on a cudaMemcpy here (rasterizer_impl.cu:287)
    # [0]: what the reference reports as num_rendered (all tiles of every rectangle); [1]: instances actually binned
    num_rendered, n_inst = int(st.nr_host[0]), int(st.nr_host[1])
    note_instances(st.P, st.W, st.H, n_inst)
    if st.binning is not None and n_inst <= st.bin_capacity:
        # binned ahead of time (launch_geometry(bin_capacity=...)): only the blend is left
        binningBuffer, cap = st.binning, st.bin_capacity
        a.flags |= _lib.FLAG_SKIP_BINNING
    else:
        cap = n_inst
        bin_bytes = L.isr_binning_bytes(st.P, cap, st.W, st.H)
        binningBuffer = torch.empty(bin_bytes, dtype=torch.uint8, device=dev)
        a.binning, a.binning_bytes = binningBuffer.data_ptr(), bin_bytes
    _lib.check(L.isr_forward_render(C.byref(a), cap, _stream()), "isr_forward_render")
    a.flags &= ~_lib.FLAG_SKIP_BINNING
    if debug:
        torch.cuda.synchronize()
    res = (num_rendered, st.out_color, st.out_others, st.radii, out_extra, st.geom, binningBuffer, st.img, st.pairs,
           st.pair_count - 1)
    if return_args:  # profiling hook (bench.py roofline leg); keeps every tensor the struct points to alive
        a._keepalive = st.keep + (extra_attrs, st.pair_count, st.nr_host) + res[1:9]
        a._n_inst = n_inst
        a._bin_capacity = cap
        return res + (a,)
    return res


def finish_binning(st: _ForwardState):
    """Phase B without the blend: after this the view is PREPARED for the sampled-pixel kernels
    (isr_forward_sparse_extra): tile ranges + instance lists exist, no image is composited.  Returns
    (num_rendered, binningBuffer)."""
    L = _require_cuda_lib()
    a = st.args
    if st.ready_event is not None:
        torch.cuda.current_stream().wait_event(st.ready_event)
        st.ready_event.synchronize()
    else:
        torch.cuda.current_stream().synchronize()
    num_rendered, n_inst = int(st.nr_host[0]), int(st.nr_host[1])
    note_instances(st.P, st.W, st.H, n_inst)
    if st.binning is not None and n_inst <= st.bin_capacity:
        return num_rendered, st.binning  # binned ahead of time (launch_geometry(bin_capacity=...))
    bin_bytes = L.isr_binning_bytes(st.P, n_inst, st.W, st.H)
    st.binning = torch.empty(bin_bytes, dtype=torch.uint8, device=st.dev)
    st.bin_capacity = n_inst
    a.binning, a.binning_bytes = st.binning.data_ptr(), bin_bytes
    flags = a.flags
    a.flags = flags | _lib.FLAG_SKIP_BLEND
    _lib.check(L.isr_forward_render(C.byref(a), n_inst, _stream()), "isr_forward_render(binning only)")
    a.flags = flags
    return num_rendered, st.binning


class _SampledFeatures(torch.autograd.Function):
    """features[n,F] = the rendered extra_attrs at (view_ids[i], pix_ids[i]) -- only those pixels are composited
    (isr_forward_sparse_extra), all views in one launch; backward = isr_backward_sparse_extra_views."""

    @staticmethod
    def forward(ctx, extra_attrs, states, pix_ids, view_ids):
        L = _require_cuda_lib()
        st0 = states[0]
        P, H, W = st0.P, st0.H, st0.W
        feats = _f32c(extra_attrs.detach(), "extra_attrs")
        F = int(feats.shape[1])
        views = (_lib.IsrSparseView * len(states))()
        for i, st in enumerate(states):
            if (st.P, st.H, st.W) != (P, H, W):
                raise RuntimeError("sampled rendering needs views of one cloud at one resolution")
            views[i].geom, views[i].image, views[i].binning = st.geom.data_ptr(), st.img.data_ptr(), st.binning.data_ptr()
        ids32 = pix_ids.to(torch.int32).contiguous()
        vid32 = None if view_ids is None else view_ids.to(torch.int32).contiguous()
        n = int(ids32.numel())
        out = torch.empty((n, F), dtype=torch.float32, device=feats.device)
        flags = st0.args.flags & _lib.FLAG_SPEC_ARITH
        _lib.check(L.isr_forward_sparse_extra(len(states), views, P, F, W, H, _ptr(feats), n, _ptr(ids32), _ptr(vid32), _ptr(out),
                                              flags, _stream()), "isr_forward_sparse_extra")
        ctx.states, ctx.views, ctx.flags = states, views, flags  # the states own the workspaces the table points to
        ctx.dims = (P, F, W, H, n)
        ctx.save_for_backward(ids32) if vid32 is None else ctx.save_for_backward(ids32, vid32)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        L = _require_cuda_lib()
        saved = ctx.saved_tensors
        ids32, vid32 = saved[0], (saved[1] if len(saved) > 1 else None)
        P, F, W, H, n = ctx.dims
        g = _f32c(grad_out, "cotangent")
        dL_dextra = torch.zeros((P, F), dtype=torch.float32, device=g.device)
        _lib.check(L.isr_backward_sparse_extra_views(len(ctx.states), ctx.views, P, F, W, H, n, _ptr(ids32), _ptr(vid32), _ptr(g),
                                                     _ptr(dL_dextra), ctx.flags, _stream()), "isr_backward_sparse_extra_views")
        return dL_dextra, None, None, None


def sampled_features(extra_attrs: torch.Tensor, states, pix_ids: torch.Tensor, view_ids: Optional[torch.Tensor] = None):
    """Rendered features [n,F] at the given pixels of PREPARED views (launch_geometry + finish_binning each)."""
    if len(states) < 1 or len(states) > _lib.MAX_SPARSE_VIEWS:
        raise RuntimeError(f"between 1 and {_lib.MAX_SPARSE_VIEWS} views per call")
    return _SampledFeatures.apply(extra_attrs, tuple(states), pix_ids, view_ids)


def c_rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, transMat_precomp,
                          extra_attrs, attr_degree, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                          image_width, sh, degree, campos, prefiltered, debug, want_pairs: bool = True,
                          return_args: bool = False, geom_state: Optional[_ForwardState] = None):
    """RasterizeGaussiansCUDA (DSR/rasterize_points.cu:39-151).  Returns the reference's 10-tuple
    (num_rendered, out_color, out_others, radii, out_extra, geomBuffer, binningBuffer, imgBuffer,
     gau_related_pixels [cap,2], gau_pixel_indices [1] = count-1).  `geom_state`: phase A already launched."""
    _require_cuda_lib()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    dev = means3D.device
    P, H, W, F = int(means3D.shape[0]), int(image_height), int(image_width), int(attr_degree)
    f32 = dict(dtype=torch.float32, device=dev)
    if P == 0:  # DSR/rasterize_points.cu:112: nothing runs, outputs stay at their fill values
        out_color = torch.zeros((3, H, W), **f32)
        out_others = torch.zeros((7, H, W), **f32)
        out_extra = torch.zeros((F, H, W), **f32) if F > 0 else torch.empty(0, **f32)
        empty_u8 = torch.empty(0, dtype=torch.uint8, device=dev)
        return (0, out_color, out_others, torch.zeros(0, dtype=torch.int32, device=dev), out_extra, empty_u8,
                empty_u8.clone(), empty_u8.clone(), torch.empty((0, 2), dtype=torch.int32, device=dev),
                torch.full((1,), -1, dtype=torch.int32, device=dev))
    st = geom_state
    if st is None:
        st = launch_geometry(background, means3D, colors, opacity, scales, rotations, scale_modifier, transMat_precomp,
                             viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, sh, degree, campos, want_pairs=want_pairs)
    return finish_render(st, extra_attrs, F, debug=debug, return_args=return_args)


def c_rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, extra_attrs, scale_modifier,
                                   transMat_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                   dL_dout_others, dL_dout_extra, sh, degree, campos, geomBuffer, R, binningBuffer,
                                   imageBuffer, debug, grad_mask: int = _lib.GRAD_ALL, image_size=None,
                                   sparse_extra=None, flags: int = _lib.FLAG_BWD_WH_QUIRK, arith=None):
    """RasterizeGaussiansBackwardCUDA (DSR/rasterize_points.cu:153-262).  Returns the reference's 9-tuple
    (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dscales, dL_drotations, dL_dextra).
    Cotangents may be None (== zeros).  `sparse_extra` = (pix_ids int32 [n], rows float32 [n,F]).
    `arith`: the forward's ISR_FLAG_SPEC_ARITH bit (default: the current global mode)."""
    L = _require_cuda_lib()
    flags = (flags & ~_lib.FLAG_SPEC_ARITH) | (_arith_flag if arith is None else arith)
    dev = means3D.device
    P = int(means3D.shape[0])
    if image_size is None:
        H, W = int(dL_dout_color.shape[1]), int(dL_dout_color.shape[2])
    else:
        H, W = image_size
    F = int(extra_attrs.shape[1]) if extra_attrs.numel() else 0
    M = int(sh.shape[1]) if sh.numel() else 0
    geo, col, opa, ext = (bool(grad_mask & m) for m in (_lib.GRAD_GEOMETRY, _lib.GRAD_COLOR, _lib.GRAD_OPACITY, _lib.GRAD_EXTRA))
    none = torch.empty(0, dtype=torch.float32, device=dev)
    # the reference zero-fills all ten buffers (304+4F bytes per Gaussian) on every backward; only requested ones here
    z = lambda on, *s: torch.zeros(s, dtype=torch.float32, device=dev) if on else none
    dL_dmeans3D, dL_dmeans2D = z(geo, P, 3), z(geo, P, 3)
    dL_dnormal, dL_dtransMat = z(geo, P, 3), z(geo, P, 9)
    dL_dscales, dL_drotations = z(geo, P, 2), z(geo, P, 4)
    dL_dcolors, dL_dsh = z(col or geo, P, 3), z(col or geo, P, M, 3)
    dL_dopacity = z(opa, P, 1)
    dL_dextra = z(ext and F > 0, P, F)
    if P == 0:
        return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dscales, dL_drotations, dL_dextra
    # Every tensor whose raw pointer crosses the C ABI is made contiguous fp32 first, like the forward does: the
    # reference's cameras are `.transpose(0, 1)` VIEWS (scene/cameras.py:81-86, strides (1, 4)), and the reference
    # binding calls .contiguous() on them (rasterize_points.cu:231-247; Q17: it forgets scales/rotations).
    fc = lambda t, name: _f32c(t, name) if (t is not None and t.numel()) else t
    means3D, background = _f32c(means3D, "means3D"), _f32c(background, "background")
    viewmatrix, projmatrix, campos = _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix"), _f32c(campos, "campos")
    scales, rotations, transMat_precomp = fc(scales, "scales"), fc(rotations, "rotations"), fc(transMat_precomp, "transMat_precomp")
    sh, colors, extra_attrs = fc(sh, "sh"), fc(colors, "colors"), fc(extra_attrs, "extra_attrs")
    if radii.dtype != torch.int32 or not radii.is_contiguous():
        radii = radii.to(torch.int32).contiguous()
    cot = [None if t is None else _f32c(t, "cotangent") for t in (dL_dout_color, dL_dout_others, dL_dout_extra)]
    dense_needed = any(t is not None for t in cot)
    stream = _stream()
    if dense_needed and (geo or col or opa or ext):
        a = _lib.IsrBackwardArgs()
        a.P, a.sh_degree, a.sh_coeffs, a.F, a.W, a.H = P, int(degree), M, F, W, H
        a.flags, a.grad_mask, a.num_rendered = flags, grad_mask, int(R)
        a.tan_fovx, a.tan_fovy, a.scale_modifier = float(tan_fovx), float(tan_fovy), float(scale_modifier)
        a.background, a.viewmatrix, a.projmatrix, a.campos = _ptr(background), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos)
        a.means3D, a.scales, a.rotations, a.transMat_precomp = _ptr(means3D), _ptr(scales), _ptr(rotations), _ptr(transMat_precomp)
        a.shs, a.colors_precomp, a.extra_attrs = _ptr(sh), _ptr(colors), _ptr(extra_attrs)
        a.radii, a.geom, a.image, a.binning = _ptr(radii), _ptr(geomBuffer), _ptr(imageBuffer), _ptr(binningBuffer)
        a.dL_dcolor, a.dL_dothers, a.dL_dextra_pix = _ptr(cot[0]), _ptr(cot[1]), (_ptr(cot[2]) if F > 0 else None)
        a.dL_dmeans2D, a.dL_dnormal, a.dL_dopacity, a.dL_dcolors = _ptr(dL_dmeans2D), _ptr(dL_dnormal), _ptr(dL_dopacity), _ptr(dL_dcolors)
        a.dL_dmeans3D, a.dL_dtransMat, a.dL_dsh = _ptr(dL_dmeans3D), _ptr(dL_dtransMat), _ptr(dL_dsh)
        a.dL_dscales, a.dL_drotations, a.dL_dextra = _ptr(dL_dscales), _ptr(dL_drotations), _ptr(dL_dextra)
        _lib.check(L.isr_backward(C.byref(a), stream), "isr_backward")
    if sparse_extra is not None and ext and F > 0:
        pix_ids, rows = sparse_extra
        pix_ids = pix_ids.to(torch.int32).contiguous()
        rows = _f32c(rows, "sparse cotangent rows")
        _lib.check(L.isr_backward_extra_sparse(P, F, W, H, _ptr(extra_attrs), _ptr(geomBuffer), _ptr(imageBuffer),
                                               _ptr(binningBuffer), int(R), int(pix_ids.numel()), _ptr(pix_ids),
                                               _ptr(rows), _ptr(dL_dextra), flags & _lib.FLAG_SPEC_ARITH, stream),
                   "isr_backward_extra_sparse")
    if debug:
        torch.cuda.synchronize()
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dscales, dL_drotations, dL_dextra


def c_mark_visible(means3D, viewmatrix, projmatrix):
    """markVisible (DSR/rasterize_points.cu:264-283)."""
    L = _require_cuda_lib()
    P = int(means3D.shape[0])
    present = torch.zeros(P, dtype=torch.bool, device=means3D.device)
    if P:
        means3D, viewmatrix, projmatrix = _f32c(means3D, "means3D"), _f32c(viewmatrix, "viewmatrix"), _f32c(projmatrix, "projmatrix")
        _lib.check(L.isr_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix), present.data_ptr(), _stream()),
                   "isr_mark_visible")
    return present


class _CNamespace:
    """Stand-in for the reference's pybind module `diff_surfel_rasterization._C`."""
    rasterize_gaussians = staticmethod(c_rasterize_gaussians)
    rasterize_gaussians_backward = staticmethod(c_rasterize_gaussians_backward)
    mark_visible = staticmethod(c_mark_visible)


_C = _CNamespace()


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        extra_attrs, raster_settings):
    out = _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                    cov3Ds_precomp, extra_attrs, raster_settings)
    color, radii, depth, extra, pairs, handle = out
    if extra.numel():
        extra._isr_handle = handle  # consumed by instascene_b200.sample_pixels
    return color, radii, depth, extra, pairs


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                extra_attrs, raster_settings):
        F = extra_attrs.shape[1] if extra_attrs.shape[0] != 0 else 0
        args = (raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, extra_attrs, F,
                raster_settings.viewmatrix, raster_settings.projmatrix, raster_settings.tanfovx,
                raster_settings.tanfovy, raster_settings.image_height, raster_settings.image_width, sh,
                raster_settings.sh_degree, raster_settings.campos, raster_settings.prefiltered, raster_settings.debug)
        want_pairs = getattr(raster_settings, "want_pairs", True)
        geom_state = getattr(raster_settings, "_geom_state", None)  # phase A launched early by render()
        if geom_state is not None:
            # ctx keeps raster_settings alive and the state owns this Function's output tensors: drop the link, or
            # ctx <-> outputs form a reference cycle that only the cyclic GC frees (hundreds of MB per view)
            raster_settings._geom_state = None
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                res = c_rasterize_gaussians(*args, want_pairs=want_pairs)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            res = c_rasterize_gaussians(*args, want_pairs=want_pairs, geom_state=geom_state)
        (num_rendered, color, depth, radii, extra, geomBuffer, binningBuffer, imgBuffer, gau_related_pixels,
         gau_pixel_indices) = res
        if want_pairs and not getattr(raster_settings, "defer_pairs", False):
            gau_related_pixels = gau_related_pixels[:(int(gau_pixel_indices.item()) + 1)]
        elif want_pairs:
            # render() slices lazily on first access (avoids a host sync right after the blend)
            gau_related_pixels._isr_count_minus_1 = gau_pixel_indices
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.arith = _arith_flag if geom_state is None else (geom_state.args.flags & _lib.FLAG_SPEC_ARITH)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, extra_attrs, sh,
                              geomBuffer, binningBuffer, imgBuffer)
        H, W = int(raster_settings.image_height), int(raster_settings.image_width)
        # zero-stride [H*W, F] handle: only its (hybrid-sparse) gradient is ever used
        handle = torch.zeros(1, dtype=torch.float32, device=means3D.device).expand(H * W, max(F, 1))
        ctx.mark_non_differentiable(radii, gau_related_pixels)
        return color, radii, depth, extra, gau_related_pixels, handle

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth, grad_out_extra, grad_pairs, grad_handle):
        num_rendered = ctx.num_rendered
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, extra_attrs, sh, geomBuffer,
         binningBuffer, imgBuffer) = ctx.saved_tensors
        # input order: means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, extra_attrs
        need = ctx.needs_input_grad
        mask = 0
        if need[0] or need[1] or need[5] or need[6] or need[7]:
            mask |= _lib.GRAD_GEOMETRY | _lib.GRAD_COLOR  # SH backward also feeds dL_dmeans3D
        if need[2] or need[3]:
            mask |= _lib.GRAD_COLOR | (_lib.GRAD_GEOMETRY if need[2] else 0)
        if need[4]:
            mask |= _lib.GRAD_OPACITY
        if need[8]:
            mask |= _lib.GRAD_EXTRA
        sparse = None
        if grad_handle is not None:
            if grad_handle.is_sparse and (mask & ~_lib.GRAD_EXTRA) == 0:
                # only the features are trainable: their gradient needs just the sampled pixels
                sparse = (grad_handle._indices()[0], grad_handle._values())
            else:  # other gradients are wanted too (or someone densified it): fold into the dense cotangent
                if grad_handle.is_sparse:
                    grad_handle = grad_handle.to_dense()
                F = extra_attrs.shape[1]
                dense = grad_handle[:, :F].t().reshape(F, rs.image_height, rs.image_width)
                grad_out_extra = dense if grad_out_extra is None else grad_out_extra + dense
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, extra_attrs, rs.scale_modifier,
                cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, grad_depth,
                grad_out_extra, sh, rs.sh_degree, rs.campos, geomBuffer, num_rendered, binningBuffer, imgBuffer,
                rs.debug)
        kw = dict(grad_mask=mask, image_size=(int(rs.image_height), int(rs.image_width)), sparse_extra=sparse,
                  arith=ctx.arith)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                res = c_rasterize_gaussians_backward(*args, **kw)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            res = c_rasterize_gaussians_backward(*args, **kw)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations, grad_extra_attrs) = res
        grads = (grad_means3D if need[0] else None, grad_means2D if need[1] else None, grad_sh if need[2] else None,
                 grad_colors_precomp if need[3] else None, grad_opacities if need[4] else None,
                 grad_scales if need[5] else None, grad_rotations if need[6] else None,
                 grad_cov3Ds_precomp if need[7] else None, grad_extra_attrs if need[8] else None, None)
        return grads


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return c_mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, extra_attrs=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        dev = means3D.device
        empty = lambda: torch.empty(0, dtype=torch.float32, device=dev)
        if shs is None:
            shs = empty()
        if colors_precomp is None:
            colors_precomp = empty()
        if scales is None:
            scales = empty()
        if rotations is None:
            rotations = empty()
        if cov3D_precomp is None:
            cov3D_precomp = empty()
        if extra_attrs is None:
            extra_attrs = empty()
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   extra_attrs, rs)


# --------------------------------------------------------------------------------------------------------------
class _SamplePixels(torch.autograd.Function):
    """features[n,F] = extra_map[:, pix_ids]; the gradient travels as a hybrid-sparse [H*W, F] tensor on the
    rasterizer's hidden handle so that the backward touches only the sampled pixels."""

    @staticmethod
    def forward(ctx, handle, extra_map, pix_ids):
        L = _require_cuda_lib()
        F = int(extra_map.shape[0])
        HW = int(extra_map.shape[1] * extra_map.shape[2])
        ids32 = pix_ids.to(torch.int32).contiguous()
        out = torch.empty((ids32.numel(), F), dtype=torch.float32, device=extra_map.device)
        src = extra_map.detach().contiguous()
        _lib.check(L.isr_gather_pixels(F, HW, src.data_ptr(), int(ids32.numel()), _ptr(ids32), _ptr(out), _stream()),
                   "isr_gather_pixels")
        ctx.save_for_backward(pix_ids)
        ctx.shape = (HW, int(handle.shape[1]))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (pix_ids,) = ctx.saved_tensors
        g = torch.sparse_coo_tensor(pix_ids.reshape(1, -1).to(torch.int64), grad_out.contiguous(), size=ctx.shape,
                                    check_invariants=False)
        return g, None, None


class _RowNorm(torch.autograd.Function):
    """y = x / (|x| + eps1) [then y / (|y| + eps2)] over the rows of [P,F], one fused kernel each way.
    `sink` (a tensor object, normally x itself): the backward does NOT run the chain rule; it parks dL/dy on
    `sink._isr_deferred_dy` (+ `_isr_deferred_cfg`) for an optimizer that applies it inside its update
    (instascene_b200.FusedAdam) and returns no gradient for x."""

    @staticmethod
    def forward(ctx, x, eps1, eps2, stages, sink=None):
        L = _require_cuda_lib()
        xc = _f32c(x.detach(), "x")
        P, F = int(xc.shape[0]), int(xc.shape[1])
        y = torch.empty_like(xc)
        _lib.check(L.isr_rownorm_forward(P, F, _ptr(xc), float(eps1), float(eps2), int(stages), _ptr(y), _stream()),
                   "isr_rownorm_forward")
        ctx.save_for_backward(xc)
        ctx.cfg = (float(eps1), float(eps2), int(stages))
        ctx.sink = sink
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _require_cuda_lib()
        (xc,) = ctx.saved_tensors
        e1, e2, stages = ctx.cfg
        dyc = _f32c(dy, "dy")
        if ctx.sink is not None:
            prev = getattr(ctx.sink, "_isr_deferred_dy", None)
            if prev is not None and getattr(ctx.sink, "_isr_deferred_cfg", None) != ctx.cfg:
                raise RuntimeError("deferred row-normalisation gradients with different settings on one parameter")
            ctx.sink._isr_deferred_dy = dyc if prev is None else prev + dyc
            ctx.sink._isr_deferred_cfg = ctx.cfg
            return None, None, None, None, None
        dx = torch.empty_like(xc)
        _lib.check(L.isr_rownorm_backward(int(xc.shape[0]), int(xc.shape[1]), _ptr(xc), _ptr(dyc), e1, e2, stages,
                                          _ptr(dx), _stream()), "isr_rownorm_backward")
        return dx, None, None, None, None


def normalize_rows(x: torch.Tensor, eps1: float, eps2: float = 0.0, stages: int = 1, defer_to=None) -> torch.Tensor:
    """Fused `x / (x.norm(dim=-1, keepdim=True) + eps1)`, optionally applied twice (second eps = eps2).
    defer_to: see _RowNorm (the gradient of x is left to an optimizer that fuses the chain rule)."""
    if x.numel() == 0:
        return x
    return _RowNorm.apply(x, eps1, eps2, stages, defer_to)


def sample_pixels(extra_map: torch.Tensor, pix_ids: torch.Tensor) -> torch.Tensor:
    """Gather rows [n, F] of a rendered [F,H,W] feature map at flat pixel ids (= W*y + x).
    Equivalent to `extra_map.reshape(F, -1)[:, pix_ids].T` (train_semantic.py:124-129 after the mask gather)."""
    handle = getattr(extra_map, "_isr_handle", None)
    if handle is None or not handle.requires_grad:
        return extra_map.reshape(extra_map.shape[0], -1)[:, pix_ids.long()].t()
    return _SamplePixels.apply(handle, extra_map, pix_ids)


def sample_labelled_pixels(labels_flat: torch.Tensor, n: int, generator=None):
    """Draw `n` pixel ids uniformly WITH replacement among the pixels whose label is > 0 -- the sampling of
    train_semantic.py:118-129 (`valid = segmap > 0; idx = randint(0, len(valid), (n,))`) without the boolean-mask
    gather of the [F,H,W] map and without a host sync: one `torch.rand` (the caller's generator) + two kernels
    (occupancy words + block counts, then per-sample rank selection; csrc/isr_sampler.cu).
    Returns (pix_ids int64 [n], labels int32 [n])."""
    L = _require_cuda_lib()
    lab = labels_flat.reshape(-1)
    if not lab.is_cuda:
        raise RuntimeError("labels must be a CUDA tensor")
    if lab.dtype not in (torch.int8, torch.int16, torch.int32, torch.int64):
        lab = lab.to(torch.int32)
    lab = lab.contiguous()
    HW = int(lab.numel())
    dev = lab.device
    u = torch.rand(n, device=dev, generator=generator)
    pix = torch.empty(n, dtype=torch.int64, device=dev)
    out_lab = torch.empty(n, dtype=torch.int32, device=dev)
    ws_bytes = L.isr_sampler_workspace_bytes(HW)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(L.isr_sample_labelled(lab.data_ptr(), lab.element_size(), HW, int(n), _ptr(u), ws.data_ptr(), ws_bytes,
                                     pix.data_ptr(), out_lab.data_ptr(), _stream()), "isr_sample_labelled")
    return pix, out_lab


def sample_labelled_pixels_torch(labels_flat: torch.Tensor, n: int, generator=None):
    """Round 1's formulation of the same draw in torch ops (inclusive scan of the mask + searchsorted); kept as the
    checker of the fused kernels: same `torch.rand` stream -> identical pixel ids."""
    mask = labels_flat > 0
    csum = torch.cumsum(mask, dim=0, dtype=torch.int32)
    n_valid = csum[-1]
    u = torch.rand(n, device=labels_flat.device, generator=generator)
    target = torch.clamp((u * n_valid).to(torch.int32), max=n_valid - 1) + 1  # rank (1-based) of the chosen valid pixel
    pix = torch.searchsorted(csum, target, right=False)
    pix = torch.clamp(pix, max=labels_flat.numel() - 1)
    return pix, labels_flat[pix]
