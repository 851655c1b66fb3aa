"""ctypes binding of libisr.so (include/isr.h).  Fails loudly when the CUDA library is missing: there is no
CPU or PyTorch fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libisr.so")

ISR_OK = 0
FLAG_BWD_WH_QUIRK = 1
FLAG_NO_PAIRS = 2
FLAG_SKIP_BINNING = 4
FLAG_SPEC_ARITH = 8
FLAG_SKIP_BLEND = 16
GRAD_GEOMETRY, GRAD_COLOR, GRAD_OPACITY, GRAD_EXTRA, GRAD_ALL = 1, 2, 4, 8, 15
MAX_EXTRA_DIMS = 32

# enum IsrField
GEOM_SPLAT, GEOM_RGB, GEOM_DEPTH, GEOM_TILES, GEOM_CLAMPED, GEOM_DEPTH_ORDER, GEOM_OFFSETS, GEOM_TILE_COUNT = 0, 1, 2, 3, 4, 5, 6, 7
IMG_FINAL_T, IMG_NCONTRIB, IMG_RANGES = 16, 17, 18
BIN_POINT_LIST = 32

_vp, _fp, _ip = C.c_void_p, C.c_void_p, C.c_void_p  # raw device pointers travel as integers


class IsrForwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("sh_degree", C.c_int), ("sh_coeffs", C.c_int), ("F", C.c_int), ("W", C.c_int), ("H", C.c_int),
        ("flags", C.c_uint),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
        ("background", _fp), ("viewmatrix", _fp), ("projmatrix", _fp), ("campos", _fp),
        ("means3D", _fp), ("opacities", _fp), ("scales", _fp), ("rotations", _fp), ("transMat_precomp", _fp),
        ("shs", _fp), ("colors_precomp", _fp), ("extra_attrs", _fp),
        ("geom", _vp), ("geom_bytes", C.c_size_t), ("image", _vp), ("image_bytes", C.c_size_t),
        ("binning", _vp), ("binning_bytes", C.c_size_t),
        ("radii", _ip), ("out_color", _fp), ("out_others", _fp), ("out_extra", _fp),
        ("pairs", _ip), ("pair_capacity", C.c_int64), ("pair_count", _ip), ("num_rendered_host", _vp),
    ]


class IsrBackwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("sh_degree", C.c_int), ("sh_coeffs", C.c_int), ("F", C.c_int), ("W", C.c_int), ("H", C.c_int),
        ("flags", C.c_uint), ("grad_mask", C.c_uint), ("num_rendered", C.c_int64),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
        ("background", _fp), ("viewmatrix", _fp), ("projmatrix", _fp), ("campos", _fp),
        ("means3D", _fp), ("scales", _fp), ("rotations", _fp), ("transMat_precomp", _fp),
        ("shs", _fp), ("colors_precomp", _fp), ("extra_attrs", _fp),
        ("radii", _ip), ("geom", _vp), ("image", _vp), ("binning", _vp),
        ("dL_dcolor", _fp), ("dL_dothers", _fp), ("dL_dextra_pix", _fp),
        ("dL_dmeans2D", _fp), ("dL_dnormal", _fp), ("dL_dopacity", _fp), ("dL_dcolors", _fp), ("dL_dmeans3D", _fp),
        ("dL_dtransMat", _fp), ("dL_dsh", _fp), ("dL_dscales", _fp), ("dL_drotations", _fp), ("dL_dextra", _fp),
    ]


class IsrSparseView(C.Structure):
    _fields_ = [("geom", _vp), ("image", _vp), ("binning", _vp)]


MAX_SPARSE_VIEWS = 8

EXPORTED_SYMBOLS = [
    "isr_version", "isr_status_string", "isr_last_cuda_error", "isr_device_sm_count", "isr_kernel_launch_count",
    "isr_geom_bytes", "isr_image_bytes", "isr_binning_bytes", "isr_field_offset",
    "isr_forward_geometry", "isr_forward_render", "isr_backward", "isr_backward_extra_sparse", "isr_forward_sparse_extra",
    "isr_backward_sparse_extra_views", "isr_mark_visible",
    "isr_gather_pixels", "isr_sampler_workspace_bytes", "isr_sample_labelled", "isr_contrastive_workspace_bytes", "isr_contrastive_forward", "isr_contrastive_backward",
    "isr_rownorm_forward", "isr_rownorm_backward", "isr_aux_maps_forward", "isr_aux_maps_backward", "isr_adam_step", "isr_adam_rownorm_step", "isr_knn_workspace_bytes", "isr_knn_mean_dist2",
    "isr_tracker_workspace_bytes", "isr_tracker_mark", "isr_tracker_fill",
    "isr_photometric_workspace_bytes", "isr_photometric_forward", "isr_photometric_backward",
    "isr_densify_stats",
]

_lib = None


class IsrError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads libisr.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IsrError(
            f"{LIB_PATH} is missing: build it with `python -m instascene_b200.build` "
            "(nvcc, sm_100a).  instascene_b200 has no CPU/PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    L.isr_status_string.restype = C.c_char_p
    L.isr_kernel_launch_count.restype = C.c_longlong
    L.isr_geom_bytes.restype = C.c_size_t
    L.isr_geom_bytes.argtypes = [C.c_int]
    L.isr_image_bytes.restype = C.c_size_t
    L.isr_image_bytes.argtypes = [C.c_int, C.c_int]
    L.isr_binning_bytes.restype = C.c_size_t
    L.isr_binning_bytes.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int]
    L.isr_field_offset.restype = C.c_int64
    L.isr_field_offset.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int]
    L.isr_forward_geometry.argtypes = [C.POINTER(IsrForwardArgs), C.c_void_p]
    L.isr_forward_render.argtypes = [C.POINTER(IsrForwardArgs), C.c_int64, C.c_void_p]
    L.isr_backward.argtypes = [C.POINTER(IsrBackwardArgs), C.c_void_p]
    L.isr_backward_extra_sparse.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _fp, _vp, _vp, _vp, C.c_int64, C.c_int,
                                            _ip, _fp, _fp, C.c_uint, C.c_void_p]
    L.isr_forward_sparse_extra.argtypes = [C.c_int, C.POINTER(IsrSparseView), C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_int,
                                           _ip, _ip, _fp, C.c_uint, C.c_void_p]
    L.isr_backward_sparse_extra_views.argtypes = [C.c_int, C.POINTER(IsrSparseView), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  _ip, _ip, _fp, _fp, C.c_uint, C.c_void_p]
    L.isr_mark_visible.argtypes = [C.c_int, _fp, _fp, _fp, _vp, C.c_void_p]
    L.isr_gather_pixels.argtypes = [C.c_int, C.c_int64, _fp, C.c_int, _ip, _fp, C.c_void_p]
    L.isr_sampler_workspace_bytes.restype = C.c_size_t
    L.isr_sampler_workspace_bytes.argtypes = [C.c_int64]
    L.isr_sample_labelled.argtypes = [_vp, C.c_int, C.c_int64, C.c_int, _fp, _vp, C.c_size_t, _ip, _ip, C.c_void_p]
    L.isr_contrastive_workspace_bytes.restype = C.c_size_t
    L.isr_contrastive_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    L.isr_contrastive_forward.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _ip, _fp, C.c_float, C.c_int, _vp, C.c_size_t, _fp,
                                          C.c_void_p]
    L.isr_contrastive_backward.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _ip, _fp, _vp, _fp, _fp, C.c_void_p]
    L.isr_rownorm_forward.argtypes = [C.c_int, C.c_int, _fp, C.c_float, C.c_float, C.c_int, _fp, C.c_void_p]
    L.isr_rownorm_backward.argtypes = [C.c_int, C.c_int, _fp, _fp, C.c_float, C.c_float, C.c_int, _fp, C.c_void_p]
    L.isr_aux_maps_forward.argtypes = [C.c_int, C.c_int, _fp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, _fp, _fp, _fp,
                                       _fp, _fp, C.c_void_p]
    L.isr_aux_maps_backward.argtypes = [C.c_int, C.c_int, _fp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, _fp, _fp, _fp,
                                        _fp, _fp, _fp, C.c_void_p]
    L.isr_adam_step.argtypes = [C.c_size_t, _fp, _fp, _fp, _fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _ip, C.c_void_p]
    L.isr_adam_rownorm_step.argtypes = [C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, C.c_float, C.c_float, C.c_int, C.c_float,
                                        C.c_float, C.c_float, C.c_float, C.c_int, _ip, C.c_void_p]
    L.isr_knn_workspace_bytes.restype = C.c_size_t
    L.isr_knn_workspace_bytes.argtypes = [C.c_int]
    L.isr_knn_mean_dist2.argtypes = [C.c_int, _fp, _fp, _vp, C.c_size_t, C.c_void_p]
    L.isr_tracker_workspace_bytes.restype = C.c_size_t
    L.isr_tracker_workspace_bytes.argtypes = [C.c_int, C.c_int]
    L.isr_tracker_mark.argtypes = [_ip, C.c_int64, _ip, C.c_int64, C.c_int, C.c_int, _vp, C.c_size_t, _ip, C.c_void_p]
    L.isr_tracker_fill.argtypes = [C.c_int, C.c_int, _vp, _ip, _ip, C.c_void_p]
    L.isr_photometric_workspace_bytes.restype = C.c_size_t
    L.isr_photometric_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    L.isr_photometric_forward.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_float, _vp, C.c_size_t, _fp, C.c_void_p]
    L.isr_photometric_backward.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_float, _vp, _fp, _fp, C.c_void_p]
    L.isr_densify_stats.argtypes = [C.c_int, _ip, _fp, _fp, _fp, _fp, C.c_void_p]
    _lib = L
    return L


def check(status: int, what: str) -> None:
    if status != ISR_OK:
        L = lib()
        msg = L.isr_status_string(status).decode()
        extra = f" (cudaError {L.isr_last_cuda_error()})" if status == -4 else ""
        raise IsrError(f"{what}: {msg}{extra}")
