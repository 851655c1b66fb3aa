"""ctypes front-end of the CPU oracle (oracle/isr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
All arrays are numpy, C-contiguous; float32 / int32 / uint32 / uint8 as in the reference buffers.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libisr_oracle.so")
_lib = None

FLAG_BWD_WH_QUIRK = 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "isr_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-B", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_bin_count.restype = C.c_int64
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a: Optional[np.ndarray]):
    if a is None:
        return C.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
    return C.c_void_p(a.ctypes.data)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(int(n)))


def mark_visible(means3D, viewmatrix) -> np.ndarray:
    means3D, viewmatrix = _f32(means3D), _f32(viewmatrix)
    P = means3D.shape[0]
    out = np.zeros(P, dtype=np.uint8)
    lib().orc_mark_visible(C.c_int(P), _p(means3D), _p(viewmatrix), _p(out))
    return out.astype(bool)


def forward(means3D, opacities, viewmatrix, projmatrix, campos, W: int, H: int, bg,
            scales=None, rotations=None, scale_modifier: float = 1.0, shs=None, sh_degree: int = 0,
            colors_precomp=None, transMat_precomp=None, extra_attrs=None,
            want_pairs: bool = True, blend: bool = True, tile_stride: int = 1) -> Dict[str, np.ndarray]:
    """Full forward (K1..K6).  Returns every public output and every intermediate buffer."""
    L = lib()
    means3D, opacities = _f32(means3D), _f32(opacities).reshape(-1)
    viewmatrix, projmatrix, campos, bg = _f32(viewmatrix), _f32(projmatrix), _f32(campos), _f32(bg)
    scales, rotations, shs = _f32(scales), _f32(rotations), _f32(shs)
    colors_precomp, transMat_precomp, extra_attrs = _f32(colors_precomp), _f32(transMat_precomp), _f32(extra_attrs)
    P = means3D.shape[0]
    M = 0 if shs is None else shs.shape[1]
    F = 0 if extra_attrs is None or extra_attrs.size == 0 else extra_attrs.shape[1]
    radii = np.zeros(P, np.int32)
    means2D = np.zeros((P, 2), np.float32)
    depths = np.zeros(P, np.float32)
    transMats = np.zeros((P, 9), np.float32)
    rgb = np.zeros((P, 3), np.float32)
    normal_opacity = np.zeros((P, 4), np.float32)
    tiles_touched = np.zeros(P, np.uint32)
    clamped = np.zeros((P, 3), np.uint8)
    L.orc_preprocess_forward(C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(scales),
                             C.c_float(scale_modifier), _p(rotations), _p(opacities), _p(shs),
                             _p(transMat_precomp), _p(colors_precomp), _p(viewmatrix), _p(projmatrix),
                             _p(campos), C.c_int(W), C.c_int(H), _p(radii), _p(means2D), _p(depths),
                             _p(transMats), _p(rgb), _p(normal_opacity), _p(tiles_touched), _p(clamped))
    point_offsets = np.zeros(P, np.uint32)
    R = int(L.orc_bin_count(C.c_int(P), _p(tiles_touched), _p(point_offsets)))
    gx, gy = (W + 15) // 16, (H + 15) // 16
    keys = np.zeros(max(R, 1), np.uint64)
    point_list = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    L.orc_bin(C.c_int(P), C.c_int(W), C.c_int(H), _p(radii), _p(means2D), _p(depths), _p(point_offsets),
              C.c_int64(R), _p(keys), _p(point_list), _p(ranges))
    out = dict(radii=radii, means2D=means2D, depths=depths, transMats=transMats, rgb=rgb,
               normal_opacity=normal_opacity, tiles_touched=tiles_touched, clamped=clamped,
               point_offsets=point_offsets, num_rendered=R, keys=keys[:R], point_list=point_list[:R],
               ranges=ranges)
    if not blend:
        return out
    colors = colors_precomp if colors_precomp is not None else rgb
    tms = transMat_precomp if transMat_precomp is not None else transMats
    HW = H * W
    final_T = np.zeros((3, H, W), np.float32)
    n_contrib = np.zeros((2, H, W), np.uint32)
    out_color = np.zeros((3, H, W), np.float32)
    out_others = np.zeros((7, H, W), np.float32)
    out_extra = np.zeros((F, H, W), np.float32)
    cap = 9 * HW if want_pairs else 0
    pairs = np.full((max(cap, 1), 2), -1, np.int32)
    cnt = C.c_int64(0)
    L.orc_blend_forward(C.c_int(W), C.c_int(H), C.c_int(F), _p(ranges), _p(point_list), _p(means2D),
                        _p(colors), _p(tms), _p(extra_attrs if F else None), _p(normal_opacity), _p(bg),
                        _p(final_T), _p(n_contrib), _p(out_color), _p(out_others), _p(out_extra),
                        _p(pairs) if want_pairs else C.c_void_p(0), C.c_int64(cap), C.byref(cnt), C.c_int(tile_stride))
    out.update(final_T=final_T, n_contrib=n_contrib, color=out_color, others=out_others, extra=out_extra,
               pairs=pairs[: min(cnt.value, cap)], pair_count=int(cnt.value))
    return out


def backward(fwd: Dict[str, np.ndarray], means3D, viewmatrix, projmatrix, campos, W: int, H: int, bg,
             tan_fovx: float, tan_fovy: float, dL_dcolor, dL_dothers, dL_dextra=None,
             scales=None, rotations=None, scale_modifier: float = 1.0, shs=None, sh_degree: int = 0,
             colors_precomp=None, transMat_precomp=None, extra_attrs=None,
             flags: int = FLAG_BWD_WH_QUIRK, tile_stride: int = 1, pixel_mask=None,
             preprocess: bool = True) -> Dict[str, np.ndarray]:
    """Full backward (K7 + K8) from the buffers returned by forward().  pixel_mask (uint8 [H,W]): walk only those pixels
    (the cotangents must be zero elsewhere); preprocess=False stops after K7 (feature-only training needs no K8)."""
    L = lib()
    means3D = _f32(means3D)
    viewmatrix, projmatrix, campos, bg = _f32(viewmatrix), _f32(projmatrix), _f32(campos), _f32(bg)
    scales, rotations, shs = _f32(scales), _f32(rotations), _f32(shs)
    colors_precomp, transMat_precomp, extra_attrs = _f32(colors_precomp), _f32(transMat_precomp), _f32(extra_attrs)
    dL_dcolor, dL_dothers, dL_dextra = _f32(dL_dcolor), _f32(dL_dothers), _f32(dL_dextra)
    P = means3D.shape[0]
    M = 0 if shs is None else shs.shape[1]
    F = 0 if extra_attrs is None or extra_attrs.size == 0 else extra_attrs.shape[1]
    colors = colors_precomp if colors_precomp is not None else fwd["rgb"]
    tms = transMat_precomp if transMat_precomp is not None else fwd["transMats"]
    g = dict(dL_dmeans2D=np.zeros((P, 3), np.float32), dL_dcolors=np.zeros((P, 3), np.float32),
             dL_dnormal=np.zeros((P, 3), np.float32), dL_dopacity=np.zeros((P, 1), np.float32),
             dL_dtransMat=np.zeros((P, 9), np.float32), dL_dsh=np.zeros((P, M, 3), np.float32),
             dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dscales=np.zeros((P, 2), np.float32),
             dL_drotations=np.zeros((P, 4), np.float32), dL_dextra=np.zeros((P, F), np.float32))
    point_list = np.ascontiguousarray(fwd["point_list"]) if fwd["num_rendered"] else np.zeros(1, np.uint32)
    L.orc_blend_backward(C.c_int(W), C.c_int(H), C.c_int(F), _p(fwd["ranges"]), _p(point_list), _p(bg),
                         _p(fwd["means2D"]), _p(fwd["normal_opacity"]), _p(tms), _p(colors),
                         _p(extra_attrs if F else None), _p(fwd["final_T"]), _p(fwd["n_contrib"]),
                         _p(dL_dcolor), _p(dL_dothers), _p(dL_dextra if F else None), _p(g["dL_dtransMat"]),
                         _p(g["dL_dmeans2D"]), _p(g["dL_dnormal"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]),
                         _p(g["dL_dextra"]), C.c_int(tile_stride),
                         _p(None if pixel_mask is None else np.ascontiguousarray(pixel_mask, dtype=np.uint8)))
    if not preprocess:
        return g
    g["dL_dmeans2D_raw"] = g["dL_dmeans2D"].copy()
    g["dL_dtransMat_raw"] = g["dL_dtransMat"].copy()
    L.orc_preprocess_backward(C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(fwd["radii"]),
                              _p(shs), _p(fwd["clamped"]), _p(scales), _p(rotations), C.c_float(scale_modifier),
                              _p(tms), _p(viewmatrix), _p(projmatrix), C.c_int(W), C.c_int(H),
                              C.c_float(tan_fovx), C.c_float(tan_fovy), _p(campos), C.c_int(flags),
                              _p(g["dL_dmeans2D"]), _p(g["dL_dnormal"]), _p(g["dL_dtransMat"]), _p(g["dL_dcolors"]),
                              _p(g["dL_dsh"]), _p(g["dL_dmeans3D"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]))
    return g


def knn_mean_dist2(points) -> np.ndarray:
    points = _f32(points)
    out = np.zeros(points.shape[0], np.float32)
    lib().orc_knn_mean_dist2(C.c_int(points.shape[0]), _p(points), _p(out))
    return out
