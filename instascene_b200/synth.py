"""Seeded synthetic Gaussian clouds, cameras and label maps.

Everything is generated on the CPU with numpy's PCG64 so that the same seed gives the same
float32 inputs here, on the GPU box, for the oracle, for the reference and for the CUDA path.

Camera matrices follow the reference conventions exactly:
  world_view_transform = getWorld2View2(R, T).T          (scene/cameras.py:81, utils/graphics_utils.py:38-49)
  projection_matrix    = getProjectionMatrix(...).T      (scene/cameras.py:82-83, utils/graphics_utils.py:51-71)
  full_proj_transform  = world_view_transform @ projection_matrix   (scene/cameras.py:84-85)
  camera_center        = world_view_transform.inverse()[3, :3]      (scene/cameras.py:86)
i.e. both 4x4 matrices are stored TRANSPOSED (row-vector convention) and are indexed m[4*col+row]
by the kernels (DSR/cuda_rasterizer/auxiliary.h:80-99).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


@dataclass
class SynthCamera:
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: np.ndarray  # [4,4] float32, transposed W2V
    full_proj_transform: np.ndarray  # [4,4] float32
    camera_center: np.ndarray  # [3] float32
    R: np.ndarray = field(default=None)  # c2w rotation (reference Camera.R)
    T: np.ndarray = field(default=None)  # w2c translation (reference Camera.T)
    znear: float = 0.01
    zfar: float = 100.0

    @property
    def tanfovx(self) -> float:
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self) -> float:
        return math.tan(self.FoVy * 0.5)


@dataclass
class SynthScene:
    xyz: np.ndarray  # [P,3]
    scaling_raw: np.ndarray  # [P,2] log-scales (GaussianModel._scaling)
    rotation_raw: np.ndarray  # [P,4] unnormalised quaternion (w,x,y,z)
    opacity_raw: np.ndarray  # [P,1] logits
    features_dc: np.ndarray  # [P,1,3]
    features_rest: np.ndarray  # [P,15,3]
    seg_feature_raw: Optional[np.ndarray]  # [P,F] or None
    active_sh_degree: int = 3

    @property
    def P(self) -> int:
        return int(self.xyz.shape[0])

    @property
    def F(self) -> int:
        return 0 if self.seg_feature_raw is None else int(self.seg_feature_raw.shape[1])

    # activations as GaussianModel applies them (scene/gaussian_model.py:109-135, :45-60)
    def scales(self) -> np.ndarray:
        return np.exp(self.scaling_raw.astype(np.float64)).astype(np.float32)

    def rotations(self) -> np.ndarray:
        q = self.rotation_raw.astype(np.float64)
        return (q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)).astype(np.float32)

    def opacities(self) -> np.ndarray:
        return (1.0 / (1.0 + np.exp(-self.opacity_raw.astype(np.float64)))).astype(np.float32)

    def shs(self) -> np.ndarray:
        return np.ascontiguousarray(np.concatenate([self.features_dc, self.features_rest], axis=1))

    def seg_features(self) -> Optional[np.ndarray]:
        """get_seg_feature (eps 1e-6) followed by render()'s re-normalisation (eps 1e-9)  -- Q9."""
        if self.seg_feature_raw is None:
            return None
        f = self.seg_feature_raw.astype(np.float32)
        f = f / (np.linalg.norm(f, axis=1, keepdims=True).astype(np.float32) + np.float32(1e-6))
        f = f / (np.linalg.norm(f, axis=1, keepdims=True).astype(np.float32) + np.float32(1e-9))
        return np.ascontiguousarray(f.astype(np.float32))


def synth_scene(P: int, F: int = 0, seed: int = 1000, extent: float = 1.5, scale_mult: float = 1.0) -> SynthScene:
    rng = np.random.Generator(np.random.PCG64(seed))
    xyz = rng.uniform(-extent, extent, size=(P, 3)).astype(np.float32)
    s0 = 0.4 * (27.0 / max(P, 1)) ** (1.0 / 3.0) * scale_mult
    scaling = (math.log(s0) + 0.5 * rng.standard_normal(size=(P, 2))).astype(np.float32)
    rot = rng.standard_normal(size=(P, 4)).astype(np.float32)
    opa = (1.5 * rng.standard_normal(size=(P, 1))).astype(np.float32)
    fdc = (0.5 * rng.standard_normal(size=(P, 1, 3))).astype(np.float32)
    frest = (0.05 * rng.standard_normal(size=(P, 15, 3))).astype(np.float32)
    seg = None
    if F > 0:
        seg = rng.uniform(0.0, 1.0, size=(P, F)).astype(np.float32)
        seg = (seg / (np.linalg.norm(seg, axis=1, keepdims=True) + 1e-9)).astype(np.float32)
    return SynthScene(xyz, scaling, rot, opa, fdc, frest, seg)


def _projection_matrix(znear: float, zfar: float, fovX: float, fovY: float) -> np.ndarray:
    tanY, tanX = math.tan(fovY / 2), math.tan(fovX / 2)
    top, right = tanY * znear, tanX * znear
    bottom, left = -top, -right
    Pm = np.zeros((4, 4), dtype=np.float64)
    Pm[0, 0] = 2.0 * znear / (right - left)
    Pm[1, 1] = 2.0 * znear / (top - bottom)
    Pm[0, 2] = (right + left) / (right - left)
    Pm[1, 2] = (top + bottom) / (top - bottom)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm.astype(np.float32)


def make_camera(R: np.ndarray, T: np.ndarray, W: int, H: int, FoVx: float, FoVy: float,
                znear: float = 0.01, zfar: float = 100.0) -> SynthCamera:
    Rt = np.zeros((4, 4), dtype=np.float64)
    Rt[:3, :3] = R.T
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    w2v = np.float32(np.linalg.inv(np.linalg.inv(Rt)))  # getWorld2View2 with translate=0, scale=1
    wvt = np.ascontiguousarray(w2v.T)
    proj_t = np.ascontiguousarray(_projection_matrix(znear, zfar, FoVx, FoVy).T)
    full = (wvt.astype(np.float64) @ proj_t.astype(np.float64)).astype(np.float32)
    center = np.linalg.inv(wvt.astype(np.float64))[3, :3].astype(np.float32)
    return SynthCamera(W, H, FoVx, FoVy, wvt, np.ascontiguousarray(full), np.ascontiguousarray(center),
                       R=np.asarray(R, dtype=np.float64), T=np.asarray(T, dtype=np.float64), znear=znear, zfar=zfar)


def ring_cameras(n_views: int, W: int, H: int, radius: float = 4.0, fovx_deg: float = 60.0) -> List[SynthCamera]:
    """View i of n on a ring of radius 4 about the origin: azimuth 2*pi*i/n, elevation 20deg*sin(4*pi*i/n)."""
    FoVx = math.radians(fovx_deg)
    FoVy = 2.0 * math.atan(math.tan(FoVx / 2) * H / W)
    cams = []
    for i in range(n_views):
        az = 2.0 * math.pi * i / n_views
        el = math.radians(20.0) * math.sin(4.0 * math.pi * i / n_views)
        C = radius * np.array([math.cos(el) * math.cos(az), math.sin(el), math.cos(el) * math.sin(az)])
        f = -C / np.linalg.norm(C)
        up = np.array([0.0, 1.0, 0.0])
        r = np.cross(f, up)
        r /= np.linalg.norm(r)
        d = np.cross(f, r)
        R = np.stack([r, d, f], axis=1)  # c2w rotation, columns = camera x(right), y(down), z(forward)
        T = -R.T @ C
        cams.append(make_camera(R, T, W, H, FoVx, FoVy))
    return cams


def label_map(W: int, H: int, seed: int, grid: int = 8, zero_frac: float = 0.1) -> np.ndarray:
    """[H,W] int16: grid x grid rectangles labelled 1..grid^2, `zero_frac` of the pixels zeroed."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ys = np.minimum((np.arange(H) * grid) // H, grid - 1)
    xs = np.minimum((np.arange(W) * grid) // W, grid - 1)
    lab = (ys[:, None] * grid + xs[None, :] + 1).astype(np.int16)
    lab[rng.uniform(size=(H, W)) < zero_frac] = 0
    return lab


def gram_schmidt_prototypes(K: int, F: int, seed: int) -> np.ndarray:
    """Fixed class prototypes of --gram_feat_3d (scene/gaussian_model.py:158-176): Gram-Schmidt over rand(K,F).
    For K > F the reference's vectors beyond the F-th are normalised residual noise; reproduced as is."""
    rng = np.random.Generator(np.random.PCG64(seed))
    vs = rng.uniform(0.0, 1.0, size=(K, F)).astype(np.float32)
    out = []
    for v in vs:
        for u in out:
            v = v - np.float32(np.dot(v, u)) * u
        out.append((v / (np.linalg.norm(v) + np.float32(1e-9))).astype(np.float32))
    return np.stack(out).astype(np.float32)


def morton_labels(xyz: np.ndarray, K: int = 64) -> np.ndarray:
    """3D instance labels for the 3D contrastive term: 1 + (morton(xyz) mod K)."""
    mn, mx = xyz.min(0), xyz.max(0)
    q = np.clip(((xyz - mn) / np.maximum(mx - mn, 1e-9) * 1023.0).astype(np.uint32), 0, 1023)

    def prep(x):
        x = (x | (x << 16)) & 0x030000FF
        x = (x | (x << 8)) & 0x0300F00F
        x = (x | (x << 4)) & 0x030C30C3
        x = (x | (x << 2)) & 0x09249249
        return x

    code = prep(q[:, 0]) | (prep(q[:, 1]) << 1) | (prep(q[:, 2]) << 2)
    return (1 + (code % K)).astype(np.int64)
