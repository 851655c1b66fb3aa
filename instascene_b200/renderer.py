"""Drop-in mirror of the reference ``gaussian_renderer.render`` (gaussian_renderer/__init__.py:20-169) and of
``utils.point_utils.depth_to_normal`` (utils/point_utils.py:10-40) on top of the B200 rasterizer.

`viewpoint_camera`, `pc` and `pipe` are duck-typed exactly like the reference uses them:
  camera: FoVx, FoVy, image_height, image_width, world_view_transform, full_proj_transform, camera_center
  pc:     get_xyz, get_opacity, get_seg_feature, get_scaling, get_rotation, get_features, active_sh_degree,
          max_sh_degree, get_covariance(scaling_modifier)
  pipe:   compute_cov3D_python, convert_SHs_python (forced False, Q10), depth_ratio
The returned dict has the reference's 13 keys.
"""
from __future__ import annotations

import math

import torch

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, normalize_rows


class _SettingsNoPairs(GaussianRasterizationSettings):
    """Same 12 fields; tells the rasterizer not to emit gau_related_pixels (no post-blend host sync)."""
    want_pairs = False


class _SettingsDeferPairs(GaussianRasterizationSettings):
    """Same 12 fields; the pair list comes back unsliced with its device-side count (sliced lazily)."""
    defer_pairs = True


class RenderPackage(dict):
    """The reference's 13-key result dict.  The six rasterizer outputs are stored eagerly; the seven derived maps
    (gaussian_renderer/__init__.py:127-167) and the sliced pair list are computed on FIRST ACCESS -- the semantic
    training loop (train_semantic.py:102-108) never reads them, the RGB loop (train.py) reads them all.  Values,
    autograd history and key set are identical to the eager reference dict."""
    _LAZY = ("rend_alpha", "rend_normal", "rend_dist", "surf_depth", "surf_normal", "rend_depth", "rend_median_depth")

    def __init__(self, eager, aux_fn, pairs_fn):
        super().__init__(eager)
        self._aux_fn, self._pairs_fn = aux_fn, pairs_fn

    def _materialise(self, key):
        if key == "gau_related_pixels" and self._pairs_fn is not None:
            fn, self._pairs_fn = self._pairs_fn, None
            super().__setitem__("gau_related_pixels", fn())
        elif key in self._LAZY and self._aux_fn is not None:
            fn, self._aux_fn = self._aux_fn, None
            for k, v in fn().items():
                super().__setitem__(k, v)

    def materialise(self):
        self._materialise("gau_related_pixels")
        self._materialise("rend_alpha")
        return self

    def __getitem__(self, key):
        self._materialise(key)
        return super().__getitem__(key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __contains__(self, key):
        return super().__contains__(key) or (key in self._LAZY and self._aux_fn is not None)

    def keys(self):
        return self.materialise() and super().keys()

    def items(self):
        return self.materialise() and super().items()

    def values(self):
        return self.materialise() and super().values()

    def __iter__(self):
        self.materialise()
        return super().__iter__()

    def __len__(self):
        return super().__len__() + (len(self._LAZY) if self._aux_fn is not None else 0)


def depths_to_points(view, depthmap):
    """utils/point_utils.py:10-26"""
    dev = depthmap.device
    c2w = (view.world_view_transform.T).inverse()
    W, H = view.image_width, view.image_height
    ndc2pix = torch.tensor([[W / 2, 0, 0, W / 2], [0, H / 2, 0, H / 2], [0, 0, 0, 1]], dtype=torch.float32, device=dev).T
    projection_matrix = c2w.T @ view.full_proj_transform
    intrins = (projection_matrix @ ndc2pix)[:3, :3].T
    grid_x, grid_y = torch.meshgrid(torch.arange(W, device=dev).float(), torch.arange(H, device=dev).float(),
                                    indexing='xy')
    points = torch.stack([grid_x, grid_y, torch.ones_like(grid_x)], dim=-1).reshape(-1, 3)
    rays_d = points @ intrins.inverse().T @ c2w[:3, :3].T
    rays_o = c2w[:3, 3]
    return depthmap.reshape(-1, 1) * rays_d + rays_o


def depth_to_normal(view, depth):
    """utils/point_utils.py:29-40"""
    points = depths_to_points(view, depth).reshape(*depth.shape[1:], 3)
    output = torch.zeros_like(points)
    dx = points[2:, 1:-1] - points[:-2, 1:-1]
    dy = points[1:-1, 2:] - points[1:-1, :-2]
    normal_map = torch.nn.functional.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    output[1:-1, 1:-1, :] = normal_map
    return output


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
           norm_seg_feat=True, want_pairs: bool = True):
    """Render the scene (reference signature + `want_pairs`).  Background tensor (bg_color) must be on GPU!"""
    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device=pc.get_xyz.device)
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    cls = _SettingsDeferPairs if want_pairs else _SettingsNoPairs
    raster_settings = cls(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx,
        tanfovy=tanfovy,
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=False,
    )
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)

    means3D = pc.get_xyz
    means2D = screenspace_points
    opacity = pc.get_opacity
    raw = getattr(pc, "_seg_feature", None)
    if norm_seg_feat and raw is not None and getattr(pipe, "fused_seg_activation", True) \
            and getattr(pc, "seg_feature_eps", 1e-6) is not None:
        # get_seg_feature (x / (|x| + 1e-6), scene/gaussian_model.py:121-125) and the renormalisation below it in the
        # reference render() (eps 1e-9) fused into one kernel each way (SURVEY.md Q9 / build-plan step 7)
        seg_feature = normalize_rows(raw, getattr(pc, "seg_feature_eps", 1e-6), 1e-9, stages=2)
    else:
        seg_feature = pc.get_seg_feature
        if seg_feature is not None and norm_seg_feat:
            seg_feature = normalize_rows(seg_feature, 1e-9)

    scales = None
    rotations = None
    cov3D_precomp = None
    if pipe.compute_cov3D_python:
        splat2world = pc.get_covariance(scaling_modifier)
        W, H = viewpoint_camera.image_width, viewpoint_camera.image_height
        near, far = viewpoint_camera.znear, viewpoint_camera.zfar
        ndc2pix = torch.tensor([[W / 2, 0, 0, (W - 1) / 2], [0, H / 2, 0, (H - 1) / 2], [0, 0, far - near, near],
                                [0, 0, 0, 1]], dtype=torch.float32, device=means3D.device).T
        world2pix = viewpoint_camera.full_proj_transform @ ndc2pix
        cov3D_precomp = (splat2world[:, [0, 1, 3]] @ world2pix[:, [0, 1, 3]]).permute(0, 2, 1).reshape(-1, 9)
    else:
        scales = pc.get_scaling
        rotations = pc.get_rotation

    pipe.convert_SHs_python = False  # Q10: the reference mutates the caller's object too
    shs = None
    colors_precomp = None
    if override_color is None:
        shs = pc.get_features
    else:
        colors_precomp = override_color

    rendered_image, radii, allmap, extra_attrs, gau_related_pixels = rasterizer(
        means3D=means3D, means2D=means2D, shs=shs, colors_precomp=colors_precomp, opacities=opacity, scales=scales,
        rotations=rotations, cov3D_precomp=cov3D_precomp, extra_attrs=seg_feature)

    eager = {"render": rendered_image, "viewspace_points": means2D, "visibility_filter": radii > 0, "radii": radii,
             "seg_feature": extra_attrs, "gau_related_pixels": gau_related_pixels}

    def aux():
        render_alpha = allmap[1:2]
        render_normal = allmap[2:5]
        render_normal = (render_normal.permute(1, 2, 0) @ (viewpoint_camera.world_view_transform[:3, :3].T)).permute(2, 0, 1)
        render_depth_median = torch.nan_to_num(allmap[5:6], 0, 0)
        render_depth_expected = torch.nan_to_num(allmap[0:1] / render_alpha, 0, 0)
        render_dist = allmap[6:7]
        surf_depth = render_depth_expected * (1 - pipe.depth_ratio) + (pipe.depth_ratio) * render_depth_median
        surf_normal = depth_to_normal(viewpoint_camera, surf_depth).permute(2, 0, 1)
        surf_normal = surf_normal * (render_alpha).detach()
        return {'rend_alpha': render_alpha, 'rend_normal': render_normal, 'rend_dist': render_dist,
                'surf_depth': surf_depth, 'surf_normal': surf_normal, "rend_depth": render_depth_expected,
                "rend_median_depth": render_depth_median}

    pairs_fn = None
    cnt = getattr(gau_related_pixels, "_isr_count_minus_1", None)
    if cnt is not None:
        pairs_fn = lambda: gau_related_pixels[:(int(cnt.item()) + 1)]
    if not getattr(pipe, "lazy_outputs", True):
        return RenderPackage(eager, aux, pairs_fn).materialise()
    return RenderPackage(eager, aux, pairs_fn)
