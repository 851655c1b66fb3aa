"""`render()` with the interface of the reference's gaussian_renderer.render (gaussian_renderer/__init__.py:20-169)
on top of the B200 rasterizer, plus `depth_to_normal` (utils/point_utils.py:10-40).

Duck-typed inputs, exactly the attributes the reference touches:
  camera  FoVx, FoVy, image_height, image_width, world_view_transform, full_proj_transform, camera_center
          (znear/zfar only with pipe.compute_cov3D_python)
  pc      get_xyz, get_opacity, get_seg_feature (or the raw parameter `_seg_feature`), get_scaling, get_rotation,
          get_features, active_sh_degree, get_covariance(scaling_modifier)
  pipe    compute_cov3D_python, convert_SHs_python (forced to False like the reference does, Q10), depth_ratio;
          optional switches of this implementation: lazy_outputs (default True), fused_seg_activation (default True)
Result: a dict-like `RenderPackage` with the reference's 13 keys.
"""
from __future__ import annotations

import math

import torch

import ctypes as C

from . import _lib
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, binning_capacity_hint, finish_binning,
                         launch_geometry, normalize_rows, sampled_features, _ptr, _require_cuda_lib, _stream)

import os

_GEOMETRY_FIRST = os.environ.get("ISR_GEOMETRY_FIRST", "1") != "0"
_AUX_KEYS = ("rend_alpha", "rend_normal", "rend_dist", "surf_depth", "surf_normal", "rend_depth", "rend_median_depth")


class _SettingsNoPairs(GaussianRasterizationSettings):
    """Same 12 fields; tells the rasterizer not to emit gau_related_pixels (no post-blend host sync)."""
    want_pairs = False


class _SettingsDeferPairs(GaussianRasterizationSettings):
    """Same 12 fields; the pair list comes back unsliced with its device-side count (sliced lazily)."""
    defer_pairs = True


class RenderPackage(dict):
    """The 13-key result of render().  The six rasterizer outputs are stored eagerly; the seven derived maps and the
    sliced pair list are produced on FIRST ACCESS: the semantic training loop (train_semantic.py:102-108) never reads
    them, the RGB loop (train.py) reads them all.  Values, autograd history and key set equal the eager dict."""

    def __init__(self, eager, aux_fn, pairs_fn):
        super().__init__(eager)
        self._aux_fn, self._pairs_fn = aux_fn, pairs_fn

    def _materialise(self, key):
        if key == "gau_related_pixels" and self._pairs_fn is not None:
            fn, self._pairs_fn = self._pairs_fn, None
            super().__setitem__("gau_related_pixels", fn())
        elif key in _AUX_KEYS and self._aux_fn is not None:
            fn, self._aux_fn = self._aux_fn, None
            for k, v in fn().items():
                super().__setitem__(k, v)

    def materialise(self):
        self._materialise("gau_related_pixels")
        self._materialise(_AUX_KEYS[0])
        return self

    def __getitem__(self, key):
        self._materialise(key)
        return super().__getitem__(key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __contains__(self, key):
        return super().__contains__(key) or (key in _AUX_KEYS and self._aux_fn is not None)

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
        return super().__len__() + (len(_AUX_KEYS) if self._aux_fn is not None else 0)


# ---- geometry helpers (utils/point_utils.py) ---------------------------------------------------------------------
def _pixel_rays(view, device):
    """Per-pixel ray directions (world space, un-normalised, z_view = 1) and the camera origin."""
    c2w = view.world_view_transform.T.inverse()
    W, H = view.image_width, view.image_height
    ndc2pix = torch.tensor([[W / 2, 0, 0, W / 2], [0, H / 2, 0, H / 2], [0, 0, 0, 1]], dtype=torch.float32, device=device).T
    intrins = ((c2w.T @ view.full_proj_transform) @ ndc2pix)[:3, :3].T
    xs = torch.arange(W, device=device, dtype=torch.float32)
    ys = torch.arange(H, device=device, dtype=torch.float32)
    gx, gy = torch.meshgrid(xs, ys, indexing="xy")
    pix_h = torch.stack([gx, gy, torch.ones_like(gx)], dim=-1).reshape(-1, 3)
    return pix_h @ intrins.inverse().T @ c2w[:3, :3].T, c2w[:3, 3]


def depths_to_points(view, depthmap):
    rays_d, rays_o = _pixel_rays(view, depthmap.device)
    return depthmap.reshape(-1, 1) * rays_d + rays_o


def depth_to_normal(view, depth):
    """Finite-difference normals of the back-projected depth map; the one-pixel border stays zero."""
    pts = depths_to_points(view, depth).reshape(*depth.shape[1:], 3)
    normals = torch.zeros_like(pts)
    d_row = pts[2:, 1:-1] - pts[:-2, 1:-1]
    d_col = pts[1:-1, 2:] - pts[1:-1, :-2]
    normals[1:-1, 1:-1, :] = torch.nn.functional.normalize(torch.cross(d_row, d_col, dim=-1), dim=-1)
    return normals


def _camera_aux_constants(camera):
    """Per-camera constants of the fused aux kernel as host arrays: M = world_view_transform[:3,:3].T (view->world
    normal rotation) and K with ray = [x, y, 1] @ K (utils/point_utils.py:10-24).  Cached on the camera object and
    keyed on the transforms' storage + version, so a static camera costs one device->host copy in its lifetime."""
    wvt, fpt = camera.world_view_transform, camera.full_proj_transform
    key = (wvt.data_ptr(), wvt._version, fpt.data_ptr(), fpt._version, int(camera.image_width), int(camera.image_height))
    cached = getattr(camera, "_isr_aux_consts", None)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    c2w = wvt.T.inverse()
    W, H = camera.image_width, camera.image_height
    ndc2pix = torch.tensor([[W / 2, 0, 0, W / 2], [0, H / 2, 0, H / 2], [0, 0, 0, 1]], dtype=torch.float32, device=wvt.device).T
    intrins = ((c2w.T @ fpt) @ ndc2pix)[:3, :3].T
    K = (intrins.inverse().T @ c2w[:3, :3].T).contiguous().cpu().reshape(-1).tolist()
    M = wvt[:3, :3].T.contiguous().cpu().reshape(-1).tolist()
    M_c, K_c = (C.c_float * 9)(*M), (C.c_float * 9)(*K)
    try:
        camera._isr_aux_consts = (key, M_c, K_c)
    except Exception:
        pass
    return M_c, K_c


class _AuxMaps(torch.autograd.Function):
    """allmap[7,H,W] -> (rend_normal, rend_depth, rend_median_depth, surf_depth, surf_normal), one kernel each way."""

    @staticmethod
    def forward(ctx, allmap, M_c, K_c, depth_ratio):
        L = _require_cuda_lib()
        am = allmap.detach().contiguous()
        H, W = int(am.shape[1]), int(am.shape[2])
        new = lambda ch: torch.empty((ch, H, W), dtype=torch.float32, device=am.device)
        rn, rd, rm, sd, sn = new(3), new(1), new(1), new(1), new(3)
        _lib.check(L.isr_aux_maps_forward(W, H, _ptr(am), M_c, K_c, float(depth_ratio), _ptr(rn), _ptr(rd), _ptr(rm), _ptr(sd),
                                          _ptr(sn), _stream()), "isr_aux_maps_forward")
        ctx.save_for_backward(am)
        ctx.consts = (M_c, K_c, float(depth_ratio))
        ctx.set_materialize_grads(False)
        return rn, rd, rm, sd, sn

    @staticmethod
    def backward(ctx, g_rn, g_rd, g_rm, g_sd, g_sn):
        L = _require_cuda_lib()
        (am,) = ctx.saved_tensors
        M_c, K_c, ratio = ctx.consts
        H, W = int(am.shape[1]), int(am.shape[2])
        gs = [None if g is None else g.contiguous().float() for g in (g_rn, g_rd, g_rm, g_sd, g_sn)]
        g_allmap = torch.empty_like(am)
        _lib.check(L.isr_aux_maps_backward(W, H, _ptr(am), M_c, K_c, ratio, _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]), _ptr(gs[3]),
                                           _ptr(gs[4]), _ptr(g_allmap), _stream()), "isr_aux_maps_backward")
        return g_allmap, None, None, None


def _derived_maps_fused(allmap, camera, depth_ratio):
    M_c, K_c = _camera_aux_constants(camera)
    rn, rd, rm, sd, sn = _AuxMaps.apply(allmap, M_c, K_c, depth_ratio)
    return {"rend_alpha": allmap[1:2], "rend_normal": rn, "rend_dist": allmap[6:7], "surf_depth": sd, "surf_normal": sn,
            "rend_depth": rd, "rend_median_depth": rm}


def _derived_maps(allmap, camera, depth_ratio):
    """allmap channels: 0 depth*w, 1 alpha, 2-4 view-space normal, 5 median depth, 6 distortion
    (DSR/cuda_rasterizer/auxiliary.h:24-28); post-processing of gaussian_renderer/__init__.py:127-156."""
    alpha = allmap[1:2]
    world_normal = (allmap[2:5].permute(1, 2, 0) @ camera.world_view_transform[:3, :3].T).permute(2, 0, 1)
    median = torch.nan_to_num(allmap[5:6], 0, 0)
    expected = torch.nan_to_num(allmap[0:1] / alpha, 0, 0)
    surf_depth = expected * (1 - depth_ratio) + depth_ratio * median
    surf_normal = depth_to_normal(camera, surf_depth).permute(2, 0, 1) * alpha.detach()
    return {"rend_alpha": alpha, "rend_normal": world_normal, "rend_dist": allmap[6:7], "surf_depth": surf_depth,
            "surf_normal": surf_normal, "rend_depth": expected, "rend_median_depth": median}


def _seg_features_for_raster(pc, pipe, norm_seg_feat):
    """get_seg_feature (x/(|x|+1e-6), scene/gaussian_model.py:121-125) followed by render()'s own renormalisation
    (eps 1e-9, gaussian_renderer/__init__.py:60-62) -- Q9.  With the raw parameter available both are one kernel."""
    raw = getattr(pc, "_seg_feature", None)
    if norm_seg_feat and raw is not None and getattr(pipe, "fused_seg_activation", True):
        # pipe.defer_seg_feature_grad: leave the chain rule of the two normalisations to instascene_b200.FusedAdam
        defer = raw if (getattr(pipe, "defer_seg_feature_grad", False) and raw.requires_grad) else None
        return normalize_rows(raw, getattr(pc, "seg_feature_eps", 1e-6), 1e-9, stages=2, defer_to=defer)
    feats = pc.get_seg_feature
    if feats is not None and norm_seg_feat:
        feats = normalize_rows(feats, 1e-9)
    return feats


def _precomputed_transmats(camera, pc, scaling_modifier, device):
    """pipe.compute_cov3D_python branch (gaussian_renderer/__init__.py:69-81)."""
    splat2world = pc.get_covariance(scaling_modifier)
    W, H = camera.image_width, camera.image_height
    near, far = camera.znear, camera.zfar
    ndc2pix = torch.tensor([[W / 2, 0, 0, (W - 1) / 2], [0, H / 2, 0, (H - 1) / 2], [0, 0, far - near, near], [0, 0, 0, 1]],
                           dtype=torch.float32, device=device).T
    world2pix = camera.full_proj_transform @ ndc2pix
    return (splat2world[:, [0, 1, 3]] @ world2pix[:, [0, 1, 3]]).permute(0, 2, 1).reshape(-1, 9)


def _raster_inputs(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, override_color, want_pairs):
    """Settings tuple + geometry / appearance arguments of one view, as render() assembles them (reference :36-100)."""
    xyz = pc.get_xyz
    settings_cls = _SettingsDeferPairs if want_pairs else _SettingsNoPairs
    settings = settings_cls(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center, prefiltered=False, debug=False)
    geometry = {}
    if pipe.compute_cov3D_python:
        geometry["cov3D_precomp"] = _precomputed_transmats(viewpoint_camera, pc, scaling_modifier, xyz.device)
    else:
        geometry["scales"], geometry["rotations"] = pc.get_scaling, pc.get_rotation
    pipe.convert_SHs_python = False  # the reference mutates the caller's object in the same way (Q10)
    appearance = {"shs": pc.get_features} if override_color is None else {"colors_precomp": override_color}
    return settings, xyz, pc.get_opacity, geometry, appearance


def _launch_phase_a(settings, xyz, opacity, geometry, appearance, viewpoint_camera, pc, bg_color, scaling_modifier, want_pairs,
                    pinned_counts=None, bin_capacity=0):
    none = torch.empty(0, dtype=torch.float32, device=xyz.device)
    return launch_geometry(
        bg_color, xyz, appearance.get("colors_precomp", none), opacity, geometry.get("scales", none),
        geometry.get("rotations", none), scaling_modifier, geometry.get("cov3D_precomp", none),
        viewpoint_camera.world_view_transform, viewpoint_camera.full_proj_transform, settings.tanfovx,
        settings.tanfovy, settings.image_height, settings.image_width, appearance.get("shs", none),
        pc.active_sh_degree, viewpoint_camera.camera_center, want_pairs=want_pairs, pinned_counts=pinned_counts,
        bin_capacity=bin_capacity)


class PrefetchedGeometry:
    """Phase A (projection, depth order, offsets, instance count) of a FUTURE view, already running on a side stream."""
    __slots__ = ("state", "key")

    def __init__(self, state, key):
        self.state, self.key = state, key


def _prefetch_key(viewpoint_camera, pc, scaling_modifier, want_pairs):
    return (viewpoint_camera.world_view_transform.data_ptr(), viewpoint_camera.full_proj_transform.data_ptr(),
            int(viewpoint_camera.image_height), int(viewpoint_camera.image_width), pc.get_xyz.data_ptr(),
            float(scaling_modifier), bool(want_pairs))


_prefetch_streams = {}
_prefetch_pinned = {}  # device index -> [ring of pinned int64[2] count buffers, next slot]: up to 4 views in flight


def _next_pinned_counts(dev_index):
    ring = _prefetch_pinned.get(dev_index)
    if ring is None:
        ring = _prefetch_pinned[dev_index] = [[torch.zeros(2, dtype=torch.int64).pin_memory() for _ in range(4)], 0]
    buf = ring[0][ring[1] % 4]
    ring[1] += 1
    return buf


def _discard_prefetched(handle) -> None:
    """A prefetched phase A that will not be consumed: its temporaries (opacity / scale / rotation activations, the SH
    concatenation) were allocated on the main stream and are still being read by the side stream, so the main stream is
    ordered behind the side stream's work before the caching allocator may hand those blocks out again."""
    st = getattr(handle, "state", None) if handle is not None else None
    if st is not None and st.ready_event is not None:
        torch.cuda.current_stream().wait_event(st.ready_event)


def prefetch_geometry(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
                      want_pairs: bool = True, stream=None, bin_ahead: bool = True) -> PrefetchedGeometry:
    """Starts phase A of render() -- and, once an instance-count high-water mark exists, the binning -- for a view that
    will be rendered LATER (typically the next training view, drawn one iteration ahead) on a side stream, so that it overlaps the latency-bound loss / backward / optimizer tail of the
    current iteration and the host never waits for the instance count in the middle of the next forward.  Valid
    whenever nothing phase A reads changes in between: positions, scales, rotations, opacities, SH -- i.e. in the
    semantic-feature training loop (train_semantic.py optimises `_seg_feature` only), NOT in RGB training.  Pass the
    handle to `render(..., prefetched=handle)` with the same camera object; results are identical to an inline render.
    The side stream first waits for everything already queued on the current stream (so every buffer it may recycle is
    no longer in use), then runs K1 + the depth sort."""
    dev = pc.get_xyz.device
    main = torch.cuda.current_stream()
    side = stream if stream is not None else _prefetch_streams.get(dev.index)
    if side is None:
        side = _prefetch_streams[dev.index] = torch.cuda.Stream(device=dev)
    with torch.no_grad():
        settings, xyz, opacity, geometry, appearance = _raster_inputs(viewpoint_camera, pc, pipe, bg_color, scaling_modifier,
                                                                      override_color, want_pairs)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            # the binning reads no features either: queue it behind phase A, into a workspace sized from the largest
            # instance count seen so far for this scene / image size (render() re-bins inline if that was too small)
            cap = binning_capacity_hint(xyz.shape[0], settings.image_width, settings.image_height) if bin_ahead else 0
            st = _launch_phase_a(settings, xyz, opacity, geometry, appearance, viewpoint_camera, pc, bg_color,
                                 scaling_modifier, want_pairs, pinned_counts=_next_pinned_counts(dev.index), bin_capacity=cap)
            if st is not None:
                st.ready_event = torch.cuda.Event()
                st.ready_event.record(side)
    return PrefetchedGeometry(st, _prefetch_key(viewpoint_camera, pc, scaling_modifier, want_pairs))


def render_sampled(viewpoint_cameras, pc, pipe, bg_color: torch.Tensor, pix_ids: torch.Tensor, view_ids=None,
                   scaling_modifier=1.0, norm_seg_feat=True, prefetched=None):
    """The `seg_feature` output of render() at SAMPLED pixels only, for 1..8 views of the same cloud in one launch.

    train_semantic.py renders the whole [F,H,W] feature map of a view (and of five more views every tenth iteration,
    :146-173) and then keeps `sample_batchsize` labelled pixels of it (:118-129); the image, depth and normal maps are
    never used by its losses.  Here the pixels are drawn first (`sample_labelled_pixels` on the label map) and only they
    are composited: per view projection + binning as in render(), then ONE kernel for the samples of all views -- a warp
    per sample walking its pixel's tile list with the arithmetic and order of the dense blend, so the rows are
    bit-identical to `render(...)["seg_feature"].reshape(F, -1)[:, pix].T`.  The gradient reaches `pc._seg_feature`
    through the sampled-pixel backward and the fused normalisation, exactly as with render() + sample_pixels().

    pix_ids: [n] flat pixel ids (= W*y + x); view_ids: [n] index into `viewpoint_cameras` (None: a single view).
    prefetched: optional list of `prefetch_geometry` handles (one per view, None where absent).
    Returns {"features": [n,F], "radii": [V,P] int32, "visibility_filter": [V,P] bool}."""
    cams = list(viewpoint_cameras) if isinstance(viewpoint_cameras, (list, tuple)) else [viewpoint_cameras]
    states = []
    for k, cam in enumerate(cams):
        handle = prefetched[k] if prefetched is not None else None
        settings, xyz, opacity, geometry, appearance = _raster_inputs(cam, pc, pipe, bg_color, scaling_modifier, None, False)
        if any(t.requires_grad for t in (xyz, opacity, *geometry.values(), *appearance.values())) and torch.is_grad_enabled():
            raise RuntimeError("render_sampled differentiates the semantic features only: freeze the geometry "
                               "(GaussianModel.training_setup of semantic training) or use render()")
        if (handle is not None and handle.state is not None
                and handle.key == _prefetch_key(cam, pc, scaling_modifier, False)):
            st = handle.state
        else:
            _discard_prefetched(handle)
            with torch.no_grad():
                cap = binning_capacity_hint(xyz.shape[0], settings.image_width, settings.image_height)
                st = _launch_phase_a(settings, xyz, opacity, geometry, appearance, cam, pc, bg_color, scaling_modifier, False,
                                     pinned_counts=_next_pinned_counts(xyz.device.index), bin_capacity=cap)
        if st is None:
            raise RuntimeError("render_sampled needs at least one Gaussian")
        states.append(st)
    ready = getattr(pc, "_isr_param_ready_event", None)
    if ready is not None:
        torch.cuda.current_stream().wait_event(ready)
    with torch.no_grad():
        for st in states:
            finish_binning(st)
    feats = sampled_features(_seg_features_for_raster(pc, pipe, norm_seg_feat), states, pix_ids, view_ids)
    radii = torch.stack([st.radii for st in states])
    return {"features": feats, "radii": radii, "visibility_filter": radii > 0}


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None,
           norm_seg_feat=True, want_pairs: bool = True, prefetched: "PrefetchedGeometry | None" = None):
    """Rasterise `pc` from `viewpoint_camera`.  `bg_color` must live on the GPU.  `prefetched`: handle of
    `prefetch_geometry` for this very view (ignored if it was made for different inputs)."""
    settings, xyz, opacity, geometry, appearance = _raster_inputs(viewpoint_camera, pc, pipe, bg_color, scaling_modifier,
                                                                  override_color, want_pairs)

    # Phase A (projection, depth order, offsets) does not read the semantic features: launch it first so that it can
    # overlap whatever still produces them (pc._isr_param_ready_event: e.g. the previous step's gradient all-reduce +
    # optimizer step running on another stream).  Same kernels, same order of results.
    # Dummy leaf that receives the densification proxy dL/dmean2D (reference :29-33).  The reference makes it require
    # grad unconditionally; here only when some geometry/appearance input does (the proxy is read by densification,
    # which optimises those).  With frozen geometry -- train_semantic.py -- that keeps the backward on the
    # feature-only path instead of computing and zero-filling ten per-Gaussian gradient buffers nobody reads.
    geom_trainable = torch.is_grad_enabled() and any(
        t.requires_grad for t in (xyz, opacity, *geometry.values(), *appearance.values()))
    screen_pts = torch.zeros_like(xyz, requires_grad=bool(geom_trainable))
    if (prefetched is not None and prefetched.state is not None and not geom_trainable
            and prefetched.key == _prefetch_key(viewpoint_camera, pc, scaling_modifier, want_pairs)):
        settings._geom_state = prefetched.state
    else:
        _discard_prefetched(prefetched)
    if getattr(settings, "_geom_state", None) is None and getattr(pipe, "geometry_first", _GEOMETRY_FIRST):
        settings._geom_state = _launch_phase_a(settings, xyz, opacity, geometry, appearance, viewpoint_camera, pc, bg_color,
                                               scaling_modifier, want_pairs)
    ready = getattr(pc, "_isr_param_ready_event", None)
    if ready is not None:
        torch.cuda.current_stream().wait_event(ready)

    image, radii, allmap, seg_map, pairs = GaussianRasterizer(raster_settings=settings)(
        means3D=xyz, means2D=screen_pts, opacities=opacity,
        extra_attrs=_seg_features_for_raster(pc, pipe, norm_seg_feat), **geometry, **appearance)

    eager = {"render": image, "viewspace_points": screen_pts, "visibility_filter": radii > 0, "radii": radii,
             "seg_feature": seg_map, "gau_related_pixels": pairs}
    count_m1 = getattr(pairs, "_isr_count_minus_1", None)
    pairs_fn = None if count_m1 is None else (lambda: pairs[:(int(count_m1.item()) + 1)])
    aux = _derived_maps_fused if getattr(pipe, "fused_aux_maps", True) else _derived_maps
    pkg = RenderPackage(eager, lambda: aux(allmap, viewpoint_camera, pipe.depth_ratio), pairs_fn)
    return pkg if getattr(pipe, "lazy_outputs", True) else pkg.materialise()
