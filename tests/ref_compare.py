"""Field-by-field comparison of this rasterizer with the UNMODIFIED reference CUDA rasterizer (baseline/_ref, built by
baseline/build_ref.sh) on the same inputs, at any size, on the GPU (torch only: no Python loops over tiles).

The reference's opaque byte buffers are decoded with the layouts of GeometryState / ImageState / BinningState
(DSR/cuda_rasterizer/rasterizer_impl.cu:155-194, `obtain` = 128-byte aligned consecutive arrays,
rasterizer_impl.h:20-27).  Used by tests/test_reference_scale_gpu.py and tools/parity_at_scale.py."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from instascene_b200 import _lib, synth  # noqa: E402


def reference_C():
    """The reference's pybind module (diff_surfel_rasterization._C), or None if baseline/_ref is not installed."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import ref_loader
        if not ref_loader.available():
            return None
        ref_loader._install_stubs()
        if ref_loader.REF_DIR not in sys.path:
            sys.path.insert(0, ref_loader.REF_DIR)
        from diff_surfel_rasterization import _C
        return _C
    except Exception:  # noqa: BLE001
        return None


def _carve(buf: torch.Tensor, specs):
    """obtain(): consecutive arrays, each aligned to 128 bytes (the torch allocation itself is 512-byte aligned)."""
    out, off = {}, buf.data_ptr() % 128
    base = off
    for name, dtype, count in specs:
        off = (off + 127) // 128 * 128
        nbytes = count * torch.empty(0, dtype=dtype).element_size()
        out[name] = buf[off - base: off - base + nbytes].view(dtype)
        off += nbytes
    return out


def ref_geom_fields(geom: torch.Tensor, P: int):
    return _carve(geom, [("depths", torch.float32, P), ("clamped", torch.uint8, 3 * P), ("radii", torch.int32, P),
                         ("means2D", torch.float32, 2 * P), ("transMat", torch.float32, 9 * P),
                         ("normal_opacity", torch.float32, 4 * P), ("rgb", torch.float32, 3 * P),
                         ("tiles_touched", torch.int32, P)])


def ref_img_fields(img: torch.Tensor, HW: int):
    return _carve(img, [("accum_alpha", torch.float32, 3 * HW), ("n_contrib", torch.int32, 2 * HW), ("ranges", torch.int32, 2 * HW)])


def ref_point_list(binning: torch.Tensor, R: int):
    return _carve(binning, [("point_list", torch.int32, R)])["point_list"]


def my_fields(geom, img, binning, P, R_emitted, W, H):
    L = _lib.lib()
    HW, tiles = W * H, ((W + 15) // 16) * ((H + 15) // 16)
    off = lambda f: int(L.isr_field_offset(f, P, R_emitted, W, H))

    def view(buf, o, dtype, count):
        return buf[o: o + count * torch.empty(0, dtype=dtype).element_size()].view(dtype)

    splat = view(geom, off(_lib.GEOM_SPLAT), torch.float32, 16 * P).view(P, 16)
    return dict(splat=splat, rgb=view(geom, off(_lib.GEOM_RGB), torch.float32, 4 * P).view(P, 4),
                depths=view(geom, off(_lib.GEOM_DEPTH), torch.float32, P),
                tiles_touched=view(geom, off(_lib.GEOM_TILES), torch.int32, P),
                clamped=view(geom, off(_lib.GEOM_CLAMPED), torch.uint8, P),
                tiles_emitted=view(geom, off(_lib.GEOM_TILE_COUNT), torch.int32, P),
                final_T=view(img, off(_lib.IMG_FINAL_T), torch.float32, 3 * HW),
                n_contrib=view(img, off(_lib.IMG_NCONTRIB), torch.int32, 2 * HW),
                ranges=view(img, off(_lib.IMG_RANGES), torch.int32, 2 * tiles).view(tiles, 2),
                point_list=view(binning, off(_lib.BIN_POINT_LIST), torch.int32, R_emitted))


def _bits_differ(a: torch.Tensor, b: torch.Tensor) -> int:
    return int((a.contiguous().view(torch.int32) != b.contiguous().view(torch.int32)).sum())


def float_stats(a, b):
    a, b = a.double(), b.double()
    d = (a - b).abs()
    scale = float(b.abs().max()) + 1e-30
    return {"bits_differ": _bits_differ(a.float(), b.float()), "max_abs": float(d.max()), "normwise_rel": float(d.max()) / scale,
            "n_abs_gt_1e-4_of_max": int((d > 1e-4 * scale).sum()), "numel": a.numel()}


def make_inputs(P, F, W, H, seed, view=7, dev="cuda:0"):
    sc = synth.synth_scene(P, F=F, seed=seed)
    cam = synth.ring_cameras(200, W, H)[view]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    e = torch.empty(0, device=dev)
    inp = dict(P=P, F=F, W=W, H=H, cam=cam, e=e, means=t(sc.xyz), opa=t(sc.opacities()).reshape(-1, 1), scales=t(sc.scales()),
               rots=t(sc.rotations()), shs=t(sc.shs()), extra=t(sc.seg_features()) if F else e,
               view=t(cam.world_view_transform), proj=t(cam.full_proj_transform), campos=t(cam.camera_center),
               bg=torch.tensor([0.1, 0.2, 0.3], device=dev))
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    inp["dcolor"] = torch.randn((3, H, W), device=dev, generator=g)
    inp["dothers"] = torch.randn((7, H, W), device=dev, generator=g)
    inp["dextra"] = torch.randn((F, H, W), device=dev, generator=g) if F else e
    return inp


def run_mine(inp, backward=True):
    from instascene_b200.rasterizer import c_rasterize_gaussians, c_rasterize_gaussians_backward
    i, e, cam = inp, inp["e"], inp["cam"]
    fw = c_rasterize_gaussians(i["bg"], i["means"], e, i["opa"], i["scales"], i["rots"], 1.0, e, i["extra"], i["F"], i["view"],
                               i["proj"], cam.tanfovx, cam.tanfovy, i["H"], i["W"], i["shs"], 3, i["campos"], False, False,
                               return_args=True)
    bw = None
    if backward:
        bw = c_rasterize_gaussians_backward(i["bg"], i["means"], fw[3], e, i["scales"], i["rots"], i["extra"], 1.0, e, i["view"],
                                            i["proj"], cam.tanfovx, cam.tanfovy, i["dcolor"], i["dothers"],
                                            i["dextra"] if i["F"] else None, i["shs"], 3, i["campos"], fw[5], fw[0], fw[6],
                                            fw[7], False, image_size=(i["H"], i["W"]))
    return fw, bw


def run_ref(ref_C, inp, extra=None, dextra=None, dcolor=None, dothers=None, backward=True):
    i, e, cam = inp, inp["e"], inp["cam"]
    extra = i["extra"] if extra is None else extra
    dextra = i["dextra"] if dextra is None else dextra
    dcolor = i["dcolor"] if dcolor is None else dcolor
    dothers = i["dothers"] if dothers is None else dothers
    Fx = extra.shape[1] if extra.numel() else 0
    fw = ref_C.rasterize_gaussians(i["bg"], i["means"], e, i["opa"], i["scales"], i["rots"], 1.0, e, extra, Fx, i["view"],
                                   i["proj"], cam.tanfovx, cam.tanfovy, i["H"], i["W"], i["shs"], 3, i["campos"], False, False)
    bw = None
    if backward:
        bw = ref_C.rasterize_gaussians_backward(i["bg"], i["means"], fw[3], e, i["scales"], i["rots"], extra, 1.0, e, i["view"],
                                                i["proj"], cam.tanfovx, cam.tanfovy, dcolor, dothers, dextra if Fx else e,
                                                i["shs"], 3, i["campos"], fw[5], fw[0], fw[6], fw[7], False)
    return fw, bw


def compare_forward(inp, mfw, rfw, r_extra=None):
    """Returns a report dict.  Every `*_mismatch` / `bits_differ` entry counts elements that differ BITWISE."""
    P, F, W, H = inp["P"], inp["F"], inp["W"], inp["H"]
    HW, tiles_x = W * H, (W + 15) // 16
    tiles = tiles_x * ((H + 15) // 16)
    R_ref = int(rfw[0])
    a = mfw[-1]
    mine = my_fields(mfw[5], mfw[7], mfw[6], P, int(a._n_inst), W, H)
    rg, ri = ref_geom_fields(rfw[5], P), ref_img_fields(rfw[7], HW)
    rep = {"P": P, "F": F, "W": W, "H": H, "num_rendered": [int(mfw[0]), R_ref], "emitted_instances": int(a._n_inst)}
    vis = rfw[3] > 0
    rep["visible"] = int(vis.sum())
    rep["radii_mismatch"] = int((mfw[3] != rfw[3]).sum())
    rep["tiles_touched_mismatch"] = int((mine["tiles_touched"][vis] != rg["tiles_touched"][vis]).sum())
    sp = mine["splat"]
    rep["depths_bits_differ"] = _bits_differ(mine["depths"][vis], rg["depths"][vis])
    rep["transMat_bits_differ"] = _bits_differ(sp[:, :9][vis], rg["transMat"].view(P, 9)[vis])
    rep["means2D_bits_differ"] = _bits_differ(sp[:, 9:11][vis], rg["means2D"].view(P, 2)[vis])
    rep["normal_opacity_bits_differ"] = _bits_differ(sp[:, 11:15][vis], rg["normal_opacity"].view(P, 4)[vis])
    rep["rgb_bits_differ"] = _bits_differ(mine["rgb"][:, :3][vis], rg["rgb"].view(P, 3)[vis])
    rc = rg["clamped"].view(P, 3).to(torch.int32)
    rep["clamped_mismatch"] = int(((rc[:, 0] | (rc[:, 1] << 1) | (rc[:, 2] << 2))[vis] != mine["clamped"].to(torch.int32)[vis]).sum())

    # ---- tile lists: mine must be an ordered subsequence of the reference's list of the same tile ------------------
    r_ranges = ri["ranges"][:2 * tiles].view(tiles, 2).long()
    r_list = ref_point_list(rfw[6], R_ref).long()
    r_len = r_ranges[:, 1] - r_ranges[:, 0]
    rep["ref_ranges_cover_list"] = int(r_len.sum()) == R_ref
    tile_ids = torch.arange(tiles, device=r_list.device)
    r_tile = torch.repeat_interleave(tile_ids, r_len)
    # instances in list order: the reference's sorted list is tile-major, so position within the tile = index - start
    r_order_start = torch.zeros(tiles, dtype=torch.long, device=r_list.device)
    nz = r_len > 0
    r_order_start[nz] = r_ranges[nz, 0]
    # (the sorted list is grouped by ascending tile id, so r_tile built from ascending ranges matches it only if the
    # ranges are ascending -- true: identifyTileRanges writes them from the sorted keys)
    r_pos = torch.arange(R_ref, device=r_list.device) - r_order_start[r_tile]
    r_key = r_tile * P + r_list
    r_key_sorted, r_perm = torch.sort(r_key)
    m_ranges = mine["ranges"].long()
    m_len = m_ranges[:, 1] - m_ranges[:, 0]
    m_raw = mine["point_list"].long()
    packed = P < (1 << 24)
    m_gid = (m_raw & 0xFFFFFF) if packed else m_raw
    rep["my_ranges_cover_list"] = int(m_len.sum()) == int(a._n_inst)
    m_tile = torch.repeat_interleave(tile_ids, m_len)
    m_key = m_tile * P + m_gid
    loc = torch.searchsorted(r_key_sorted, m_key).clamp(max=max(R_ref - 1, 0))
    found = r_key_sorted[loc] == m_key
    rep["emitted_not_in_reference_list"] = int((~found).sum())
    m_refpos = r_pos[r_perm[loc]]                       # position of each of my entries in the reference's tile list
    same_tile = m_tile[1:] == m_tile[:-1]
    rep["tile_list_order_violations"] = int(((m_refpos[1:] <= m_refpos[:-1]) & same_tile).sum())

    # ---- per-pixel state: final_T / M1 / M2 bitwise, contributors mapped to reference list positions -----------------
    rep["final_T_bits_differ"] = _bits_differ(mine["final_T"], ri["accum_alpha"])
    ys = torch.arange(H, device=r_list.device).view(H, 1).expand(H, W)
    xs = torch.arange(W, device=r_list.device).view(1, W).expand(H, W)
    pix_tile = ((ys // 16) * tiles_x + xs // 16).reshape(-1)
    m_start = m_ranges[:, 0][pix_tile]
    for k, name in ((0, "last_contributor"), (1, "median_contributor")):
        mc = mine["n_contrib"][k * HW:(k + 1) * HW].long()
        rc_ = ri["n_contrib"][k * HW:(k + 1) * HW].long()
        idx = (m_start + mc - 1).clamp(min=0, max=max(int(a._n_inst) - 1, 0))
        mapped = torch.where(mc > 0, m_refpos[idx] + 1, torch.zeros_like(mc)) if int(a._n_inst) else torch.zeros_like(mc)
        bad = mapped != rc_
        rep[f"{name}_mismatch_px"] = int(bad.sum())
        if k == 1:
            # Pixels nothing contributed to: the reference stores (uint32_t)(float)-1 there (forward.cu:322,449; undefined
            # in C++, whatever the compiler made of it) and never reads it (the backward loop is empty for them)
            r_last = ri["n_contrib"][:HW].long()
            rep["median_contributor_mismatch_px_with_contributors"] = int((bad & (r_last > 0)).sum())
            if int(bad.sum()):
                ids = torch.nonzero(bad).flatten()[:6]
                rep["median_mismatch_samples(pix,mine,mapped,ref,ref_last)"] = [
                    [int(i), int(mc[i]), int(mapped[i]), int(rc_[i]), int(r_last[i])] for i in ids]

    # ---- outputs ----------------------------------------------------------------------------------------------------
    rep["color"] = float_stats(mfw[1], rfw[1])
    rep["others"] = float_stats(mfw[2], rfw[2])
    chan = ["depth_w", "alpha", "normal_x", "normal_y", "normal_z", "median_depth", "distortion"]
    rep["others_bits_differ_per_channel"] = {c: _bits_differ(mfw[2][i], rfw[2][i]) for i, c in enumerate(chan)}
    if F:
        rep["extra"] = float_stats(mfw[4], rfw[4] if r_extra is None else r_extra)
    n_m, n_r = int(mfw[9].item()) + 1, int(rfw[9].item()) + 1
    pm, pr = mfw[8][:n_m].long(), rfw[8][:n_r].long()
    km = torch.sort(pm[:, 0] * HW + pm[:, 1])[0]
    kr = torch.sort(pr[:, 0] * HW + pr[:, 1])[0]
    if n_m == n_r:
        sym = int((km != kr).sum())
    else:
        both = torch.cat([torch.unique(km), torch.unique(kr)]).unique(return_counts=True)[1]
        sym = int((both == 1).sum())
    rep["pairs"] = {"mine": n_m, "ref": n_r, "symmetric_difference": sym}
    return rep


GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dtransMat", "dL_dsh", "dL_dscales", "dL_drotations"]


def compare_backward(inp, mbw, rbw, r_dextra=None):
    rep = {}
    for k, a, b in zip(GRAD_NAMES, mbw[:8], rbw[:8]):
        b = b.reshape(a.shape)
        st = float_stats(a, b)
        d = (a.double() - b.double()).abs().reshape(a.shape[0], -1).amax(1)
        rep[k] = {"normwise_rel": st["normwise_rel"], "gaussians_abs_gt_1e-4_of_max": int((d > 1e-4 * (float(b.abs().max()) + 1e-30)).sum()),
                  "max": float(b.abs().max())}
    if inp["F"]:
        b = rbw[8] if r_dextra is None else r_dextra
        st = float_stats(mbw[8], b)
        rep["dL_dextra"] = {"normwise_rel": st["normwise_rel"], "gaussians_abs_gt_1e-4_of_max": st["n_abs_gt_1e-4_of_max"], "max": float(b.abs().max())}
    return rep


INT_KEYS = ["radii_mismatch", "tiles_touched_mismatch", "emitted_not_in_reference_list", "tile_list_order_violations",
            "last_contributor_mismatch_px", "median_contributor_mismatch_px", "clamped_mismatch"]
BIT_KEYS = ["depths_bits_differ", "transMat_bits_differ", "means2D_bits_differ", "normal_opacity_bits_differ", "rgb_bits_differ",
            "final_T_bits_differ"]
