"""Shared test helpers: run the CUDA path (through the reference-shaped API / C ABI) and the oracle on the same
seeded synthetic inputs and unpack the CUDA workspaces for field-by-field comparison."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from instascene_b200 import synth  # noqa: E402


def scene_inputs(P, F, W, H, seed, view=1, n_views=4, sh_degree=3, scale_mult=1.0):
    sc = synth.synth_scene(P, F=F, seed=seed, scale_mult=scale_mult)
    cam = synth.ring_cameras(n_views, W, H)[view]
    return dict(means3D=sc.xyz, opacities=sc.opacities(), scales=sc.scales(), rotations=sc.rotations(), shs=sc.shs(),
                extra_attrs=sc.seg_features(), sh_degree=sh_degree, viewmatrix=cam.world_view_transform,
                projmatrix=cam.full_proj_transform, campos=cam.camera_center, W=W, H=H,
                tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=np.array([0.1, 0.2, 0.3], np.float32), cam=cam, scene=sc)


def _variant_kwargs(inp):
    cp, tp = inp.get("colors_precomp"), inp.get("transMat_precomp")
    return dict(scales=None if tp is not None else inp["scales"], rotations=None if tp is not None else inp["rotations"],
                shs=None if cp is not None else inp["shs"], colors_precomp=cp, transMat_precomp=tp,
                scale_modifier=inp.get("scale_modifier", 1.0), sh_degree=inp["sh_degree"], extra_attrs=inp["extra_attrs"])


def oracle_forward(orc, inp, **kw):
    return orc.forward(inp["means3D"], inp["opacities"], inp["viewmatrix"], inp["projmatrix"], inp["campos"],
                       inp["W"], inp["H"], inp["bg"], **_variant_kwargs(inp), **kw)


def oracle_backward(orc, inp, fwd, dcolor, dothers, dextra, flags=1):
    return orc.backward(fwd, inp["means3D"], inp["viewmatrix"], inp["projmatrix"], inp["campos"], inp["W"], inp["H"],
                        inp["bg"], inp["tanfovx"], inp["tanfovy"], dcolor, dothers, dextra, flags=flags,
                        **_variant_kwargs(inp))


def cuda_forward(inp, want_pairs=True, device="cuda:0"):
    """Runs c_rasterize_gaussians (the reference-shaped binding over the C ABI) and unpacks everything to numpy.
    Optional keys of `inp`: colors_precomp [P,3], transMat_precomp [P,9], scale_modifier."""
    import torch
    from instascene_b200 import _lib
    from instascene_b200.rasterizer import c_rasterize_gaussians
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    e = torch.empty(0, dtype=torch.float32, device=device)
    F = 0 if inp["extra_attrs"] is None else inp["extra_attrs"].shape[1]
    cp, tp = inp.get("colors_precomp"), inp.get("transMat_precomp")
    tens = dict(bg=t(inp["bg"]), means3D=t(inp["means3D"]), opacities=t(inp["opacities"]),
                scales=e if tp is not None else t(inp["scales"]), rotations=e if tp is not None else t(inp["rotations"]),
                shs=e if cp is not None else t(inp["shs"]), extra=t(inp["extra_attrs"]) if F else e,
                colors=t(cp) if cp is not None else e, transmat=t(tp) if tp is not None else e,
                view=t(inp["viewmatrix"]), proj=t(inp["projmatrix"]), campos=t(inp["campos"]))
    res = c_rasterize_gaussians(tens["bg"], tens["means3D"], tens["colors"], tens["opacities"], tens["scales"],
                                tens["rotations"], inp.get("scale_modifier", 1.0), tens["transmat"], tens["extra"], F,
                                tens["view"], tens["proj"], inp["tanfovx"], inp["tanfovy"], inp["H"], inp["W"], tens["shs"],
                                inp["sh_degree"], tens["campos"], False, False, want_pairs=want_pairs)
    (R, color, others, radii, extra, geom, binning, img, pairs, pidx) = res
    torch.cuda.synchronize()
    L = _lib.lib()
    P, W, H = inp["means3D"].shape[0], inp["W"], inp["H"]
    HW, tiles = W * H, ((W + 15) // 16) * ((H + 15) // 16)
    off = lambda f: int(L.isr_field_offset(f, P, R, W, H))
    g, im, b = geom.cpu().numpy(), img.cpu().numpy(), binning.cpu().numpy()

    def view(buf, o, dtype, count):
        return np.frombuffer(buf.tobytes()[o:o + count * np.dtype(dtype).itemsize], dtype=dtype).copy()

    splat = view(g, off(_lib.GEOM_SPLAT), np.float32, P * 16).reshape(P, 16)
    rgb = view(g, off(_lib.GEOM_RGB), np.float32, P * 4).reshape(P, 4)[:, :3]
    out = dict(num_rendered=R, color=color.cpu().numpy(), others=others.cpu().numpy(), radii=radii.cpu().numpy(),
               extra=extra.cpu().numpy() if F else np.zeros((0, H, W), np.float32),
               transMats=splat[:, :9].copy(), means2D=splat[:, 9:11].copy(),
               normal_opacity=splat[:, 11:15].copy(), rgb=rgb.copy(),
               depths=view(g, off(_lib.GEOM_DEPTH), np.float32, P),
               tiles_touched=view(g, off(_lib.GEOM_TILES), np.uint32, P),
               clamped_mask=view(g, off(_lib.GEOM_CLAMPED), np.uint8, P),
               final_T=view(im, off(_lib.IMG_FINAL_T), np.float32, 3 * HW).reshape(3, H, W),
               n_contrib=view(im, off(_lib.IMG_NCONTRIB), np.uint32, 2 * HW).reshape(2, H, W),
               tiles_emitted=view(g, off(_lib.GEOM_TILE_COUNT), np.uint32, P),
               ranges=view(im, off(_lib.IMG_RANGES), np.uint32, 2 * tiles).reshape(tiles, 2))
    # the CUDA path emits only the tiles a Gaussian's footprint can reach: the list is shorter than num_rendered
    n_inst = int(out["tiles_emitted"].sum())
    raw = view(b, off(_lib.BIN_POINT_LIST), np.uint32, n_inst) if n_inst else np.zeros(0, np.uint32)
    # entries carry 8 per-block footprint bits above the 24-bit Gaussian id (P < 2^24)
    out["point_list"], out["block_bits"] = raw & np.uint32(0xFFFFFF), raw >> np.uint32(24)
    if want_pairs:
        n = int(pidx.item()) + 1
        out["pairs"] = pairs[:n].cpu().numpy()
        out["pair_count"] = n
    out["_torch"] = dict(tens=tens, geom=geom, binning=binning, img=img, radii=radii, R=R)
    return out


def check_tile_lists(c, o, W, H):
    """The CUDA path never emits a (tile, Gaussian) instance whose footprint provably misses the tile; the oracle (like
    the reference) emits every tile of the getRect rectangle.  Checks that
    (1) the tile ranges partition the CUDA instance list, whose length is the sum of the per-Gaussian emitted counts,
        and no Gaussian is emitted into more tiles than the reference touches,
    (2) every CUDA tile list is an ordered SUBSEQUENCE of the oracle's list of the same tile (same (depth, id) order),
    (3) the per-pixel last / median contributor indices, mapped from CUDA-list to oracle-list positions, are
        bit-identical to the oracle's n_contrib (both stop at the same Gaussian for every pixel)."""
    tiles_x, tiles_y = (W + 15) // 16, (H + 15) // 16
    cr, cl = c["ranges"].astype(np.int64), c["point_list"].astype(np.int64)
    orr, ol = o["ranges"].astype(np.int64), o["point_list"].astype(np.int64)
    assert np.all(c["tiles_emitted"] <= c["tiles_touched"])
    assert int((cr[:, 1] - cr[:, 0]).sum()) == len(cl), "ranges partition the list"
    mapped = np.zeros_like(c["n_contrib"])
    for t in range(tiles_x * tiles_y):
        tl = ol[orr[t, 0]:orr[t, 1]]
        pos = {int(g): i for i, g in enumerate(tl)}
        assert len(pos) == len(tl)
        sub = cl[cr[t, 0]:cr[t, 1]]
        idx = np.array([pos.get(int(g), -1) for g in sub], dtype=np.int64)
        assert np.all(idx >= 0), "emitted instance missing from the reference's tile list"
        assert np.all(np.diff(idx) > 0), "tile list is not in the reference's (depth, id) order"
        ty, tx = divmod(t, tiles_x)
        y0, x0 = ty * 16, tx * 16
        nb = c["n_contrib"][:, y0:y0 + 16, x0:x0 + 16].astype(np.int64)
        assert nb.max(initial=0) <= len(sub)
        lut = np.concatenate([[0], idx + 1]).astype(np.int64)  # CUDA position (1-based, 0 = none) -> oracle position
        mapped[:, y0:y0 + 16, x0:x0 + 16] = lut[nb]
    assert np.array_equal(mapped, o["n_contrib"]), "last/median contributor (mapped to the reference's list positions)"
    # (4) per-block footprint bits: every (Gaussian, pixel) pair that contributed with weight >= 0.1 (the pair list)
    #     must have the bit of the pixel's 8x4 block set in that Gaussian's entry of the pixel's tile
    if "pairs" in c and len(c["pairs"]) and c["block_bits"].any():  # (all zero: plain-entry fallback, no bits)
        bits = c["block_bits"].astype(np.int64)
        entry = {}
        for t in range(tiles_x * tiles_y):
            for j in range(cr[t, 0], cr[t, 1]):
                entry[(t, int(cl[j]))] = int(bits[j])
        for gid, pix in np.asarray(c["pairs"], np.int64).reshape(-1, 2)[:20000]:
            y, x = divmod(int(pix), W)
            t, blk = (y // 16) * tiles_x + x // 16, ((y % 16) // 4) * 2 + (x % 16) // 8
            assert (entry[(t, int(gid))] >> blk) & 1, "contributing pair outside the entry's block bits"


def cuda_backward(inp, fwd, dcolor, dothers, dextra, grad_mask=15, sparse=None, flags=1):
    import torch
    from instascene_b200.rasterizer import c_rasterize_gaussians_backward
    st = fwd["_torch"]
    tens = st["tens"]
    dev = tens["means3D"].device
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    e = torch.empty(0, dtype=torch.float32, device=dev)
    sp = None
    if sparse is not None:
        sp = (torch.from_numpy(sparse[0]).to(dev), torch.from_numpy(sparse[1]).to(dev))
    res = c_rasterize_gaussians_backward(tens["bg"], tens["means3D"], st["radii"], tens["colors"], tens["scales"],
                                         tens["rotations"], tens["extra"], inp.get("scale_modifier", 1.0), tens["transmat"],
                                         tens["view"], tens["proj"], inp["tanfovx"], inp["tanfovy"],
                                         t(dcolor), t(dothers), t(dextra), tens["shs"], inp["sh_degree"], tens["campos"],
                                         st["geom"], st["R"], st["binning"], st["img"], False, grad_mask=grad_mask,
                                         image_size=(inp["H"], inp["W"]), sparse_extra=sp, flags=flags)
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dtransMat", "dL_dsh", "dL_dscales",
             "dL_drotations", "dL_dextra"]
    return {k: v.cpu().numpy() for k, v in zip(names, res)}


def rel_err(a, b):
    """max |a-b| / (max|b| + tiny): norm-wise relative error used for float gradients."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def pair_set(pairs):
    p = np.asarray(pairs, np.int64).reshape(-1, 2)
    return set((p[:, 0] << 32 | p[:, 1]).tolist())


def tracker_label_map(W, H, seed):
    """Masks of very different sizes so that the 50-Gaussian threshold keeps some and drops others; ids are not
    contiguous (the reference sorts `torch.unique`)."""
    rng = np.random.default_rng(seed)
    lab = np.zeros((H, W), np.int16)
    lab[: H // 2, : W // 2] = 3
    lab[: H // 2, W // 2:] = 7
    lab[H // 2:, : W // 3] = 12
    lab[H // 2:, W // 3: W // 3 + 3] = 40       # 3-pixel-wide sliver: few Gaussians
    lab[H - 4:, W - 4:] = 41                    # 4x4 corner
    lab[H // 2 + 5: H // 2 + 25, W // 2: W // 2 + 30] = 300
    lab[rng.random((H, W)) < 0.1] = 0
    return lab


def knn_fixture_points(P, seed):
    """Point cloud of the kNN fixtures: a uniform cube, a dense cluster (1/4 of the points, 100x denser), a planar sheet
    (z constant) and a few exact duplicates -- the cases the reference's Morton-box search and the uniform-grid search
    could disagree on."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-1.5, 1.5, size=(P, 3)).astype(np.float32)
    q = P // 4
    pts[:q] = (rng.standard_normal((q, 3)) * 0.01 + np.array([0.3, -0.2, 0.5])).astype(np.float32)
    pts[q:q + P // 8, 2] = np.float32(0.25)
    d = max(P // 100, 1)
    pts[-d:] = pts[q + P // 8: q + P // 8 + d]
    return np.ascontiguousarray(pts)
