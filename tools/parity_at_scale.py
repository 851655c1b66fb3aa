"""Full-size parity report against the UNMODIFIED reference CUDA rasterizer (baseline/_ref) on a B200.

    gpurun -- 'python tools/parity_at_scale.py'      ->  gpurun_out/parity_scale.json  (copied to profiles/)

cfg3 (2M Gaussians, F=16, 1080p) and cfg2 (500k, F=0): forward outputs, integer buffers, pair list, dense backward
with seeded random cotangents.  cfg5 (5M, F=32, 1600x1200): the reference cannot run F>24 (MAX_EXTRA_DIMS); its
feature map / feature gradient are assembled from two 16-channel reference passes.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

from instascene_b200 import synth  # noqa: E402
from instascene_b200.rasterizer import c_rasterize_gaussians, c_rasterize_gaussians_backward  # noqa: E402


def stats(a, b):
    a, b = a.double(), b.double()
    d = (a - b).abs()
    scale = b.abs().max().item() + 1e-30
    rel = d / (b.abs() + 1e-4 * scale)
    return {"max_abs": d.max().item(), "normwise_rel": d.max().item() / scale,
            "frac_elems_rel_gt_1e-4": (rel > 1e-4).double().mean().item(),
            "frac_elems_abs_gt_1e-4_of_max": (d > 1e-4 * scale).double().mean().item(),
            "median_abs": d.median().item(), "max": scale}


def run_case(name, P, F, W, H, seed, ref_C, out):
    dev = "cuda:0"
    sc = synth.synth_scene(P, F=F, seed=seed)
    cam = synth.ring_cameras(200, W, H)[7]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    e = torch.empty(0, device=dev)
    means, opa, scales, rots, shs = t(sc.xyz), t(sc.opacities()).reshape(-1, 1), t(sc.scales()), t(sc.rotations()), t(sc.shs())
    extra = t(sc.seg_features()) if F else e
    view, proj, campos, bg = t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.camera_center), torch.zeros(3, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    dcolor = torch.randn((3, H, W), device=dev, generator=g)
    dothers = torch.randn((7, H, W), device=dev, generator=g)
    dextra = torch.randn((F, H, W), device=dev, generator=g) if F else e

    def mine():
        fw = c_rasterize_gaussians(bg, means, e, opa, scales, rots, 1.0, e, extra, F, view, proj, cam.tanfovx, cam.tanfovy, H, W,
                                   shs, 3, campos, False, False)
        bw = c_rasterize_gaussians_backward(bg, means, fw[3], e, scales, rots, extra, 1.0, e, view, proj, cam.tanfovx,
                                            cam.tanfovy, dcolor, dothers, dextra if F else None, shs, 3, campos, fw[5], fw[0],
                                            fw[6], fw[7], False, image_size=(H, W))
        return fw, bw

    def ref(ex, dex, dcolor=dcolor, dothers=dothers):
        Fx = ex.shape[1] if ex.numel() else 0
        fw = ref_C.rasterize_gaussians(bg, means, e, opa, scales, rots, 1.0, e, ex, Fx, view, proj, cam.tanfovx, cam.tanfovy,
                                       H, W, shs, 3, campos, False, False)
        bw = ref_C.rasterize_gaussians_backward(bg, means, fw[3], e, scales, rots, ex, 1.0, e, view, proj, cam.tanfovx,
                                                cam.tanfovy, dcolor, dothers, dex if Fx else e, shs, 3, campos, fw[5], fw[0],
                                                fw[6], fw[7], False)
        return fw, bw

    mfw, mbw = mine()
    torch.cuda.synchronize()
    rep = {"P": P, "F": F, "W": W, "H": H}
    if F <= 24:
        rfw, rbw = ref(extra, dextra)
        r_extra, r_dextra = rfw[4], rbw[8]
    else:
        h = F // 2
        rfw, rbw = ref(extra[:, :h].contiguous(), dextra[:h].contiguous())
        # second pass: remaining channels, ZERO colour/aux cotangents; K7/K8 are linear in the cotangents, so all
        # per-Gaussian gradients of the two passes add (except the densification proxy dL_dmeans2D, which K8 overwrites
        # from dL_dtransMat and is therefore additive too)
        rfw2, rbw2 = ref(extra[:, h:].contiguous(), dextra[h:].contiguous(), torch.zeros_like(dcolor), torch.zeros_like(dothers))
        r_extra = torch.cat([rfw[4], rfw2[4]], 0)
        r_dextra = torch.cat([rbw[8], rbw2[8]], 1)
        rbw = tuple(a + b for a, b in zip(rbw[:8], rbw2[:8])) + (r_dextra,)
        rep["reference_F32"] = "two 16-channel passes"
    torch.cuda.synchronize()
    rep["num_rendered"] = [int(mfw[0]), int(rfw[0])]
    rep["radii_mismatch"] = int((mfw[3] != rfw[3]).sum())
    rep["visible"] = int((rfw[3] > 0).sum())
    rep["color"] = stats(mfw[1], rfw[1])
    rep["others"] = stats(mfw[2], rfw[2])
    chan = ["depth_w", "alpha", "normal_x", "normal_y", "normal_z", "median_depth", "distortion"]
    rep["others_per_channel"] = {c: stats(mfw[2][i], rfw[2][i]) for i, c in enumerate(chan)}
    rep["n_contrib_last_mismatch_px"] = None
    if F:
        rep["extra"] = stats(mfw[4], r_extra)
    n_m, n_r = int(mfw[9].item()) + 1, int(rfw[9].item()) + 1
    pm = mfw[8][:n_m].long()
    pr = rfw[8][:n_r].long()
    km = torch.unique(pm[:, 0] * (W * H) + pm[:, 1])
    kr = torch.unique(pr[:, 0] * (W * H) + pr[:, 1])
    both = torch.cat([km, kr]).unique(return_counts=True)[1]
    rep["pairs"] = {"mine": n_m, "ref": n_r, "symmetric_difference": int((both == 1).sum())}
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dtransMat", "dL_dsh", "dL_dscales", "dL_drotations"]
    def gstats(a, b):
        b = b.reshape(a.shape)
        st = stats(a, b)
        rows = ((a.double() - b.double()).abs().reshape(a.shape[0], -1).amax(1) > 1e-4 * st["max"]).double().mean().item()
        return {"normwise_rel": st["normwise_rel"], "frac_gaussians_abs_gt_1e-4_of_max": rows, "median_abs": st["median_abs"],
                "max": st["max"]}
    rep["grads"] = {k: gstats(a, b) for k, a, b in zip(names, mbw[:8], rbw[:8])}
    if F:
        rep["grads"]["dL_dextra"] = gstats(mbw[8], r_dextra)
    out[name] = rep
    print(name, json.dumps(rep)[:600], flush=True)


def main():
    import ref_loader
    ref_loader._install_stubs()
    from diff_surfel_rasterization import _C as ref_C
    out = {}
    run_case("cfg2", 500_000, 0, 1920, 1080, 1002, ref_C, out)
    run_case("cfg3", 2_000_000, 16, 1920, 1080, 1003, ref_C, out)
    run_case("cfg5", 5_000_000, 32, 1600, 1200, 1005, ref_C, out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_scale.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
