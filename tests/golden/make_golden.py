"""Generates tests/golden/*.npz by running the UNMODIFIED reference CUDA rasterizer (baseline/_ref, built by
baseline/build_ref.sh from /root/reference for sm_100a) on seeded synthetic inputs.  Must run on a GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'        (then copy the .npz into tests/golden/)

Each fixture stores the inputs, every public output, the gradients for seeded random cotangents and the reference's
internal buffers (parsed from its opaque geometry / binning / image byte tensors, layout per
DSR/cuda_rasterizer/rasterizer_impl.cu:155-194) so that integer intermediates can be pinned too.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

from helpers import scene_inputs  # noqa: E402

GOLDEN_CASES = {
    # name: (P, F, W, H, seed, scale_mult)
    "g0_rgb": (3000, 0, 96, 64, 101, 1.0),
    "g1_feat16": (5000, 16, 128, 80, 102, 1.0),
    "g2_feat24_odd": (2500, 24, 70, 50, 103, 1.8),
    "g3_big_splats": (400, 8, 48, 32, 104, 5.0),
}


def _align(o, a=128):
    return (o + a - 1) // a * a


def parse_ref_buffers(geom, binning, img, P, R, HW, base_geom, base_bin, base_img):
    """obtain(chunk, ptr, count, 128) walks with absolute-address alignment."""
    out = {}

    def take(buf, state, dtype, count):
        o = _align(state["addr"]) - state["base"]
        n = count * np.dtype(dtype).itemsize
        arr = np.frombuffer(buf[o:o + n].tobytes(), dtype=dtype).copy()
        state["addr"] = state["base"] + o + n
        return arr

    g = {"addr": base_geom, "base": base_geom}
    gb = geom.cpu().numpy()
    out["depths"] = take(gb, g, np.float32, P)
    out["clamped"] = take(gb, g, np.uint8, 3 * P).reshape(P, 3)
    take(gb, g, np.int32, P)  # internal_radii (unused: radii are passed in)
    out["means2D"] = take(gb, g, np.float32, 2 * P).reshape(P, 2)
    out["transMats"] = take(gb, g, np.float32, 9 * P).reshape(P, 9)
    out["normal_opacity"] = take(gb, g, np.float32, 4 * P).reshape(P, 4)
    out["rgb"] = take(gb, g, np.float32, 3 * P).reshape(P, 3)
    out["tiles_touched"] = take(gb, g, np.uint32, P)
    i = {"addr": base_img, "base": base_img}
    ib = img.cpu().numpy()
    out["final_T"] = take(ib, i, np.float32, 3 * HW)
    out["n_contrib"] = take(ib, i, np.uint32, 2 * HW)
    out["ranges_perpixel"] = take(ib, i, np.uint32, 2 * HW)
    if R > 0:
        b = {"addr": base_bin, "base": base_bin}
        bb = binning.cpu().numpy()
        out["point_list"] = take(bb, b, np.uint32, R)
    else:
        out["point_list"] = np.zeros(0, np.uint32)
    return out


def run_reference(inp, dcolor, dothers, dextra):
    from diff_surfel_rasterization import _C as ref_C  # the unmodified reference extension
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    e = torch.empty(0, dtype=torch.float32, device=dev)
    F = 0 if inp["extra_attrs"] is None else inp["extra_attrs"].shape[1]
    P, W, H = inp["means3D"].shape[0], inp["W"], inp["H"]
    extra = t(inp["extra_attrs"]) if F else e
    args = (t(inp["bg"]), t(inp["means3D"]), e, t(inp["opacities"]).reshape(-1, 1), t(inp["scales"]), t(inp["rotations"]),
            1.0, e, extra, F, t(inp["viewmatrix"]), t(inp["projmatrix"]), inp["tanfovx"], inp["tanfovy"], H, W,
            t(inp["shs"]), inp["sh_degree"], t(inp["campos"]), False, False)
    (R, color, others, radii, out_extra, geom, binning, img, pairs, pidx) = ref_C.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    n_pairs = int(pidx.item()) + 1
    bufs = parse_ref_buffers(geom, binning, img, P, R, H * W, geom.data_ptr(), binning.data_ptr() if R else 0,
                             img.data_ptr())
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    res = dict(num_rendered=np.int64(R), color=color.cpu().numpy(), others=others.cpu().numpy(),
               radii=radii.cpu().numpy(), extra=out_extra.cpu().numpy() if F else np.zeros((0, H, W), np.float32),
               pairs=np.unique(pairs[:n_pairs].cpu().numpy().astype(np.int64), axis=0).astype(np.int32),
               pair_count=np.int64(n_pairs), depths=bufs["depths"], clamped=bufs["clamped"], means2D=bufs["means2D"],
               transMats=bufs["transMats"], normal_opacity=bufs["normal_opacity"], rgb=bufs["rgb"],
               tiles_touched=bufs["tiles_touched"], final_T=bufs["final_T"].reshape(3, H, W),
               n_contrib=bufs["n_contrib"].reshape(2, H, W), ranges=bufs["ranges_perpixel"][:2 * tiles].reshape(tiles, 2),
               point_list=bufs["point_list"])
    bargs = (args[0], args[1], radii, e, args[4], args[5], extra, 1.0, e, args[10], args[11], inp["tanfovx"],
             inp["tanfovy"], t(dcolor), t(dothers), t(dextra) if F else e, args[16], inp["sh_degree"], args[18], geom, R,
             binning, img, False)
    g = ref_C.rasterize_gaussians_backward(*bargs)
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dtransMat", "dL_dsh", "dL_dscales",
             "dL_drotations", "dL_dextra"]
    for k, v in zip(names, g):
        res[k] = v.cpu().numpy()
    return res


def cotangents(F, W, H, seed):
    rng = np.random.default_rng(seed + 1)
    return (rng.standard_normal((3, H, W)).astype(np.float32), rng.standard_normal((7, H, W)).astype(np.float32),
            rng.standard_normal((F, H, W)).astype(np.float32) if F else None)


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, (P, F, W, H, seed, sm) in GOLDEN_CASES.items():
        inp = scene_inputs(P, F, W, H, seed, scale_mult=sm)
        dcolor, dothers, dextra = cotangents(F, W, H, seed)
        res = run_reference(inp, dcolor, dothers, dextra)
        ins = {("in_" + k): v for k, v in inp.items() if isinstance(v, np.ndarray)}
        ins.update(in_W=np.int64(W), in_H=np.int64(H), in_sh_degree=np.int64(inp["sh_degree"]),
                   in_tanfovx=np.float64(inp["tanfovx"]), in_tanfovy=np.float64(inp["tanfovy"]))
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **ins, **{("ref_" + k): v for k, v in res.items()})
        print(name, "R", int(res["num_rendered"]), "pairs", int(res["pair_count"]), "V", int((res["radii"] > 0).sum()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
