"""Binning-only timing on the GPU box (development aid): phase B without the blend (ISR_FLAG_SKIP_BLEND) on the cfg3
scene, plus the footprint statistics that drive it.  Usage: python tools/bin_bench.py [P] [F] [W] [H] [views]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instascene_b200 import _lib, synth  # noqa: E402
from instascene_b200 import rasterizer as rz  # noqa: E402


def main():
    a = [int(x) for x in sys.argv[1:]]
    P, F, W, H, nv = (a + [2000000, 16, 1920, 1080, 4][len(a):])[:5]
    dev = "cuda:0"
    L = _lib.lib()
    sc = synth.synth_scene(P, F=F, seed=1003)
    cams = synth.ring_cameras(200, W, H)[:nv]
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    e = torch.empty(0, device=dev)
    xyz, opa, scl, rot, shs = t(sc.xyz), t(sc.opacities()).reshape(-1, 1), t(sc.scales()), t(sc.rotations()), t(sc.shs())
    bg = torch.zeros(3, device=dev)
    out = {"P": P, "W": W, "H": H, "views": []}
    for cam in cams:
        view, proj, cpos = t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.camera_center)
        st = rz.launch_geometry(bg, xyz, e, opa, scl, rot, 1.0, e, view, proj, cam.tanfovx, cam.tanfovy, H, W, shs, 3, cpos,
                                want_pairs=False)
        torch.cuda.synchronize()
        n_ref, n_inst = int(st.nr_host[0]), int(st.nr_host[1])
        cap = n_inst
        bin_bytes = L.isr_binning_bytes(P, cap, W, H)
        binning = torch.empty(bin_bytes, dtype=torch.uint8, device=dev)
        ar = st.args
        ar.binning, ar.binning_bytes = binning.data_ptr(), bin_bytes
        ar.flags |= _lib.FLAG_SKIP_BLEND
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            _lib.check(L.isr_forward_render(C.byref(ar), cap, stream), "render")
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(L.isr_forward_render(C.byref(ar), cap, stream), "render")
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        g = st.geom
        off = lambda f: int(L.isr_field_offset(f, P, cap, W, H))
        tiles = g[off(_lib.GEOM_TILES):off(_lib.GEOM_TILES) + 4 * P].view(torch.int32)
        tcnt = g[off(_lib.GEOM_TILE_COUNT):off(_lib.GEOM_TILE_COUNT) + 4 * P].view(torch.int32)
        vis = tiles > 0
        big = tiles > 64
        hist = torch.bincount(tcnt[vis & ~big].clamp(max=64), minlength=65).cpu().numpy()
        out["views"].append({
            "binning_ms_med": float(np.median(ts)), "binning_ms_min": float(np.min(ts)),
            "num_rendered_ref": n_ref, "instances": n_inst, "visible": int(vis.sum()), "big": int(big.sum()),
            "big_instances": int(tcnt[big].sum()), "max_tiles": int(tiles.max()),
            "emitted_mean": float(tcnt[vis].float().mean()), "emitted_p99": float(np.searchsorted(np.cumsum(hist) / max(1, hist.sum()), 0.99)),
            "hist_0_1_2_4_8_16_32_64": [int(hist[0]), int(hist[1]), int(hist[2]), int(hist[3:5].sum()), int(hist[5:9].sum()),
                                          int(hist[9:17].sum()), int(hist[17:33].sum()), int(hist[33:].sum())]})
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bin_bench.json"), "w"))


if __name__ == "__main__":
    main()
