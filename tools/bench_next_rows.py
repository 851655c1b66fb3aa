"""Times the §8(f) kernels on the GPU with CUDA events (not part of bench.py's headline line):
  photometric loss fwd+bwd @1080p vs the reference's formula in torch (conv2d + autograd),
  tracker extraction on the pair list of one cfg3 view vs the reference's Python-set loop (init_tracker.py:26-47)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def ev_time(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    import instascene_b200 as isr
    from instascene_b200 import synth
    from instascene_b200.tracker import segmap_gaussians
    from test_losses import _torch_reference
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    out = {}
    gen = torch.Generator(device="cuda").manual_seed(1)
    gt = torch.rand((3, 1080, 1920), device="cuda", generator=gen)
    img = (gt + 0.1 * torch.randn((3, 1080, 1920), device="cuda", generator=gen)).clamp(0, 1)

    def ours():
        a = img.detach().requires_grad_(True)
        isr.photometric_loss(a, gt, 0.2).backward()

    def ref():
        a = img.detach().requires_grad_(True)
        _torch_reference(a, gt, 0.2).backward()

    chw = img.numel()
    t_o, t_r = ev_time(ours), ev_time(ref)
    # the two kernels alone, through the C ABI (no autograd / Python between launches)
    from instascene_b200 import _lib
    L = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream
    ws_bytes = L.isr_photometric_workspace_bytes(3, 1080, 1920)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    out3 = torch.empty(3, device="cuda")
    dimg = torch.empty_like(img)
    fwd = lambda: _lib.check(L.isr_photometric_forward(3, 1080, 1920, img.data_ptr(), gt.data_ptr(), 0.2, ws.data_ptr(), ws_bytes,
                                                       out3.data_ptr(), stream), "fwd")
    bwd = lambda: _lib.check(L.isr_photometric_backward(3, 1080, 1920, img.data_ptr(), gt.data_ptr(), 0.2, ws.data_ptr(), None,
                                                        dimg.data_ptr(), stream), "bwd")
    t_f, t_b = ev_time(fwd, n=50), ev_time(bwd, n=50)
    alg_f, alg_b = chw * 4 * (2 + 3), chw * 4 * (3 + 2 + 1)
    out["photometric_1080p"] = {"autograd_fwd_bwd_ms": t_o, "torch_formula_ms": t_r, "speedup": t_r / t_o,
                                "fwd_kernel_ms": t_f, "bwd_kernel_ms": t_b, "fwd_GBs": alg_f / t_f / 1e6,
                                "bwd_GBs": alg_b / t_b / 1e6, "algorithmic_bytes_fwd": alg_f, "algorithmic_bytes_bwd": alg_b,
                                "hbm_peak_GBs": peaks.get("hbm_gbs")}
    # tracker on a cfg3-sized pair list
    P, W, H = 2_000_000, 1920, 1080
    scene = synth.synth_scene(P, F=0, seed=1003)
    cam = synth.ring_cameras(200, W, H)[0]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

    class PC:
        active_sh_degree, max_sh_degree = 3, 3
        get_xyz, get_opacity = t(scene.xyz), t(scene.opacities()).reshape(-1, 1)
        get_scaling, get_rotation, get_features = t(scene.scales()), t(scene.rotations()), t(scene.shs())
        get_seg_feature = None

    from bench import _Cam, _Pipe
    view = _Cam(cam, t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.camera_center))
    with torch.no_grad():
        pairs = isr.render(view, PC(), _Pipe, torch.zeros(3, device="cuda"))["gau_related_pixels"]
    seg = t(synth.label_map(W, H, 7))
    torch.cuda.synchronize()
    t_dev = ev_time(lambda: segmap_gaussians(pairs, seg, P), n=5, warm=2)
    t0 = time.time()
    ts = segmap_gaussians(pairs, seg, P)
    torch.cuda.synchronize()
    wall = time.time() - t0
    # the reference's loop on the same pair list (python sets, init_tracker.py:26-47), timed once
    t0 = time.time()
    gaus_ids, pixel_ids = pairs[:, 0], pairs[:, 1]
    mask_image = seg.reshape(-1)
    ids = torch.unique(mask_image).cpu().numpy()
    info = {}
    frame = set(gaus_ids.tolist())
    for mask_id in ids:
        if mask_id == 0:
            continue
        valid = (mask_image == mask_id)[pixel_ids.long()]
        s = set(gaus_ids[valid].tolist())
        if len(s) >= 50:
            info[mask_id] = s
    t_ref = time.time() - t0
    assert sorted(info) == [int(m) for m in ts.mask_ids] and len(frame) == ts.frame_ids.numel()
    G = int(pairs.shape[0])
    out["tracker_cfg3_view"] = {"pairs": G, "masks": len(info), "ours_ms_device_events": t_dev, "ours_ms_wall": wall * 1e3,
                                "reference_python_sets_ms": t_ref * 1e3, "speedup_wall": t_ref / wall}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
