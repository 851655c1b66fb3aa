"""Quick stage timing on the GPU box (development aid; bench.py is the contract)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instascene_b200 as isr  # noqa: E402
from instascene_b200 import synth  # noqa: E402


def make_pc(sc, dev):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    class PC:
        active_sh_degree, max_sh_degree = 3, 3
        get_xyz = t(sc.xyz)
        get_opacity = t(sc.opacities()).reshape(-1, 1)
        get_scaling = t(sc.scales())
        get_rotation = t(sc.rotations())
        get_features = t(sc.shs())
        _seg_feature = t(sc.seg_feature_raw).requires_grad_(True) if sc.seg_feature_raw is not None else None

        @property
        def get_seg_feature(self):
            if self._seg_feature is None:
                return None
            return self._seg_feature / (torch.norm(self._seg_feature, p=2, dim=1, keepdim=True) + 1e-6)

    return PC()


def make_cam(cam, dev):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    class Cam:
        FoVx, FoVy, image_width, image_height = cam.FoVx, cam.FoVy, cam.image_width, cam.image_height
        world_view_transform, full_proj_transform, camera_center = t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.camera_center)
        znear, zfar = 0.01, 100.0

    return Cam()


class Pipe:
    compute_cov3D_python, convert_SHs_python, depth_ratio = False, False, 1.0


def main():
    P = int(os.environ.get("QB_P", 2000000))
    F = int(os.environ.get("QB_F", 16))
    W, H = int(os.environ.get("QB_W", 1920)), int(os.environ.get("QB_H", 1080))
    nviews = int(os.environ.get("QB_VIEWS", 8))
    with_ref = os.environ.get("QB_REF", "1") == "1"
    dev = "cuda:0"
    sc = synth.synth_scene(P, F=F, seed=1003)
    cams = [make_cam(c, dev) for c in synth.ring_cameras(max(nviews, 4), W, H)[:nviews]]
    pc = make_pc(sc, dev)
    bg = torch.zeros(3, device=dev)
    lab = synth.label_map(W, H, 1005)
    valid = np.flatnonzero(lab.reshape(-1) > 0)
    rng = np.random.default_rng(1006)
    pix = torch.from_numpy(valid[rng.integers(0, valid.size, size=32768)]).to(dev)
    labels = torch.from_numpy(lab.reshape(-1)[pix.cpu().numpy()].astype(np.int64)).to(dev)
    out = {"P": P, "F": F, "W": W, "H": H}

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def run_mine(cam, want_pairs=True):
        pkg = isr.render(cam, pc, Pipe, bg, want_pairs=want_pairs)
        feats = isr.sample_pixels(pkg["seg_feature"], pix)
        loss = isr.contrastive_loss(feats, labels, num_labels=64) * 1e-6
        loss.backward()
        g = pc._seg_feature.grad
        pc._seg_feature.grad = None
        return pkg, loss, g

    for wp in (True, False):
        for _ in range(2):
            run_mine(cams[0], wp)
        torch.cuda.synchronize()
        ts = []
        for c in cams:
            e0, e1 = ev(), ev()
            e0.record()
            pkg, loss, g = run_mine(c, wp)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out["mine_fwdbwd_ms_pairs%d" % wp] = ts
    # raw forward only (rasterizer binding, no glue)
    from instascene_b200.rasterizer import c_rasterize_gaussians
    with torch.no_grad():
        seg = pc.get_seg_feature
        seg = (seg / (seg.norm(dim=-1, keepdim=True) + 1e-9)).contiguous()
        e = torch.empty(0, device=dev)
        ts = []
        for c in cams:
            e0, e1 = ev(), ev()
            e0.record()
            res = c_rasterize_gaussians(bg, pc.get_xyz, e, pc.get_opacity, pc.get_scaling, pc.get_rotation, 1.0, e, seg, F,
                                        c.world_view_transform, c.full_proj_transform, np.tan(c.FoVx / 2), np.tan(c.FoVy / 2),
                                        H, W, pc.get_features, 3, c.camera_center, False, False, want_pairs=False)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out["mine_raster_fwd_ms"] = ts
        out["R"] = int(res[0])
        out["V"] = int((res[3] > 0).sum())
    if with_ref:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import ref_loader
        if ref_loader.available():
            rrender, rloss, _ = ref_loader.load()

            def run_ref(cam):
                pkg = rrender(cam, pc, Pipe, bg)
                segf = pkg["seg_feature"]
                feats = segf.reshape(F, -1)[:, pix].T
                loss = rloss(feats, labels) * 1e-6
                loss.backward()
                g = pc._seg_feature.grad
                pc._seg_feature.grad = None
                return pkg, loss, g

            for _ in range(2):
                run_ref(cams[0])
            torch.cuda.synchronize()
            ts = []
            for c in cams[:4]:
                e0, e1 = ev(), ev()
                e0.record()
                pkg_r, loss_r, g_r = run_ref(c)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            out["ref_fwdbwd_ms"] = ts
            # agreement on the last timed view
            pkg_m, loss_m, g_m = run_mine(cams[3], True)
            out["loss_mine"], out["loss_ref"] = float(loss_m), float(loss_r)
            out["grad_rel_err"] = float((g_m - g_r).abs().max() / g_r.abs().max())
            out["seg_rel_err"] = float((pkg_m["seg_feature"] - pkg_r["seg_feature"]).abs().max() / pkg_r["seg_feature"].abs().max())
            out["radii_mismatch"] = int((pkg_m["radii"] != pkg_r["radii"]).sum())
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "quick_bench.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
