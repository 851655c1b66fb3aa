"""Generates tests/golden/tracker_g1.npz by running the UNMODIFIED reference on a GPU box: the reference's own
`get_segmap_gaussians` (spatial_track/modules/init_tracker.py:16-47, imported from baseline/_ref) driving the reference's
own render() and CUDA rasterizer on a seeded synthetic scene.

    gpurun -- 'python tests/golden/make_tracker_golden.py gpurun_out/golden'     (then copy the .npz into tests/golden/)

Stored: the pair list the reference produced, the segmentation map, P, and the reference's result (per kept mask the
sorted Gaussian ids, and the sorted frame ids)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "baseline"))


def main(out_dir):
    import ref_loader
    from helpers import scene_inputs, tracker_label_map
    from bench import _Cam, _Pipe
    rrender, _, _ = ref_loader.load()
    sys.path.insert(0, ref_loader.REF_DIR)
    import importlib
    it = importlib.import_module("spatial_track.modules.init_tracker")
    P, F, W, H, seed = 5000, 16, 128, 80, 102
    inp = scene_inputs(P, F, W, H, seed)
    dev = torch.device("cuda")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    class PC:
        active_sh_degree, max_sh_degree = 3, 3
        pipelineparams = _Pipe

        def __init__(self):
            self.get_xyz = t(inp["means3D"])
            self.get_opacity = t(inp["opacities"]).reshape(-1, 1)
            self.get_scaling = t(inp["scales"])
            self.get_rotation = t(inp["rotations"])
            self.get_features = t(inp["shs"])
            self.get_seg_feature = t(inp["extra_attrs"])

    cam = inp["cam"]
    view = _Cam(cam, t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.camera_center))
    lab = tracker_label_map(W, H, seed)
    view.segmap = torch.from_numpy(lab)
    pc = PC()
    with torch.no_grad():
        # the pair list first (stored as the fixture's input), then the reference function end to end
        pairs = rrender(view, pc, _Pipe, torch.zeros(3, device=dev))["gau_related_pixels"].cpu().numpy()
        mask_info, frame_ids = it.get_segmap_gaussians(pc, view)
    kept = sorted(int(k) for k in mask_info)
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, "tracker_g1.npz"), pairs=pairs.astype(np.int32), segmap=lab, P=P, W=W, H=H,
                        kept_mask_ids=np.array(kept, np.int64),
                        frame_ids=np.array(sorted(frame_ids), np.int64),
                        **{f"mask_{k}": np.array(sorted(mask_info[k]), np.int64) for k in kept})
    print("tracker golden: G=%d pairs, masks kept %s of %s, frame ids %d" % (
        len(pairs), kept, sorted(set(np.unique(lab).tolist()) - {0}), len(frame_ids)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
