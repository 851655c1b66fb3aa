"""Generates tests/golden/ply_g1.ply (+ ply_g1_noseg_crop.ply) with the reference's UNMODIFIED GaussianModel.save_ply
(scene/gaussian_model.py:285-320, imported from baseline/_ref).  Runs on the CPU:

    python tests/golden/make_ply_golden.py tests/golden

The third-party `plyfile` package is absent from this image; tests/plyfile_stub.py stands in for it (same API subset,
plyfile's default binary_little_endian layout); open3d & co. are stubbed by baseline/ref_loader.py (save_ply only uses
them for two colour previews written after the parameter file)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "baseline"))


def fixture_model_tensors(P=64, F=16, seed=31):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(xyz=r(P, 3), features_dc=r(P, 1, 3), features_rest=r(P, 15, 3), opacity=r(P, 1), scaling=r(P, 2),
                rotation=r(P, 4), seg_feature=r(P, F))


def reference_model(t, with_seg=True):
    import plyfile_stub
    sys.modules["plyfile"] = plyfile_stub
    import ref_loader
    ref_loader._install_stubs()
    if ref_loader.REF_DIR not in sys.path:
        sys.path.insert(0, ref_loader.REF_DIR)
    import scene.gaussian_model as gm
    m = gm.GaussianModel(3)
    m._xyz, m._features_dc, m._features_rest = t["xyz"], t["features_dc"], t["features_rest"]
    m._opacity, m._scaling, m._rotation = t["opacity"], t["scaling"], t["rotation"]
    m._seg_feature = t["seg_feature"] if with_seg else None
    return m


def main(out_dir):
    t = fixture_model_tensors()
    os.makedirs(out_dir, exist_ok=True)
    import tempfile
    with tempfile.TemporaryDirectory() as d:  # save_ply also writes *_color.ply / *_feat.ply previews next to the file
        reference_model(t, True).save_ply(os.path.join(d, "a", "ply_g1.ply"))
        mask = torch.arange(t["xyz"].shape[0]) % 3 != 0
        reference_model(t, False).save_ply(os.path.join(d, "b", "ply_g1_noseg_crop.ply"), crop_mask=mask)
        for sub, name in (("a", "ply_g1.ply"), ("b", "ply_g1_noseg_crop.ply")):
            data = open(os.path.join(d, sub, name), "rb").read()
            open(os.path.join(out_dir, name), "wb").write(data)
            print(name, len(data), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden"))
