"""point_cloud.ply and COLMAP .bin formats (instascene_b200/io.py), CPU only.
Where /root/reference exists (this container), the files we write are also parsed by the reference's own COLMAP
readers (scene/colmap_loader.py) -- on the GPU box that leg is skipped."""
import importlib.util
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from instascene_b200 import io as isr_io
from instascene_b200 import synth


def _scene(P=257, F=16, seed=5):
    sc = synth.synth_scene(P, F=F, seed=seed)
    return sc


def test_ply_layout_and_round_trip(tmp_path):
    sc = _scene()
    path = str(tmp_path / "pc" / "point_cloud.ply")
    isr_io.save_ply(path, sc.xyz, sc.features_dc, sc.features_rest, sc.opacity_raw, sc.scaling_raw, sc.rotation_raw,
                    sc.seg_feature_raw)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().strip().split("\n")
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 257"]
    names = [l.split()[-1] for l in lines[3:]]
    assert all(l.startswith("property float ") for l in lines[3:])
    # scene/gaussian_model.py:263-283
    assert names == (["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(45)]
                     + ["opacity", "scale_0", "scale_1", "rot_0", "rot_1", "rot_2", "rot_3"] + [f"segfeat_{i}" for i in range(16)])
    assert len(body) == 257 * len(names) * 4
    tab = np.frombuffer(body, "<f4").reshape(257, len(names))
    assert np.array_equal(tab[:, 0:3], sc.xyz) and not tab[:, 3:6].any()
    # SH blocks are channel-major on disk: f_rest_k = features_rest[:, k % 15, k // 15]
    assert np.array_equal(tab[:, 9 + 17], sc.features_rest[:, 2, 1])
    g = isr_io.load_ply(path, max_sh_degree=3, seg_feat_dim=16)
    for got, want in zip(g, (sc.xyz, sc.features_dc, sc.features_rest, sc.opacity_raw, sc.scaling_raw, sc.rotation_raw,
                             sc.seg_feature_raw)):
        assert got.dtype == np.float32 and np.array_equal(got, want)
    # the reference ignores the seg feature when the column count differs from seg_feat_dim (gaussian_model.py:399-404)
    assert isr_io.load_ply(path, seg_feat_dim=8).seg_feature is None


def test_ply_crop_mask_no_segfeat_and_ascii_reader(tmp_path):
    sc = _scene(P=40, F=0)
    mask = np.arange(40) % 3 == 0
    path = str(tmp_path / "a.ply")
    isr_io.save_ply(path, sc.xyz, sc.features_dc, sc.features_rest, sc.opacity_raw, sc.scaling_raw, sc.rotation_raw, None,
                    crop_mask=mask)
    g = isr_io.load_ply(path)
    assert g.seg_feature is None and np.array_equal(g.xyz, sc.xyz[mask]) and np.array_equal(g.rotation, sc.rotation_raw[mask])
    # an ascii file with double columns, shuffled f_rest order and a third scale column (3DGS export) parses the same
    v = isr_io.read_ply_vertices(path)
    names = list(v)
    order = names[:9] + names[9:54][::-1] + names[54:57] + ["scale_2"] + names[57:]
    v["scale_2"] = np.zeros(len(v["x"]), np.float32)
    apath = str(tmp_path / "b.ply")
    with open(apath, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by a test\nelement vertex %d\n" % len(v["x"]))
        f.write("".join(f"property double {n}\n" for n in order) + "element face 0\nproperty list uchar int vertex_indices\nend_header\n")
        for i in range(len(v["x"])):
            f.write(" ".join(repr(float(v[n][i])) for n in order) + "\n")
    g2 = isr_io.load_ply(apath)
    for a, b in zip(g[:6], g2[:6]):
        assert np.array_equal(a, b)


def test_colmap_bin_round_trip_and_camera_reconstruction(tmp_path):
    cams = synth.ring_cameras(7, 640, 360)
    d = str(tmp_path / "sparse" / "0")
    isr_io.write_synthetic_colmap(d, cams)
    intr, extr = isr_io.read_cameras_binary(d + "/cameras.bin"), isr_io.read_images_binary(d + "/images.bin")
    assert len(intr) == 1 and intr[1].model == "PINHOLE" and (intr[1].width, intr[1].height) == (640, 360)
    assert [extr[i + 1].name for i in range(7)] == [f"view_{i:05d}.png" for i in range(7)]
    # byte layout (scene/colmap_loader.py:215-241): u64 count | i32 id, i32 model, u64 w, u64 h | 4 f64
    raw = open(d + "/cameras.bin", "rb").read()
    assert len(raw) == 8 + 24 + 32 and struct.unpack("<QiiQQ", raw[:32]) == (1, 1, 1, 640, 360)
    back = isr_io.load_colmap_cameras(d)
    for a, b in zip(cams, back):
        assert np.allclose(a.world_view_transform, b.world_view_transform, atol=2e-6)
        assert np.allclose(a.full_proj_transform, b.full_proj_transform, atol=2e-6)
        assert np.allclose(a.camera_center, b.camera_center, atol=1e-5)
        assert abs(a.FoVx - b.FoVx) < 1e-12 and abs(a.FoVy - b.FoVy) < 1e-12
    # 2D observations survive too
    im = extr[3]._replace(xys=np.array([[1.5, 2.5], [3.0, 4.0]]), point3D_ids=np.array([7, -1]))
    isr_io.write_images_binary(d + "/one.bin", {3: im})
    r = isr_io.read_images_binary(d + "/one.bin")[3]
    assert np.array_equal(r.xys, im.xys) and r.point3D_ids.tolist() == [7, -1] and np.array_equal(r.qvec, im.qvec)


def test_qvec_rotmat_inverse_pair():
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        q = -q if q[0] < 0 else q
        R = isr_io.qvec2rotmat(q)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
        assert np.allclose(isr_io.rotmat2qvec(R), q, atol=1e-9)


@pytest.mark.skipif(not os.path.isdir("/root/reference/scene"), reason="reference tree only exists in the build container")
def test_reference_colmap_reader_parses_our_files(tmp_path):
    spec = importlib.util.spec_from_file_location("ref_colmap_loader", "/root/reference/scene/colmap_loader.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cams = synth.ring_cameras(5, 320, 200)
    d = str(tmp_path)
    isr_io.write_synthetic_colmap(d, cams)
    rin, rex = ref.read_intrinsics_binary(d + "/cameras.bin"), ref.read_extrinsics_binary(d + "/images.bin")
    oin, oex = isr_io.read_cameras_binary(d + "/cameras.bin"), isr_io.read_images_binary(d + "/images.bin")
    assert set(rin) == set(oin) and set(rex) == set(oex)
    for k in rin:
        assert (rin[k].model, rin[k].width, rin[k].height) == (oin[k].model, oin[k].width, oin[k].height)
        assert np.array_equal(rin[k].params, oin[k].params)
    for k in rex:
        assert rex[k].name == oex[k].name and rex[k].camera_id == oex[k].camera_id
        assert np.array_equal(rex[k].qvec, oex[k].qvec) and np.array_equal(rex[k].tvec, oex[k].tvec)
        assert np.allclose(ref.qvec2rotmat(rex[k].qvec), isr_io.qvec2rotmat(oex[k].qvec), atol=1e-15)
        assert np.allclose(ref.rotmat2qvec(ref.qvec2rotmat(rex[k].qvec)), isr_io.rotmat2qvec(isr_io.qvec2rotmat(oex[k].qvec)), atol=1e-9)


def _ply_fixture_tensors():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_ply_golden import fixture_model_tensors
    return fixture_model_tensors()


def test_ply_writer_matches_reference_fixture(tmp_path):
    """tests/golden/ply_g1*.ply were written by the reference's UNMODIFIED GaussianModel.save_ply
    (tests/golden/make_ply_golden.py); our writer must produce the same bytes from the same tensors, and our reader must
    return the tensors from the reference's file."""
    import torch
    t = _ply_fixture_tensors()
    gold = os.path.join(ROOT, "tests", "golden")
    isr_io.save_ply(str(tmp_path / "a.ply"), t["xyz"], t["features_dc"], t["features_rest"], t["opacity"], t["scaling"],
                    t["rotation"], t["seg_feature"])
    assert open(tmp_path / "a.ply", "rb").read() == open(os.path.join(gold, "ply_g1.ply"), "rb").read()
    mask = torch.arange(t["xyz"].shape[0]) % 3 != 0
    isr_io.save_ply(str(tmp_path / "b.ply"), t["xyz"], t["features_dc"], t["features_rest"], t["opacity"], t["scaling"],
                    t["rotation"], None, crop_mask=mask)
    assert open(tmp_path / "b.ply", "rb").read() == open(os.path.join(gold, "ply_g1_noseg_crop.ply"), "rb").read()
    g = isr_io.load_ply(os.path.join(gold, "ply_g1.ply"), max_sh_degree=3, seg_feat_dim=16)
    assert np.array_equal(g.xyz, t["xyz"].numpy()) and np.array_equal(g.seg_feature, t["seg_feature"].numpy())
    assert np.array_equal(g.features_dc, t["features_dc"].numpy()) and np.array_equal(g.features_rest, t["features_rest"].numpy())
    assert np.array_equal(g.opacity, t["opacity"].numpy()) and np.array_equal(g.scaling, t["scaling"].numpy())
    assert np.array_equal(g.rotation, t["rotation"].numpy())


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "scene")), reason="baseline/_ref not installed")
def test_reference_save_ply_live_equals_ours(tmp_path):
    """The reference's own save_ply, run here, against our writer on fresh random tensors (not just the fixture)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_ply_golden import fixture_model_tensors, reference_model
    for P, F, seed in ((5, 4, 1), (257, 24, 2), (1000, 32, 3)):  # (the reference's PCA preview needs >= 3 rows)
        t = fixture_model_tensors(P, F, seed)
        reference_model(t, True).save_ply(str(tmp_path / f"r{seed}" / "pc.ply"))
        isr_io.save_ply(str(tmp_path / f"m{seed}.ply"), t["xyz"], t["features_dc"], t["features_rest"], t["opacity"],
                        t["scaling"], t["rotation"], t["seg_feature"])
        assert open(tmp_path / f"r{seed}" / "pc.ply", "rb").read() == open(tmp_path / f"m{seed}.ply", "rb").read()
