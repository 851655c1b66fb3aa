"""GPU parity tests (run with -m gpu on the B200 box): CUDA path through the C ABI vs the CPU oracle on the same
seeded inputs.  Forward outputs -- floats AND every thresholded integer -- must be BIT-EXACT (the spec'd fp32
arithmetic is identical on both sides); gradients are compared at 1e-4 norm-wise relative error (the summation
order over pixels differs; the reference itself uses unordered float atomics)."""
import numpy as np
import pytest

from helpers import (check_tile_lists, cuda_backward, cuda_forward, oracle_backward, oracle_forward, pair_set, rel_err, scene_inputs)

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-4

CASES = [
    # P, F, W, H, seed, scale_mult
    (3000, 0, 96, 64, 11, 1.0),
    (6000, 16, 160, 96, 12, 1.0),
    (2000, 3, 70, 50, 13, 2.0),     # non-multiple-of-16 image, odd F
    (20000, 32, 128, 128, 14, 0.7),  # F = 32 (beyond the reference's MAX_EXTRA_DIMS 24)
    (500, 8, 33, 17, 15, 6.0),      # huge splats, long per-tile lists, tiny image
]


def _check_forward_exact(c, o, F):
    for k in ["radii", "tiles_touched", "num_rendered"]:
        assert np.array_equal(c[k], o[k]), k
    vis = o["radii"] > 0
    for k in ["depths", "transMats", "means2D", "normal_opacity", "rgb"]:
        assert np.array_equal(c[k][vis].view(np.uint32), o[k][vis].view(np.uint32)), k
    cm = np.packbits(o["clamped"][vis].astype(bool), axis=1, bitorder="little")[:, 0]
    assert np.array_equal(c["clamped_mask"][vis], cm)
    check_tile_lists(c, o, c["color"].shape[2], c["color"].shape[1])
    for k in ["final_T", "color", "others"] + (["extra"] if F else []):
        assert np.array_equal(c[k].view(np.uint32), o[k].view(np.uint32)), k
    assert c["pair_count"] == o["pair_count"]
    assert pair_set(c["pairs"]) == pair_set(o["pairs"])


@pytest.mark.parametrize("P,F,W,H,seed,sm", CASES)
def test_forward_bit_exact(oracle, P, F, W, H, seed, sm):
    inp = scene_inputs(P, F, W, H, seed, scale_mult=sm)
    o = oracle_forward(oracle, inp)
    c = cuda_forward(inp)
    _check_forward_exact(c, o, F)


@pytest.mark.parametrize("P,F,W,H,seed,sm", CASES)
def test_backward_dense(oracle, P, F, W, H, seed, sm):
    inp = scene_inputs(P, F, W, H, seed, scale_mult=sm)
    rng = np.random.default_rng(seed + 1)
    dcolor = rng.standard_normal((3, H, W)).astype(np.float32)
    dothers = rng.standard_normal((7, H, W)).astype(np.float32)
    dextra = rng.standard_normal((F, H, W)).astype(np.float32) if F else None
    o = oracle_forward(oracle, inp)
    og = oracle_backward(oracle, inp, o, dcolor, dothers, dextra)
    c = cuda_forward(inp)
    cg = cuda_backward(inp, c, dcolor, dothers, dextra)
    for k in ["dL_dcolors", "dL_dopacity", "dL_dtransMat", "dL_dmeans2D", "dL_dmeans3D", "dL_dscales", "dL_drotations",
              "dL_dsh"] + (["dL_dextra"] if F else []):
        assert np.isfinite(cg[k]).all(), k
        assert rel_err(cg[k], og[k].reshape(cg[k].shape)) < GRAD_TOL, (k, rel_err(cg[k], og[k].reshape(cg[k].shape)))


def test_backward_sparse_equals_dense(oracle):
    P, F, W, H, seed = 6000, 16, 160, 96, 21
    inp = scene_inputs(P, F, W, H, seed)
    rng = np.random.default_rng(seed)
    n = 700
    pix = rng.integers(0, W * H, size=n).astype(np.int32)  # with replacement -> duplicates
    rows = rng.standard_normal((n, F)).astype(np.float32)
    dense = np.zeros((F, H * W), np.float32)
    np.add.at(dense.T, pix, rows)
    dense = dense.reshape(F, H, W)
    o = oracle_forward(oracle, inp)
    og = oracle_backward(oracle, inp, o, np.zeros((3, H, W), np.float32), np.zeros((7, H, W), np.float32), dense)
    c = cuda_forward(inp, want_pairs=False)
    g_sparse = cuda_backward(inp, c, None, None, None, grad_mask=8, sparse=(pix, rows))
    g_dense = cuda_backward(inp, c, None, None, dense, grad_mask=8)
    assert rel_err(g_sparse["dL_dextra"], og["dL_dextra"]) < GRAD_TOL
    assert rel_err(g_dense["dL_dextra"], og["dL_dextra"]) < GRAD_TOL


def test_wh_quirk_flag(oracle):
    """Q5: backward's W,H = int(focal*tan*2) can be W-1; both settings must match the oracle."""
    P, F, W, H, seed = 1500, 0, 64, 48, 31
    inp = scene_inputs(P, F, W, H, seed)
    rng = np.random.default_rng(seed)
    dcolor = rng.standard_normal((3, H, W)).astype(np.float32)
    dothers = rng.standard_normal((7, H, W)).astype(np.float32)
    o = oracle_forward(oracle, inp)
    c = cuda_forward(inp, want_pairs=False)
    for flags in (0, 1):
        og = oracle_backward(oracle, inp, o, dcolor, dothers, None, flags=flags)
        cg = cuda_backward(inp, c, dcolor, dothers, None, flags=flags)
        for k in ["dL_dmeans2D", "dL_dmeans3D", "dL_dscales", "dL_drotations"]:
            assert rel_err(cg[k], og[k]) < GRAD_TOL, (flags, k)


def test_empty_and_culled(oracle):
    import torch
    import instascene_b200 as isr
    dev = "cuda:0"
    # P = 0
    s = isr.GaussianRasterizationSettings(32, 48, 0.5, 0.4, torch.zeros(3, device=dev), 1.0, torch.eye(4, device=dev),
                                          torch.eye(4, device=dev), 0, torch.zeros(3, device=dev), False, False)
    r = isr.GaussianRasterizer(s)
    z = torch.zeros((0, 3), device=dev)
    color, radii, allmap, extra, pairs = r(z, z, torch.zeros((0, 1), device=dev), colors_precomp=z,
                                            scales=torch.zeros((0, 2), device=dev), rotations=torch.zeros((0, 4), device=dev))
    assert color.shape == (3, 32, 48) and float(color.abs().max()) == 0.0 and pairs.shape[0] == 0
    # everything behind the camera: R = 0, colour = background
    inp = scene_inputs(300, 4, 48, 32, 41)
    inp["means3D"] = inp["means3D"] + np.array([100.0, 0, 0], np.float32) * 0  # keep
    inp["viewmatrix"] = inp["viewmatrix"].copy()
    inp["viewmatrix"][3, 2] -= 50.0  # push the scene far behind the camera (z_view <= 0.2)
    o = oracle_forward(oracle, inp)
    c = cuda_forward(inp)
    assert o["num_rendered"] == 0 and c["num_rendered"] == 0
    assert np.array_equal(c["color"], o["color"]) and np.array_equal(c["others"], o["others"])
    with pytest.raises(Exception):
        r(z, z, torch.zeros((0, 1), device=dev))  # neither SHs nor colours


def test_mark_visible(oracle):
    import torch
    import instascene_b200 as isr
    inp = scene_inputs(5000, 0, 64, 64, 51)
    inp["viewmatrix"] = inp["viewmatrix"].copy()
    inp["viewmatrix"][3, 2] -= 4.0  # camera plane cuts through the cloud
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(a).to(dev)
    s = isr.GaussianRasterizationSettings(64, 64, 0.5, 0.5, torch.zeros(3, device=dev), 1.0, t(inp["viewmatrix"]),
                                          t(inp["projmatrix"]), 0, t(inp["campos"]), False, False)
    vis = isr.GaussianRasterizer(s).markVisible(t(inp["means3D"])).cpu().numpy()
    ref = oracle.mark_visible(inp["means3D"], inp["viewmatrix"])
    assert np.array_equal(vis, ref) and 0 < vis.sum() < vis.size


def test_knn(oracle):
    import torch
    import instascene_b200 as isr
    rng = np.random.default_rng(5)
    for P in (3, 100, 5000):
        pts = rng.standard_normal((P, 3)).astype(np.float32)
        pts[: P // 3] *= 0.01  # a dense cluster
        got = isr.distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
        want = oracle.knn_mean_dist2(pts)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), P


def test_knn_matches_reference_fixture():
    """tests/golden/knn_g1.npz = distCUDA2 of the UNMODIFIED reference simple_knn (tests/golden/make_knn_golden.py) on a
    cloud with a dense cluster, a planar sheet and exact duplicates: bit-identical (both are exact 3-NN searches and
    the squared distance / mean use the reference's FMA pattern, simple_knn.cu:123-124,183)."""
    import os
    import torch
    import instascene_b200 as isr
    from helpers import knn_fixture_points
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "knn_g1.npz"))
    pts = knn_fixture_points(int(z["P"]), int(z["seed"]))
    assert np.uint32(np.bitwise_xor.reduce(pts.view(np.uint32).reshape(-1))) == z["points_crc"]
    got = isr.distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), z["ref_mean_dist2"].view(np.uint32))


def test_knn_bit_identical_to_reference_simple_knn_at_1M():
    """Live against baseline/_ref/simple_knn (the unmodified reference build) at P = 1M: uniform + clustered."""
    import os
    import sys
    import torch
    import instascene_b200 as isr
    from helpers import knn_fixture_points
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_dir = os.path.join(root, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "simple_knn")):
        pytest.skip("baseline/_ref/simple_knn is not installed (baseline/build_ref.sh)")
    sys.path.insert(0, ref_dir)
    from simple_knn._C import distCUDA2 as ref_dist
    for P, seed in ((1_000_000, 9), (300_001, 10)):
        pts = torch.from_numpy(knn_fixture_points(P, seed)).cuda()
        want = ref_dist(pts)
        got = isr.distCUDA2(pts)
        torch.cuda.synchronize()
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), (P, int((got != want).sum()))


def test_knn_degenerate_clouds(oracle):
    """Planar / collinear / all-identical clouds (a grid axis with no extent): exact result, and no O(P * cells) scan --
    a 200k-point planar cloud must finish in well under a second."""
    import time
    import torch
    import instascene_b200 as isr
    rng = np.random.default_rng(3)
    for kind in ("planar", "collinear", "identical"):
        pts = rng.standard_normal((4000, 3)).astype(np.float32)
        if kind == "planar":
            pts[:, 1] = 0.5
        elif kind == "collinear":
            pts[:, 1:] = 0.0
        else:
            pts[:] = pts[0]
        got = isr.distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
        want = oracle.knn_mean_dist2(pts)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), kind
    big = rng.standard_normal((200_000, 3)).astype(np.float32)
    big[:, 2] = -1.0
    t = torch.from_numpy(big).cuda()
    isr.distCUDA2(t)
    torch.cuda.synchronize()
    t0 = time.time()
    out = isr.distCUDA2(t)
    torch.cuda.synchronize()
    assert time.time() - t0 < 1.0 and bool(torch.isfinite(out).all())


@pytest.mark.parametrize("predef,consider_negative", [(False, False), (True, False), (False, True)])
def test_contrastive_loss(predef, consider_negative):
    import torch
    import instascene_b200 as isr
    from instascene_b200 import synth
    from oracle.contrastive_ref import contrastive_loss_ref
    rng = np.random.default_rng(7)
    N, F, K = 4096, 16, 37
    feats = rng.standard_normal((N, F)).astype(np.float32)
    labels = rng.integers(0, K + 1, size=N).astype(np.int64)  # 0 = unlabelled
    labels[labels == 5] = 6                                   # an absent id in the middle
    proto = synth.gram_schmidt_prototypes(K + 1, F, 3) if predef else None
    f64 = torch.tensor(feats, dtype=torch.float64, requires_grad=True)
    want = contrastive_loss_ref(f64, torch.tensor(labels), None if proto is None else torch.tensor(proto, dtype=torch.float64),
                                consider_negative=consider_negative)
    want.backward()
    fc = torch.tensor(feats, device="cuda", requires_grad=True)
    got = isr.contrastive_loss(fc, torch.tensor(labels, device="cuda"),
                               None if proto is None else torch.tensor(proto, device="cuda"),
                               consider_negative=consider_negative)
    (got * 0.5).backward()
    assert abs(float(got) - float(want)) / abs(float(want)) < 1e-4
    assert rel_err(fc.grad.cpu().numpy() * 2.0, f64.grad.numpy()) < 1e-4


@pytest.mark.parametrize("predef", [False, True])
def test_contrastive_loss_min_pixnum(predef):
    """utils/contrastive_utils.py:33-35: clusters with <= min_pixnum samples are dropped together with their samples --
    inside the kernels (no bincount / where prefilter on the host)."""
    import torch
    import instascene_b200 as isr
    from oracle.contrastive_ref import contrastive_loss_ref
    rng = np.random.default_rng(77)
    N, F, K = 4000, 16, 40
    feats = rng.standard_normal((N, F)).astype(np.float32)
    # very uneven cluster sizes: some below, some above the threshold
    p = rng.random(K + 1) ** 4
    labels = rng.choice(K + 1, size=N, p=p / p.sum()).astype(np.int64)
    counts = np.bincount(labels, minlength=K + 1)
    thr = int(np.sort(counts[1:])[K // 2])
    assert (counts[1:] <= thr).any() and (counts[1:] > thr).any()
    proto = None
    if predef:
        proto = rng.standard_normal((K + 1, F)).astype(np.float32)
        proto /= np.linalg.norm(proto, axis=1, keepdims=True)
    for mp in (thr, 1, 0):
        f64 = torch.tensor(feats, dtype=torch.float64, requires_grad=True)
        want = contrastive_loss_ref(f64, torch.tensor(labels), None if proto is None else torch.tensor(proto, dtype=torch.float64),
                                    min_pixnum=mp)
        want.backward()
        fc = torch.tensor(feats, device="cuda", requires_grad=True)
        got = isr.contrastive_loss(fc, torch.tensor(labels, device="cuda"), None if proto is None else torch.tensor(proto, device="cuda"),
                                   min_pixnum=mp)
        got.backward()
        assert abs(float(got) - float(want)) / abs(float(want)) < 1e-4, mp
        assert rel_err(fc.grad.cpu().numpy(), f64.grad.numpy()) < 1e-4, mp


def test_autograd_render_sample_loss(oracle):
    """render() -> sample_pixels -> contrastive_loss -> backward: gradient w.r.t. the raw seg feature parameter
    through the sparse path equals the oracle's dense backward composed with torch autograd on the CPU."""
    import torch
    import instascene_b200 as isr
    from instascene_b200 import synth
    from oracle.contrastive_ref import contrastive_loss_ref
    P, F, W, H, seed = 5000, 16, 128, 80, 61
    inp = scene_inputs(P, F, W, H, seed)
    sc, cam = inp["scene"], inp["cam"]
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    class PC:
        active_sh_degree, max_sh_degree = 3, 3
        get_xyz = t(sc.xyz)
        get_opacity = t(sc.opacities()).reshape(-1, 1)
        get_scaling = t(sc.scales())
        get_rotation = t(sc.rotations())
        get_features = t(sc.shs())
        _seg_feature = t(sc.seg_feature_raw).requires_grad_(True)

        @property
        def get_seg_feature(self):
            return self._seg_feature / (torch.norm(self._seg_feature, p=2, dim=1, keepdim=True) + 1e-6)

    class Cam:
        FoVx, FoVy, image_width, image_height = cam.FoVx, cam.FoVy, W, H
        world_view_transform, full_proj_transform, camera_center = t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.camera_center)
        znear, zfar = 0.01, 100.0

    class Pipe:
        compute_cov3D_python, convert_SHs_python, depth_ratio = False, False, 1.0

    pc = PC()
    pkg = isr.render(Cam(), pc, Pipe(), t(inp["bg"]))
    assert set(pkg) == {"render", "viewspace_points", "visibility_filter", "radii", "seg_feature", "gau_related_pixels",
                        "rend_alpha", "rend_normal", "rend_dist", "surf_depth", "surf_normal", "rend_depth",
                        "rend_median_depth"}
    lab = synth.label_map(W, H, seed + 2)
    valid = np.flatnonzero(lab.reshape(-1) > 0)
    rng = np.random.default_rng(seed + 3)
    pix = valid[rng.integers(0, valid.size, size=2048)]
    labels = lab.reshape(-1)[pix].astype(np.int64)
    feats = isr.sample_pixels(pkg["seg_feature"], t(pix.astype(np.int64)))
    loss = isr.contrastive_loss(feats, t(labels)) * 1e-3
    loss.backward()
    got = pc._seg_feature.grad.cpu().numpy()

    # oracle: CPU forward, torch-CPU loss, oracle dense backward, torch-CPU activation chain
    # the oracle gets the feature rows exactly as the GPU normalised them (torch-CUDA and numpy norms differ by ulps)
    with torch.no_grad():
        segn = isr.normalize_rows(pc._seg_feature.detach(), 1e-6, 1e-9, stages=2)
        ref_n = pc.get_seg_feature
        ref_n = ref_n / (ref_n.norm(dim=-1, keepdim=True) + 1e-9)
        assert float((segn - ref_n).abs().max()) < 1e-6   # fused double normalisation == the reference's two torch ones
    inp["extra_attrs"] = segn.cpu().numpy()
    o = oracle_forward(oracle, inp)
    assert np.array_equal(pkg["seg_feature"].detach().cpu().numpy().view(np.uint32), o["extra"].view(np.uint32))
    emap = torch.tensor(o["extra"], requires_grad=True)
    f_o = emap.reshape(F, -1)[:, torch.tensor(pix)].t()
    l_o = contrastive_loss_ref(f_o, torch.tensor(labels)) * 1e-3
    l_o.backward()
    assert abs(float(loss) - float(l_o)) / abs(float(l_o)) < 1e-4
    og = oracle_backward(oracle, inp, o, np.zeros((3, H, W), np.float32), np.zeros((7, H, W), np.float32), emap.grad.numpy())
    raw = torch.tensor(sc.seg_feature_raw, requires_grad=True)
    a = raw / (torch.norm(raw, p=2, dim=1, keepdim=True) + 1e-6)
    a = a / (a.norm(dim=-1, keepdim=True) + 1e-9)
    a.backward(torch.tensor(og["dL_dextra"]))
    assert rel_err(got, raw.grad.numpy()) < 2e-4


def _golden_paths():
    import glob
    import os
    return sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g[0-9]_*.npz")))


@pytest.mark.parametrize("path", _golden_paths(), ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_matches_reference_goldens(path):
    """CUDA path vs outputs of the unmodified reference CUDA rasterizer (tests/golden/*.npz, generated on a B200 by
    tests/golden/make_golden.py): integers exact, floats / gradients within 1e-4."""
    from test_golden_cpu import SEEDS, check_against_reference, cotangents, load_fixture
    inp, ref = load_fixture(path)
    name = path.split("/")[-1][:-4]
    F = 0 if inp["extra_attrs"] is None else inp["extra_attrs"].shape[1]
    c = cuda_forward(inp)
    dcolor, dothers, dextra = cotangents(F, inp["W"], inp["H"], SEEDS[name])
    g = cuda_backward(inp, c, dcolor, dothers, dextra)
    c["clamped"] = None
    check_against_reference(c, ref, F, g, culled_lists=True)
    # Product default = the reference's arithmetic (no ISR_FLAG_SPEC_ARITH): every forward float is BIT-IDENTICAL to what
    # the unmodified reference CUDA rasterizer produced for this fixture.
    vis = ref["radii"] > 0
    for k in ("depths", "transMats", "means2D", "normal_opacity", "rgb"):
        assert np.array_equal(c[k][vis].view(np.uint32), np.asarray(ref[k], np.float32)[vis].view(np.uint32)), k
    for k in ("color", "others", "final_T") + (("extra",) if F else ()):
        assert np.array_equal(c[k].view(np.uint32), np.asarray(ref[k], np.float32).view(np.uint32)), k


def test_normalize_rows_matches_torch():
    import torch
    import instascene_b200 as isr
    torch.manual_seed(0)
    for F in (3, 16, 32):
        x = torch.rand(1000, F, device="cuda") + 0.01
        x[5] = 0.0  # zero row: 0 / (0 + eps)
        for stages in (1, 2):
            a = x.clone().requires_grad_(True)
            b = x.clone().requires_grad_(True)
            ya = isr.normalize_rows(a, 1e-6, 1e-9, stages)
            yb = b / (torch.norm(b, p=2, dim=1, keepdim=True) + 1e-6)
            if stages == 2:
                yb = yb / (yb.norm(dim=-1, keepdim=True) + 1e-9)
            g = torch.randn_like(x)
            ya.backward(g)
            yb.backward(g)
            assert float((ya - yb).abs().max()) < 1e-6
            assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < 1e-4


def test_lazy_render_package_equals_eager():
    import torch
    import instascene_b200 as isr
    inp = scene_inputs(3000, 8, 96, 64, 71)
    sc, cam = inp["scene"], inp["cam"]
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    class PC:
        active_sh_degree, max_sh_degree = 3, 3
        get_xyz, get_opacity = t(sc.xyz), t(sc.opacities()).reshape(-1, 1)
        get_scaling, get_rotation, get_features = t(sc.scales()), t(sc.rotations()), t(sc.shs())
        get_seg_feature = t(sc.seg_features())

    class Cam:
        FoVx, FoVy, image_width, image_height = cam.FoVx, cam.FoVy, 96, 64
        world_view_transform, full_proj_transform, camera_center = t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.camera_center)
        znear, zfar = 0.01, 100.0

    class Pipe:
        compute_cov3D_python, convert_SHs_python, depth_ratio = False, False, 0.3

    class PipeEager(Pipe):
        lazy_outputs = False

    lazy = isr.render(Cam(), PC(), Pipe(), t(inp["bg"]))
    assert "surf_normal" in lazy and len(lazy) == 13
    eager = isr.render(Cam(), PC(), PipeEager(), t(inp["bg"]))
    for k in eager:
        a, b = lazy[k], eager[k]
        if k == "gau_related_pixels":  # emission order is unspecified (Q7): compare as a set
            assert a.shape == b.shape and a.shape[0] > 0
            assert pair_set(a.cpu().numpy()) == pair_set(b.cpu().numpy())
            continue
        assert a.shape == b.shape and torch.equal(torch.nan_to_num(a.float()), torch.nan_to_num(b.float())), k


def test_sample_labelled_pixels_is_uniform_over_valid():
    import torch
    import instascene_b200 as isr
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    lab = torch.zeros(10000, dtype=torch.int16, device="cuda")
    valid = torch.tensor([3, 17, 4000, 9999], device="cuda")
    lab[valid] = torch.tensor([5, 6, 7, 8], dtype=torch.int16, device="cuda")
    pix, labels = isr.sample_labelled_pixels(lab, 40000, generator=g)
    assert set(pix.cpu().tolist()) == set(valid.cpu().tolist())
    assert torch.equal(labels.long(), lab[pix].long())
    counts = torch.bincount(pix, minlength=10000)[valid].float()
    assert float((counts / 10000.0 - 1.0).abs().max()) < 0.05


@pytest.mark.parametrize("dtype", ["int16", "int32", "int64"])
def test_fused_sampler_equals_torch_formulation(dtype):
    """The two-kernel sampler against round 1's torch formulation (mask, cumsum, rand, searchsorted, index) on the same
    `torch.rand` stream: identical pixel ids and labels -- label maps with a ragged size (not a multiple of the
    4096-pixel count block), a fully unlabelled stretch longer than a block, and a map with a single labelled pixel."""
    import torch
    import instascene_b200 as isr
    from instascene_b200.rasterizer import sample_labelled_pixels_torch
    dt = getattr(torch, dtype)
    rng = np.random.default_rng(5)
    H, W = 1080, 1917
    lab = rng.integers(-1, 60, size=H * W)
    lab[rng.random(H * W) < 0.4] = 0
    lab[100000:140000] = 0
    maps = [torch.tensor(lab, dtype=dt, device="cuda")]
    single = torch.zeros(5000, dtype=dt, device="cuda")
    single[4321] = 9
    maps.append(single)
    for m in maps:
        g1, g2 = torch.Generator(device="cuda"), torch.Generator(device="cuda")
        g1.manual_seed(11)
        g2.manual_seed(11)
        pix, labels = isr.sample_labelled_pixels(m, 32768, generator=g1)
        pix_t, labels_t = sample_labelled_pixels_torch(m, 32768, generator=g2)
        assert pix.dtype == torch.int64 and torch.equal(pix, pix_t)
        assert torch.equal(labels.long(), labels_t.long())
        assert bool((m[pix] > 0).all())


def test_fused_aux_maps_match_torch_glue():
    """isr_aux_maps_forward/backward vs the reference's torch post-processing (values and gradients wrt allmap)."""
    import torch
    from instascene_b200 import synth
    from instascene_b200.renderer import _derived_maps, _derived_maps_fused
    torch.manual_seed(0)
    W, H = 97, 61
    cam = synth.ring_cameras(5, W, H)[2]
    dev = "cuda:0"

    class Cam:
        image_width, image_height = W, H
        world_view_transform = torch.from_numpy(cam.world_view_transform).to(dev)
        full_proj_transform = torch.from_numpy(cam.full_proj_transform).to(dev)

    base = torch.rand(7, H, W, device=dev)
    base[0] = base[0] * 3 + 0.5          # depth*w
    base[1] = base[1] * 0.9 + 0.05       # alpha
    base[5] = base[5] * 3 + 0.5          # median depth
    base[1, 5:9, 7:12] = 0.0             # uncovered pixels: D = 0, alpha = 0 -> 0/0
    base[0, 5:9, 7:12] = 0.0
    base[5, 5:9, 7:12] = 0.0
    for ratio in (0.0, 1.0, 0.3):
        a = base.clone().requires_grad_(True)
        b = base.clone().requires_grad_(True)
        fa = _derived_maps_fused(a, Cam, ratio)
        fb = _derived_maps(b, Cam, ratio)
        loss_a = loss_b = 0.0
        gen = torch.Generator(device=dev)
        gen.manual_seed(1)
        for k in ("rend_alpha", "rend_normal", "rend_dist", "surf_depth", "surf_normal", "rend_depth", "rend_median_depth"):
            assert fa[k].shape == fb[k].shape, k
            assert float((fa[k] - fb[k]).abs().max()) <= 2e-5 * (float(fb[k].abs().max()) + 1e-6), (k, ratio)
            wgt = torch.randn(fa[k].shape, device=dev, generator=gen)
            loss_a = loss_a + (fa[k] * wgt).sum()
            loss_b = loss_b + (fb[k] * wgt).sum()
        loss_a.backward()
        loss_b.backward()
        ga, gb = a.grad, torch.nan_to_num(b.grad, 0.0, 0.0, 0.0)  # torch yields 0/0 on the uncovered pixels
        covered = (base[1] > 0).expand_as(ga)
        assert rel_err(ga[covered].cpu().numpy(), gb[covered].cpu().numpy()) < 2e-4, ratio


def _fwd_bwd_variant(oracle, inp, F, grads):
    H, W = inp["H"], inp["W"]
    rng = np.random.default_rng(5)
    dcolor = rng.standard_normal((3, H, W)).astype(np.float32)
    dothers = rng.standard_normal((7, H, W)).astype(np.float32)
    dextra = rng.standard_normal((F, H, W)).astype(np.float32) if F else None
    o = oracle_forward(oracle, inp)
    c = cuda_forward(inp)
    assert np.array_equal(c["radii"], o["radii"]) and c["num_rendered"] == o["num_rendered"] > 0
    check_tile_lists(c, o, W, H)
    for k in ["color", "others"] + (["extra"] if F else []):
        assert np.array_equal(c[k].view(np.uint32), o[k].view(np.uint32)), k
    og = oracle_backward(oracle, inp, o, dcolor, dothers, dextra)
    cg = cuda_backward(inp, c, dcolor, dothers, dextra)
    for k in grads:
        assert rel_err(cg[k], og[k].reshape(cg[k].shape)) < GRAD_TOL, (k, rel_err(cg[k], og[k].reshape(cg[k].shape)))
    return o, c


def test_variant_colors_precomp_and_low_sh_degree(oracle):
    inp = scene_inputs(2500, 4, 80, 64, 81)
    rng = np.random.default_rng(1)
    inp["colors_precomp"] = rng.uniform(0, 1, size=(2500, 3)).astype(np.float32)   # override_color path
    _fwd_bwd_variant(oracle, inp, 4, ["dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dextra"])
    for deg in (0, 1, 2):                                                           # active_sh_degree < max
        inp2 = scene_inputs(2500, 0, 80, 64, 82, sh_degree=deg)
        _fwd_bwd_variant(oracle, inp2, 0, ["dL_dsh", "dL_dmeans3D", "dL_dopacity"])


def test_variant_scale_modifier(oracle):
    inp = scene_inputs(2500, 0, 80, 64, 83)
    inp["scale_modifier"] = 0.6   # Q6: forward uses it, backward ignores it -- in the oracle and here alike
    _fwd_bwd_variant(oracle, inp, 0, ["dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dscales", "dL_drotations"])


def test_variant_transmat_precomp(oracle):
    """cov3D_precomp path (pipe.compute_cov3D_python): T given, normals (0,0,1), gradient returned in dL_dtransMat."""
    base = scene_inputs(2000, 0, 80, 64, 84)
    o = oracle_forward(oracle, base, blend=False)
    inp = dict(base)
    inp["transMat_precomp"] = o["transMats"].copy()
    inp["transMat_precomp"][o["radii"] == 0] = 0.0   # culled rows were never written by the oracle
    _fwd_bwd_variant(oracle, inp, 0, ["dL_dcolors", "dL_dopacity", "dL_dtransMat", "dL_dsh"])


def test_module_api_noncontiguous_inputs_and_flags(oracle):
    """GaussianRasterizer with non-contiguous / strided inputs (Q17), F = 1, debug=True, prefiltered=True."""
    import torch
    import instascene_b200 as isr
    inp = scene_inputs(1500, 1, 64, 48, 85)
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    s = isr.GaussianRasterizationSettings(48, 64, inp["tanfovx"], inp["tanfovy"], t(inp["bg"]), 1.0, t(inp["viewmatrix"]),
                                          t(inp["projmatrix"]), 3, t(inp["campos"]), True, True)
    wide = torch.zeros((1500, 6), device=dev)
    wide[:, ::2] = t(inp["means3D"])
    means_nc = wide[:, ::2]                                   # stride-2 view
    assert not means_nc.is_contiguous()
    scales_nc = t(inp["scales"]).t().contiguous().t()         # column-major
    out = isr.GaussianRasterizer(s)(means_nc, torch.zeros_like(means_nc), t(inp["opacities"]).reshape(-1, 1), shs=t(inp["shs"]),
                                    scales=scales_nc, rotations=t(inp["rotations"]), extra_attrs=t(inp["extra_attrs"]))
    o = oracle_forward(oracle, inp)
    assert np.array_equal(out[0].cpu().numpy().view(np.uint32), o["color"].view(np.uint32))
    assert np.array_equal(out[3].cpu().numpy().view(np.uint32), o["extra"].view(np.uint32))
    assert np.array_equal(out[1].cpu().numpy(), o["radii"]) and out[4].shape[0] == o["pair_count"]


def test_fused_adam_matches_torch():
    import torch
    import instascene_b200 as isr
    torch.manual_seed(0)
    for shape in ((1000, 16), (333, 3)):
        a = torch.randn(shape, device="cuda").requires_grad_(True)
        b = a.detach().clone().requires_grad_(True)
        oa = isr.FusedAdam([a], lr=0.025, eps=1e-15)
        ob = torch.optim.Adam([b], lr=0.025, eps=1e-15)
        for it in range(5):
            g = torch.randn(shape, device="cuda") * (0.0 if it == 2 else 1.0)
            a.grad, b.grad = g.clone(), g.clone()
            oa.step(); ob.step()
        assert float((a - b).abs().max()) < 1e-5 * float(b.abs().max())


@pytest.mark.parametrize("F,stages", [(16, 2), (7, 2), (32, 1)])
def test_fused_adam_with_deferred_rownorm_equals_unfused(F, stages):
    """normalize_rows(..., defer_to=param) + FusedAdam (chain rule of the normalisation inside the Adam kernel, device
    step counter) against the unfused path (rownorm backward kernel -> param.grad -> torch.optim.Adam), several steps,
    including a step with an extra ordinary gradient on the same parameter."""
    import torch
    import instascene_b200 as isr
    torch.manual_seed(3)
    P = 5000
    a = (torch.randn(P, F, device="cuda") * 2).requires_grad_(True)
    b = a.detach().clone().requires_grad_(True)
    oa = isr.FusedAdam([a], lr=0.025, eps=1e-15, capturable=True)
    ob = torch.optim.Adam([b], lr=0.025, eps=1e-15)
    for it in range(4):
        w = torch.randn(P, F, device="cuda")
        ya = isr.normalize_rows(a, 1e-6, 1e-9, stages=stages, defer_to=a)
        yb = isr.normalize_rows(b, 1e-6, 1e-9, stages=stages)
        la, lb = (ya * w).sum(), (yb * w).sum()
        if it == 2:  # another use of the raw parameter: its ordinary gradient is added after the chain rule
            la = la + (a[::7] ** 2).sum() * 0.01
            lb = lb + (b[::7] ** 2).sum() * 0.01
        la.backward(); lb.backward()
        assert getattr(a, "_isr_deferred_dy", None) is not None and (a.grad is None) == (it != 2)
        oa.step(); ob.step()
        oa.zero_grad(); ob.zero_grad()
        assert getattr(a, "_isr_deferred_dy", None) is None
    assert float((a - b).abs().max()) < 2e-5 * float(b.abs().max())


def test_render_geometry_first_and_param_ready_event():
    """render() launches the geometry phase before it touches the features and honours pc._isr_param_ready_event:
    a parameter update enqueued on a side stream is visible to the render that follows."""
    import torch
    import instascene_b200 as isr
    inp = scene_inputs(4000, 8, 96, 64, 91)
    sc, cam = inp["scene"], inp["cam"]
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    class PC:
        active_sh_degree, max_sh_degree = 3, 3
        get_xyz, get_opacity = t(sc.xyz), t(sc.opacities()).reshape(-1, 1)
        get_scaling, get_rotation, get_features = t(sc.scales()), t(sc.rotations()), t(sc.shs())
        _seg_feature = t(sc.seg_feature_raw).clone()

    class Cam:
        FoVx, FoVy, image_width, image_height = cam.FoVx, cam.FoVy, 96, 64
        world_view_transform, full_proj_transform, camera_center = t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.camera_center)

    class Pipe:
        compute_cov3D_python, convert_SHs_python, depth_ratio = False, False, 1.0

    pc = PC()
    new_feat = torch.rand_like(pc._seg_feature) + 0.1
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        big = torch.randn(4096, 4096, device=dev)
        for _ in range(20):
            big = big @ big * 1e-4          # keep the side stream busy so that a missing wait would be visible
        pc._seg_feature.copy_(new_feat)
        ev = torch.cuda.Event()
        ev.record(side)
    pc._isr_param_ready_event = ev
    got = isr.render(Cam(), pc, Pipe(), t(inp["bg"]), want_pairs=False)["seg_feature"].clone()
    torch.cuda.synchronize()
    pc._isr_param_ready_event = None
    want = isr.render(Cam(), pc, Pipe(), t(inp["bg"]), want_pairs=False)["seg_feature"]
    assert torch.equal(got, want)


def test_sampled_loss_with_trainable_geometry_uses_dense_backward():
    """A loss on sampled feature pixels (hybrid-sparse gradient on the hidden handle) while the geometry is trainable
    too: every gradient must equal the plain dense formulation `seg_map.reshape(F,-1)[:, ids]`; with frozen geometry the
    means2D proxy does not require grad and the feature gradient is the same."""
    import torch
    import instascene_b200 as isr
    P, F, W, H, seed = 3000, 8, 96, 64, 77
    inp = scene_inputs(P, F, W, H, seed)
    sc, cam = inp["scene"], inp["cam"]
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    class Cam:
        FoVx, FoVy, image_width, image_height = cam.FoVx, cam.FoVy, W, H
        world_view_transform, full_proj_transform, camera_center = t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.camera_center)
        znear, zfar = 0.01, 100.0

    class Pipe:
        compute_cov3D_python, convert_SHs_python, depth_ratio = False, False, 1.0

    def make_pc(train_geometry):
        class PC:
            active_sh_degree, max_sh_degree = 3, 3
            get_xyz = t(sc.xyz).requires_grad_(train_geometry)
            get_opacity = t(sc.opacities()).reshape(-1, 1).requires_grad_(train_geometry)
            get_scaling = t(sc.scales()).requires_grad_(train_geometry)
            get_rotation = t(sc.rotations())
            get_features = t(sc.shs())
            _seg_feature = t(sc.seg_feature_raw).requires_grad_(True)
            get_seg_feature = _seg_feature
        return PC()

    g = torch.Generator(device=dev)
    g.manual_seed(3)
    ids = torch.randint(0, W * H, (4096,), device=dev, generator=g)
    coef = torch.randn((4096, F), device=dev, generator=g)
    bg = torch.zeros(3, device=dev)
    grads = {}
    for mode in ("sampled", "dense", "sampled_frozen"):
        pc = make_pc(mode != "sampled_frozen")
        pkg = isr.render(Cam(), pc, Pipe(), bg, norm_seg_feat=False, want_pairs=False)
        seg = pkg["seg_feature"]
        feats = isr.sample_pixels(seg, ids) if mode != "dense" else seg.reshape(F, -1)[:, ids].t()
        (feats * coef).sum().backward()
        assert pkg["viewspace_points"].requires_grad == (mode != "sampled_frozen")
        grads[mode] = {k: getattr(pc, k).grad for k in ("get_xyz", "get_opacity", "get_scaling", "_seg_feature")}
        if mode != "sampled_frozen":
            grads[mode]["means2D"] = pkg["viewspace_points"].grad
    for k, ref in grads["dense"].items():
        got = grads["sampled"][k]
        assert got is not None and ref is not None, k
        assert ref.abs().max() > 0, k
        assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < 1e-4, k
    assert grads["sampled_frozen"]["get_xyz"] is None
    assert rel_err(grads["sampled_frozen"]["_seg_feature"].cpu().numpy(), grads["dense"]["_seg_feature"].cpu().numpy()) < 1e-4


def test_radix_sort_binning_fallback_matches_counting_partition():
    """The counting partition (default) and the radix-sort fallback (more tiles than fit the shared-memory table;
    forced here with ISR_BIN_SORT=1, read once per process) must produce the same lists: run the bit-exact forward test
    of one case in a fresh interpreter with the fallback forced."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ISR_BIN_SORT="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(root, "tests", "test_parity_gpu.py"),
                        "-k", "test_forward_bit_exact or test_backward_sparse_equals_dense"], env=env, cwd=root,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_unordered_scatter_with_segment_sort_matches_ordered_ranking():
    """The binning's shipped scatter (in-order ranking) and the experimental one (unordered slots + sort of every
    (run, tile) segment by depth rank; ISR_BIN_UNORDERED=1, read once per process) must produce the same lists: the
    bit-exact forward tests and the crowded-tile order tests in a fresh interpreter with the experimental kernel."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ISR_BIN_UNORDERED="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(root, "tests", "test_parity_gpu.py"),
                        "-k", "test_forward_bit_exact or test_backward_sparse_equals_dense or test_crowded"], env=env, cwd=root,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("P,W,H", [(400_000, 64, 48), (120_000, 160, 96)])
def test_crowded_tiles_keep_depth_order(P, W, H):
    """Hundreds of thousands of Gaussians on a dozen tiles: every (run, tile) segment of the partition holds many
    instances (with ISR_BIN_UNORDERED=1 the long-segment sort, > 32 entries, one CTA each, is exercised).  Every tile list must be ascending in (depth bits, Gaussian id) -- the reference's order -- and the tile
    ranges must partition the list."""
    inp = scene_inputs(P, 0, W, H, 97)
    c = cuda_forward(inp, want_pairs=False)
    ranges, pl = c["ranges"].astype(np.int64), c["point_list"].astype(np.int64)
    assert int((ranges[:, 1] - ranges[:, 0]).sum()) == len(pl) == int(c["tiles_emitted"].sum())
    key = c["depths"].view(np.uint32).astype(np.int64) << 32
    longest = 0
    for t in range(len(ranges)):
        ids = pl[ranges[t, 0]:ranges[t, 1]]
        k = key[ids] | ids
        assert np.all(np.diff(k) > 0), f"tile {t}: list not in (depth, id) order"
        longest = max(longest, len(ids))
    if W == 64:
        assert longest > 740 * 32  # pigeonhole over at most 740 runs: some (run, tile) segment is longer than 32 entries


def test_plain_entry_fallback_path():
    """Scenes with >= 2^24 Gaussians cannot carry the per-block footprint bits in the list entries; the blend kernels
    then test the footprint arithmetically.  ISR_PLAIN_ENTRIES=1 forces that path: forward parity + dense/sparse
    backward tests must pass unchanged (run in a subprocess: the switch is read once per process)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, ISR_PLAIN_ENTRIES="1")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_parity_gpu.py"), "-q", "-x", "-m", "gpu", "-k",
                        "test_forward_bit_exact or test_backward_dense or test_backward_sparse_equals_dense or test_empty"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("N,F,K,predef", [(2048, 8, 300, True), (1000, 4, 5, False), (3000, 32, 270, False), (257, 24, 33, True),
                                          # beyond round 1's shared-memory bound (K <= ~890 at F = 16, ~385 at F = 32): the
                                          # reference has no limit on the number of clusters
                                          (20000, 16, 2000, False), (9000, 32, 1100, True), (5000, 16, 5000, False)])
def test_contrastive_loss_shapes(N, F, K, predef):
    """Cluster counts beyond one 256-column MMA chunk, F below one 8-wide K step, ragged N (not a multiple of the
    128-row tile), absent clusters."""
    import torch
    import instascene_b200 as isr
    from instascene_b200 import synth
    from oracle.contrastive_ref import contrastive_loss_ref
    rng = np.random.default_rng(N + K)
    feats = rng.standard_normal((N, F)).astype(np.float32)
    labels = rng.integers(0, K + 1, size=N).astype(np.int64)
    labels[labels == 2] = 3
    proto = None
    if predef:
        proto = rng.standard_normal((K + 1, F)).astype(np.float32)
        proto /= np.linalg.norm(proto, axis=1, keepdims=True)
    f64 = torch.tensor(feats, dtype=torch.float64, requires_grad=True)
    want = contrastive_loss_ref(f64, torch.tensor(labels), None if proto is None else torch.tensor(proto, dtype=torch.float64))
    want.backward()
    fc = torch.tensor(feats, device="cuda", requires_grad=True)
    got = isr.contrastive_loss(fc, torch.tensor(labels, device="cuda"), None if proto is None else torch.tensor(proto, device="cuda"))
    got.backward()
    assert abs(float(got) - float(want)) / abs(float(want)) < 1e-4
    assert rel_err(fc.grad.cpu().numpy(), f64.grad.numpy()) < 1e-4
