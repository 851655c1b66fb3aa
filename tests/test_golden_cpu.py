"""CPU tests (no GPU): the oracle is PINNED against outputs of the unmodified reference CUDA rasterizer, generated on
a B200 by tests/golden/make_golden.py (fixtures tests/golden/*.npz carry inputs + reference outputs + reference
internal buffers).  Integers must be exact; floats within 1e-4 (rtol, with a small atol for values near zero)."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "g[0-9]_*.npz")))
RTOL, ATOL = 1e-4, 1e-5


def load_fixture(path):
    z = np.load(path)
    inp = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    for k in ("W", "H", "sh_degree"):
        inp[k] = int(inp[k])
    for k in ("tanfovx", "tanfovy"):
        inp[k] = float(inp[k])
    if "extra_attrs" not in inp:
        inp["extra_attrs"] = None
    return inp, ref


def cotangents(F, W, H, seed):
    rng = np.random.default_rng(seed + 1)
    return (rng.standard_normal((3, H, W)).astype(np.float32), rng.standard_normal((7, H, W)).astype(np.float32),
            rng.standard_normal((F, H, W)).astype(np.float32) if F else None)


SEEDS = {"g0_rgb": 101, "g1_feat16": 102, "g2_feat24_odd": 103, "g3_big_splats": 104}


def check_against_reference(got, ref, F, grads=None, culled_lists=False):
    """Shared by the CPU (oracle) and GPU (CUDA) golden tests.  culled_lists: `got` comes from the CUDA path, whose
    tile lists omit instances that provably miss the tile; they are compared with the reference's lists through
    helpers.check_tile_lists."""
    assert int(got["num_rendered"]) == int(ref["num_rendered"])
    for k in ("radii", "tiles_touched") + (() if culled_lists else ("point_list", "ranges", "n_contrib")):
        assert np.array_equal(np.asarray(got[k]).astype(np.int64), np.asarray(ref[k]).astype(np.int64)), k
    if culled_lists:
        from helpers import check_tile_lists
        refd = {"point_list": np.asarray(ref["point_list"]), "ranges": np.asarray(ref["ranges"]).reshape(-1, 2),
                "n_contrib": np.asarray(ref["n_contrib"]).reshape(np.asarray(got["n_contrib"]).shape)}
        check_tile_lists(got, refd, np.asarray(got["color"]).shape[2], np.asarray(got["color"]).shape[1])
    vis = ref["radii"] > 0
    # per-Gaussian intermediates: 1e-4 relative, with an absolute floor of 1e-5 of the field's magnitude (the AABB
    # centre of a huge splat is a difference of large terms: a 3e-5 px deviation on a 0.01 px value is rounding)
    for k in ("depths", "means2D", "transMats", "normal_opacity", "rgb"):
        floor = max(ATOL, 1e-5 * float(np.abs(ref[k][vis]).max())) if vis.any() else ATOL
        np.testing.assert_allclose(got[k][vis], ref[k][vis], rtol=RTOL, atol=floor, err_msg=k)
    for k in ("color", "others", "final_T") + (("extra",) if F else ()):
        np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, atol=ATOL, err_msg=k)
    gp = np.unique(np.asarray(got["pairs"], np.int64).reshape(-1, 2), axis=0)
    assert int(got["pair_count"]) == int(ref["pair_count"])
    assert np.array_equal(gp, ref["pairs"].astype(np.int64)), "gau_related_pixels as a set"
    if grads is not None:
        for k in ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dtransMat", "dL_dsh", "dL_dscales",
                  "dL_drotations") + (("dL_dextra",) if F else ()):
            a, b = np.asarray(grads[k], np.float64), np.asarray(ref[k], np.float64).reshape(np.asarray(grads[k]).shape)
            err = np.abs(a - b).max() / (np.abs(b).max() + 1e-30)
            assert err < 1e-4, (k, err)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_matches_reference(oracle, path):
    inp, ref = load_fixture(path)
    name = os.path.basename(path)[:-4]
    F = 0 if inp["extra_attrs"] is None else inp["extra_attrs"].shape[1]
    W, H = inp["W"], inp["H"]
    o = oracle.forward(inp["means3D"], inp["opacities"], inp["viewmatrix"], inp["projmatrix"], inp["campos"], W, H,
                       inp["bg"], scales=inp["scales"], rotations=inp["rotations"], shs=inp["shs"],
                       sh_degree=inp["sh_degree"], extra_attrs=inp["extra_attrs"])
    dcolor, dothers, dextra = cotangents(F, W, H, SEEDS[name])
    g = oracle.backward(o, inp["means3D"], inp["viewmatrix"], inp["projmatrix"], inp["campos"], W, H, inp["bg"],
                        inp["tanfovx"], inp["tanfovy"], dcolor, dothers, dextra, scales=inp["scales"],
                        rotations=inp["rotations"], shs=inp["shs"], sh_degree=inp["sh_degree"],
                        extra_attrs=inp["extra_attrs"])
    check_against_reference(o, ref, F, g)


def test_fixtures_present():
    assert len(FIXTURES) >= 4


def test_oracle_knn_matches_reference_simple_knn(oracle):
    """The brute-force kNN oracle is pinned by distCUDA2 of the unmodified reference simple_knn
    (tests/golden/knn_g1.npz, tests/golden/make_knn_golden.py): bit-identical."""
    from helpers import knn_fixture_points
    path = os.path.join(HERE, "golden", "knn_g1.npz")
    z = np.load(path)
    pts = knn_fixture_points(int(z["P"]), int(z["seed"]))
    assert np.uint32(np.bitwise_xor.reduce(pts.view(np.uint32).reshape(-1))) == z["points_crc"]
    got = oracle.knn_mean_dist2(pts)
    assert np.array_equal(got.view(np.uint32), z["ref_mean_dist2"].view(np.uint32))
