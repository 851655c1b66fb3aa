"""device-side tracker extraction (isr_tracker_mark / isr_tracker_fill, instascene_b200.tracker)
vs the CPU restatement of get_segmap_gaussians (oracle/tracker_ref.py) and vs the reference's own output
(tests/golden/tracker_g1.npz, produced by tests/golden/make_tracker_golden.py from the unmodified reference)."""
import os

import numpy as np
import pytest

from helpers import cuda_forward, scene_inputs, tracker_label_map
from oracle.tracker_ref import segmap_gaussians_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tracker_g1.npz")


def _golden():
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/tracker_g1.npz not generated yet")
    return np.load(GOLDEN)


def _check_against_ref(ts, ref_info, ref_frame):
    assert [int(m) for m in ts.mask_ids] == sorted(ref_info)
    ids = ts.gaussian_ids.cpu().numpy()
    for j, m in enumerate(ts.mask_ids):
        assert np.array_equal(ids[ts.offsets[j]:ts.offsets[j + 1]].astype(np.int64), ref_info[int(m)]), m
    assert np.array_equal(ts.frame_ids.cpu().numpy().astype(np.int64), ref_frame)


# ---- CPU: the oracle itself ---------------------------------------------------------------------------------------
def test_tracker_oracle_small_known_answer():
    seg = np.array([0, 5, 5, 9], np.int16)
    pairs = np.array([[7, 1], [7, 2], [3, 1], [3, 3], [8, 0], [3, 3]], np.int32)
    info, frame = segmap_gaussians_ref(pairs, seg, min_gaussians=2)
    assert list(info) == [5] and info[5].tolist() == [3, 7]      # mask 9 has one Gaussian (< 2): dropped
    assert frame.tolist() == [3, 7, 8]
    info, _ = segmap_gaussians_ref(pairs, seg, min_gaussians=1)
    assert sorted(info) == [5, 9] and info[9].tolist() == [3]
    info, frame = segmap_gaussians_ref(np.zeros((0, 2), np.int32), seg)
    assert info == {} and frame.size == 0


def test_tracker_oracle_matches_reference_golden():
    g = _golden()
    info, frame = segmap_gaussians_ref(g["pairs"], g["segmap"].reshape(-1))
    assert sorted(info) == g["kept_mask_ids"].tolist()
    for m in info:
        assert np.array_equal(info[m], g[f"mask_{m}"])
    assert np.array_equal(frame, g["frame_ids"])


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("P,W,H,seed,min_g", [(5000, 128, 80, 21, 50), (20000, 160, 96, 22, 50), (3000, 70, 50, 23, 1),
                                              (40, 48, 32, 24, 3)])
def test_tracker_matches_oracle(P, W, H, seed, min_g):
    import torch
    from instascene_b200.tracker import segmap_gaussians
    inp = scene_inputs(P, 0, W, H, seed)
    c = cuda_forward(inp)
    pairs = np.ascontiguousarray(c["pairs"][:c["pair_count"]]).astype(np.int32)
    lab = tracker_label_map(W, H, seed)
    ref_info, ref_frame = segmap_gaussians_ref(pairs, lab.reshape(-1), min_gaussians=min_g)
    ts = segmap_gaussians(torch.from_numpy(pairs).cuda(), torch.from_numpy(lab).cuda(), P, min_gaussians=min_g)
    _check_against_ref(ts, ref_info, ref_frame)
    # counts of every mask (kept or not) are the oracle's set sizes
    full, _ = segmap_gaussians_ref(pairs, lab.reshape(-1), min_gaussians=0)
    assert [int(m) for m in ts.all_mask_ids] == sorted(full)
    assert ts.counts.tolist() == [len(full[int(m)]) for m in ts.all_mask_ids]


@pytest.mark.gpu
def test_tracker_edge_cases():
    import torch
    from instascene_b200.tracker import segmap_gaussians
    seg = torch.zeros(64, dtype=torch.int16, device="cuda")
    # empty pair list, all-background map
    ts = segmap_gaussians(torch.zeros((0, 2), dtype=torch.int32, device="cuda"), seg, 100)
    assert ts.mask_ids.size == 0 and ts.frame_ids.numel() == 0 and ts.gaussian_ids.numel() == 0
    # Gaussian ids at the edges of the bitmap words / rows, duplicates, one huge mask id
    seg[10:20] = 32000
    seg[20:30] = 2
    P = 97
    pairs = torch.tensor([[0, 10], [31, 11], [32, 12], [96, 13], [96, 13], [64, 25], [0, 0], [5, 63]], dtype=torch.int32,
                         device="cuda")
    ts = segmap_gaussians(pairs, seg, P, min_gaussians=1)
    ref_info, ref_frame = segmap_gaussians_ref(pairs.cpu().numpy(), seg.cpu().numpy(), min_gaussians=1)
    _check_against_ref(ts, ref_info, ref_frame)
    assert ts.mask_ids.tolist() == [2, 32000]


@pytest.mark.gpu
def test_tracker_matches_reference_golden():
    import torch
    from instascene_b200.tracker import segmap_gaussians
    g = _golden()
    ts = segmap_gaussians(torch.from_numpy(g["pairs"]).cuda(), torch.from_numpy(g["segmap"]).cuda(), int(g["P"]))
    _check_against_ref(ts, {int(m): g[f"mask_{m}"] for m in g["kept_mask_ids"]}, g["frame_ids"])


@pytest.mark.gpu
def test_get_segmap_gaussians_mirror_returns_reference_types():
    """The drop-in signature: (gaussian, view) -> ({mask_id: set}, list), through our own render()."""
    import torch
    from bench import _Cam, _Pipe
    from instascene_b200.tracker import get_segmap_gaussians
    P, W, H, seed = 5000, 128, 80, 31
    inp = scene_inputs(P, 16, W, H, seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

    class PC:
        active_sh_degree, max_sh_degree, pipelineparams = 3, 3, _Pipe
        get_xyz, get_opacity = t(inp["means3D"]), t(inp["opacities"]).reshape(-1, 1)
        get_scaling, get_rotation, get_features = t(inp["scales"]), t(inp["rotations"]), t(inp["shs"])
        get_seg_feature = t(inp["extra_attrs"])

    cam = inp["cam"]
    view = _Cam(cam, t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.camera_center))
    view.segmap = torch.from_numpy(tracker_label_map(W, H, seed))
    with torch.no_grad():
        mask_info, frame_ids = get_segmap_gaussians(PC(), view)
    assert isinstance(frame_ids, list) and all(isinstance(v, set) for v in mask_info.values())
    c = cuda_forward(inp)
    pairs = c["pairs"][:c["pair_count"]]
    ref_info, ref_frame = segmap_gaussians_ref(pairs, view.segmap.numpy().reshape(-1))
    assert sorted(mask_info) == sorted(ref_info) and len(ref_info) >= 2
    for m in ref_info:
        assert mask_info[m] == set(ref_info[m].tolist())
    assert sorted(frame_ids) == ref_frame.tolist()
