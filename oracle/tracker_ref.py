"""TEST INFRASTRUCTURE ONLY (checker for tests/): CPU restatement of the reference's per-view tracker extraction,
``get_segmap_gaussians`` -- spatial_track/modules/init_tracker.py:26-47 -- on numpy arrays.

Pinning: pure integer set arithmetic; `tests/golden/tracker_g1.npz` holds the output of the reference's own loop
(init_tracker.py:26-47 executed on the stored pair list, see tests/golden/make_tracker_golden.py)."""
import numpy as np


def segmap_gaussians_ref(pairs: np.ndarray, segmap_flat: np.ndarray, min_gaussians: int = 50):
    """pairs [G,2] (gaussian id, pixel id); returns ({mask_id: ascending unique ids}, ascending unique frame ids)."""
    gaus_ids = pairs[:, 0].astype(np.int64)        # init_tracker.py:26
    pixel_ids = pairs[:, 1].astype(np.int64)       # init_tracker.py:27
    mask_of_pair = segmap_flat.reshape(-1)[pixel_ids] if len(pixel_ids) else np.zeros(0, segmap_flat.dtype)
    mask_info = {}
    for mask_id in np.unique(segmap_flat):         # init_tracker.py:30-35 (sorted distinct ids)
        if mask_id == 0:                           # :36-37
            continue
        members = np.unique(gaus_ids[mask_of_pair == mask_id])   # :38-39,44 as a sorted array instead of a set
        if len(members) < min_gaussians:           # :41-42
            continue
        mask_info[int(mask_id)] = members
    return mask_info, np.unique(gaus_ids)          # :34
