"""Generates tests/golden/knn_g1.npz with the UNMODIFIED reference simple_knn (submodules/simple-knn, built into
baseline/_ref/simple_knn by baseline/build_ref.sh; CUDA only, so this runs on a GPU box):

    gpurun -- 'python tests/golden/make_knn_golden.py gpurun_out/golden'   (then copy the .npz into tests/golden/)

The fixture stores the seed and the reference's distCUDA2 output; the points are regenerated from the seed by
tests/helpers.knn_fixture_points."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))


def main(out_dir):
    from helpers import knn_fixture_points
    from simple_knn._C import distCUDA2
    P, seed = 20000, 4242
    pts = knn_fixture_points(P, seed)
    ref = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, "knn_g1.npz"), P=P, seed=seed, ref_mean_dist2=ref,
                        points_crc=np.uint32(np.bitwise_xor.reduce(pts.view(np.uint32).reshape(-1))))
    print("knn_g1: P", P, "mean", float(ref.mean()), "zeros", int((ref == 0).sum()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
