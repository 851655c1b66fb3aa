"""Generates tests/golden/contrastive_g1.npz with the UNMODIFIED reference `contrastive_loss`
(utils/contrastive_utils.py, imported from baseline/_ref; it hard-codes .cuda(), so this runs on a GPU box):

    gpurun -- 'python tests/golden/make_contrastive_golden.py gpurun_out/golden'   (then copy the .npz into tests/golden/)

Cases: cluster means / predefined prototypes / consider_negative, with an absent label id in the middle; fp32 forward
and autograd gradient w.r.t. the features."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))


def main(out_dir):
    import ref_loader
    from instascene_b200 import synth
    _, rloss, _ = ref_loader.load()
    rng = np.random.default_rng(2024)
    N, F, K = 3000, 16, 29
    feats = rng.standard_normal((N, F)).astype(np.float32)
    labels = rng.integers(0, K + 1, size=N).astype(np.int64)   # 0 = unlabelled
    labels[labels == 7] = 8                                     # an absent id in the middle
    proto = synth.gram_schmidt_prototypes(K + 1, F, 11)
    out = {"feats": feats, "labels": labels, "proto": proto}
    for tag, predef, neg in (("means", False, False), ("predef", True, False), ("negative", False, True)):
        x = torch.tensor(feats, device="cuda", requires_grad=True)
        loss = rloss(x, torch.tensor(labels, device="cuda"),
                     predef_u_list=torch.tensor(proto, device="cuda") if predef else None, consider_negative=neg)
        loss.backward()
        out[f"loss_{tag}"] = np.float32(loss.item())
        out[f"grad_{tag}"] = x.grad.cpu().numpy()
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, "contrastive_g1.npz"), **out)
    print({k: (v.shape if getattr(v, "ndim", 0) else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
