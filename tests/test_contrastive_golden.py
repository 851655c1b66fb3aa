"""ProtoNCE loss vs the UNMODIFIED reference `contrastive_loss` (tests/golden/contrastive_g1.npz, produced on a B200 by
tests/golden/make_contrastive_golden.py): pins oracle/contrastive_ref.py (CPU, float64) and the CUDA kernels (GPU).
Tolerance 1e-4 relative on the loss, 1e-4 norm-wise on the gradient (the reference computes in fp32 on cuBLAS)."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "contrastive_g1.npz")
CASES = [("means", False, False), ("predef", True, False), ("negative", False, True)]


def _golden():
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/contrastive_g1.npz not generated yet")
    return np.load(GOLDEN)


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))


@pytest.mark.parametrize("tag,predef,neg", CASES)
def test_oracle_matches_reference_contrastive_loss(tag, predef, neg):
    import torch
    from oracle.contrastive_ref import contrastive_loss_ref
    g = _golden()
    x = torch.tensor(g["feats"], dtype=torch.float64, requires_grad=True)
    loss = contrastive_loss_ref(x, torch.tensor(g["labels"]), torch.tensor(g["proto"], dtype=torch.float64) if predef else None,
                                consider_negative=neg)
    loss.backward()
    assert abs(float(loss) - float(g[f"loss_{tag}"])) / abs(float(g[f"loss_{tag}"])) < 1e-4
    assert _rel(x.grad.numpy(), g[f"grad_{tag}"]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("tag,predef,neg", CASES)
def test_cuda_matches_reference_contrastive_loss(tag, predef, neg):
    import torch
    import instascene_b200 as isr
    g = _golden()
    x = torch.tensor(g["feats"], device="cuda", requires_grad=True)
    loss = isr.contrastive_loss(x, torch.tensor(g["labels"], device="cuda"),
                                torch.tensor(g["proto"], device="cuda") if predef else None, consider_negative=neg)
    loss.backward()
    assert abs(float(loss) - float(g[f"loss_{tag}"])) / abs(float(g[f"loss_{tag}"])) < 1e-4
    assert _rel(x.grad.cpu().numpy(), g[f"grad_{tag}"]) < 1e-4
