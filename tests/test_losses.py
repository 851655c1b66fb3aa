"""fused photometric loss (isr_photometric_forward/_backward, instascene_b200.losses) vs the numpy
oracle (oracle/loss_ref.py), the reference's own outputs (tests/golden/ssim_g1.npz from the unmodified
utils/loss_utils.py) and, at full size, the reference's formula evaluated by torch fp32 conv2d on the GPU.
Tolerances: fp32 kernel vs fp64 oracle / fp32 torch -> 2e-5 relative on the loss, 1e-4 norm-wise on the gradient."""
import os

import numpy as np
import pytest

from oracle.loss_ref import gaussian_window, photometric_loss_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ssim_g1.npz")
LOSS_TOL, GRAD_TOL = 2e-5, 1e-4


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))


# ---- CPU ------------------------------------------------------------------------------------------------------------
def test_window_matches_reference_formula():
    from oracle.loss_ref import gaussian_window_formula
    g = gaussian_window()
    assert len(g) == 11 and abs(g.sum() - 1) < 1e-6 and np.array_equal(g, g[::-1]) and abs(g[5] - 0.2660117149) < 1e-9
    assert np.allclose(g, gaussian_window_formula(), rtol=3e-7, atol=0)


@pytest.mark.parametrize("lam", [0.2, 1.0, 0.0])
def test_oracle_matches_reference_golden(lam):
    g = np.load(GOLDEN)
    tag = str(lam).replace(".", "p")
    loss, l1, ssim, grad = photometric_loss_ref(g["img"], g["gt"], lam)
    assert abs(loss - g[f"loss_{tag}"]) / g[f"loss_{tag}"] < LOSS_TOL
    assert abs(l1 - g["l1"]) / g["l1"] < LOSS_TOL and abs(ssim - g["ssim"]) / g["ssim"] < LOSS_TOL
    assert _rel(grad, g[f"grad_{tag}"].astype(np.float64)) < GRAD_TOL
    assert not grad[:, 5:9, 7:11].any() or lam > 0  # exact matches: zero L1 subgradient (torch.abs)


def test_oracle_gradient_is_the_finite_difference():
    rng = np.random.default_rng(3)
    x, y = rng.uniform(0.2, 0.8, (2, 14, 17)), rng.uniform(0.2, 0.8, (2, 14, 17))
    _, _, _, grad = photometric_loss_ref(x, y, 0.35)
    for (c, i, j) in [(0, 0, 0), (1, 7, 9), (0, 13, 16), (1, 3, 0)]:
        e = np.zeros_like(x)
        e[c, i, j] = 1e-6
        fd = (photometric_loss_ref(x + e, y, 0.35, False)[0] - photometric_loss_ref(x - e, y, 0.35, False)[0]) / 2e-6
        assert abs(fd - grad[c, i, j]) < 1e-6 * max(1.0, abs(fd) * 1e3), (c, i, j, fd, grad[c, i, j])


# ---- GPU ------------------------------------------------------------------------------------------------------------
def _torch_reference(img, gt, lam):
    """utils/loss_utils.py:18-19, 39-83 + train.py:76-77 restated with torch ops (fp32, conv2d) for the full-size check."""
    import torch
    import torch.nn.functional as F
    C = img.shape[0]
    g1 = torch.tensor(gaussian_window(), dtype=torch.float32, device=img.device).unsqueeze(1)
    win = (g1 @ g1.t()).expand(C, 1, 11, 11).contiguous()
    blur = lambda t: F.conv2d(t[None], win, padding=5, groups=C)[0]
    mu1, mu2 = blur(img), blur(gt)
    s1, s2, s12 = blur(img * img) - mu1 * mu1, blur(gt * gt) - mu2 * mu2, blur(img * gt) - mu1 * mu2
    m = ((2 * mu1 * mu2 + 0.01 ** 2) * (2 * s12 + 0.03 ** 2)) / ((mu1 * mu1 + mu2 * mu2 + 0.01 ** 2) * (s1 + s2 + 0.03 ** 2))
    return (1.0 - lam) * (img - gt).abs().mean() + lam * (1.0 - m.mean())


@pytest.mark.gpu
@pytest.mark.parametrize("lam", [0.2, 1.0, 0.0])
def test_cuda_matches_reference_golden(lam):
    import torch
    from instascene_b200.losses import photometric_loss
    g = np.load(GOLDEN)
    tag = str(lam).replace(".", "p")
    x = torch.from_numpy(g["img"]).cuda().requires_grad_(True)
    loss, parts = photometric_loss(x, torch.from_numpy(g["gt"]).cuda(), lam, return_parts=True)
    (loss * 3.0).backward()  # a non-trivial upstream gradient
    assert abs(loss.item() - g[f"loss_{tag}"]) / g[f"loss_{tag}"] < LOSS_TOL
    assert abs(parts[1].item() - g["l1"]) / g["l1"] < LOSS_TOL and abs(parts[2].item() - g["ssim"]) / g["ssim"] < LOSS_TOL
    assert _rel(x.grad.cpu().numpy() / 3.0, g[f"grad_{tag}"].astype(np.float64)) < GRAD_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("C,H,W,seed", [(3, 64, 96, 1), (1, 5, 7, 2), (3, 33, 31, 3), (4, 16, 32, 4)])
def test_cuda_matches_oracle(C, H, W, seed):
    import torch
    from instascene_b200.losses import l1_loss, photometric_loss, ssim
    rng = np.random.default_rng(seed)
    gt = rng.uniform(0, 1, (C, H, W)).astype(np.float32)
    img = np.clip(gt + 0.2 * rng.standard_normal((C, H, W)), 0, 1).astype(np.float32)
    loss_r, l1_r, ssim_r, grad_r = photometric_loss_ref(img, gt, 0.2)
    x = torch.from_numpy(img).cuda().requires_grad_(True)
    y = torch.from_numpy(gt).cuda()
    loss = photometric_loss(x, y, 0.2)
    loss.backward()
    assert abs(loss.item() - loss_r) / loss_r < LOSS_TOL
    assert _rel(x.grad.cpu().numpy(), grad_r) < GRAD_TOL
    assert abs(l1_loss(x, y).item() - l1_r) / l1_r < LOSS_TOL and abs(ssim(x, y).item() - ssim_r) / abs(ssim_r) < LOSS_TOL
    # reproducible bit for bit (fixed reduction order)
    assert photometric_loss(x, y, 0.2).item() == loss.item()


@pytest.mark.gpu
def test_cuda_matches_torch_at_1080p():
    import torch
    from instascene_b200.losses import photometric_loss
    gen = torch.Generator(device="cuda").manual_seed(5)
    gt = torch.rand((3, 1080, 1920), device="cuda", generator=gen)
    img = (gt + 0.1 * torch.randn((3, 1080, 1920), device="cuda", generator=gen)).clamp(0, 1)
    a = img.clone().requires_grad_(True)
    b = img.clone().requires_grad_(True)
    la = photometric_loss(a, gt, 0.2)
    lb = _torch_reference(b, gt, 0.2)
    la.backward()
    lb.backward()
    assert abs(la.item() - lb.item()) / lb.item() < LOSS_TOL
    assert _rel(a.grad.cpu().numpy(), b.grad.double().cpu().numpy()) < GRAD_TOL


@pytest.mark.gpu
def test_densification_stats_match_reference_ops():
    import torch
    from instascene_b200.losses import add_densification_stats
    gen = torch.Generator(device="cuda").manual_seed(9)
    P = 100_003
    radii = torch.randint(-2, 40, (P,), device="cuda", generator=gen, dtype=torch.int32).clamp(min=0)
    grad = torch.randn((P, 3), device="cuda", generator=gen)
    mr = torch.rand(P, device="cuda", generator=gen) * 30
    acc = torch.rand((P, 1), device="cuda", generator=gen)
    den = torch.randint(0, 5, (P, 1), device="cuda", generator=gen).float()
    mr_r, acc_r, den_r = mr.clone(), acc.clone(), den.clone()
    vis = radii > 0
    mr_r[vis] = torch.max(mr_r[vis], radii[vis])                       # train.py:140-141
    acc_r[vis] += torch.norm(grad[vis], dim=-1, keepdim=True)          # scene/gaussian_model.py:602-604
    den_r[vis] += 1
    add_densification_stats(mr, acc, den, radii, grad)
    assert torch.equal(mr, mr_r) and torch.equal(den, den_r)
    assert torch.allclose(acc, acc_r, rtol=1e-6, atol=0)
