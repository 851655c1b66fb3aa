"""Generates tests/golden/ssim_g1.npz with the UNMODIFIED reference losses (utils/loss_utils.py imported from
/root/reference, CPU torch, fp32 forward + autograd):  python tests/golden/make_loss_golden.py
Only runs where /root/reference exists (the build container)."""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_loss_utils", "/root/reference/utils/loss_utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(77)
    C, H, W = 3, 37, 53   # not multiples of the CUDA tile, smaller than two windows in neither direction
    gt = rng.uniform(0, 1, (C, H, W)).astype(np.float32)
    # a rendered image = smoothed ground truth + noise, so that SSIM is neither ~0 nor ~1
    img = np.clip(gt + 0.15 * rng.standard_normal((C, H, W)), 0, 1).astype(np.float32)
    img[:, 5:9, 7:11] = gt[:, 5:9, 7:11]  # exact matches: |x-y| has a zero subgradient there
    out = {"img": img, "gt": gt}
    for lam in (0.2, 1.0, 0.0):
        x = torch.from_numpy(img).clone().requires_grad_(True)
        y = torch.from_numpy(gt)
        l1 = ref.l1_loss(x, y)
        ss = ref.ssim(x, y)
        loss = (1.0 - lam) * l1 + lam * (1.0 - ss)     # train.py:76-77
        loss.backward()
        tag = str(lam).replace(".", "p")
        out[f"loss_{tag}"] = np.float32(loss.item())
        out[f"grad_{tag}"] = x.grad.numpy().copy()
        out["l1"], out["ssim"] = np.float32(l1.item()), np.float32(ss.item())
    np.savez_compressed(os.path.join(HERE, "ssim_g1.npz"), **out)
    print({k: (v.shape if getattr(v, "ndim", 0) else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
