"""Generates tests/golden/loss.npz by importing the REFERENCE'S OWN utils/loss_utils.py (read from /root/reference,
never copied) on small seeded images, on the CPU.  Run in the build container only:

    python tests/golden/make_golden_loss.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from utils.loss_utils import l1_loss, ssim  # noqa: E402  (the reference's functions, unmodified)

out = {}
for name, (H, W, seed) in {"a": (37, 53, 0), "b": (16, 16, 1), "c": (7, 40, 2), "d": (64, 48, 3)}.items():
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(3, H, W, generator=g)
    img = (gt + 0.15 * torch.randn(3, H, W, generator=g)).clamp(0, 1).requires_grad_(True)
    Ll1, s = l1_loss(img, gt), ssim(img, gt)
    loss = 0.8 * Ll1 + 0.2 * (1.0 - s)       # train.py:204 with lambda_dssim = 0.2
    loss.backward()
    out[f"{name}_img"], out[f"{name}_gt"] = img.detach().numpy(), gt.numpy()
    out[f"{name}_l1"], out[f"{name}_ssim"] = np.float32(Ll1.item()), np.float32(s.item())
    out[f"{name}_grad"] = img.grad.numpy()
np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)
print("wrote", os.path.join(HERE, "loss.npz"), {k: v.shape for k, v in out.items() if k.endswith("_img")})
