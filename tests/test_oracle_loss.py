"""The CPU restatement of the photometric loss (oracle/loss_ref.py) against golden vectors produced by the reference's
own utils/loss_utils.py (tests/golden/make_golden_loss.py)."""
import os

import numpy as np
import torch

from oracle import loss_ref

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss.npz"))


def test_oracle_loss_matches_reference_golden():
    for name in "abcd":
        img = torch.from_numpy(G[f"{name}_img"]).requires_grad_(True)
        gt = torch.from_numpy(G[f"{name}_gt"])
        l1, s = loss_ref.l1_loss(img, gt), loss_ref.ssim(img, gt)
        assert abs(float(l1) - float(G[f"{name}_l1"])) < 1e-7
        assert abs(float(s) - float(G[f"{name}_ssim"])) < 1e-6
        (0.8 * l1 + 0.2 * (1.0 - s)).backward()
        ref = G[f"{name}_grad"]
        assert np.linalg.norm(img.grad.numpy() - ref) / np.linalg.norm(ref) < 1e-5
