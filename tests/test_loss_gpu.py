"""Fused L1 + SSIM kernels (csrc/loss.cu) against the reference's own outputs (golden vectors) and the CPU oracle.
Tolerances: 1e-5 absolute on the two means, 1e-4 relative L2 on the image gradient (north_star's float budget)."""
import os

import numpy as np
import pytest
import torch

from contextgs_b200.loss_utils import l1_loss, l1_ssim, ssim
from oracle import loss_ref

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss.npz"))


def _rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("name", list("abcd"))
def test_against_reference_golden(name):
    img = torch.from_numpy(G[f"{name}_img"]).cuda().requires_grad_(True)
    gt = torch.from_numpy(G[f"{name}_gt"]).cuda()
    l1, s = l1_ssim(img, gt)
    assert abs(float(l1) - float(G[f"{name}_l1"])) < 1e-5
    assert abs(float(s) - float(G[f"{name}_ssim"])) < 1e-5
    (0.8 * l1 + 0.2 * (1.0 - s)).backward()       # train.py:204
    assert _rel(img.grad.cpu().numpy(), G[f"{name}_grad"]) < 1e-4


def test_drop_in_functions_and_full_hd_against_oracle():
    g = torch.Generator().manual_seed(5)
    gt = torch.rand(3, 1080, 1920, generator=g)
    img = (gt + 0.1 * torch.randn(3, 1080, 1920, generator=g)).clamp(0, 1)
    a = img.cuda().requires_grad_(True)
    loss = 0.8 * l1_loss(a, gt.cuda()) + 0.2 * (1.0 - ssim(a, gt.cuda()))
    loss.backward()
    b = img.clone().requires_grad_(True)
    ref = 0.8 * loss_ref.l1_loss(b, gt) + 0.2 * (1.0 - loss_ref.ssim(b, gt))
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5
    assert _rel(a.grad.cpu().numpy(), b.grad.numpy()) < 1e-4
    # no gradient requested: the derivative maps are not even written
    with torch.no_grad():
        l1, s = l1_ssim(img.cuda(), gt.cuda())
    assert abs(float(l1) - float(loss_ref.l1_loss(img, gt))) < 1e-6 and abs(float(s) - float(loss_ref.ssim(img, gt))) < 1e-5


def test_identical_images_and_argument_checks():
    x = torch.rand(3, 20, 33).cuda()
    l1, s = l1_ssim(x, x.clone())
    assert float(l1) == 0.0 and abs(float(s) - 1.0) < 1e-6
    with pytest.raises(ValueError):
        l1_ssim(x, torch.rand(3, 20, 32).cuda())
    with pytest.raises(TypeError):
        l1_ssim(x.cpu(), x.cpu())
    with pytest.raises(NotImplementedError):
        ssim(x, x, window_size=7)


def test_images_smaller_than_the_window():
    g = torch.Generator().manual_seed(7)
    for H, W in ((1, 1), (3, 5), (11, 4)):
        gt = torch.rand(3, H, W, generator=g)
        img = torch.rand(3, H, W, generator=g)
        a = img.cuda().requires_grad_(True)
        l1, s = l1_ssim(a, gt.cuda())
        (l1 - s).backward()
        b = img.clone().requires_grad_(True)
        r1, rs = loss_ref.l1_loss(b, gt), loss_ref.ssim(b, gt)
        (r1 - rs).backward()
        assert abs(float(l1) - float(r1)) < 1e-6 and abs(float(s) - float(rs)) < 1e-5
        assert _rel(a.grad.cpu().numpy(), b.grad.numpy()) < 1e-4
