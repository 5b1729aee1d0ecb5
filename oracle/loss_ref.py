"""TEST INFRASTRUCTURE ONLY: CPU restatement (torch fp32) of the reference's photometric loss,
utils/loss_utils.py `l1_loss` (:17-18) and `ssim` / `_ssim` (:33-64).  Pinned against golden vectors produced by
importing the reference's own module (tests/golden/make_golden_loss.py -> tests/golden/loss.npz)."""
from math import exp

import torch
import torch.nn.functional as F


def window(window_size=11, sigma=1.5, channel=3):
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(channel, 1, window_size, window_size).contiguous()


def l1_loss(a, b):
    return torch.abs(a - b).mean()


def ssim(img1, img2, window_size=11):
    c = img1.size(-3)
    w = window(window_size, 1.5, c).type_as(img1)
    pad = window_size // 2
    mu1, mu2 = F.conv2d(img1, w, padding=pad, groups=c), F.conv2d(img2, w, padding=pad, groups=c)
    s11 = F.conv2d(img1 * img1, w, padding=pad, groups=c) - mu1 * mu1
    s22 = F.conv2d(img2 * img2, w, padding=pad, groups=c) - mu2 * mu2
    s12 = F.conv2d(img1 * img2, w, padding=pad, groups=c) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s11 + s22 + C2))
    return m.mean()
