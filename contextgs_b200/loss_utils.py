"""Drop-in for the hot part of the reference's utils/loss_utils.py: `l1_loss` (:17-18) and `ssim` (:33-64) as ONE
fused CUDA kernel each way (csrc/loss.cu), plus `l1_ssim`, which returns both terms of train.py:200-204 from a
single pass.  Differentiable w.r.t. the rendered image; the ground truth is data."""
import torch

from . import _lib


def _check(img, gt):
    if img.shape != gt.shape or img.dim() != 3 or img.shape[0] != 3:
        raise ValueError(f"expected two [3,H,W] images, got {tuple(img.shape)} and {tuple(gt.shape)}")
    if not img.is_cuda or img.dtype != torch.float32 or gt.dtype != torch.float32:
        raise TypeError("l1 / ssim: float32 CUDA tensors required (contextgs_b200 has no CPU path)")


class _L1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, gt):
        _check(img, gt)
        if gt.requires_grad:
            raise ValueError("l1 / ssim: the ground-truth image must not require a gradient")
        L = _lib.lib()
        x, y = img.detach().contiguous(), gt.detach().contiguous()
        _, H, W = x.shape
        need_grad = img.requires_grad
        maps = [torch.empty_like(x) for _ in range(3)] if need_grad else [None, None, None]
        sums = torch.empty(2, dtype=torch.float64, device=x.device)
        _lib.check(L.cgs_l1_ssim_forward(_lib.ptr(x), _lib.ptr(y), H, W, _lib.ptr(maps[0]), _lib.ptr(maps[1]),
                                         _lib.ptr(maps[2]), _lib.ptr(sums), _lib.stream_ptr()), "cgs_l1_ssim_forward")
        out = (sums / (3.0 * H * W)).float()
        if need_grad:
            ctx.save_for_backward(x, y, *maps)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_l1, g_ssim):
        x, y, dm, dp, dq = ctx.saved_tensors
        _, H, W = x.shape
        d_img = torch.empty_like(x)
        f = lambda g: None if g is None else g.detach().reshape(1).float().contiguous()
        g1, g2 = f(g_l1), f(g_ssim)
        _lib.check(_lib.lib().cgs_l1_ssim_backward(_lib.ptr(x), _lib.ptr(y), H, W, _lib.ptr(dm), _lib.ptr(dp), _lib.ptr(dq),
                                                   _lib.ptr(g1), _lib.ptr(g2), _lib.ptr(d_img), _lib.stream_ptr()),
                   "cgs_l1_ssim_backward")
        return d_img, None


def l1_ssim(network_output, gt):
    """(mean |a - b|, mean SSIM) in one pass: both terms of train.py:200-204."""
    return _L1SSIM.apply(network_output, gt)


def l1_loss(network_output, gt):
    """utils/loss_utils.py:17-18."""
    if network_output.dim() == 3 and network_output.shape[0] == 3 and network_output.is_cuda:
        return _L1SSIM.apply(network_output, gt)[0]
    return torch.abs(network_output - gt).mean()   # other shapes (train.py:358 calls it on batches) stay in torch


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:33-64 for [3,H,W] images (the only form train.py uses), window 11, size_average=True."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("contextgs_b200.ssim implements the reference's training call: window 11, mean over the map")
    return _L1SSIM.apply(img1, img2)[1]
