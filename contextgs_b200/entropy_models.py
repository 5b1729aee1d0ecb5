"""Drop-in for the hot-path part of the reference's utils/entropy_models.py: `Entropy_gaussian`
(:30-50) with `Low_bound` (:141-156) folded in.  One fused CUDA kernel forward (the reference
launches ~15 elementwise kernels), one backward (the reference's Low_bound.backward does two
device->host copies + numpy + a host->device copy per call)."""
import torch
import torch.nn as nn

from . import _lib


class _GaussianBits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mean, scale, Q, x_mean):
        shape = x.shape
        D = shape[-1]
        n = x.numel() // D
        xc = x.contiguous()
        mean = mean.expand(shape).contiguous()
        scale = scale.expand(shape).contiguous()
        per_elem = 0
        if Q.numel() == x.numel():
            Qc, per_elem = Q.expand(shape).contiguous(), 1
        elif Q.numel() == n:
            Qc = Q.reshape(n).contiguous()
        elif Q.numel() == 1:
            Qc = Q.reshape(1).expand(n).contiguous()
        else:
            raise ValueError("Entropy_gaussian: Q must be scalar, one per row, or elementwise")
        bits = torch.empty_like(xc)
        _lib.check(_lib.lib().cgs_gaussian_bits_forward(_lib.ptr(xc), _lib.ptr(mean), _lib.ptr(scale), _lib.ptr(Qc),
                                                        per_elem, float(x_mean), n, D, _lib.ptr(bits),
                                                        _lib.stream_ptr()), "cgs_gaussian_bits_forward")
        ctx.save_for_backward(xc, mean, scale, Qc)
        ctx.meta = (per_elem, float(x_mean), n, D, Q.shape)
        return bits

    @staticmethod
    def backward(ctx, g):
        xc, mean, scale, Qc = ctx.saved_tensors
        per_elem, x_mean, n, D, qshape = ctx.meta
        g = g.contiguous()
        dx, dm, ds = torch.empty_like(xc), torch.empty_like(xc), torch.empty_like(xc)
        dQ = torch.empty_like(Qc)
        _lib.check(_lib.lib().cgs_gaussian_bits_backward(_lib.ptr(xc), _lib.ptr(mean), _lib.ptr(scale), _lib.ptr(Qc),
                                                         per_elem, x_mean, n, D, _lib.ptr(g), _lib.ptr(dx),
                                                         _lib.ptr(dm), _lib.ptr(ds), _lib.ptr(dQ), _lib.stream_ptr()),
                   "cgs_gaussian_bits_backward")
        if len(qshape) == 0 or dQ.numel() != int(torch.Size(qshape).numel()):
            dQ = dQ.sum().reshape(qshape)
        else:
            dQ = dQ.reshape(qshape)
        return dx, dm, ds, dQ, None


class Entropy_gaussian(nn.Module):
    def __init__(self, Q=1):
        super().__init__()
        self.Q = Q

    def forward(self, x, mean, scale, Q=None, x_mean=None):
        if Q is None:
            Q = self.Q
        if not torch.is_tensor(Q):
            Q = torch.tensor(float(Q), dtype=x.dtype, device=x.device)
        if x_mean is None:
            x_mean = x.mean()
        x_mean = float(x_mean.detach()) if torch.is_tensor(x_mean) else float(x_mean)
        return _GaussianBits.apply(x, mean, scale, Q, x_mean)
