"""Drop-in for the hot-path part of the reference's utils/encodings.py: `STE_multistep` (:203-216),
`Quantize_anchor` (:219-231), `get_binary_vxl_size` (:15-32) and the constants (:10-13).  Forward
passes run as CUDA kernels through the C ABI; the autograd rules are the reference's
straight-through estimators."""
import ctypes

import torch

from . import _lib

anchor_round_digits = 16
Q_anchor = 1 / (2 ** anchor_round_digits - 1)
use_clamp = True
use_multiprocessor = False


def _cuda_f32(t, name):
    if not t.is_cuda or t.dtype != torch.float32:
        raise TypeError(f"{name} must be a float32 CUDA tensor (contextgs_b200 has no CPU path)")
    return t.contiguous()


class STE_multistep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, Q):
        x = _cuda_f32(input, "input")
        if not torch.is_tensor(Q):
            Q = torch.full((1,), float(Q), dtype=torch.float32, device=x.device)
        Qf = _cuda_f32(Q, "Q")
        # Q broadcasts over the trailing dims of `input` ([n,1] or [n,1,1] in the reference)
        n = Qf.numel()
        if n == 1:
            Qf = Qf.reshape(1).expand(x.shape[0] if x.dim() > 0 else 1).contiguous()
            n = Qf.numel()
        if x.numel() % n != 0 or x.shape[0] != n:
            raise ValueError("STE_multistep: Q must hold one step per leading row of input")
        out = torch.empty_like(x)
        _lib.check(_lib.lib().cgs_ste_multistep(_lib.ptr(x), _lib.ptr(Qf), n, x.numel() // n, _lib.ptr(out),
                                                _lib.stream_ptr()), "cgs_ste_multistep")
        return out

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None


class Quantize_anchor(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchors, min_v, max_v):
        a = _cuda_f32(anchors, "anchors")
        from .rasterizer import _host_floats
        mn = (ctypes.c_float * 3)(*_host_floats(min_v, 3))
        mx = (ctypes.c_float * 3)(*_host_floats(max_v, 3))
        out, qv = torch.empty_like(a), torch.empty_like(a)
        _lib.check(_lib.lib().cgs_quantize_anchor(_lib.ptr(a), mn, mx, a.shape[0], _lib.ptr(out), _lib.ptr(qv),
                                                  _lib.stream_ptr()), "cgs_quantize_anchor")
        ctx.mark_non_differentiable(qv)
        return out, qv

    @staticmethod
    def backward(ctx, grad_output, tmp):
        return grad_output, None, None


def get_binary_vxl_size(binary_vxl):
    """utils/encodings.py:15-32: ideal Bernoulli code length of the offset masks (+32 bits for Pg)."""
    ttl_num = binary_vxl.numel()
    pos_num = torch.sum(binary_vxl)
    neg_num = ttl_num - pos_num
    Pg = torch.clamp(pos_num / ttl_num, min=1e-6, max=1 - 1e-6)
    ttl_bit = pos_num * (-torch.log2(Pg)) + neg_num * (-torch.log2(1 - Pg)) + 32
    return Pg, ttl_bit, ttl_bit.item() / 8.0 / 1024 / 1024, ttl_num
