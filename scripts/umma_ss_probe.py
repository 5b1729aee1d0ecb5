"""Probe (NOT part of the test suite): does the SS form of
tcgen05.mma kind::tf32 with the tile's row index as K reproduce D = P^T Q?  Prints the relative L2 error against an
fp64 product for several shapes and leading-dimension skews.  Run on a B200:  python scripts/umma_ss_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from contextgs_b200 import _lib

L = _lib.lib()
for M, N in ((110, 64), (128, 48), (16, 16), (75, 32)):
    for skew in (0, 1):
        g = torch.Generator().manual_seed(M * 100 + N)
        P = torch.randn(128, M, generator=g).cuda()
        Q = torch.randn(128, N, generator=g).cuda()
        D = torch.zeros(M, N, device="cuda")
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(L.cgs_umma_selftest_ss(_lib.ptr(P), _lib.ptr(Q), M, N, skew, _lib.ptr(D), _lib.ptr(err),
                                          _lib.stream_ptr()), "cgs_umma_selftest_ss")
        torch.cuda.synchronize()
        ref = P.double().t() @ Q.double()
        rel = float((D.double() - ref).norm() / ref.norm())
        print(f"M={M} N={N} skew={skew}: timeout={int(err)} rel-L2={rel:.3e}")

for M, N in ((110, 48), (54, 32), (128, 48), (16, 16)):
    for variant in (0, 1):
        g = torch.Generator().manual_seed(M * 100 + N)
        P = torch.randn(128, M, generator=g).cuda()
        Q = torch.randn(128, N, generator=g).cuda()
        D = torch.zeros(M, N, device="cuda")
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(L.cgs_umma_selftest_ss_mn(_lib.ptr(P), _lib.ptr(Q), M, N, variant, _lib.ptr(D), _lib.ptr(err),
                                             _lib.stream_ptr()), "cgs_umma_selftest_ss_mn")
        torch.cuda.synchronize()
        ref = P.double().t() @ Q.double()
        rel = float((D.double() - ref).norm() / ref.norm())
        print(f"MN-major M={M} N={N} variant={variant}: timeout={int(err)} rel-L2={rel:.3e}")
