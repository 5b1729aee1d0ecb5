"""GPU check of the tcgen05 building blocks (contextgs_b200/csrc/umma.cuh) through the diagnostic
C-ABI entry cgs_umma_selftest: D[128,N] = A[128,K] W[N,K]^T against an fp64 product."""
import numpy as np
import pytest
import torch

from contextgs_b200 import _lib
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def run(A, W, mode):
    L = _lib.lib()
    N, K = W.shape
    D = torch.full((128, N), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(L.cgs_umma_selftest(_lib.ptr(A), _lib.ptr(W), N, K, mode, _lib.ptr(D), _lib.ptr(err),
                                   _lib.stream_ptr()), "cgs_umma_selftest")
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "tcgen05 completion barrier timed out"
    return D


@pytest.mark.parametrize("N,K", [(16, 8), (32, 56), (80, 56), (176, 56), (112, 64), (256, 64)])
def test_tile_gemm_matches_fp64(N, K):
    g = torch.Generator().manual_seed(N * 100 + K)
    A = torch.randn(128, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    ref = (A.double() @ W.double().t()).cpu().numpy()
    d1 = run(A, W, 0).cpu().numpy()
    d3 = run(A, W, 1).cpu().numpy()
    assert np.isfinite(d1).all() and np.isfinite(d3).all()
    assert rel_l2(d1, ref) < 2e-3          # plain TF32: 10-bit mantissas
    assert rel_l2(d3, ref) < 2e-6          # 3xTF32: fp32-grade
    assert np.array_equal(run(A, W, 2).cpu().numpy(), d3)   # tcgen05.ld at unaligned column offsets
    # structure: each output element depends on its own row / column only
    A2 = A.clone()
    A2[5] = 0
    d = run(A2, W, 1).cpu().numpy()
    assert np.abs(d[5]).max() == 0 and np.array_equal(d[6], d3[6])
