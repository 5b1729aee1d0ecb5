"""Diagnostic: measured cost of tcgen05.mma.kind::tf32 (M = 128, K = 8) chains on B200 -- cycles per instruction by operand
form (A in TMEM / shared memory), N and the number of independent accumulators.  Run on a B200."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from contextgs_b200 import _lib

L = _lib.lib()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
iters = 512
for form in (2, 4, 6, 3, 5, 7):
    for N in (16, 64) if form >= 4 else (16, 64, 256):
        for n_acc in (1,):
            if n_acc * N > 256:
                continue
            for rep in range(2):
                _lib.check(L.cgs_umma_mma_rate(form, N, n_acc, iters, _lib.ptr(out), _lib.stream_ptr()), "cgs_umma_mma_rate")
                torch.cuda.synchronize()
            issue, total = out.tolist()
            print(f"{('TS tid==0', 'SS tid==0', 'TS elect', 'SS elect', 'TS 2 warps', 'SS 2 warps', 'TS 4 warps', 'SS 4 warps')[form]} N={N:3d} accumulators={n_acc}: issue {issue / iters:6.1f} cyc/mma, "
                  f"issue+complete {total / iters:6.1f} cyc/mma (floor 128*N/256 = {N / 2:.0f})")
