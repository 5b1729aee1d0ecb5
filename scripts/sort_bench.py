"""Micro-benchmark + correctness check of cgs_sort_pairs_u32 (onesweep) on the two shapes of the frame:
tile sort (R = 17.9 M pairs, 13 bits) and depth sort (P = 3.7 M, 32 bits, implicit iota values).
CGS_SORT_VARIANT selects the tile shape.  python scripts/sort_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from contextgs_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda")


def run(n, bits, with_vals, iters=20):
    g = torch.Generator(device="cuda").manual_seed(1)
    if bits == 32:
        keys = (torch.rand(n, generator=g, device=dev) * 50 + 0.2).view(torch.int32)
    else:
        keys = torch.randint(0, 8160, (n,), generator=g, device=dev, dtype=torch.int32)
    vals = torch.randperm(n, device=dev, dtype=torch.int32) if with_vals else None
    ko, vo, kt, vt = (torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4))
    nd = torch.tensor([n], dtype=torch.int32, device=dev)
    ws = torch.empty(L.cgs_sort_workspace_bytes(n, 0, bits), dtype=torch.uint8, device=dev)
    call = lambda: _lib.check(L.cgs_sort_pairs_u32(_lib.ptr(keys), _lib.ptr(vals), _lib.ptr(ko), _lib.ptr(vo), _lib.ptr(kt),
                                                   _lib.ptr(vt), _lib.ptr(nd), n, 0, bits, _lib.ptr(ws), ws.numel(),
                                                   _lib.stream_ptr()), "sort")
    call()
    torch.cuda.synchronize()
    sk, order = torch.sort(keys.to(torch.int64) & 0xffffffff, stable=True)
    ok = torch.equal(ko.to(torch.int64) & 0xffffffff, sk)
    ref_v = (vals if with_vals else torch.arange(n, device=dev, dtype=torch.int32))[order]
    ok = ok and torch.equal(vo, ref_v)
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    return ok, e0.elapsed_time(e1) / iters


for name, n, bits, wv in (("tile_sort", 17_864_826, 13, True), ("depth_sort", 3_705_176, 32, False),
                          ("ragged", 1_000_003, 20, True)):
    ok, ms = run(n, bits, wv)
    print(f"variant {os.environ.get('CGS_SORT_VARIANT', '0')} {name}: n={n} bits={bits} ok={ok} {ms:.4f} ms")
