"""Kernel-level summary of one training iteration (forward + backward) in both regimes (torch.profiler)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from contextgs_b200.renderer import prefilter_voxel, render

dev = torch.device("cuda", 0)
scene, dec, cams_cpu = bench.make_inputs(bench.N_ANCHORS)
pc = bench.make_model(scene, dev)
pc.train()
cams = [bench.cam_to(c, dev) for c in cams_cpu]
pipe = type("Pipe", (), {"debug": False})()
bg = torch.zeros(3, device=dev)
H, W = cams[0].image_height, cams[0].image_width
gt = torch.rand(3, H, W, device=dev)

def step(i, st):
    cam = cams[i % 16]
    with torch.no_grad():
        vis = prefilter_voxel(cam, pc, pipe, bg)
    out = render(cam, pc, pipe, bg, visible_mask=vis, retain_grad=False, step=st)
    loss = (out["render"] - gt).abs().mean() + 0.01 * out["scaling"].prod(dim=1).mean()
    if out["bit_per_param"] is not None:
        loss = loss + 0.004 * out["bit_per_param"]
    loss.backward()
    for p in list(pc.parameters()) + [pc._anchor, pc._anchor_feat, pc._offset, pc._scaling, pc._mask, pc._hyper_latent]:
        p.grad = None

from torch.profiler import profile, ProfilerActivity
for st in (100, 20000):
    for i in range(4):
        step(i, st)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for i in range(8):
        step(i, st)
    torch.cuda.synchronize()
    print(f"step={st}: {(time.perf_counter() - t0) / 8 * 1e3:.2f} ms / iteration")
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step(0, st)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = collections.defaultdict(lambda: [0.0, 0])
    for e in evs:
        agg[e.name[:80]][0] += e.time_range.end - e.time_range.start
        agg[e.name[:80]][1] += 1
    t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
    print("  span us", round(t1 - t0), "busy us", round(sum(v[0] for v in agg.values())), "kernels", len(evs))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:22]:
        print(f"   {v[0]:10.1f} us  x{v[1]:4d}  {k}")
    ks = sorted(evs, key=lambda e: e.time_range.start)
    gaps = []
    for a, b in zip(ks, ks[1:]):
        g = b.time_range.start - a.time_range.end
        if g > 25:
            gaps.append((g, a.time_range.end - t0, a.name[:50], b.name[:50]))
    print("  largest gaps (us, at, after kernel, before kernel); total of gaps > 25 us:", round(sum(g[0] for g in gaps)))
    for g in sorted(gaps, reverse=True)[:18]:
        print(f"   {g[0]:8.1f} at {g[1]:9.1f}  {g[2]}  ->  {g[3]}")

