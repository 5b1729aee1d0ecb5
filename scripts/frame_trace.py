"""Kernel-level timeline of inference frames (prefilter_voxel + render) on the bench scene: where the GPU idles."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from contextgs_b200.renderer import prefilter_voxel, render

dev = torch.device("cuda", 0)
scene, dec, cams_cpu = bench.make_inputs(bench.N_ANCHORS)
pc = bench.make_model(scene, dev).replace_with_decoded(**{k: v.to(dev) for k, v in dec.items()})
pc.eval()
cams = [bench.cam_to(c, dev) for c in cams_cpu]
pipe = type("Pipe", (), {"debug": False})()
bg = torch.zeros(3, device=dev)

def frame(i):
    with torch.no_grad():
        vis = prefilter_voxel(cams[i % 16], pc, pipe, bg)
        return render(cams[i % 16], pc, pipe, bg, visible_mask=vis)

for i in range(10):
    frame(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50):
    frame(i)
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / 50 * 1e3:.3f} ms / frame")
# host time of the two calls (the GPU is idle while the host prepares a frame: render() ends with a read-back)
tp = tr = 0.0
for i in range(50):
    torch.cuda.synchronize()
    a = time.perf_counter()
    with torch.no_grad():
        vis = prefilter_voxel(cams[i % 16], pc, pipe, bg)
    b = time.perf_counter()
    with torch.no_grad():
        out = render(cams[i % 16], pc, pipe, bg, visible_mask=vis)
    c = time.perf_counter()
    tp += b - a; tr += c - b
print(f"host: prefilter_voxel returns after {tp / 50 * 1e6:.0f} us, render after {tr / 50 * 1e6:.0f} us more")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        frame(i)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
prev = None
for e in evs:
    gap = e.time_range.start - prev if prev is not None else 0.0
    print(f"   {e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:8.1f}  gap {gap:7.1f}  {e.name[:80]}")
    prev = e.time_range.end
