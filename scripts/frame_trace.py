"""Timeline of one inference frame (torch.profiler / CUPTI): kernel names, start offsets, durations, gaps.
Run on the GPU box: python scripts/frame_trace.py > gpurun_out/frame_trace.txt"""
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

def frame(cam):
    with torch.no_grad():
        vis = prefilter_voxel(cam, pc, pipe, bg)
        return render(cam, pc, pipe, bg, visible_mask=vis)

for i in range(8):
    frame(cams[i % 16])
torch.cuda.synchronize()
# host-side enqueue time per frame
t0 = time.perf_counter()
for i in range(20):
    frame(cams[i % 16])
torch.cuda.synchronize()
print("wall ms/frame", (time.perf_counter() - t0) / 20 * 1e3)

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        frame(cams[i % 16])
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t_prev_end = None
t_first = evs[0].time_range.start
for e in evs:
    s, d = e.time_range.start, e.time_range.end - e.time_range.start
    gap = 0 if t_prev_end is None else s - t_prev_end
    print(f"{(s - t_first):10.1f} us  dur {d:8.1f}  gap {gap:8.1f}  {e.name[:90]}")
    t_prev_end = e.time_range.end
