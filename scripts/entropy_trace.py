"""Kernel-level timeline of one estimate_final_bits scoring pass (cached level plan) on the bench scene."""
import os, sys, collections, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda", 0)
scene, dec, cams_cpu = bench.make_inputs(bench.N_ANCHORS)
pc = bench.make_model(scene, dev)
pc.eval()
for _ in range(3):
    pc.estimate_final_bits(return_values=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    pc.estimate_final_bits(return_values=True)
torch.cuda.synchronize()
print(f"pass {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    pc.estimate_final_bits(return_values=True)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
print("span us", t1 - t0, "busy us", sum(e.time_range.end - e.time_range.start for e in evs), "launches", len(evs))
for e in sorted(evs, key=lambda e: e.time_range.start):
    print(f"   {e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}")
