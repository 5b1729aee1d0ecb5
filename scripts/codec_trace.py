"""Kernel-level timeline of one encode_model / decode_model call on the bench scene (torch.profiler)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from contextgs_b200 import codec
from contextgs_b200.gaussian_model import GaussianModel

dev = torch.device("cuda", 0)
scene, dec, cams_cpu = bench.make_inputs(bench.N_ANCHORS)
pc = bench.make_model(scene, dev)
pc.eval()
enc = codec.encode_model(pc)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
def summarize(prof, title):
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = collections.defaultdict(lambda: [0.0, 0])
    for e in evs:
        agg[e.name[:70]][0] += e.time_range.end - e.time_range.start
        agg[e.name[:70]][1] += 1
    t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
    print(title, "span us", t1 - t0, "busy us", sum(v[0] for v in agg.values()))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"   {v[0]:10.1f} us  x{v[1]:4d}  {k}")
    for e in sorted(evs, key=lambda e: e.time_range.start):
        if "cgs::" in e.name:
            print(f"      {e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:8.1f}  {e.name[:60]}")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    enc = codec.encode_model(pc)
    torch.cuda.synchronize()
summarize(prof, "encode")
d = GaussianModel(device=dev)
d.load_state_dict({k: v for k, v in pc.state_dict().items() if not k.startswith("_")}, strict=False)
codec.decode_model(d, enc.meta, enc.anchor_q, enc.mask_bytes, enc.mask_lens, enc.hyper_bytes, enc.hyper_lens, enc.levels)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    codec.decode_model(d, enc.meta, enc.anchor_q, enc.mask_bytes, enc.mask_lens, enc.hyper_bytes, enc.hyper_lens, enc.levels)
    torch.cuda.synchronize()
summarize(prof, "decode")
for lv in enc.levels:
    print("level", lv.level, "rows", lv.n, {k: (st.minmax.tolist(), st.bytes.numel()) for k, st in lv.streams.items()})
