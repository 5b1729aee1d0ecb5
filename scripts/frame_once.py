"""A few inference frames on the bench scene (target for ncu)."""
import os, sys
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
for i in range(4):
    with torch.no_grad():
        vis = prefilter_voxel(cams[i], pc, pipe, bg)
        out = render(cams[i], pc, pipe, bg, visible_mask=vis)
torch.cuda.synchronize()
print("ok", out["radii"].shape[0])
