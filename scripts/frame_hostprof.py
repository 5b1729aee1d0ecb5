"""cProfile of the host side of inference frames (what the GPU waits for between frames)."""
import os, sys, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from contextgs_b200.renderer import prefilter_voxel, render

dev = torch.device("cuda", 0)
scene, dec, cams_cpu = bench.make_inputs(200000)
pc = bench.make_model(scene, dev).replace_with_decoded(**{k: v.to(dev) for k, v in dec.items()})
pc.eval()
cams = [bench.cam_to(c, dev) for c in cams_cpu]
pipe = type("Pipe", (), {"debug": False})()
bg = torch.zeros(3, device=dev)

def frame(i):
    with torch.no_grad():
        vis = prefilter_voxel(cams[i % 16], pc, pipe, bg)
        return render(cams[i % 16], pc, pipe, bg, visible_mask=vis)

for i in range(20):
    frame(i)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(300):
    frame(i)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
