"""cProfile of the host side of inference frames (prefilter_voxel + render) on the bench scene."""
import os, sys, cProfile, pstats
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
for i in range(10):
    frame(cams[i % 16])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(50):
    frame(cams[i % 16])
pr.disable()
st = pstats.Stats(pr).sort_stats("tottime")
st.print_stats(22)
