import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from contextgs_b200.renderer import prefilter_voxel, render
dev = torch.device("cuda", 0)
scene, dec, cams_cpu = bench.make_inputs(int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_ANCHORS)
pc = bench.make_model(scene, dev)
pc.train()
cams = [bench.cam_to(c, dev) for c in cams_cpu]
pipe = type("Pipe", (), {"debug": False})()
bg = torch.zeros(3, device=dev)
gt = torch.rand(3, cams[0].image_height, cams[0].image_width, device=dev)
for i in range(2):
    cam = cams[i]
    with torch.no_grad():
        vis = prefilter_voxel(cam, pc, pipe, bg)
    out = render(cam, pc, pipe, bg, visible_mask=vis, retain_grad=False, step=20000)
    loss = (out["render"] - gt).abs().mean() + 0.004 * out["bit_per_param"]
    loss.backward()
torch.cuda.synchronize()
print("ok")
