"""Timing experiments on the G1 weight-gradient kernel (cgs_debug_set key 1): which part bounds it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from contextgs_b200 import _lib
from contextgs_b200.renderer import prefilter_voxel, render
dev = torch.device("cuda", 0)
scene, dec, cams_cpu = bench.make_inputs(bench.N_ANCHORS)
pc = bench.make_model(scene, dev)
pc.train()
cams = [bench.cam_to(c, dev) for c in cams_cpu]
pipe = type("Pipe", (), {"debug": False})()
bg = torch.zeros(3, device=dev)
gt = torch.rand(3, cams[0].image_height, cams[0].image_width, device=dev)
def step(i):
    cam = cams[i % 16]
    with torch.no_grad():
        vis = prefilter_voxel(cam, pc, pipe, bg)
    out = render(cam, pc, pipe, bg, visible_mask=vis, retain_grad=False, step=100)
    (out["render"] - gt).abs().mean().backward()
    for p in list(pc.parameters()) + [pc._anchor, pc._anchor_feat, pc._offset, pc._scaling, pc._mask, pc._hyper_latent]:
        p.grad = None
for mode in (0, 1, 2, 4, 3, 5, 6, 7):
    _lib.lib().cgs_debug_set(1, mode)
    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    _lib.stage_timing(True)
    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    ms, n = _lib.stage_timing_read()
    _lib.stage_timing(False)
    print(f"mode {mode} (skip: {'mma ' if mode & 1 else ''}{'convert ' if mode & 2 else ''}{'copies' if mode & 4 else ''}): "
          f"neural_gaussians_bwd stage {ms.get('neural_gaussians_bwd', 0) / 4:.3f} ms per iteration")
_lib.lib().cgs_debug_set(1, 0)
