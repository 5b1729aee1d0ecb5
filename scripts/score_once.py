import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
scene, dec, cams_cpu = bench.make_inputs(bench.N_ANCHORS)
pc = bench.make_model(scene, torch.device("cuda", 0))
pc.eval()
for _ in range(2):
    r = pc.estimate_final_bits(return_values=True)
torch.cuda.synchronize()
print("ok", r[2])
