"""Small driver for ncu: a few whole-scene scoring passes (estimate_final_bits) on the bench scene."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000
scene, dec, cams = bench.make_inputs(n)
pc = bench.make_model(scene, torch.device("cuda", 0)).eval()
for i in range(3):
    sums = pc.estimate_final_bits(return_values=True)
torch.cuda.synchronize()
print("bits", sum(sums[1:5]) / 1e6, "Mbit")
