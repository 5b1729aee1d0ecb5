"""Where does the host time of codec.encode_model go?  cProfile of one call on the bench scene (after warm-up)."""
import os, sys, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from contextgs_b200 import codec
scene, dec, cams_cpu = bench.make_inputs(bench.N_ANCHORS)
pc = bench.make_model(scene, torch.device("cuda", 0))
pc.eval()
for _ in range(2):
    enc = codec.encode_model(pc)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
enc = codec.encode_model(pc)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
