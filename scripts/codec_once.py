import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from contextgs_b200 import codec
scene, dec, cams_cpu = bench.make_inputs(int(sys.argv[1]) if len(sys.argv) > 1 else 400000)
pc = bench.make_model(scene, torch.device("cuda", 0))
pc.eval()
enc = codec.encode_model(pc)
torch.cuda.synchronize()
print("ok", codec.encoded_bits(enc)["total"])
