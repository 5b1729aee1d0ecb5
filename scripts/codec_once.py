"""One encode_model + decode_model on the bench scene (target for ncu: -k regex:gauss_level)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from contextgs_b200 import codec
from contextgs_b200.gaussian_model import GaussianModel

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_ANCHORS
scene, dec, cams_cpu = bench.make_inputs(n)
pc = bench.make_model(scene, dev)
pc.eval()
enc = codec.encode_model(pc)
d = GaussianModel(device=dev)
d.load_state_dict({k: v for k, v in pc.state_dict().items() if not k.startswith("_")}, strict=False)
out = codec.decode_model(d, enc.meta, enc.anchor_q, enc.mask_bytes, enc.mask_lens, enc.hyper_bytes, enc.hyper_lens, enc.levels)
torch.cuda.synchronize()
print("ok", bool(torch.equal(out["feat"], enc.quantised["feat"])))
