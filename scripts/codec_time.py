"""Wall time of encode_model / decode_model on the bench scene (CUDA events around whole calls, like bench.py)."""
import os, sys, time
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
d = GaussianModel(device=dev)
d.load_state_dict({k: v for k, v in pc.state_dict().items() if not k.startswith("_")}, strict=False)

def timed(fn, k=8, w=3):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3

box = {}
def enc():
    box["e"] = codec.encode_model(pc)
ms_e = timed(enc)
e = box["e"]
def enc_plain():
    box["p"] = codec.encode_model(pc, estimate_bits=False)
ms_p = timed(enc_plain)
for lv, lp in zip(e.levels, box["p"].levels):
    for k in lv.streams:
        assert torch.equal(lv.streams[k].bytes, lp.streams[k].bytes) and torch.equal(lv.streams[k].lens, lp.streams[k].lens), k
print(f"encode without the entropy estimate {ms_p:.2f} ms (same bytes)")
def decf():
    box["o"] = codec.decode_model(d, e.meta, e.anchor_q, e.mask_bytes, e.mask_lens, e.hyper_bytes, e.hyper_lens, e.levels)
ms_d = timed(decf)
bits = codec.encoded_bits(e)["total"]
print(f"anchors {n} encode {ms_e:.2f} ms decode {ms_d:.2f} ms bits {bits/1e6:.1f} Mbit  -> {bits/ms_e/1e3:.0f} / {bits/ms_d/1e3:.0f} Mbit/s")
for k in ("feat", "scaling", "offsets"):
    assert torch.equal(box["o"][k].reshape(-1), e.quantised[k].reshape(-1)), k
print("round trip exact")
