"""Diagnostic (not part of the test suite): G1 backward on tcgen05 vs the fp32-FMA kernel for both descriptor variants
of the weight-gradient kernel's MN-major operands, with timings.  Run on a B200."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from contextgs_b200 import _lib, synthetic
from contextgs_b200.gaussian_model import GaussianModel
from contextgs_b200.neural_gaussians import generate_neural_gaussians

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
scene = synthetic.make_scene("bicycle", N, seed=4)
cam = synthetic.make_cameras("bicycle", 4, device="cuda")[3]
vis = (torch.rand(N, generator=torch.Generator().manual_seed(6)) < 0.77).cuda()
L = _lib.lib()


def run(impl, variant=0, reps=1):
    os.environ["CGS_G1_BWD_IMPL"] = impl
    L.cgs_debug_set(0, variant)
    torch.manual_seed(6)
    model = GaussianModel.from_tensors(scene, device="cuda").train()
    t = []
    for r in range(reps):
        for p in model.parameters():
            p.grad = None
        out = generate_neural_gaussians(cam, model, vis, is_training=True, step=0)
        g = torch.Generator(device="cuda").manual_seed(3)
        ws = [torch.randn(x.shape, generator=g, device="cuda") for x in out[:5]]
        loss = sum((x * w).sum() for x, w in zip(out[:5], ws))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss.backward()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    _lib.raise_deferred()
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}, min(t), out[0].shape[0]


ref, t_ref, P = run("simt", reps=3)
print(f"N={N} visible={int(vis.sum())} P={P}  simt backward (autograd incl. glue): {t_ref:.3f} ms")
for variant in (0, 1):
    try:
        got, t_u, _ = run("umma", variant, reps=3)
    except Exception as e:
        print(f"variant {variant}: FAILED {type(e).__name__}: {e}")
        break
    rel = {n: float((got[n].double() - ref[n].double()).norm() / (ref[n].double().norm() + 1e-30)) for n in ref}
    print(f"variant {variant}: umma backward {t_u:.3f} ms;", {k: f"{v:.1e}" for k, v in rel.items()})
