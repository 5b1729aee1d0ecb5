"""Generates tests/golden/statis.npz by exec'ing the REFERENCE'S OWN `GaussianModel.training_statis`
(scene/gaussian_model.py:696-713, read from /root/reference, never copied; the module itself cannot be imported here
because of its compressai / plyfile / simple_knn imports) on a seeded CPU example.  Run in the build container only:

    python tests/golden/make_golden_statis.py
"""
import os
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
src = open("/root/reference/scene/gaussian_model.py").read()
a = src.index("    def training_statis(")
b = src.index("    def _prune_anchor_optimizer(")
ns = dict(torch=torch)
exec(compile(textwrap.dedent(src[a:b]), "reference:scene/gaussian_model.py[696:713]", "exec"), ns)
training_statis = ns["training_statis"]

g = torch.Generator().manual_seed(0)
N, K = 500, 10
out = {}
self = types.SimpleNamespace(n_offsets=K, opacity_accum=torch.zeros(N, 1), anchor_demon=torch.zeros(N, 1),
                             offset_gradient_accum=torch.zeros(N * K, 1), offset_denom=torch.zeros(N * K, 1))
for it in range(3):   # three consecutive iterations accumulate into the same state
    vis = torch.rand(N, generator=g) < 0.6
    nv = int(vis.sum())
    opacity = torch.randn(nv * K, 1, generator=g)
    keep = (opacity.view(-1) > 0) & (torch.rand(nv * K, generator=g) < 0.9)
    P = int(keep.sum())
    vsp = types.SimpleNamespace(grad=torch.randn(P, 3, generator=g))
    upd = torch.rand(P, generator=g) < 0.7
    training_statis(self, vsp, opacity, upd, keep, vis)
    for k, v in dict(vis=vis, opacity=opacity, keep=keep, grad=vsp.grad, upd=upd).items():
        out[f"it{it}_{k}"] = v.numpy()
    for k in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom"):
        out[f"it{it}_{k}"] = getattr(self, k).numpy().copy()
np.savez_compressed(os.path.join(HERE, "statis.npz"), **out)
print("wrote statis.npz", {k: v.shape for k, v in out.items() if k.startswith("it0_")})
