"""Generates tests/golden/growing.npz by exec'ing the REFERENCE'S OWN densification methods
(scene/gaussian_model.py:673-910: cat_tensors_to_optimizer, _prune_anchor_optimizer, prune_anchor, anchor_growing,
adjust_anchor; read from /root/reference, never copied) on a seeded CPU example.  Run in the build container only:

    python tests/golden/make_golden_growing.py

The module itself cannot be imported here (compressai / plyfile / simple_knn / torch_scatter are absent), so the method
sources are exec'd into a duck-typed class.  Three things are NOT the reference's own, all forced by this container:
  * device placement: `.cuda()` is dropped and `device='cuda'` becomes `device='cpu'` in the exec'd text (no GPU here);
  * `torch_scatter.scatter_max` is replaced by `Tensor.scatter_reduce_(..., 'amax', include_self=False)` (same values:
    every output row has at least one contributor on this path);
  * `torch.rand_like` is wrapped to RECORD its draws, which are stored in the fixture and replayed by the CUDA path
    (the CPU and CUDA generators produce different streams).
`get_anchor` goes through the reference's own `Quantize_anchor` (utils/encodings.py:219-231).
"""
import os
import sys
import textwrap
import types
from functools import reduce

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)
sys.modules.setdefault("torchac", types.ModuleType("torchac"))
import utils.encodings as enc  # noqa: E402  (the reference's)

src = open(os.path.join(REF, "scene/gaussian_model.py")).read()


def method_source(name, until):
    a = src.index(f"    def {name}(")
    b = src.index(f"    def {until}(")
    text = textwrap.dedent(src[a:b])
    return text.replace(".cuda()", "").replace("device='cuda'", "device='cpu'").replace('device="cuda"', 'device="cpu"')


def scatter_max(src_t, index, dim=0):
    n = int(index.max()) + 1 if index.numel() else 0
    out = torch.zeros((n,) + tuple(src_t.shape[1:]), dtype=src_t.dtype)
    out.scatter_reduce_(dim, index, src_t, "amax", include_self=False)
    return out, None


class RecordingTorch:
    """`torch` for the exec'd code: identical except that rand_like keeps a copy of what it returned."""

    def __init__(self):
        self.draws = []

    def __getattr__(self, k):
        return getattr(torch, k)

    def rand_like(self, t, *a, **kw):
        r = torch.rand_like(t, *a, **kw)
        self.draws.append(r.clone())
        return r


rt = RecordingTorch()
ns = dict(torch=rt, nn=nn, reduce=reduce, scatter_max=scatter_max,
          inverse_sigmoid=lambda x: torch.log(x / (1 - x)))       # utils/general_utils.py:20-21
pieces = [("cat_tensors_to_optimizer", "training_statis"), ("_prune_anchor_optimizer", "prune_anchor"),
          ("prune_anchor", "anchor_growing"), ("anchor_growing", "adjust_anchor"), ("adjust_anchor", "save_mlp_checkpoints")]
for name, until in pieces:
    exec(compile(method_source(name, until), f"reference:scene/gaussian_model.py:{name}", "exec"), ns)


class RefModel:
    n_offsets, feat_dim, hyper_divisor = 10, 50, 4
    update_depth, update_init_factor, update_hierachy_factor = 3, 16, 4        # arguments/__init__.py:53-55

    get_anchor = property(lambda self: enc.Quantize_anchor.apply(self._anchor, self.x_bound_min, self.x_bound_max)[0])
    get_scaling = property(lambda self: torch.exp(self._scaling))                # scene/gaussian_model.py:288 (1.0 * exp)


for name, _ in pieces:
    setattr(RefModel, name, ns[name])

NAMES = ("anchor", "offset", "mask", "anchor_feat", "hyper_latent", "opacity", "scaling", "rotation")


def make_state(N, seed, voxel):
    g = torch.Generator().manual_seed(seed)
    K = 10
    # anchors on the voxel grid inside a small box, so that grown cells collide with existing anchors at every depth
    anchor = torch.unique(torch.randint(-28, 28, (N, 3), generator=g), dim=0).float() * voxel
    anchor = anchor[torch.randperm(anchor.shape[0], generator=g)]
    N = anchor.shape[0]
    st = dict(anchor=anchor, offset=torch.randn(N, K, 3, generator=g) * 1.5, mask=torch.ones(N, K, 1),
              anchor_feat=torch.randn(N, 50, generator=g), hyper_latent=torch.randn(N, 12, generator=g),
              opacity=torch.zeros(N, 1), scaling=torch.log(voxel * (1 + 8 * torch.rand(N, 6, generator=g))),
              rotation=torch.cat([torch.ones(N, 1), torch.zeros(N, 3)], 1))
    st["scaling"][::17, 4] = 0.3                                       # exercises the clamp of _prune_anchor_optimizer (:729-733)
    stats = dict(opacity_accum=torch.rand(N, 1, generator=g) * 2.0,
                 anchor_demon=torch.randint(60, 120, (N, 1), generator=g).float(),
                 offset_gradient_accum=torch.rand(N * K, 1, generator=g) * 0.05,
                 offset_denom=torch.randint(0, 100, (N * K, 1), generator=g).float())
    stats["opacity_accum"][torch.rand(N, generator=g) < 0.2] = 0.0      # some anchors never contributed -> pruned
    return st, stats


def build(st, stats, voxel):
    m = RefModel()
    m.voxel_size = voxel
    groups = []
    for k in NAMES:
        p = nn.Parameter(st[k].clone().requires_grad_(True))
        setattr(m, "_" + k, p)
        groups.append({"params": [p], "lr": 1e-3, "name": k})
    m.mlp = nn.Linear(4, 4)
    groups.append({"params": list(m.mlp.parameters()), "lr": 1e-3, "name": "mlp_opacity"})
    m.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    # Adam state for every per-anchor group (to be extended / pruned), defined by exact arithmetic so that the test can
    # rebuild it bit for bit: exp_avg = p / 2, exp_avg_sq = p * p
    for k in NAMES:
        p = getattr(m, "_" + k)
        m.optimizer.state[p] = {"step": torch.tensor(1.0), "exp_avg": p.detach() * 0.5, "exp_avg_sq": p.detach() * p.detach()}
    lo, hi = m._anchor.min(0, keepdim=True)[0].detach(), m._anchor.max(0, keepdim=True)[0].detach()
    m.x_bound_min, m.x_bound_max = lo - 0.2 * lo.abs() - voxel, hi + 0.2 * hi.abs() + voxel
    for k, v in stats.items():
        setattr(m, k, v.clone())
    return m


def snapshot(m, prefix, out, adam):
    for k in NAMES:
        p = getattr(m, "_" + k)
        out[f"{prefix}_{k}"] = p.detach().numpy().copy()
        if adam:
            s = m.optimizer.state.get(p)
            out[f"{prefix}_{k}_exp_avg"] = s["exp_avg"].numpy().copy()
            if k in ("anchor", "scaling", "hyper_latent"):      # the others follow the same row selection (size)
                out[f"{prefix}_{k}_exp_avg_sq"] = s["exp_avg_sq"].numpy().copy()
    for k in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom"):
        out[f"{prefix}_{k}"] = getattr(m, k).numpy().copy()


out = {}
VOXEL = 0.01
torch.manual_seed(1234)
for case, (N, seed) in enumerate([(1000, 0), (250, 5), (150, 9)]):
    st, stats = make_state(N, seed, VOXEL)
    if case == 2:      # gradients far below the threshold: nothing grows at depth 0, so the finer depths are skipped (:774-776)
        stats["offset_gradient_accum"] *= 1e-4
    m = build(st, stats, VOXEL)
    out[f"c{case}_voxel"] = np.float64(VOXEL)
    out[f"c{case}_x_bound_min"], out[f"c{case}_x_bound_max"] = m.x_bound_min.numpy(), m.x_bound_max.numpy()
    snapshot(m, f"c{case}_before", out, adam=False)
    rt.draws.clear()
    with torch.no_grad():
        # train.py:247 -> adjust_anchor(check_interval=100, success_threshold=0.8, grad_threshold=0.0002, min_opacity=0.005)
        m.adjust_anchor(check_interval=100, success_threshold=0.8, grad_threshold=0.0002, min_opacity=0.005)
    for i, r in enumerate(rt.draws):
        out[f"c{case}_rand{i}"] = r.numpy()
    out[f"c{case}_n_rand"] = np.int64(len(rt.draws))
    snapshot(m, f"c{case}_after", out, adam=True)
    print(f"case {case}: {st['anchor'].shape[0]} anchors -> {m._anchor.shape[0]} after growing + pruning "
          f"({len(rt.draws)} rand draws)")
np.savez_compressed(os.path.join(HERE, "growing.npz"), **out)
print("wrote growing.npz", os.path.getsize(os.path.join(HERE, "growing.npz")) // 1024, "KiB")
