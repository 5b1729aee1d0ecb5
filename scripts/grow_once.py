"""Times one depth of anchor growing at BASELINE configs[2] size (1.5 M anchors) on cuda:0:
the library call (GaussianModel.grow_cells) against the reference's torch expression of the same step
(scene/gaussian_model.py:778-816 restated with torch ops on the GPU; scatter_max -> scatter_reduce amax).
Usage: python scripts/grow_once.py [n_anchors] [candidate_fraction]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_ANCHORS
FRAC = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
scene, dec, cams_cpu = bench.make_inputs(N)
pc = bench.make_model(scene, torch.device("cuda", 0))
K = pc.n_offsets
cand = (torch.rand(N * K, generator=torch.Generator().manual_seed(1)) < FRAC).cuda()


def torch_expression(cur_size, true_division=False):
    """true_division: divide by a device tensor (IEEE division, what torch's CPU kernel and this library do) instead of
    by a Python scalar (torch's CUDA kernel multiplies by the reciprocal: cells differ at a handful of rounding ties)."""
    with torch.no_grad():
        anchor = pc.get_anchor
        all_xyz = anchor.unsqueeze(1) + pc._offset * pc.get_scaling[:, :3].unsqueeze(1)
        div = torch.full((), cur_size, device="cuda") if true_division else cur_size
        grid = torch.round(anchor / div).int()
        sel = torch.round(all_xyz.view(-1, 3)[cand] / div).int()
        uniq, inv = torch.unique(sel, return_inverse=True, dim=0)
        dup = torch.zeros(uniq.shape[0], dtype=torch.bool, device="cuda")
        for i in range(0, grid.shape[0], 4096):
            dup |= (uniq.unsqueeze(1) == grid[i:i + 4096]).all(-1).any(-1)
        keep = ~dup
        feat = pc._anchor_feat.unsqueeze(1).expand(-1, K, -1).reshape(-1, pc.feat_dim)[cand]
        out = torch.zeros(uniq.shape[0], pc.feat_dim, device="cuda")
        out.scatter_reduce_(0, inv.unsqueeze(1).expand(-1, pc.feat_dim), feat, "amax", include_self=False)
        return uniq[keep] * cur_size, out[keep]


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        r = fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps, r


res = {"n_anchors": N, "candidates": int(cand.sum())}
for cell in (16 * pc.voxel_size, 4 * pc.voxel_size, pc.voxel_size):
    ms, (na, nf, nh) = timed(lambda: pc.grow_cells(cand, cell), 5)
    entry = {"cell": cell, "new_anchors": int(na.shape[0]), "cgs_ms": round(ms, 3)}
    if os.environ.get("GROW_TORCH", "1") == "1" and cell == 16 * pc.voxel_size:
        ms_t, (ta, tf) = timed(lambda: torch_expression(cell), 1)
        entry["torch_expression_ms"] = round(ms_t, 1)
        entry["torch_expression_new_anchors"] = int(ta.shape[0])
        tb, tg = torch_expression(cell, true_division=True)
        entry["identical_to_true_division_expression"] = bool(torch.equal(tb, na) and torch.equal(tg, nf))
    res[f"depth_cell_{cell:g}"] = entry
print(json.dumps(res))
