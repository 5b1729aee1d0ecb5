"""TEST INFRASTRUCTURE ONLY: CPU restatement (numpy, explicit index arithmetic) of the reference's anchor densification
and pruning, scene/gaussian_model.py:673-910:

    cat_tensors_to_optimizer :673-694   _prune_anchor_optimizer :715-745   prune_anchor :747-759
    anchor_growing           :762-854   adjust_anchor           :856-910

Pinned against golden vectors produced by the reference's own methods run on the CPU (tests/golden/make_golden_growing.py
-> tests/golden/growing.npz; that script lists the three forced stand-ins: device placement, torch_scatter.scatter_max,
recorded rand draws).  The product (contextgs_b200/densify.py + csrc/anchor_growing.cu) never imports this file.

State layout: `params[name]`, `exp_avg[name]`, `exp_avg_sq[name]` for the eight per-anchor Adam groups
(anchor, offset, mask, anchor_feat, hyper_latent, opacity, scaling, rotation) and `stats[...]` for the four accumulators.
"""
import numpy as np

from oracle import entropy_ref

NAMES = ("anchor", "offset", "mask", "anchor_feat", "hyper_latent", "opacity", "scaling", "rotation")
F32 = np.float32


def get_anchor(st):
    """get_anchor (:340-345) -> Quantize_anchor (utils/encodings.py:219-231)."""
    import torch
    a = entropy_ref.quantize_anchor(torch.from_numpy(st["params"]["anchor"]), torch.from_numpy(st["x_bound_min"]),
                                    torch.from_numpy(st["x_bound_max"]))[0]
    return a.numpy()


def grow_cells(anchor_q, offset, scaling, feat, hyper, candidate, cur_size):
    """One depth of anchor_growing, :778-801 and :812-816, for the candidate (anchor, offset) slots:
    returns (new_anchor[M,3], new_feat[M,50], new_hyper[M,12]) in the order of the lexicographically sorted unique cells."""
    cs = F32(cur_size)
    N, K = offset.shape[:2]
    all_xyz = anchor_q[:, None, :] + offset * scaling[:, None, :3]                     # :778 (fp32: mul, then add)
    grid = np.rint(anchor_q / cs).astype(np.int32)                                     # :783
    slots = np.nonzero(candidate)[0]
    sel = np.rint(all_xyz.reshape(-1, 3)[slots] / cs).astype(np.int32)                 # :785-786
    if sel.shape[0] == 0:
        return np.zeros((0, 3), F32), np.zeros((0, feat.shape[1]), F32), np.zeros((0, hyper.shape[1]), F32)
    uniq, inverse = np.unique(sel, axis=0, return_inverse=True)                        # :788 (sorted rows)
    inverse = inverse.reshape(-1)
    existing = set(map(tuple, grid.tolist()))
    fresh = np.array([tuple(r) not in existing for r in uniq.tolist()], bool)          # :790-802
    new_anchor = uniq[fresh].astype(F32) * cs                                          # :803
    src = slots // K
    mf = np.full((uniq.shape[0], feat.shape[1]), -np.inf, F32)
    mh = np.full((uniq.shape[0], hyper.shape[1]), -np.inf, F32)
    np.maximum.at(mf, inverse, feat[src])                                              # :812-813 scatter_max
    np.maximum.at(mh, inverse, hyper[src])                                             # :815-816
    return new_anchor, mf[fresh], mh[fresh]


def _append(st, new):
    """cat_tensors_to_optimizer :673-694 (+ the accumulator padding of :835-841)."""
    m = new["anchor"].shape[0]
    for k in NAMES:
        st["params"][k] = np.concatenate([st["params"][k], new[k]], 0)
        for s in ("exp_avg", "exp_avg_sq"):
            if st[s].get(k) is not None:
                st[s][k] = np.concatenate([st[s][k], np.zeros_like(new[k])], 0)
    for k in ("anchor_demon", "opacity_accum"):
        st["stats"][k] = np.concatenate([st["stats"][k], np.zeros((m, 1), F32)], 0)


def anchor_growing(st, grads, threshold, offset_mask, rands, voxel_size, update_depth=3, update_init_factor=16,
                   update_hierachy_factor=4):
    """:762-854.  rands[i]: the uniform draw of depth i (recorded / injected)."""
    K = st["params"]["offset"].shape[1]
    init_length = st["params"]["anchor"].shape[0] * K
    for i in range(update_depth):
        cur_threshold = threshold * ((update_hierachy_factor // 2) ** i)
        cand = (grads >= F32(cur_threshold)) & offset_mask
        cand &= rands[i] > F32(0.5 ** (i + 1))
        length_inc = st["params"]["anchor"].shape[0] * K - init_length
        if length_inc == 0:
            if i > 0:
                continue                                                                  # :774-776 (quirk kept)
        else:
            cand = np.concatenate([cand, np.zeros(length_inc, bool)])
        size_factor = update_init_factor // (update_hierachy_factor ** i)
        cur_size = voxel_size * size_factor
        p = st["params"]
        na, nf, nh = grow_cells(get_anchor(st), p["offset"], np.exp(p["scaling"]), p["anchor_feat"], p["hyper_latent"],
                                cand, cur_size)
        m = na.shape[0]
        if m == 0:
            continue
        new = dict(anchor=na, scaling=np.log(np.full((m, 6), F32(cur_size), F32)),
                   rotation=np.concatenate([np.ones((m, 1), F32), np.zeros((m, 3), F32)], 1), anchor_feat=nf,
                   hyper_latent=nh, offset=np.zeros((m, K, 3), F32), mask=np.ones((m, K, 1), F32),
                   opacity=np.full((m, 1), np.log(F32(0.1) / (F32(1) - F32(0.1))), F32))
        _append(st, new)


def prune_anchor(st, prune_mask):
    """prune_anchor :747-759 with _prune_anchor_optimizer :715-745 (incl. the clamp of scaling[:, 3:] to <= 0.05)."""
    keep = ~prune_mask
    for k in NAMES:
        st["params"][k] = st["params"][k][keep]
        for s in ("exp_avg", "exp_avg_sq"):
            if st[s].get(k) is not None:
                st[s][k] = st[s][k][keep]
        if k == "scaling":
            t = st["params"][k][:, 3:]
            t[t > 0.05] = 0.05


def adjust_anchor(st, rands, voxel_size, check_interval=100, success_threshold=0.8, grad_threshold=0.0002,
                  min_opacity=0.005):
    """:856-910."""
    s = st["stats"]
    K = st["params"]["offset"].shape[1]
    with np.errstate(invalid="ignore", divide="ignore"):
        grads = s["offset_gradient_accum"] / s["offset_denom"]
    grads[np.isnan(grads)] = 0.0
    grads_norm = np.abs(grads[:, 0])                                                  # norm over a size-1 axis
    offset_mask = s["offset_denom"][:, 0] > check_interval * success_threshold * 0.5
    anchor_growing(st, grads_norm, grad_threshold, offset_mask, rands, voxel_size)
    n_now = st["params"]["anchor"].shape[0]
    for k in ("offset_denom", "offset_gradient_accum"):
        s[k][offset_mask] = 0
        s[k] = np.concatenate([s[k], np.zeros((n_now * K - s[k].shape[0], 1), F32)], 0)
    prune_mask = s["opacity_accum"][:, 0] < F32(min_opacity) * s["anchor_demon"][:, 0]
    anchors_mask = s["anchor_demon"][:, 0] > check_interval * success_threshold
    prune_mask &= anchors_mask
    for k in ("offset_denom", "offset_gradient_accum"):
        s[k] = s[k].reshape(-1, K)[~prune_mask].reshape(-1, 1)
    s["opacity_accum"][anchors_mask] = 0
    s["anchor_demon"][anchors_mask] = 0
    s["opacity_accum"] = s["opacity_accum"][~prune_mask]
    s["anchor_demon"] = s["anchor_demon"][~prune_mask]
    if prune_mask.shape[0] > 0:
        prune_anchor(st, prune_mask)
    return prune_mask


def state_from_golden(g, case):
    """The `before` state of a fixture case, with the generator's exact Adam-state rule (exp_avg = p/2, exp_avg_sq = p*p)."""
    pre = f"c{case}_before_"
    params = {k: g[pre + k].copy() for k in NAMES}
    return dict(params=params, exp_avg={k: v * F32(0.5) for k, v in params.items()},
                exp_avg_sq={k: v * v for k, v in params.items()},
                stats={k: g[pre + k].copy() for k in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom")},
                x_bound_min=g[f"c{case}_x_bound_min"], x_bound_max=g[f"c{case}_x_bound_max"])
