"""Anchor-level autoregressive context / entropy model: drop-in for `multi_scale_generating`,
`find_divide_scale`, `divide_levels` (scene/gaussian_model.py:1541-1793) on top of the fused CUDA
kernels in contextgs_b200/csrc/context_model.cu.

Per call: 1 EntropyBottleneck kernel + 3 fused level kernels (coarse -> fine).  The level index
bookkeeping (which rows each level codes, where each row's context comes from -- including the
reference's mid-level row misalignment) is flattened once into int32 index
arrays (`LevelPlan`) and cached while anchors and masks are unchanged.
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .encodings import get_binary_vxl_size

N_CODED = 86  # 50 feat + 6 scaling + 30 offsets

def pack_grid_weights(pc, level):
    """W1[in][100] | b1[100] | W2[100][176] | b2[176] (k-major) for `pc.get_grid_mlp[level]`."""
    m = pc.get_grid_mlp[level]
    params = (m[0].weight, m[0].bias, m[2].weight, m[2].bias)
    key = tuple((p.data_ptr(), p._version) for p in params)
    cache = _lib.object_cache(pc)
    ent = cache.get(("grid_pack", level))
    if ent is not None and ent[0] == key:
        return ent[1], ent[2]
    with torch.no_grad():
        dev = params[0].device
        in_dim = m[0].weight.shape[1]
        W2 = torch.zeros(100, 176, device=dev)
        W2[:, :175] = m[2].weight.t()
        b2 = torch.zeros(176, device=dev)
        b2[:175] = m[2].bias
        packed = torch.cat([m[0].weight.t().reshape(-1), m[0].bias, W2.reshape(-1), b2]).float().contiguous()
    assert packed.numel() == _lib.lib().cgs_context_level_packed_floats(in_dim)
    cache[("grid_pack", level)] = (key, packed, in_dim)
    return packed, in_dim


def pack_grid_weights_umma(pc, level):
    """tcgen05 block of `pc.get_grid_mlp[level]` (csrc/context_model_umma.cu `cmu::Layout`):
    W1 hi | W1 lo ([K1p/4][112][4]) | W2 hi | W2 lo ([26][176][4]) | b1[112] | b2[176]."""
    from .neural_gaussians import umma_b_operand
    m = pc.get_grid_mlp[level]
    params = (m[0].weight, m[0].bias, m[2].weight, m[2].bias)
    key = tuple((p.data_ptr(), p._version) for p in params)
    cache = _lib.object_cache(pc)
    ent = cache.get(("grid_pack_umma", level))
    if ent is not None and ent[0] == key:
        return ent[1], ent[2]
    with torch.no_grad():
        dev = params[0].device
        in_dim = m[0].weight.shape[1]
        k1p = (in_dim + 7) // 8 * 8
        b1 = torch.zeros(112, device=dev)
        b1[:100] = m[0].bias
        b2 = torch.zeros(176, device=dev)
        b2[:175] = m[2].bias
        packed = torch.cat([*umma_b_operand(m[0].weight, 112, k1p), *umma_b_operand(m[2].weight, 176, 104), b1, b2])
        packed = packed.float().contiguous()
    assert packed.numel() == _lib.lib().cgs_context_level_umma_packed_floats(in_dim)
    cache[("grid_pack_umma", level)] = (key, packed, in_dim)
    return packed, in_dim


def pack_grid_weights_bwd_umma(pc, level):
    """Transposed B operands of the tcgen05 data-gradient kernel (csrc/context_model_bwd_umma.cu `cbu::DLayout`):
    W2^T as [hidden 112][K = 176 outputs], W1^T as [input 80 | 16][K = 104 hidden], each split into TF32 hi / lo.
    Cached per parameter version."""
    from .neural_gaussians import umma_b_operand
    m = pc.get_grid_mlp[level]
    params = (m[0].weight, m[2].weight)
    key = tuple((p.data_ptr(), p._version) for p in params)
    cache = _lib.object_cache(pc)
    ent = cache.get(("grid_pack_bwd_umma", level))
    if ent is not None and ent[0] == key:
        return ent[1]
    with torch.no_grad():
        in_dim = m[0].weight.shape[1]
        n1 = 80 if in_dim == 71 else 16
        packed = torch.cat([*umma_b_operand(m[2].weight.t().contiguous(), 112, 176),      # B[j][n] = W2[n][j]
                            *umma_b_operand(m[0].weight.t().contiguous(), n1, 104)])      # B[i][h] = W1[h][i]
        packed = packed.float().contiguous()
    assert packed.numel() == _lib.lib().cgs_context_level_bwd_umma_packed_floats(in_dim)
    cache[("grid_pack_bwd_umma", level)] = (key, packed)
    return packed


def ctx_bwd_impl():
    """'umma' (tcgen05 kernels fed by what the training-mode forward saves; default with the tcgen05 forward) or 'simt'
    (fp32 FMA kernel that recomputes the forward; kept for cross-checks)."""
    import os
    v = os.environ.get("CGS_CTX_BWD_IMPL", "umma" if ctx_impl() == "umma" else "simt")
    if v not in ("umma", "simt"):
        raise ValueError("CGS_CTX_BWD_IMPL must be 'umma' or 'simt'")
    if v == "umma" and ctx_impl() != "umma":
        raise ValueError("CGS_CTX_BWD_IMPL=umma needs the tcgen05 forward (CGS_CTX_IMPL=umma)")
    return v


def ctx_impl():
    """'umma' (tcgen05 tensor cores, default) or 'simt' (fp32 FMA tiles; kept for cross-checks)."""
    import os
    v = os.environ.get("CGS_CTX_IMPL", "umma")
    if v not in ("umma", "simt"):
        raise ValueError("CGS_CTX_IMPL must be 'umma' or 'simt'")
    return v


# ----------------------------------------------------------------------------- level division (E1-E3)

def unique_voxels(points, voxel_size, scale, keep=None):
    """rows = round(points / voxel_size / scale) -> (count, inverse[n] int64, first[count] int64) with the
    semantics of utils/multi_level.py:3-31 (sorted unique rows, minimum source index).  CUDA tensors go
    through cgs_unique_voxels (own radix sort + chained scan, no host sync besides reading `count`);
    CPU tensors (host-logic tests only) use the same torch calls as the reference."""
    n = points.shape[0]
    if not points.is_cuda:
        rows = torch.round((points if keep is None else points * keep.unsqueeze(1)) / voxel_size / scale)
        uniq, inverse = torch.unique(rows, return_inverse=True, dim=0)
        first = torch.full((uniq.shape[0],), n, dtype=torch.long)
        first.scatter_reduce_(0, inverse, torch.arange(n), reduce="amin")
        return uniq.shape[0], inverse, first
    L = _lib.lib()
    dev = points.device
    pts = points.detach().contiguous().float()
    inverse = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    first = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    status = torch.empty(2, dtype=torch.int32, device=dev)
    ws = torch.empty(L.cgs_unique_voxels_workspace_bytes(n), dtype=torch.uint8, device=dev)
    k = None if keep is None else keep.contiguous().view(torch.uint8)
    _lib.check(L.cgs_unique_voxels(_lib.ptr(pts), _lib.ptr(k), n, float(voxel_size), float(scale), _lib.ptr(inverse),
                                   _lib.ptr(first), _lib.ptr(status), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
               "cgs_unique_voxels")
    count, overflow = status.tolist()
    if overflow:
        raise _lib.CgsError("cgs_unique_voxels: |round(anchor / voxel / scale)| exceeds 2^20")
    return count, inverse[:n].long(), first[:count].long()


def find_divide_scale(pc, anchor, target_ratio, level_num):
    """scene/gaussian_model.py:1726-1749.  The reference bisects with 0-dim float32 TENSORS (scale_upper comes from the
    bounds, `(scale_upper + scale_lower) / 2` stays a tensor, `.item()` at the end), so the arithmetic of the search is
    float32: reproduced here with numpy float32 scalars on the host (the scales are saved in checkpoints and define
    the level division, i.e. the bitstream)."""
    import numpy as np
    f32 = np.float32
    upper0 = f32(float(((pc.x_bound_max - pc.x_bound_min) / pc.voxel_size).max()))
    cur, lower, scales = anchor, f32(1.0), []
    for _ in range(level_num - 1):
        hi, lo = upper0, lower
        while True:
            scale = f32(f32(hi + lo) / f32(2.0))
            count, _, first = unique_voxels(cur, pc.voxel_size, float(scale))
            ratio = count / cur.shape[0]
            if abs(ratio - target_ratio) < 0.01 or abs(f32(hi - lo)) < 1:
                break
            if ratio < target_ratio:
                hi = scale
            else:
                lo = scale
        # the reference continues with the unique voxel centres; their rows at any coarser scale are
        # what matters for the next search
        cur = torch.round(cur[first] / pc.voxel_size / float(scale)) * pc.voxel_size * float(scale)
        lower = scale
        scales.append(float(scale))
    return scales


def divide_levels(pc, anchor, mask_anchor_bool=None):
    """scene/gaussian_model.py:1751-1765 -> (hybrid_anchor_list, inverse_indices_list, mapping_list, last)."""
    level_anchor, inverse, first = [anchor], [], []
    cur = anchor
    for i in range(1, pc.level_num):
        keep = mask_anchor_bool if (i == 1 and mask_anchor_bool is not None) else None
        _, inv, fst = unique_voxels(cur, pc.voxel_size, pc.level_scale[i - 1], keep)
        if keep is not None:
            cur = cur * keep.unsqueeze(1)
        cur = cur[fst]
        level_anchor.append(cur)
        inverse.append(inv)
        first.append(fst)
    return level_anchor, inverse, first, cur


def build_level_plan(pc, anchor, mask_anchor_bool=None):
    """Flatten the 3-level coding order into int32 index arrays (see oracle/entropy_ref.level_plan
    for the derivation from scene/gaussian_model.py:1562-1652,1711-1793)."""
    N, dev = anchor.shape[0], anchor.device
    level_anchor, inverse, first, _ = divide_levels(pc, anchor, mask_anchor_bool)
    map1, map2 = first
    inv1, inv2 = inverse
    o1, o2 = map1, map1[map2]
    i32 = lambda t: t.to(torch.int32).contiguous()
    levels = [SimpleNamespace(level=2, orig=i32(o2), ctx_src=None, level_anchor=level_anchor[2].contiguous(),
                              n=int(o2.shape[0]))]
    to_code1 = torch.ones(map1.shape[0], dtype=torch.bool, device=dev)
    to_code1[map2] = False
    coded = torch.zeros(N, dtype=torch.bool, device=dev)
    coded[o2] = True
    member1 = torch.zeros(N, dtype=torch.bool, device=dev)
    member1[o1] = True
    gather1 = torch.nonzero(member1 & ~coded)[:, 0]          # ascending ORIGINAL index (quirk Q1)
    levels.append(SimpleNamespace(level=1, orig=i32(o1[to_code1]), ctx_src=i32(o2[inv2[inv1[gather1]]]),
                                  level_anchor=None, n=int(gather1.shape[0])))
    coded[o1] = True
    gather0 = torch.nonzero(~coded)[:, 0]
    levels.append(SimpleNamespace(level=0, orig=i32(gather0), ctx_src=i32(o1[inv1[gather0]]), level_anchor=None,
                                  n=int(gather0.shape[0])))
    return SimpleNamespace(N=N, levels=levels, inverse=inverse, first=first, level_anchor=level_anchor)


def _plan_content_hash(anchor, mask_anchor_bool):
    """Position-weighted integer checksums of the (quantised) anchor bit patterns and of the anchor mask."""
    q = anchor.detach().contiguous().view(torch.int32).reshape(-1).to(torch.int64)
    w = torch.arange(1, q.numel() + 1, device=q.device, dtype=torch.int64)
    parts = [q.sum(), (q * (w % 65521 + 1)).sum()]
    if mask_anchor_bool is not None:
        m = mask_anchor_bool.reshape(-1).to(torch.int64)
        parts += [m.sum(), (m * (w[:m.numel()] % 8191 + 7)).sum()]
    return tuple(torch.stack(parts).tolist())


def get_level_plan(pc, anchor, mask_anchor_bool):
    """Plan cache.  The plan depends on the quantised anchors (`_anchor`, the bounds, the voxel size) and on which
    anchors still have a live offset.  Fast path: the source tensors are the SAME objects at the SAME version
    (evaluation / scoring passes).  Otherwise -- every training iteration, because Adam bumps `_version` of `_anchor`
    (position lr is 0, arguments/__init__.py:86-87, so the values do not move) and of `_mask` (whose binarised
    any-offset-alive reduction flips rarely) -- a content checksum of the quantised anchors and of the anchor mask
    decides (two small reductions and one read-back instead of two voxel sorts and a dozen index kernels)."""
    src_a, src_m = getattr(pc, "_anchor", anchor), getattr(pc, "_mask", None)
    fast = (anchor.shape[0], anchor.data_ptr() if anchor is src_a else None, src_a.data_ptr(), src_a._version,
            None if src_m is None else (src_m.data_ptr(), src_m._version), mask_anchor_bool is None,
            getattr(pc, "_cgs_bound_version", 0),
            tuple((t.data_ptr(), t._version) for t in (pc.x_bound_min, pc.x_bound_max)))
    static = (anchor.shape[0], mask_anchor_bool is None, tuple(pc.level_scale), float(pc.voxel_size))
    ent = getattr(pc, "_cgs_level_plan", None)
    if ent is not None and ent[0] == fast and ent[2] == static:
        return ent[1]
    content = _plan_content_hash(anchor, mask_anchor_bool)
    if ent is not None and ent[2] == static and ent[3] == content:
        plan = ent[1]
    else:
        plan = build_level_plan(pc, anchor, mask_anchor_bool)
    try:
        pc._cgs_level_plan = (fast, plan, static, content)
    except Exception:
        pass
    return plan


def global_means(pc):
    """The three global means the reference passes as x_mean (gaussian_model.py:1667-1669), as host floats.  They
    only change when the parameters do, so they are cached per parameter version (three 100 MB reductions and a
    host synchronisation per call otherwise)."""
    srcs = (pc._anchor_feat, pc._scaling, pc._offset)
    key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in srcs) + (bool(pc.decoded_version),)
    ent = getattr(pc, "_cgs_means", None)
    if ent is not None and ent[0] == key:
        return ent[1]
    with torch.no_grad():
        means = tuple(torch.stack([pc._anchor_feat.mean(), pc.get_scaling.mean(), pc._offset.mean()]).tolist())
    try:
        pc._cgs_means = (key, means)
    except Exception:
        pass
    return means


# ----------------------------------------------------------------------------- forward core + training autograd

def _forward_levels(pc, plan, anchor, hyper, feat, scaling, offsets, masks, choose_u8, noise, training, return_details,
                    save=False):
    """EntropyBottleneck kernel + one fused kernel per level (coarse -> fine).  All inputs are detached,
    contiguous [N, .] tensors.  Returns a dict with the quantised attributes, the fp64 bit sums
    ([4i..4i+3] level i: feat, scaling, offsets, chosen rows; [12] hyper) and what the backward needs."""
    L = _lib.lib()
    dev, N = anchor.device, anchor.shape[0]
    sums = torch.zeros(16, dtype=torch.float64, device=dev)
    hyper_q, lik = pc.latent_codec(hyper, training=training, noise=None if noise is None else noise["eb"].to(dev),
                                   choose=choose_u8, bit_sum=sums[12:13])
    if getattr(pc, "disable_hyper", False):
        hyper_q = hyper_q * 0
    # every anchor is coded by exactly one level, so a full plan overwrites every row: no 516 MB zero-fill per pass
    # (a shard of the plan leaves the other ranks' rows at zero: the assembled result is a sum)
    full_plan = sum(lv.n for lv in plan.levels) == N
    alloc = torch.empty_like if full_plan else torch.zeros_like
    feat_q, scaling_q, offsets_q = alloc(feat), alloc(scaling), alloc(offsets)
    bits_out = torch.zeros((N, N_CODED), dtype=torch.float32, device=dev) if return_details else None
    means = global_means(pc)
    stream = _lib.stream_ptr()
    level_noise = []
    saved = [None] * len(plan.levels)
    umma = ctx_impl() == "umma"
    # the tensor-core time-out flag lives in the last slot of the bit sums (as an int32 view): ONE read-back serves both
    err = sums[15:16].view(torch.int32)[:1] if umma else None
    for li, lv in enumerate(plan.levels):
        nz = None
        if training and lv.n:
            nz = noise["levels"][li].to(dev).contiguous() if noise is not None else \
                torch.empty((lv.n, N_CODED), device=dev).uniform_(-0.5, 0.5)
        level_noise.append(nz)
        if lv.n == 0:
            continue
        args = (_lib.ptr(lv.orig), _lib.ptr(lv.ctx_src), _lib.ptr(lv.level_anchor), lv.n, _lib.ptr(anchor),
                _lib.ptr(hyper_q), _lib.ptr(feat), _lib.ptr(scaling), _lib.ptr(offsets), _lib.ptr(masks),
                _lib.ptr(choose_u8), _lib.ptr(nz), means[0], means[1], means[2], _lib.ptr(feat_q), _lib.ptr(scaling_q),
                _lib.ptr(offsets_q), _lib.ptr(bits_out), _lib.ptr(sums[4 * li:4 * li + 4]))
        if umma and save:
            # training with the tcgen05 backward: the forward leaves (mean, scale, Q), the hidden layer and its signs behind
            packed, in_dim = pack_grid_weights_umma(pc, lv.level)
            sv = dict(params=torch.empty((lv.n, 176), dtype=torch.float32, device=dev),
                      h=torch.empty((lv.n, 112), dtype=torch.float32, device=dev),
                      hmask=torch.empty((lv.n, 4), dtype=torch.int32, device=dev))
            saved[li] = sv
            _lib.check(L.cgs_context_level_umma_forward_train(in_dim, _lib.ptr(packed), *args, _lib.ptr(err), _lib.ptr(sv["params"]),
                                                              _lib.ptr(sv["h"]), _lib.ptr(sv["hmask"]), stream),
                       "cgs_context_level_umma_forward_train")
        elif umma:
            packed, in_dim = pack_grid_weights_umma(pc, lv.level)
            _lib.check(L.cgs_context_level_umma_forward(in_dim, _lib.ptr(packed), *args, _lib.ptr(err), stream),
                       "cgs_context_level_umma_forward")
        else:
            packed, in_dim = pack_grid_weights(pc, lv.level)
            _lib.check(L.cgs_context_level_forward(in_dim, _lib.ptr(packed), *args, stream), "cgs_context_level_forward")
    return dict(feat_q=feat_q, scaling_q=scaling_q, offsets_q=offsets_q, hyper_q=hyper_q, lik=lik, sums=sums,
                bits_out=bits_out, level_noise=level_noise, means=means, err=err, saved=saved)


def pack_grid_weights_bwd(m):
    """DIFFERENTIABLE packing of one context MLP into the layout of the backward kernel
    (csrc/context_model_bwd.cu: W1[in][101] | b1[100] | W2[100][177] | b2[176], odd leading dimensions).
    The kernel returns the gradient in the same layout; autograd un-packs it into the parameters."""
    F = torch.nn.functional
    W1 = F.pad(m[0].weight.t(), (0, 1))
    W2 = F.pad(m[2].weight.t(), (0, 2))
    flat = torch.cat([W1.reshape(-1), m[0].bias, W2.reshape(-1), F.pad(m[2].bias, (0, 1))])
    return F.pad(flat, (0, (-flat.numel()) % 4)).float()


class _ContextModelTrain(torch.autograd.Function):
    """Training-mode context model (x_q = x + U * Q, bit_per_param over the chosen anchors) with the
    fused backward kernels: replaces the autograd graph of scene/gaussian_model.py:1556-1707."""

    @staticmethod
    def forward(ctx, pc, plan, choose_u8, noise, rate, return_details, info, anchor, hyper, feat, offsets, scaling,
                masks, eb_packed, *w_bwd):
        d = lambda t: t.detach().contiguous()
        anchor, hyper, feat, offsets, scaling, masks = (d(t) for t in (anchor, hyper, feat, offsets, scaling, masks))
        bwd_impl = ctx_bwd_impl()
        if bwd_impl == "umma":
            # The rows of a level are independent, so the training pass is free to order them: rows chosen for the bit-rate
            # term first.  Everything the forward saves is then in that order and the backward treats the two CONTIGUOUS
            # ranges differently (the ~85 % not chosen have an output gradient with three non-zero columns).
            levels, lv_noise = [], []
            for li, lv in enumerate(plan.levels):
                if lv.n == 0:
                    levels.append(lv)
                    lv_noise.append(None)
                    continue
                perm = torch.argsort(choose_u8[lv.orig.long()] == 0, stable=True)
                levels.append(SimpleNamespace(
                    level=lv.level, n=lv.n, orig=lv.orig[perm].contiguous(),
                    ctx_src=None if lv.ctx_src is None else lv.ctx_src[perm].contiguous(),
                    level_anchor=None if lv.level_anchor is None else lv.level_anchor[perm].contiguous()))
                lv_noise.append(None if noise is None else noise["levels"][li].to(anchor.device)[perm].contiguous())
            plan = SimpleNamespace(N=plan.N, levels=levels)
            if noise is not None:
                noise = dict(noise, levels=lv_noise)
        out = _forward_levels(pc, plan, anchor, hyper, feat, scaling, offsets, masks, choose_u8, noise, True,
                              return_details, save=(bwd_impl == "umma"))
        s = (out["sums"] if rate is None else torch.cat([out["sums"], rate])).tolist()  # the one host read-back of the forward
        rate = 1.0 if rate is None else s.pop() / info["rate_den"]
        out["sums_host"], out["rate"] = s, rate
        info.update(out)
        n_chosen = sum(s[4 * i + 3] for i in range(3))
        total_bits = sum(s[4 * i + c] for i in range(3) for c in range(3)) + s[12]
        factor = rate / (max(n_chosen, 1e-30) * N_CODED)
        per_param = torch.tensor(total_bits * factor, dtype=torch.float32, device=anchor.device)
        ctx.pc, ctx.plan, ctx.factor, ctx.means = pc, plan, factor, out["means"]
        ctx.level_noise = out["level_noise"]
        # Row lists of the backward, without a host synchronisation: a stable sort moves the rows chosen for the
        # bit-rate term to the front (their count per level came back with the bit sums above).
        ctx.row_lists = []
        ctx.bwd_impl, ctx.saved_act = bwd_impl, out["saved"]
        for li, lv in enumerate(plan.levels):
            if bwd_impl == "umma":      # rows [0, n_full) are the chosen ones (ordered above): no row lists
                ctx.row_lists.append((None, int(round(s[4 * li + 3]))))
                continue
            if lv.n == 0:
                ctx.row_lists.append(None)
                continue
            not_chosen = choose_u8[lv.orig.long()] == 0
            perm = torch.argsort(not_chosen, stable=True).to(torch.int32)
            ctx.row_lists.append((perm, int(round(s[4 * li + 3]))))
        ctx.save_for_backward(anchor, out["hyper_q"], out["feat_q"], out["scaling_q"], out["offsets_q"], masks, choose_u8,
                              eb_packed.detach(), *[w.detach() for w in w_bwd])
        return out["feat_q"], out["scaling_q"], out["offsets_q"], per_param

    @staticmethod
    def backward(ctx, g_feat, g_scaling, g_offsets, g_bpp):
        L = _lib.lib()
        anchor, hyper_q, feat_q, scaling_q, offsets_q, masks, choose_u8, eb_packed, *w_bwd = ctx.saved_tensors
        plan, dev, N = ctx.plan, anchor.device, anchor.shape[0]
        G = lambda g, like: torch.zeros_like(like) if g is None else g.contiguous().clone().float()
        G_f, G_s, G_o = G(g_feat, feat_q), G(g_scaling, scaling_q), G(g_offsets, offsets_q)
        d_mask, d_hyper, d_anchor = torch.zeros_like(masks), torch.zeros_like(hyper_q), torch.zeros_like(anchor)
        d_w = [torch.zeros_like(w) for w in w_bwd]
        d_eb = torch.zeros_like(eb_packed)
        ticket = torch.zeros(4, dtype=torch.int32, device=dev)
        g_bpp_c = None if g_bpp is None else g_bpp.contiguous().float()   # named: must outlive the launches below
        g_ptr = _lib.ptr(g_bpp_c)
        stream = _lib.stream_ptr()
        err_flag = None
        for li in reversed(range(len(plan.levels))):       # fine -> coarse
            lv = plan.levels[li]
            if lv.n == 0:
                continue
            in_dim = 15 if lv.ctx_src is None else 71
            if ctx.bwd_impl == "umma":
                sv = ctx.saved_act[li]
                d_out = torch.empty((lv.n, 176), dtype=torch.float32, device=dev)
                d_pre = torch.empty((lv.n, 112), dtype=torch.float32, device=dev)
                if err_flag is None:
                    err_flag = torch.zeros(1, dtype=torch.int32, device=dev)
                _lib.check(L.cgs_context_level_backward_umma(
                    in_dim, _lib.ptr(pack_grid_weights_bwd_umma(ctx.pc, lv.level)), _lib.ptr(lv.orig), _lib.ptr(lv.ctx_src),
                    _lib.ptr(lv.level_anchor), lv.n, _lib.ptr(anchor), _lib.ptr(hyper_q), _lib.ptr(feat_q),
                    _lib.ptr(scaling_q), _lib.ptr(offsets_q), _lib.ptr(masks), _lib.ptr(choose_u8),
                    _lib.ptr(ctx.level_noise[li]), ctx.means[0], ctx.means[1], ctx.means[2], g_ptr, ctx.factor,
                    _lib.ptr(sv["params"]), _lib.ptr(sv["h"]), _lib.ptr(sv["hmask"]), _lib.ptr(G_f), _lib.ptr(G_s),
                    _lib.ptr(G_o), _lib.ptr(d_mask), _lib.ptr(d_hyper), _lib.ptr(d_anchor), _lib.ptr(d_w[li]),
                    _lib.ptr(d_out), _lib.ptr(d_pre), _lib.ptr(err_flag), 0 if g_ptr is None else ctx.row_lists[li][1], stream),
                    "cgs_context_level_backward_umma")
                ctx.saved_act[li] = None
                continue
            # rows chosen for the bit-rate term take the full kernel, the other ~85 % the 3-output one
            perm, n_full = ctx.row_lists[li]
            row_lists = ((perm[:n_full], 0), (perm[n_full:], 1))
            for rows, lite in row_lists:
                if rows.numel() == 0:
                    continue
                _lib.check(L.cgs_context_level_backward_rows(
                    in_dim, _lib.ptr(w_bwd[li]), _lib.ptr(lv.orig), _lib.ptr(lv.ctx_src), _lib.ptr(lv.level_anchor),
                    _lib.ptr(rows), rows.numel(), lite, _lib.ptr(anchor), _lib.ptr(hyper_q), _lib.ptr(feat_q),
                    _lib.ptr(scaling_q), _lib.ptr(offsets_q), _lib.ptr(masks), _lib.ptr(choose_u8),
                    _lib.ptr(ctx.level_noise[li]), ctx.means[0], ctx.means[1], ctx.means[2], g_ptr, ctx.factor,
                    _lib.ptr(G_f), _lib.ptr(G_s), _lib.ptr(G_o), _lib.ptr(d_mask), _lib.ptr(d_hyper), _lib.ptr(d_anchor),
                    _lib.ptr(d_w[li]), _lib.ptr(ticket), stream), "cgs_context_level_backward_rows")
        if g_ptr is not None:
            _lib.check(L.cgs_eb_backward(_lib.ptr(eb_packed), eb_packed.shape[0], _lib.ptr(hyper_q), N,
                                         _lib.ptr(choose_u8), g_ptr, ctx.factor, _lib.ptr(d_hyper), _lib.ptr(d_eb),
                                         stream), "cgs_eb_backward")
        if err_flag is not None:
            _lib.deferred_error_check(err_flag, "cgs_context_level_backward_umma: a tensor-core completion barrier timed out")
        return (None,) * 7 + (d_anchor, d_hyper, G_f, G_o, G_s, d_mask, d_eb, *d_w)


# ----------------------------------------------------------------------------- the fused forward

def multi_scale_generating(pc, anchor, hyper, feat, grid_offsets, grid_scaling, binary_grid_masks,
                           mask_anchor_bool=None, training=False, predict_bpp=False, return_sum_bits=False,
                           noise=None, return_details=False, plan=None, group=None):
    """Same signature and the same three return arities as scene/gaussian_model.py:1541-1707.
    `noise` (optional, for reproducible parity runs): dict(eb=[N,12], levels=[[n_i,86] x3 coarse->fine],
    choose=bool[N]); otherwise drawn with the CUDA generator.
    `plan` (optional): a shard of the level plan (contextgs_b200.distributed.shard_level_plan) -- only
    its rows are coded and scored, and the bit sums are all-reduced over `group`, so that every rank
    returns the whole-scene totals (anchors sharded, scalar all-reduce only)."""
    L = _lib.lib()
    dev = anchor.device
    N, K = anchor.shape[0], pc.n_offsets
    anchor = anchor.contiguous()
    feat = feat.contiguous()
    scaling = grid_scaling.contiguous()
    offsets = grid_offsets.reshape(N, 3 * K).contiguous()
    masks = binary_grid_masks.reshape(N, K).contiguous()
    hyper = hyper.contiguous()
    sharded = plan is not None
    with torch.no_grad():
        if pc.level_scale is None:
            # (the reference indexes with mask_anchor_bool even when it is None, scene/gaussian_model.py:1559;
            #  it only ever reaches this line from training, where the mask is given)
            pc.level_scale = find_divide_scale(pc, anchor if mask_anchor_bool is None else anchor[mask_anchor_bool],
                                               pc.target_ratio, pc.level_num)
        if plan is None:
            plan = get_level_plan(pc, anchor, mask_anchor_bool)

    if predict_bpp:
        if noise is not None and "choose" in noise:
            choose = noise["choose"].to(dev)
        else:
            thresh = 1 if return_sum_bits else 0.15
            choose = torch.rand(N, device=dev) <= thresh
        if mask_anchor_bool is not None:
            choose = choose & mask_anchor_bool
    else:
        choose = torch.zeros(N, dtype=torch.bool, device=dev)
    if sharded:  # hyper bits are summed over the anchors this shard codes
        owned = torch.zeros(N, dtype=torch.bool, device=dev)
        for lv in plan.levels:
            owned[lv.orig.long()] = True
        choose = choose & owned
    choose_u8 = choose.contiguous().view(torch.uint8)

    # mask_anchor_rate of the reference (scene/gaussian_model.py:1660): the count stays on the device and reaches the host with
    # the bit sums (one read-back instead of two)
    rate = mask_anchor_bool.sum(dtype=torch.float64).view(1) if mask_anchor_bool is not None else None
    rate_den = mask_anchor_bool.numel() if mask_anchor_bool is not None else 1
    differentiable = (training and predict_bpp and not return_sum_bits and not sharded and torch.is_grad_enabled()
                      and any(t.requires_grad for t in (hyper, feat, grid_offsets, grid_scaling, binary_grid_masks,
                                                        *pc.get_grid_mlp.parameters())))
    info = {"rate_den": rate_den}
    if differentiable:
        wb = [pack_grid_weights_bwd(pc.get_grid_mlp[lv.level]) for lv in plan.levels]
        feat_q, scaling_q, offsets_q, per_param_t = _ContextModelTrain.apply(
            pc, plan, choose_u8, noise, rate, return_details, info, anchor, hyper, feat, offsets, scaling, masks,
            pc.latent_codec.packed_diff(), *wb)
    else:
        info = _forward_levels(pc, plan, anchor, hyper, feat, scaling, offsets, masks, choose_u8, noise, training,
                               return_details)
        feat_q, scaling_q, offsets_q, per_param_t = info["feat_q"], info["scaling_q"], info["offsets_q"], None
    sums, hyper_q, lik, bits_out = info["sums"], info["hyper_q"], info["lik"], info["bits_out"]
    offsets_q3 = offsets_q.view(N, K, 3)
    if not predict_bpp:
        return feat_q, scaling_q, offsets_q3

    if sharded:
        from .distributed import all_reduce_sums
        all_reduce_sums(sums, group)
    # one host read-back (the reference: several .item()); the mask count of the size report rides on it
    pos_num = None
    if "sums_host" in info:
        s, rate = info["sums_host"], info["rate"]
    else:
        extra = ([] if rate is None else [rate]) + ([binary_grid_masks.sum(dtype=torch.float64).view(1)] if return_sum_bits else [])
        vals = torch.cat([sums] + extra).tolist() if extra else sums.tolist()
        s = vals[:sums.numel()]
        tail = vals[sums.numel():]
        rate = 1.0 if rate is None else tail.pop(0) / rate_den
        if return_sum_bits:
            pos_num = tail.pop(0)
    if info.get("err") is not None and s[15] != 0.0:   # the int32 flag aliases the low word of slot 15
        raise _lib.CgsError("cgs_context_level_umma_forward: a tensor-core completion barrier timed out")
    bit_feat, bit_scaling, bit_offsets = (sum(s[4 * i + c] for i in range(3)) for c in range(3))
    n_chosen = sum(s[4 * i + 3] for i in range(3))
    bit_hyper = s[12]
    details = dict(feat_q=feat_q, scaling_q=scaling_q, offsets_q=offsets_q3, bits=bits_out, hyper_q=hyper_q,
                   lik_hyper=lik, choose=choose, sums=s, plan=plan)
    if return_sum_bits:
        bit_anchor = int(n_chosen) * 3 * 16
        if pos_num is None:
            bit_masks = get_binary_vxl_size(binary_grid_masks)[1].item()
        else:   # utils/encodings.py:15-32 on the host, in the fp32 arithmetic of the reference's tensors (the clamp bound
            # 1 - 1e-6 is not representable: its fp32 value decides the result when every offset is kept)
            f32 = np.float32
            ttl = binary_grid_masks.numel()
            pos, neg = f32(pos_num), f32(ttl) - f32(pos_num)
            pg = np.clip(pos / f32(ttl), f32(1e-6), f32(1 - 1e-6))
            bit_masks = float(pos * -np.log2(pg) + neg * -np.log2(f32(1) - pg) + f32(32))
        res = (bit_anchor, bit_hyper, bit_feat, bit_scaling, bit_offsets, bit_masks)
        return (res, details) if return_details else res
    nc = max(n_chosen, 1e-30)
    t = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
    per_hyper = bit_hyper / (nc * 12) * rate
    per_feat = t(bit_feat / (nc * 50) * rate)
    per_scaling = t(bit_scaling / (nc * 6) * rate)
    per_offsets = t(bit_offsets / (nc * 30) * rate)
    per_param = per_param_t if per_param_t is not None else \
        t((bit_feat + bit_scaling + bit_offsets + bit_hyper) / (nc * N_CODED) * rate)
    level_bpp = [1 - rate, per_hyper]
    for li, lv in enumerate(plan.levels):
        cnt = s[4 * li + 3]
        bpp = (s[4 * li] + s[4 * li + 1] + s[4 * li + 2]) / cnt / N_CODED if cnt > 0 else float("nan")
        level_bpp.append([lv.n / N, bpp])
    res = (feat_q, scaling_q, offsets_q3, per_param, per_feat, per_scaling, per_offsets, level_bpp)
    return (res, details) if return_details else res
