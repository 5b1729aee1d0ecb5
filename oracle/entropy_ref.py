"""CPU restatement (PyTorch-on-CPU, fp32) of ContextGS's anchor-level context / entropy model and of
`generate_neural_gaussians` -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(),
bench.py cpu_baseline / --impl reference).  Never imported by the product path.

Pinned: tests/test_oracle_entropy.py checks every function here against golden vectors produced by
running the REFERENCE'S OWN Python (utils/entropy_models.py, utils/encodings.py,
utils/multi_level.py, scene/gaussian_model.py:1541-1793, gaussian_renderer/__init__.py:25-150) in
the build container; generator: tests/golden/make_golden.py.

PARITY UNPINNED for one sub-part: `EntropyBottleneckRef` restates CompressAI's
`compressai.entropy_models.EntropyBottleneck` (third-party, version unpinned in the reference's
environment.yml:21 / setup_env.sh:25, not installed here, source not in /root/reference) from its
published algorithm: filters (3,3,3,3), init_scale 10, tail_mass 1e-9, likelihood bound 1e-9.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

Q_FEAT0, Q_SCALING0, Q_OFFSETS0 = 1.0, 0.001, 0.2  # scene/gaussian_model.py:1564-1566
ANCHOR_ROUND_DIGITS = 16                            # utils/encodings.py:10
Q_ANCHOR = 1 / (2 ** ANCHOR_ROUND_DIGITS - 1)       # utils/encodings.py:11
CLAMP_STEPS = 15_000                                # utils/encodings.py:207-208, entropy_models.py:39-40


# ----------------------------------------------------------------------------- small pieces

def ste_multistep(x, Q):
    """utils/encodings.py:203-213 (forward)."""
    x = torch.clamp(x, min=-CLAMP_STEPS * Q, max=CLAMP_STEPS * Q)
    return torch.round(x / Q) * Q


def quantize_anchor(anchors, min_v, max_v):
    """utils/encodings.py:219-227 (forward)."""
    interval = (max_v - min_v) * Q_ANCHOR + 1e-6
    q = torch.div(anchors - min_v, interval, rounding_mode="floor")
    q = torch.clamp(q, 0, 2 ** ANCHOR_ROUND_DIGITS - 1)
    return q * interval + min_v, q


def binary_mask_bits(binary_vxl):
    """utils/encodings.py:15-32 -> (Pg, total bits)."""
    ttl = binary_vxl.numel()
    pos = torch.sum(binary_vxl)
    neg = ttl - pos
    Pg = torch.clamp(pos / ttl, min=1e-6, max=1 - 1e-6)
    bits = pos * (-torch.log2(Pg)) + neg * (-torch.log2(1 - Pg)) + 32
    return Pg, bits


def gaussian_bits(x, mean, scale, Q, x_mean):
    """utils/entropy_models.py:34-50 (`Entropy_gaussian.forward`, use_clamp=True)."""
    x = torch.clamp(x, min=x_mean - CLAMP_STEPS * Q, max=x_mean + CLAMP_STEPS * Q)
    scale = torch.clamp(scale, min=1e-9)

    def cdf(v):  # torch.distributions.Normal.cdf
        return 0.5 * (1 + torch.erf((v - mean) * scale.reciprocal() / math.sqrt(2)))

    likelihood = torch.abs(cdf(x + 0.5 * Q) - cdf(x - 0.5 * Q))
    likelihood = torch.clamp(likelihood, min=1e-6)  # Low_bound, entropy_models.py:141-147
    return -torch.log2(likelihood)


def low_bound_backward(x, g):
    """utils/entropy_models.py:149-156: effectively g * (x >= 1e-6) (quirk Q2 in SURVEY.md)."""
    grad1 = g.clone()
    grad1[x < 1e-6] = 0
    t = ((x >= 1e-6) | (g < 0.0)).to(g.dtype)
    return grad1 * t


class EntropyBottleneckRef:
    """CompressAI EntropyBottleneck.forward restated (PARITY UNPINNED, see module docstring).
    Parameters per channel: matrices [C,f_{i+1},f_i], biases [C,f_{i+1},1], factors [C,f_{i+1},1],
    quantiles [C,1,3]."""

    def __init__(self, channels, seed=0, filters=(3, 3, 3, 3), init_scale=10.0):
        g = torch.Generator().manual_seed(seed)
        self.channels, self.filters = channels, tuple(filters)
        f = (1,) + self.filters + (1,)
        scale = init_scale ** (1 / (len(self.filters) + 1))
        self.matrices, self.biases, self.factors = [], [], []
        for i in range(len(self.filters) + 1):
            init = float(np.log(np.expm1(1 / scale / f[i + 1])))
            self.matrices.append(torch.full((channels, f[i + 1], f[i]), init))
            self.biases.append(torch.rand(channels, f[i + 1], 1, generator=g) - 0.5)
            if i < len(self.filters):
                self.factors.append(torch.zeros(channels, f[i + 1], 1))
        self.quantiles = torch.tensor([-init_scale, 0.0, init_scale]).repeat(channels, 1, 1)
        self.likelihood_bound = 1e-9

    def randomize(self, seed=1, amount=0.3):
        """Perturb the parameters so that tests exercise non-initial weights."""
        g = torch.Generator().manual_seed(seed)
        for lst in (self.matrices, self.biases, self.factors):
            for t in lst:
                t.add_(torch.randn(t.shape, generator=g) * amount)
        self.quantiles[:, :, 1] += torch.randn(self.channels, 1, generator=g) * amount
        return self

    def logits_cumulative(self, v):
        for i in range(len(self.filters) + 1):
            v = torch.matmul(F.softplus(self.matrices[i]), v) + self.biases[i]
            if i < len(self.filters):
                v = v + torch.tanh(self.factors[i]) * torch.tanh(v)
        return v

    def forward(self, x, training, noise=None):
        """x [N,C] -> (x_hat [N,C], likelihood [N,C]).  Training adds U(-.5,.5) drawn on the
        permuted [C,1,N] tensor (that is the order CompressAI consumes the RNG in)."""
        v = x.permute(1, 0).contiguous().reshape(self.channels, 1, -1)
        if training:
            if noise is None:
                noise = torch.empty_like(v).uniform_(-0.5, 0.5)
            out = v + noise.reshape(v.shape)
        else:
            med = self.quantiles[:, :, 1:2]
            out = torch.round(v - med) + med
        lower = self.logits_cumulative(out - 0.5)
        upper = self.logits_cumulative(out + 0.5)
        sign = -torch.sign(lower + upper)
        lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
        lik = torch.clamp(lik, min=self.likelihood_bound)
        N = x.shape[0]
        return out.reshape(self.channels, N).permute(1, 0), lik.reshape(self.channels, N).permute(1, 0)

    __call__ = forward

    def packed(self):
        """Flat per-channel parameter block consumed by the CUDA kernel: for each channel
        [softplus(M0)(3) b0(3) tanh(f0)(3) | 3x{softplus(M)(9) b(3) tanh(f)(3)} | softplus(M4)(3) b4(1) | median]."""
        C = self.channels
        rows = []
        for c in range(C):
            r = []
            for i in range(len(self.filters) + 1):
                r.append(F.softplus(self.matrices[i][c]).reshape(-1))
                r.append(self.biases[i][c].reshape(-1))
                if i < len(self.filters):
                    r.append(torch.tanh(self.factors[i][c]).reshape(-1))
            r.append(self.quantiles[c, 0, 1:2])
            rows.append(torch.cat(r))
        return torch.stack(rows).contiguous()


# ----------------------------------------------------------------------------- MLPs

def mlp2(x, W1, b1, W2, b2):
    """nn.Sequential(Linear, ReLU, Linear) of scene/gaussian_model.py:153-188."""
    return F.linear(F.relu(F.linear(x, W1, b1)), W2, b2)


def make_mlps(feat_dim=50, n_offsets=10, hyper_divisor=4, level_num=3, seed=6):
    """Default nn.Linear initialisation of the 3 decoder MLPs and the 3 context MLPs
    (shapes: scene/gaussian_model.py:153-188)."""
    g = torch.Generator().manual_seed(seed)

    def linear(i, o):
        bound = 1 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * bound, (torch.rand(o, generator=g) * 2 - 1) * bound

    def two(i, h, o):
        W1, b1 = linear(i, h)
        W2, b2 = linear(h, o)
        return [W1, b1, W2, b2]

    d_in = feat_dim + 3 + 1
    mlps = dict(opacity=two(d_in, feat_dim, n_offsets), cov=two(d_in, feat_dim, 7 * n_offsets),
                color=two(d_in, feat_dim, 3 * n_offsets), grid=[])
    H = feat_dim // hyper_divisor
    out = (feat_dim + 6 + 3 * n_offsets) * 2 + 3
    for i in range(level_num):
        d = H + 3 if i == level_num - 1 else feat_dim + 6 + 3 + H
        mlps["grid"].append(two(d, feat_dim * 2, out))
    return mlps


# ----------------------------------------------------------------------------- level division

def unique_rows_first_index(rows):
    """utils/multi_level.py:3-31: lexicographically sorted unique rows, inverse map and the
    SMALLEST source index of every unique row."""
    uniq, inverse = torch.unique(rows, return_inverse=True, dim=0)
    first = torch.full((uniq.shape[0],), rows.shape[0], dtype=torch.long)
    first.scatter_reduce_(0, inverse, torch.arange(rows.shape[0]), reduce="amin", include_self=True)
    return uniq, inverse, first


def find_divide_scale(anchor, voxel_size, x_bound_min, x_bound_max, target_ratio, level_num):
    """scene/gaussian_model.py:1726-1749: per-level bisection on the voxel scale."""
    upper0 = ((x_bound_max - x_bound_min) / voxel_size).max()
    cur, lower, scales = anchor, 1, []
    for _ in range(level_num - 1):
        hi, lo = upper0, lower
        while True:
            scale = (hi + lo) / 2
            uniq = torch.unique(torch.round(cur / voxel_size / scale), dim=0) * voxel_size * scale
            ratio = uniq.shape[0] / cur.shape[0]
            if abs(ratio - target_ratio) < 0.01 or abs(hi - lo) < 1:
                break
            if ratio < target_ratio:
                hi = scale
            else:
                lo = scale
        cur, lower = uniq, scale
        scales.append(float(scale))
    return scales


def divide_levels(anchor, voxel_size, level_scale, mask_anchor_bool=None):
    """scene/gaussian_model.py:1751-1765 -> per level: anchors, inverse (level i-1 -> i), first-index map."""
    level_anchor, inverse, first = [anchor], [], []
    cur = anchor
    for i in range(1, len(level_scale) + 1):
        if i == 1 and mask_anchor_bool is not None:
            cur = cur * mask_anchor_bool.unsqueeze(1)
        _, inv, fst = unique_rows_first_index(torch.round(cur / voxel_size / level_scale[i - 1]))
        cur = cur[fst]
        level_anchor.append(cur)
        inverse.append(inv)
        first.append(fst)
    return level_anchor, inverse, first


def level_plan(N, inverse, first):
    """Index bookkeeping of the 3-level coding order (scene/gaussian_model.py:1562-1652,1711-1793)
    flattened into explicit arrays.  Returns, per level i (coarse -> fine): `orig` = original index of
    every row the level codes (in the row order the reference uses), `ctx_src` = original index of the
    representative whose quantised attributes are that ROW's context (None for the coarsest level).
    Quirk Q1 is reproduced: for the middle level the context rows are ordered by ascending original
    index while the coded rows are ordered by level index (sorted voxel key)."""
    assert len(first) == 2, "restated for level_num == 3 (the reference hard-codes 3 at :1674-1678)"
    map1, map2 = first
    inv1, inv2 = inverse
    o1 = map1                       # original index of level-1 members
    o2 = map1[map2]                 # original index of level-2 members
    plan = []
    plan.append(SimpleNamespace(level=2, orig=o2, ctx_src=None, level_rows=torch.arange(o2.shape[0])))
    # level 1
    to_code1 = torch.ones(map1.shape[0], dtype=torch.bool)
    to_code1[map2] = False
    coded = torch.zeros(N, dtype=torch.bool)
    coded[o2] = True
    member1 = torch.zeros(N, dtype=torch.bool)
    member1[o1] = True
    gather1 = torch.nonzero(member1 & ~coded)[:, 0]           # ascending ORIGINAL index
    ctx1 = o2[inv2[inv1[gather1]]]
    plan.append(SimpleNamespace(level=1, orig=o1[to_code1], ctx_src=ctx1,
                                level_rows=torch.nonzero(to_code1)[:, 0]))
    # level 0
    coded[o1] = True
    gather0 = torch.nonzero(~coded)[:, 0]
    ctx0 = o1[inv1[gather0]]
    to_code0 = torch.ones(N, dtype=torch.bool)
    to_code0[map1] = False
    plan.append(SimpleNamespace(level=0, orig=torch.nonzero(to_code0)[:, 0], ctx_src=ctx0, level_rows=None))
    return plan


# ----------------------------------------------------------------------------- the context model

def multi_scale_generating(pc, anchor, hyper, feat, grid_offsets, grid_scaling, binary_grid_masks,
                           mask_anchor_bool=None, training=False, predict_bpp=False, return_sum_bits=False,
                           return_details=False, choose_override=None):
    """scene/gaussian_model.py:1541-1707.  `pc` needs: latent_codec, level_scale (or None), target_ratio,
    level_num, voxel_size, x_bound_min/max, mlps['grid'], n_offsets, feat_dim, feat_mean, scaling_mean,
    offset_mean (the three global means the reference takes at :1667-1669).
    RNG is consumed in the reference's call order (EB noise, per level feat/scaling/offsets noise,
    choose_mask), so torch.manual_seed(s) before this call and before the reference gives equal noise."""
    N, K, Fd = anchor.shape[0], pc.n_offsets, pc.feat_dim
    feat_q = torch.zeros_like(feat)
    scaling_q = torch.zeros_like(grid_scaling)
    offsets_q = torch.zeros_like(grid_offsets)
    mean_f, std_f, Qf_all = torch.zeros(N, Fd), torch.zeros(N, Fd), torch.zeros(N, 1)
    mean_s, std_s, Qs_all = torch.zeros(N, 6), torch.zeros(N, 6), torch.zeros(N, 1)
    mean_o, std_o, Qo_all = torch.zeros(N, 3 * K), torch.zeros(N, 3 * K), torch.zeros(N, 1)

    hyper_q, lik_hyper = pc.latent_codec(hyper, training=training)
    if pc.level_scale is None:
        pc.level_scale = find_divide_scale(anchor[mask_anchor_bool], pc.voxel_size, pc.x_bound_min, pc.x_bound_max,
                                           pc.target_ratio, pc.level_num)
    level_anchor, inverse, first = divide_levels(anchor, pc.voxel_size, pc.level_scale, mask_anchor_bool)
    plan = level_plan(N, inverse, first)

    for lv in plan:
        o = lv.orig
        if o.numel() == 0:
            continue
        if lv.ctx_src is None:
            head = level_anchor[lv.level][lv.level_rows]
        else:
            s = lv.ctx_src
            head = torch.cat([anchor[s], feat_q[s], scaling_q[s]], dim=1)
        out = mlp2(torch.cat([head, hyper_q[o]], dim=1), *pc.mlps["grid"][lv.level])
        m_f, s_f, m_s, s_s, m_o, s_o, a_f, a_s, a_o = torch.split(out, [Fd, Fd, 6, 6, 3 * K, 3 * K, 1, 1, 1], dim=-1)
        Qf = (Q_FEAT0 * (1 + torch.tanh(a_f))).clamp(1e-9)
        Qs = (Q_SCALING0 * (1 + torch.tanh(a_s))).clamp(1e-9)
        Qo = (Q_OFFSETS0 * (1 + torch.tanh(a_o))).clamp(1e-9)
        f, sc, of = feat[o], grid_scaling[o], grid_offsets[o]
        if training:
            f = f + torch.empty_like(f).uniform_(-0.5, 0.5) * Qf
            sc = sc + torch.empty_like(sc).uniform_(-0.5, 0.5) * Qs
            of = of + torch.empty_like(of).uniform_(-0.5, 0.5) * Qo.unsqueeze(1)
        else:
            f, sc, of = ste_multistep(f, Qf), ste_multistep(sc, Qs), ste_multistep(of, Qo.unsqueeze(1))
        feat_q[o], scaling_q[o], offsets_q[o] = f, sc, of
        mean_f[o], std_f[o], Qf_all[o] = m_f, s_f, Qf
        mean_s[o], std_s[o], Qs_all[o] = m_s, s_s, Qs
        mean_o[o], std_o[o], Qo_all[o] = m_o, s_o, Qo

    if not predict_bpp:
        return feat_q, scaling_q, offsets_q

    thresh = 1 if return_sum_bits else 0.15
    choose = torch.rand_like(anchor[:, 0]) <= thresh
    if choose_override is not None:   # test hook: a caller-chosen subset instead of the random draw (:1658-1659)
        choose = choose_override.clone()
    if mask_anchor_bool is not None:
        choose = choose & mask_anchor_bool
        rate = mask_anchor_bool.sum() / mask_anchor_bool.numel()
    else:
        rate = 1
    bit_hyper = -torch.log2(lik_hyper[choose])
    bit_feat = gaussian_bits(feat_q[choose], mean_f[choose], std_f[choose], Qf_all[choose], pc.feat_mean)
    bit_scaling = gaussian_bits(scaling_q[choose], mean_s[choose], std_s[choose], Qs_all[choose], pc.scaling_mean)
    bit_offsets = gaussian_bits(offsets_q[choose].view(-1, 3 * K), mean_o[choose], std_o[choose], Qo_all[choose],
                                pc.offset_mean)
    bit_offsets = bit_offsets * binary_grid_masks[choose].repeat(1, 1, 3).view(-1, 3 * K)
    details = dict(feat_q=feat_q, scaling_q=scaling_q, offsets_q=offsets_q, choose=choose, bit_hyper=bit_hyper,
                   bit_feat=bit_feat, bit_scaling=bit_scaling, bit_offsets=bit_offsets, hyper_q=hyper_q,
                   lik_hyper=lik_hyper, Qf=Qf_all, Qs=Qs_all, Qo=Qo_all, mean_f=mean_f, std_f=std_f,
                   mean_s=mean_s, std_s=std_s, mean_o=mean_o, std_o=std_o,
                   plan=plan, inverse=inverse, first=first)
    if return_sum_bits:
        bit_anchor = bit_hyper.shape[0] * 3 * 16
        bit_masks = binary_mask_bits(binary_grid_masks)[1].item()
        res = (bit_anchor, bit_hyper.sum().item(), bit_feat.sum().item(), bit_scaling.sum().item(),
               bit_offsets.sum().item(), bit_masks)
        return (res, details) if return_details else res
    per_feat = bit_feat.sum() / bit_feat.numel() * rate
    per_scaling = bit_scaling.sum() / bit_scaling.numel() * rate
    per_offsets = bit_offsets.sum() / bit_offsets.numel() * rate
    per_hyper = bit_hyper.sum() / bit_hyper.numel() * rate
    per_param = (bit_feat.sum() + bit_scaling.sum() + bit_offsets.sum() + bit_hyper.sum()) / \
                (bit_feat.numel() + bit_scaling.numel() + bit_offsets.numel()) * rate
    bpp_map = bit_offsets.sum(dim=1) + bit_scaling.sum(dim=1) + bit_feat.sum(dim=1)
    dim = Fd + 6 + 3 * K
    level_bpp = [1 - mask_anchor_bool.float().mean().item(), per_hyper.item()]
    for lv in plan:
        lm = torch.zeros(N, dtype=torch.bool)
        lm[lv.orig] = True
        level_bpp.append([lv.orig.shape[0] / N, bpp_map[lm[choose]].mean().item() / dim])
    res = (feat_q, scaling_q, offsets_q, per_param, per_feat, per_scaling, per_offsets, level_bpp)
    return (res, details) if return_details else res


# ----------------------------------------------------------------------------- anchor -> Gaussians

def generate_neural_gaussians(pc, camera_center, anchor, feat, grid_offsets, grid_scaling, binary_grid_masks):
    """gaussian_renderer/__init__.py:106-150 (the part after the anchor attributes have been chosen).
    Inputs are already restricted to the visible anchors."""
    K = pc.n_offsets
    ob_view = anchor - camera_center
    ob_dist = ob_view.norm(dim=1, keepdim=True)
    ob_view = ob_view / ob_dist
    x = torch.cat([feat, ob_view, ob_dist], dim=1)
    pre_opacity = mlp2(x, *pc.mlps["opacity"])
    neural_opacity = torch.tanh(pre_opacity).reshape(-1, 1) * binary_grid_masks.view(-1, 1)
    mask = (neural_opacity > 0.0).view(-1)
    opacity = neural_opacity[mask]
    color = torch.sigmoid(mlp2(x, *pc.mlps["color"])).reshape(-1, 3)
    scale_rot = mlp2(x, *pc.mlps["cov"]).reshape(-1, 7)
    offsets = grid_offsets.view(-1, 3)
    rep = torch.cat([grid_scaling, anchor], dim=-1).repeat_interleave(K, dim=0)
    allv = torch.cat([rep, color, scale_rot, offsets], dim=-1)[mask]
    scaling_rep, anchor_rep, color, scale_rot, offsets = allv.split([6, 3, 3, 7, 3], dim=-1)
    scaling = scaling_rep[:, 3:] * torch.sigmoid(scale_rot[:, :3])
    rot = F.normalize(scale_rot[:, 3:7])
    xyz = anchor_rep + offsets * scaling_rep[:, :3]
    return dict(xyz=xyz, color=color, opacity=opacity, scaling=scaling, rot=rot, neural_opacity=neural_opacity,
                mask=mask, pre_opacity=pre_opacity)


# ----------------------------------------------------------------------------- model container

def make_model(scene, mlp_seed=6, eb_seed=9, randomize_eb=True, target_ratio=0.2, level_num=3):
    """Duck-typed stand-in for the attributes of GaussianModel that the hot path reads."""
    feat_dim = scene["feat"].shape[1]
    K = scene["offset"].shape[1]
    H = scene["hyper"].shape[1]
    pc = SimpleNamespace()
    pc.feat_dim, pc.n_offsets, pc.voxel_size = feat_dim, K, scene["voxel_size"]
    pc.level_num, pc.target_ratio, pc.level_scale = level_num, target_ratio, None
    pc.disable_hyper, pc.adaptQ_per_channel, pc.decoded_version = False, False, False
    pc._anchor, pc._anchor_feat, pc._hyper_latent = scene["anchor"], scene["feat"], scene["hyper"]
    pc._offset, pc._scaling, pc._mask = scene["offset"], scene["scaling"], scene["mask"]
    a = pc._anchor
    mn, mx = a.min(dim=0, keepdim=True)[0], a.max(dim=0, keepdim=True)[0]  # update_anchor_bound, :352-360
    pc.x_bound_min = torch.where(mn < 0, mn * 1.2, mn * 0.8)
    pc.x_bound_max = torch.where(mx > 0, mx * 1.2, mx * 0.8)
    pc.mlps = make_mlps(feat_dim, K, feat_dim // H, level_num, seed=mlp_seed)
    pc.latent_codec = EntropyBottleneckRef(H, seed=eb_seed)
    if randomize_eb:
        pc.latent_codec.randomize()
    pc.get_anchor = quantize_anchor(pc._anchor, pc.x_bound_min, pc.x_bound_max)[0]
    pc.get_scaling = torch.exp(pc._scaling)
    sig = torch.sigmoid(pc._mask)
    pc.get_mask = (sig > 0.01).float()
    pc.get_mask_anchor = pc.get_mask.sum(dim=1)[:, 0] > 0
    pc.feat_mean, pc.scaling_mean, pc.offset_mean = pc._anchor_feat.mean(), pc.get_scaling.mean(), pc._offset.mean()
    return pc
