"""Anchor -> neural Gaussian generation: drop-in for `generate_neural_gaussians`
(gaussian_renderer/__init__.py:25-150) on top of the fused CUDA kernel
`cgs_neural_gaussians_forward` (contextgs_b200/csrc/neural_gaussians.cu)."""
import ctypes

import torch

from . import _lib

Q_FEAT, Q_SCALING, Q_OFFSETS = 1, 0.001, 0.2  # gaussian_renderer/__init__.py:40-42

def pack_decoder_weights(pc):
    """Packed weight block of the three decoder MLPs (layout: include/contextgs_b200.h), cached per
    parameter version so that inference frames do not repack."""
    mods = (pc.get_opacity_mlp, pc.get_color_mlp, pc.get_cov_mlp)
    params = [p for m in mods for p in (m[0].weight, m[0].bias, m[2].weight, m[2].bias)]
    key = tuple((p.data_ptr(), p._version) for p in params)
    cache = _lib.object_cache(pc)
    ent = cache.get("decoder_pack")
    if ent is not None and ent[0] == key:
        return ent[1]
    with torch.no_grad():
        dev = params[0].device
        W1 = torch.zeros(54, 152, device=dev)
        b1 = torch.zeros(152, device=dev)
        for i, m in enumerate(mods):
            W1[:, 50 * i:50 * i + 50] = m[0].weight.t()
            b1[50 * i:50 * i + 50] = m[0].bias
        parts = [W1.reshape(-1), b1]
        for m, ld in zip(mods, (12, 32, 72)):
            n = m[2].weight.shape[0]
            W2 = torch.zeros(50, ld, device=dev)
            W2[:, :n] = m[2].weight.t()
            b2 = torch.zeros(ld, device=dev)
            b2[:n] = m[2].bias
            parts += [W2.reshape(-1), b2]
        packed = torch.cat(parts).float().contiguous()
    assert packed.numel() == _lib.lib().cgs_neural_gaussians_packed_floats()
    cache["decoder_pack"] = (key, packed)
    return packed


def tf32_split(w):
    """w = hi + lo with both parts exactly representable in TF32 (10 explicit mantissa bits):
    round-to-nearest on the magnitude, as cvt.rna.tf32.f32 does (contextgs_b200/csrc/umma.cuh)."""
    def rna(x):
        return ((x.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    hi = rna(w)
    return hi, rna(w - hi)


def umma_b_operand(W, n_pad, k_pad):
    """nn.Linear weight W[n, k] -> tcgen05 B operand [k_pad/4][n_pad][4] (K-major core matrices,
    see umma.cuh), zero padded, split into TF32 hi / lo parts.  Returns (hi.flat, lo.flat)."""
    n, k = W.shape
    full = torch.zeros(n_pad, k_pad, device=W.device)
    full[:n, :k] = W
    hi, lo = tf32_split(full)
    lay = lambda t: t.view(n_pad, k_pad // 4, 4).permute(1, 0, 2).contiguous().reshape(-1)
    return lay(hi), lay(lo)


def _mlp_params(mods):
    """(weight, bias) of layers 0 and 2 of each `nn.Sequential(Linear, ReLU, Linear[, act])`, read from the module
    dictionaries: this runs once per frame as a cache key, and `m[0].weight` costs a Sequential.__getitem__ plus a
    Module.__getattr__ per access (~30 us per frame for the twelve tensors)."""
    out = []
    for m in mods:
        sub = m._modules
        for name in ("0", "2"):
            prm = sub[name]._parameters
            out.append(prm["weight"])
            out.append(prm["bias"])
    return out


def pack_decoder_weights_umma(pc):
    """Packed block of the tcgen05 G1 kernel (layout: csrc/neural_gaussians_umma.cu `ngu::kOff*`).
    Layer-1 rows: head h (opacity, color, cov) occupies rows 56h .. 56h+49.  Layer-2 rows are placed
    where the epilogue reads them: opacity of offset k at row 8(k/5)+k%5, colour (k, c) at 4k+c,
    covariance (k, i) at 8k+i."""
    mods = (pc.get_opacity_mlp, pc.get_color_mlp, pc.get_cov_mlp)
    params = _mlp_params(mods)
    key = tuple((p.data_ptr(), p._version) for p in params)
    cache = _lib.object_cache(pc)
    ent = cache.get("decoder_pack_umma")
    if ent is not None and ent[0] == key:
        return ent[1]
    with torch.no_grad():
        dev = params[0].device
        K = pc.n_offsets
        W1 = torch.zeros(176, 54, device=dev)
        b1 = torch.zeros(176, device=dev)
        for h, m in enumerate(mods):
            W1[56 * h:56 * h + 50] = m[0].weight
            b1[56 * h:56 * h + 50] = m[0].bias
        ko = torch.arange(K, device=dev)
        if K != 10:
            raise NotImplementedError("the tcgen05 G1 kernel is specialised for n_offsets = 10")
        rows_o = 8 * (ko // 5) + ko % 5
        rows_c = (4 * ko.view(K, 1) + torch.arange(3, device=dev).view(1, 3)).reshape(-1)
        rows_v = (8 * ko.view(K, 1) + torch.arange(7, device=dev).view(1, 7)).reshape(-1)
        parts, biases = list(umma_b_operand(W1, 176, 56)), [b1]
        for m, rows, n_pad in zip(mods, (rows_o, rows_c, rows_v), (16, 48, 80)):
            W2 = torch.zeros(n_pad, 50, device=dev)
            b2 = torch.zeros(n_pad, device=dev)
            W2[rows] = m[2].weight
            b2[rows] = m[2].bias
            parts += list(umma_b_operand(W2, n_pad, 56))
            biases.append(b2)
        packed = torch.cat(parts + biases).float().contiguous()
    assert packed.numel() == _lib.lib().cgs_neural_gaussians_umma_packed_floats()
    cache["decoder_pack_umma"] = (key, packed)
    return packed


def g1_impl():
    """'umma' (tcgen05 tensor cores, default) or 'simt' (fp32 FMA tiles; kept for cross-checks)."""
    import os
    v = os.environ.get("CGS_G1_IMPL", "umma")
    if v not in ("umma", "simt"):
        raise ValueError("CGS_G1_IMPL must be 'umma' or 'simt'")
    return v


def g1_bwd_impl():
    """'umma' (tcgen05 data- and weight-gradient kernels fed by what the training-mode forward saves; default with the
    tcgen05 forward) or 'simt' (fp32 FMA kernel that recomputes the forward; kept for cross-checks)."""
    import os
    v = os.environ.get("CGS_G1_BWD_IMPL", "umma" if g1_impl() == "umma" else "simt")
    if v not in ("umma", "simt"):
        raise ValueError("CGS_G1_BWD_IMPL must be 'umma' or 'simt'")
    if v == "umma" and g1_impl() != "umma":
        raise ValueError("CGS_G1_BWD_IMPL=umma needs the tcgen05 forward (CGS_G1_IMPL=umma): it consumes its saved activations")
    return v


def pack_decoder_weights_bwd_umma(pc):
    """Transposed B operands of the tcgen05 data-gradient kernel (csrc/neural_gaussians_bwd_umma.cu `ngbu::kOff*`):
    per head W2^T as [hidden 64][K = padded output layout of the forward], then W1^T as [input 64][K = 192 hidden
    columns (head h at 64h)], each split into TF32 hi / lo.  Cached per parameter version."""
    mods = (pc.get_opacity_mlp, pc.get_color_mlp, pc.get_cov_mlp)
    params = [p for m in mods for p in (m[0].weight, m[2].weight)]
    key = tuple((p.data_ptr(), p._version) for p in params)
    cache = _lib.object_cache(pc)
    ent = cache.get("decoder_pack_bwd_umma")
    if ent is not None and ent[0] == key:
        return ent[1]
    with torch.no_grad():
        dev = params[0].device
        K = pc.n_offsets
        ko = torch.arange(K, device=dev)
        rows_o = 8 * (ko // 5) + ko % 5
        rows_c = (4 * ko.view(K, 1) + torch.arange(3, device=dev).view(1, 3)).reshape(-1)
        rows_v = (8 * ko.view(K, 1) + torch.arange(7, device=dev).view(1, 7)).reshape(-1)
        parts = []
        for m, rows, k_pad in zip(mods, (rows_o, rows_c, rows_v), (16, 48, 80)):
            B = torch.zeros(50, k_pad, device=dev)
            B[:, rows] = m[2].weight.t()                      # B[j][padded(n)] = W2[n][j]
            parts += list(umma_b_operand(B, 64, k_pad))
        B1 = torch.zeros(54, 192, device=dev)
        for h, m in enumerate(mods):
            B1[:, 64 * h:64 * h + 50] = m[0].weight.t()       # B[i][64h + u] = W1_h[u][i] (TMEM head stride 64)
        parts += list(umma_b_operand(B1, 64, 192))
        packed = torch.cat(parts).float().contiguous()
    assert packed.numel() == _lib.lib().cgs_neural_gaussians_bwd_umma_packed_floats()
    cache["decoder_pack_bwd_umma"] = (key, packed)
    return packed


def compact_indices(mask):
    """Order-preserving device-side `nonzero` of a bool/uint8 mask -> (idx[int32, capacity N], count_dev)."""
    L = _lib.lib()
    m = mask.contiguous().view(torch.uint8)
    N = m.numel()
    idx = torch.empty((max(N, 1),), dtype=torch.int32, device=m.device)
    cnt = torch.empty((1,), dtype=torch.int32, device=m.device)
    ws = torch.empty((L.cgs_compact_workspace_bytes(N),), dtype=torch.uint8, device=m.device)
    _lib.check(L.cgs_compact_indices(_lib.ptr(m), N, _lib.ptr(idx), _lib.ptr(cnt), _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr()), "cgs_compact_indices")
    return idx, cnt


def generate_raw(pc, camera_center, anchor, feat, grid_offsets, grid_scaling, binary_grid_masks, vis_idx=None,
                 n_vis=None, impl=None, nv_dev=None, out_cap=None, save=False):
    """Launch the fused kernel.  Inputs are the FULL per-anchor arrays plus an optional visible-index
    list; returns capacity-sized outputs and the device-side Gaussian count.
    nv_dev / out_cap (tcgen05 kernel only): the visible-anchor count stays on the device (n_vis is then the
    capacity of vis_idx), at most out_cap Gaussians are written and the training-side outputs are skipped."""
    L = _lib.lib()
    dev = anchor.device
    Nv = int(anchor.shape[0] if vis_idx is None else (n_vis if n_vis is not None else vis_idx.shape[0]))
    f32 = torch.float32
    if nv_dev is not None:
        cap = max(int(out_cap), 1)
        out = dict(
            xyz=torch.empty((cap, 3), dtype=f32, device=dev), color=torch.empty((cap, 3), dtype=f32, device=dev),
            opacity=torch.empty((cap, 1), dtype=f32, device=dev), scaling=torch.empty((cap, 3), dtype=f32, device=dev),
            rot=torch.empty((cap, 4), dtype=f32, device=dev), count=torch.empty((1,), dtype=torch.int32, device=dev))
        ws = torch.empty((L.cgs_neural_gaussians_umma_workspace_bytes(Nv),), dtype=torch.uint8, device=dev)
        from .rasterizer import _host_floats
        campos = (ctypes.c_float * 3)(*_host_floats(camera_center, 3))
        # named, so that a copy made by .contiguous() outlives the launch (a pointer taken from a temporary dangles)
        anchor, feat, grid_offsets = anchor.contiguous(), feat.contiguous(), grid_offsets.contiguous()
        grid_scaling, binary_grid_masks = grid_scaling.contiguous(), binary_grid_masks.contiguous()
        _lib.check(L.cgs_neural_gaussians_umma_forward_dev(
            _lib.ptr(pack_decoder_weights_umma(pc)), _lib.ptr(vis_idx), Nv, _lib.ptr(nv_dev), cap,
            _lib.ptr(anchor), _lib.ptr(feat), _lib.ptr(grid_offsets),
            _lib.ptr(grid_scaling), _lib.ptr(binary_grid_masks), campos, _lib.ptr(out["xyz"]),
            _lib.ptr(out["color"]), _lib.ptr(out["opacity"]), _lib.ptr(out["scaling"]), _lib.ptr(out["rot"]), None, None,
            _lib.ptr(out["count"]), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "cgs_neural_gaussians_umma_forward_dev")
        out["n_vis"] = Nv
        return out
    cap = max(Nv * pc.n_offsets, 1)
    out = dict(
        xyz=torch.empty((cap, 3), dtype=f32, device=dev), color=torch.empty((cap, 3), dtype=f32, device=dev),
        opacity=torch.empty((cap, 1), dtype=f32, device=dev), scaling=torch.empty((cap, 3), dtype=f32, device=dev),
        rot=torch.empty((cap, 4), dtype=f32, device=dev), neural_opacity=torch.empty((cap, 1), dtype=f32, device=dev),
        mask=torch.empty((cap,), dtype=torch.uint8, device=dev), count=torch.empty((1,), dtype=torch.int32, device=dev))
    impl = impl or g1_impl()
    if impl == "umma":
        ws_bytes, fn, packed = L.cgs_neural_gaussians_umma_workspace_bytes(Nv), L.cgs_neural_gaussians_umma_forward, \
            pack_decoder_weights_umma(pc)
    else:
        ws_bytes, fn, packed = L.cgs_neural_gaussians_workspace_bytes(Nv), L.cgs_neural_gaussians_forward, \
            pack_decoder_weights(pc)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    from .rasterizer import _host_floats
    campos = (ctypes.c_float * 3)(*_host_floats(camera_center, 3))
    anchor, feat, grid_offsets = anchor.contiguous(), feat.contiguous(), grid_offsets.contiguous()   # (named: see above)
    grid_scaling, binary_grid_masks = grid_scaling.contiguous(), binary_grid_masks.contiguous()
    if save:
        # training mode with the tcgen05 backward: the forward leaves its activations behind (1.3 kB / visible anchor)
        if impl != "umma":
            raise ValueError("saved activations are produced by the tcgen05 forward only")
        tiles = (max(Nv, 1) + 127) // 128
        i32 = torch.int32
        sv = dict(h=torch.empty((max(Nv, 1), 176), dtype=f32, device=dev), hmask=torch.empty((max(Nv, 1), 6), dtype=i32, device=dev),
                  pre2=torch.empty((max(Nv, 1), 144), dtype=f32, device=dev),
                  rowpos=torch.empty((max(Nv, 1), 2), dtype=i32, device=dev), tilebase=torch.empty((tiles,), dtype=i32, device=dev))
        _lib.check(L.cgs_neural_gaussians_umma_forward_train(
            _lib.ptr(packed), _lib.ptr(vis_idx), Nv, _lib.ptr(anchor), _lib.ptr(feat),
            _lib.ptr(grid_offsets), _lib.ptr(grid_scaling), _lib.ptr(binary_grid_masks),
            campos, _lib.ptr(out["xyz"]), _lib.ptr(out["color"]), _lib.ptr(out["opacity"]), _lib.ptr(out["scaling"]),
            _lib.ptr(out["rot"]), _lib.ptr(out["neural_opacity"]), _lib.ptr(out["mask"]), _lib.ptr(out["count"]),
            _lib.ptr(sv["h"]), _lib.ptr(sv["hmask"]), _lib.ptr(sv["pre2"]), _lib.ptr(sv["rowpos"]), _lib.ptr(sv["tilebase"]),
            _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "cgs_neural_gaussians_umma_forward_train")
        out["n_vis"] = Nv
        out["save"] = sv
        return out
    _lib.check(fn(
        _lib.ptr(packed), _lib.ptr(vis_idx), Nv, _lib.ptr(anchor),
        _lib.ptr(feat), _lib.ptr(grid_offsets), _lib.ptr(grid_scaling),
        _lib.ptr(binary_grid_masks), campos, _lib.ptr(out["xyz"]), _lib.ptr(out["color"]),
        _lib.ptr(out["opacity"]), _lib.ptr(out["scaling"]), _lib.ptr(out["rot"]), _lib.ptr(out["neural_opacity"]),
        _lib.ptr(out["mask"]), _lib.ptr(out["count"]), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
        "cgs_neural_gaussians_forward")
    out["n_vis"] = Nv
    return out


def pack_decoder_weights_transposed(pc):
    """Backward-GEMM block: W1T[150][56] (= the three Linear(54,50).weight stacked), then the
    Linear(50,n).weight of each head as [n][52] (layout: csrc/neural_gaussians_bwd.cu `ngb::kOff*T`)."""
    mods = (pc.get_opacity_mlp, pc.get_color_mlp, pc.get_cov_mlp)
    with torch.no_grad():
        dev = mods[0][0].weight.device
        W1T = torch.zeros(150, 56, device=dev)
        for i, m in enumerate(mods):
            W1T[50 * i:50 * i + 50, :54] = m[0].weight
        parts = [W1T.reshape(-1)]
        for m in mods:
            n = m[2].weight.shape[0]
            W2T = torch.zeros(n, 52, device=dev)
            W2T[:, :50] = m[2].weight
            parts.append(W2T.reshape(-1))
        packed = torch.cat(parts).float().contiguous()
    assert packed.numel() == _lib.lib().cgs_neural_gaussians_backward_packed_floats()
    return packed


def unpack_decoder_weight_grads(d_packed):
    """Gradient block in the forward layout (`pack_decoder_weights`) -> 12 tensors shaped like
    (opacity, color, cov) x (l0.weight, l0.bias, l2.weight, l2.bias)."""
    o = 0
    W1 = d_packed[o:o + 54 * 152].view(54, 152); o += 54 * 152
    b1 = d_packed[o:o + 152]; o += 152
    out = []
    heads = []
    for n, ld in ((10, 12), (30, 32), (70, 72)):
        W2 = d_packed[o:o + 50 * ld].view(50, ld); o += 50 * ld
        b2 = d_packed[o:o + ld]; o += ld
        heads.append((W2[:, :n].t().contiguous(), b2[:n].contiguous()))
    for i in range(3):
        out += [W1[:, 50 * i:50 * i + 50].t().contiguous(), b1[50 * i:50 * i + 50].contiguous(), heads[i][0], heads[i][1]]
    return out


def _decoder_params(pc):
    mods = (pc.get_opacity_mlp, pc.get_color_mlp, pc.get_cov_mlp)
    return [p for m in mods for p in (m[0].weight, m[0].bias, m[2].weight, m[2].bias)]


class _NeuralGaussians(torch.autograd.Function):
    """Differentiable fused G1: forward = cgs_neural_gaussians_(umma_)forward, backward =
    cgs_neural_gaussians_backward.  Replaces the autograd graph of
    gaussian_renderer/__init__.py:106-145."""

    @staticmethod
    def forward(ctx, pc, campos, vis_idx, n_vis, anchor, feat, offsets, scaling, mask, *params):
        N = anchor.shape[0]
        K = pc.n_offsets
        a, f, o = anchor.detach().contiguous(), feat.detach().contiguous(), offsets.detach().reshape(N, -1).contiguous()
        sc, m = scaling.detach().contiguous(), mask.detach().reshape(N, -1).contiguous()
        need_grad = any(ctx.needs_input_grad)
        ctx.bwd_impl = g1_bwd_impl() if need_grad else "simt"
        raw = generate_raw(pc, campos, a, f, o, sc, m, vis_idx=vis_idx, n_vis=n_vis,
                           save=(ctx.bwd_impl == "umma" and n_vis > 0))
        P = int(raw["count"].item())  # the reference synchronises here too (boolean indexing, :119,136)
        if P < 0:
            raise _lib.CgsError("cgs_neural_gaussians_umma_forward: a tensor-core completion barrier timed out")
        _lib.raise_deferred()   # error flags of earlier backward kernels (the read-back above has synchronised)
        ctx.pc, ctx.campos, ctx.n_vis, ctx.P = pc, campos, n_vis, P
        ctx.shapes = (offsets.shape, mask.shape)
        ctx.saved_act = raw.get("save")
        ctx.save_for_backward(a, f, o, sc, m, vis_idx, raw["mask"])
        nop = raw["neural_opacity"][:n_vis * K]
        keep = raw["mask"][:n_vis * K]
        ctx.mark_non_differentiable(nop, keep)
        return (raw["xyz"][:P], raw["color"][:P], raw["opacity"][:P], raw["scaling"][:P], raw["rot"][:P], nop, keep)

    @staticmethod
    def backward(ctx, g_xyz, g_color, g_opacity, g_scaling, g_rot, _g_nop, _g_keep):
        L = _lib.lib()
        a, f, o, sc, m, vis_idx, keep = ctx.saved_tensors
        pc, P, n_vis = ctx.pc, ctx.P, ctx.n_vis
        N, dev = a.shape[0], a.device
        z = lambda t: torch.zeros_like(t)
        d_a, d_f, d_o, d_sc, d_m = z(a), z(f), z(o), z(sc), z(m)
        d_w = torch.zeros(L.cgs_neural_gaussians_packed_floats(), device=dev)
        if P > 0 and n_vis > 0:
            c = lambda g, shape: (torch.zeros(shape, device=dev) if g is None else g.contiguous().float())
            g_xyz, g_color, g_scaling = c(g_xyz, (P, 3)), c(g_color, (P, 3)), c(g_scaling, (P, 3))
            g_opacity, g_rot = c(g_opacity, (P, 1)), c(g_rot, (P, 4))
            from .rasterizer import _host_floats
            campos = (ctypes.c_float * 3)(*_host_floats(ctx.campos, 3))
            sv = ctx.saved_act
        if P > 0 and n_vis > 0 and sv is not None:
            # tcgen05 path: data gradients + weight gradients from the saved activations (csrc/neural_gaussians_bwd_umma.cu)
            d_out = torch.empty((n_vis, 144), dtype=torch.float32, device=dev)
            d_pre = torch.empty((n_vis, 176), dtype=torch.float32, device=dev)
            err = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.check(L.cgs_neural_gaussians_backward_umma(
                _lib.ptr(pack_decoder_weights_bwd_umma(pc)), _lib.ptr(vis_idx), n_vis, _lib.ptr(a), _lib.ptr(f), _lib.ptr(o),
                _lib.ptr(sc), _lib.ptr(m), campos, _lib.ptr(keep), _lib.ptr(sv["h"]), _lib.ptr(sv["hmask"]),
                _lib.ptr(sv["pre2"]), _lib.ptr(sv["rowpos"]), _lib.ptr(sv["tilebase"]), _lib.ptr(g_xyz), _lib.ptr(g_color),
                _lib.ptr(g_opacity), _lib.ptr(g_scaling), _lib.ptr(g_rot), _lib.ptr(d_a), _lib.ptr(d_f), _lib.ptr(d_o),
                _lib.ptr(d_sc), _lib.ptr(d_m), _lib.ptr(d_w), _lib.ptr(d_out), _lib.ptr(d_pre), _lib.ptr(err),
                _lib.stream_ptr()), "cgs_neural_gaussians_backward_umma")
            ctx.saved_act = None
            _lib.deferred_error_check(err, "cgs_neural_gaussians_backward_umma: a tensor-core completion barrier timed out")
        elif P > 0 and n_vis > 0:
            ws = torch.empty((L.cgs_neural_gaussians_backward_workspace_bytes(n_vis),), dtype=torch.uint8, device=dev)
            w_t = pack_decoder_weights_transposed(pc)   # named: not cached, must outlive the launch
            _lib.check(L.cgs_neural_gaussians_backward(
                _lib.ptr(pack_decoder_weights(pc)), _lib.ptr(w_t), _lib.ptr(vis_idx),
                n_vis, _lib.ptr(a), _lib.ptr(f), _lib.ptr(o), _lib.ptr(sc), _lib.ptr(m), campos, _lib.ptr(keep),
                _lib.ptr(g_xyz), _lib.ptr(g_color), _lib.ptr(g_opacity), _lib.ptr(g_scaling), _lib.ptr(g_rot),
                _lib.ptr(d_a), _lib.ptr(d_f), _lib.ptr(d_o), _lib.ptr(d_sc), _lib.ptr(d_m), _lib.ptr(d_w), _lib.ptr(ws),
                ws.numel(), _lib.stream_ptr()), "cgs_neural_gaussians_backward")
        off_shape, mask_shape = ctx.shapes
        return (None, None, None, None, d_a, d_f, d_o.view(off_shape), d_sc, d_m.view(mask_shape),
                *unpack_decoder_weight_grads(d_w))


def neural_gaussians(pc, camera_center, anchor, feat, grid_offsets, grid_scaling, binary_grid_masks, vis_idx, n_vis):
    """Differentiable anchor -> Gaussian generation on the visible anchors `vis_idx[:n_vis]`.
    Returns (xyz, color, opacity, scaling, rot, neural_opacity, selection_mask[bool])."""
    out = _NeuralGaussians.apply(pc, camera_center, vis_idx, n_vis, anchor, feat, grid_offsets, grid_scaling,
                                 binary_grid_masks, *_decoder_params(pc))
    xyz, color, opacity, scaling, rot, nop, keep = out
    return xyz, color, opacity, scaling, rot, nop, keep.bool()


def select_attributes(pc, is_training, step):
    """The per-anchor attributes `generate_neural_gaussians` feeds to the decoder MLPs, by training stage
    (gaussian_renderer/__init__.py:31-104) -> (anchor, feat, offsets, scaling, masks, bit outputs ...)."""
    from .context_model import multi_scale_generating
    anchor_all = pc.get_anchor
    bits = (None, None, None, None, None)
    feat, grid_offsets, grid_scaling = pc._anchor_feat, pc._offset, pc.get_scaling
    binary_grid_masks = pc.get_mask
    if is_training:
        if 3000 < step <= 10000:
            feat = feat + torch.empty_like(feat).uniform_(-0.5, 0.5) * Q_FEAT
            grid_scaling = grid_scaling + torch.empty_like(grid_scaling).uniform_(-0.5, 0.5) * Q_SCALING
            grid_offsets = grid_offsets + torch.empty_like(grid_offsets).uniform_(-0.5, 0.5) * Q_OFFSETS
        if step == 10000:
            pc.update_anchor_bound()
        if step > 10000:
            mask_anchor_bool = pc.get_mask_anchor.to(torch.bool)
            feat, grid_scaling, grid_offsets, *bits = multi_scale_generating(
                pc, anchor_all, pc._hyper_latent, feat, grid_offsets, grid_scaling, binary_grid_masks,
                mask_anchor_bool, predict_bpp=True, training=True)
    elif not pc.decoded_version:
        mask_anchor_bool = pc.get_mask_anchor.to(torch.bool)
        feat, grid_scaling, grid_offsets = multi_scale_generating(
            pc, anchor_all, pc._hyper_latent, feat, grid_offsets, grid_scaling, binary_grid_masks, mask_anchor_bool,
            predict_bpp=False, training=False)
    return anchor_all, feat, grid_offsets, grid_scaling, binary_grid_masks, tuple(bits)


def generate_neural_gaussians(viewpoint_camera, pc, visible_mask=None, is_training=False, step=0):
    """Same signature and return tuples as gaussian_renderer/__init__.py:25-150.  Differentiable
    w.r.t. the per-anchor parameters and the decoder MLPs through `_NeuralGaussians` (autograd is
    recorded whenever torch.is_grad_enabled()); for step > 10000 the context model is differentiable
    too (`context_model._ContextModelTrain`), so `loss.backward()` of train.py:199-211 reaches every
    parameter the reference trains."""
    anchor_all, feat, grid_offsets, grid_scaling, binary_grid_masks, bits = select_attributes(pc, is_training, step)
    bit_per_param, bit_per_feat_param, bit_per_scaling_param, bit_per_offsets_param, bpp_per_level = bits
    N = anchor_all.shape[0]
    if visible_mask is None:
        visible_mask = torch.ones(N, dtype=torch.bool, device=anchor_all.device)

    with torch.no_grad():
        vis_idx, cnt = compact_indices(visible_mask)
        n_vis = int(cnt.item())  # the reference synchronises here too (boolean indexing, :44-50)
    xyz, color, opacity, scaling, rot, neural_opacity, mask = neural_gaussians(
        pc, viewpoint_camera.camera_center, anchor_all, feat, grid_offsets, grid_scaling, binary_grid_masks, vis_idx,
        n_vis)
    if is_training:
        return (xyz, color, opacity, scaling, rot, neural_opacity, mask, bit_per_param, 16, bit_per_feat_param,
                bit_per_scaling_param, bit_per_offsets_param, bpp_per_level)
    return xyz, color, opacity, scaling, rot, 0
