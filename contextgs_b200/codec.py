"""GPU bitstream codec: the `conduct_encoding` / `conduct_decoding` pair of the reference
(scene/gaussian_model.py:1005-1300 and :1302-1538) on top of csrc/entropy_codec.cu.

Same structure as the reference -- anchors as raw 16-bit grid indices, offset masks as one Bernoulli
stream, hyper latents under the factorised prior, then feat / scaling / masked offsets level by level
(coarse -> fine), each level predicted by the context MLP from what has been decoded so far -- and the
same file names in the output directory (anchor.npy, masks.b, hyper.b, feat{L}.b, scaling{L}.b,
offsets{L}.b, meta.b, mlp.pt).  The byte format of the .b streams is this library's own range coder
(torchac is not part of the reference tree and cannot be pinned): what is guaranteed and tested is that
decoding returns the encoder's quantised tensors bit for bit.

Everything heavy runs on the GPU: per level one kernel (cgs_context_level_umma_forward_ex) produces the
(mean, scale, Q) of every coded value and the alphabets of the level's three streams, a parallel pass turns every
value into its 16-bit coding interval and one thread per <= 480-symbol chunk carries the range coder's state over
them (cgs_codec_gauss_level_encode); decoding inverts the tabulated normal CDF per symbol
(cgs_codec_gauss_level_decode).  The host reads back ONE block of scalars per encode (stream sizes, error flags) and
slices the packed byte strings.
"""
import ctypes
import os
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .context_model import build_level_plan, find_divide_scale, global_means, pack_grid_weights_umma
from .encodings import Q_anchor, Quantize_anchor

# Level rows per independently coded chunk of a feat stream (the reference: 1000 anchors, coded one after the
# other on the host).  One GPU thread codes one chunk and the kernel time is the serial latency of one coder, so
# the chunk size sets the speed: 8 rows = 400 feat symbols give ~190 k concurrent coders per million anchors.
# Side information per chunk: a 16-bit length (+ the coder's single termination byte); the alphabet bounds are per
# (level, attribute) stream.  Streams with fewer values per row use more rows per chunk.
CHUNK_ROWS = 8
ATTRS = (("feat", 50), ("scaling", 6), ("offsets", 30))
ATTR_CHUNK_MULT = (1, 8, 2)     # rows per chunk = CHUNK_ROWS * mult: 400 / 384 / 480 symbols per chunk
TABLE_CHUNK_MULT = 4            # hyper (12 per row) and mask (10 per row) streams: 32 rows per chunk
PARAM_LD = 176
# A level with few rows has few chunks, i.e. few coder threads, and then the SERIAL length of a chunk is the whole kernel
# time (decode of the 60 k-row coarsest level of a 1.5 M-anchor model: 12 k threads walking 400 symbols each took as long
# as the 1.2 M-row level).  Levels therefore halve their chunk rows until they have MIN_FEAT_CHUNKS feat chunks (or reach
# MIN_CHUNK_ROWS); the rows actually used travel with the level (`chunk_rows` of its entry, `meta.b`).  A chunk costs 3
# bytes of side information (16-bit length + the coder's single termination byte).
STREAM_VERSION = 2              # 2: one termination byte per chunk, no leading zero, per-level chunk rows, tabulated normal CDF
MIN_FEAT_CHUNKS = 100_000
MIN_CHUNK_ROWS = 2


def dequantize_anchor(q, x_bound_min, x_bound_max):
    """utils/encodings.py:224-227: grid index -> position (the exact expression the encoder's anchors come from)."""
    interval = (x_bound_max - x_bound_min) * Q_anchor + 1e-6
    return q.float() * interval + x_bound_min


def frequency_tables(pmf):
    """pmf [T, L] (any positive weights) -> int32 [T, L + 1] cumulative 16-bit frequencies with every symbol
    codable: C(i) = floor(cum_i / cum_L * (65536 - L)) + i, C(0) = 0, C(L) = 65536.  Computed in float64 on the
    device of `pmf` (no host round trip when the weights are already on the GPU)."""
    pmf = pmf.detach().to(torch.float64).clamp_min(0)
    T, L = pmf.shape
    if L >= 32768:
        raise ValueError("alphabet too large for 16-bit frequencies")
    dev = pmf.device
    cum = torch.cumsum(pmf, dim=1)
    total = cum[:, -1:].clamp_min(1e-300)
    body = torch.floor(cum / total * (65536 - L)).to(torch.int64) + torch.arange(1, L + 1, device=dev).view(1, L)
    body[:, -1] = 65536
    return torch.cat([torch.zeros(T, 1, dtype=torch.int64, device=dev), body], dim=1).to(torch.int32)


def mask_table(p1):
    """The Bernoulli table of the offset-mask stream from P(mask = 1) (utils/encodings.py:147-165).  `p1` is a python
    float (decoder: from the metadata) or a 0-dim tensor (encoder: stays on the device); both go through the same
    float64 arithmetic, so the two sides build identical tables."""
    p = p1.detach().to(torch.float64).reshape(()) if torch.is_tensor(p1) else torch.tensor(float(p1), dtype=torch.float64)
    return frequency_tables(torch.stack([(1.0 - p).clamp_min(1e-9), p.clamp_min(1e-9)]).view(1, 2))


def phi_table():
    """(T float32 [4097], z0, inv_h): the tabulated normal CDF of the Gaussian streams (include/contextgs_b200.h)."""
    T = np.empty(4097, dtype=np.float32)
    z0, inv_h = ctypes.c_float(), ctypes.c_float()
    _lib.check(_lib.lib().cgs_codec_phi_table(T.ctypes.data_as(ctypes.c_void_p), T.size, ctypes.byref(z0), ctypes.byref(inv_h)),
               "cgs_codec_phi_table")
    return T, float(z0.value), float(inv_h.value)


def level_chunk_rows(n_rows, chunk_rows, adaptive=True):
    """Rows per chunk of the (feat, scaling, offsets) streams of a level with `n_rows` rows."""
    rows = int(chunk_rows)
    while adaptive and rows > MIN_CHUNK_ROWS and rows % 2 == 0 and n_rows // rows < MIN_FEAT_CHUNKS:
        rows //= 2
    return [rows * m for m in ATTR_CHUNK_MULT]


def _offsets(lens):
    """exclusive prefix sum of the chunk lengths, with the total appended: int64 [n + 1]"""
    off = torch.zeros(lens.numel() + 1, dtype=torch.int64, device=lens.device)
    torch.cumsum(lens, 0, dtype=torch.int64, out=off[1:])
    return off


def _table_encode(symbols, tables, chunk_rows, err):
    """Codes a static-table stream into fixed-stride scratch slots.  Returns (scratch, cap, lens, off): packing into one
    byte string needs the total on the host and is deferred to _table_pack so that an encode has ONE read-back."""
    L = _lib.lib()
    dev = symbols.device
    n, C = symbols.shape
    n_chunks = (n + chunk_rows - 1) // chunk_rows
    cap = (2 * C * chunk_rows + 16 + 3) // 4 * 4
    scratch = torch.empty(max(n_chunks, 1) * cap // 4, dtype=torch.int32, device=dev)
    lens = torch.zeros(max(n_chunks, 1), dtype=torch.int32, device=dev)
    tb = tables.to(dev).contiguous()
    _lib.check(L.cgs_codec_table_encode(_lib.ptr(symbols), n, C, chunk_rows, _lib.ptr(tb), tb.shape[0], tb.shape[1],
                                        _lib.ptr(scratch), cap, _lib.ptr(lens), _lib.ptr(err), _lib.stream_ptr()),
               "cgs_codec_table_encode")
    lens = lens[:n_chunks]
    return scratch, cap, lens, _offsets(lens)


def _table_pack(pending, total):
    scratch, cap, lens, off = pending
    packed = torch.empty(max(total, 1), dtype=torch.uint8, device=lens.device)
    _lib.check(_lib.lib().cgs_codec_pack_streams(_lib.ptr(scratch), cap, _lib.ptr(lens), _lib.ptr(off), lens.numel(),
                                                 _lib.ptr(packed), _lib.stream_ptr()), "cgs_codec_pack_streams")
    return packed[:total]


def _table_decode(packed, lens, n, C, tables, chunk_rows):
    L = _lib.lib()
    dev = packed.device
    lens = lens.to(torch.int32).contiguous()
    off = _offsets(lens)
    tb = tables.to(dev).contiguous()
    tlen = torch.full((tb.shape[0],), tb.shape[1] - 1, dtype=torch.int32, device=dev)
    sym = torch.empty((n, C), dtype=torch.int16, device=dev)
    _lib.check(L.cgs_codec_table_decode(_lib.ptr(packed), _lib.ptr(off), _lib.ptr(lens), n, C, chunk_rows,
                                        _lib.ptr(tb), _lib.ptr(tlen), tb.shape[0], tb.shape[1], _lib.ptr(sym),
                                        _lib.stream_ptr()), "cgs_codec_table_decode")
    return sym


def _hyper_tables(pc, smin, smax):
    """Factorised-prior pmf of the integer symbols smin..smax per channel (EntropyBottleneck likelihood at
    median + s, the quantity `latent_codec.update()` tabulates) -> 16-bit cumulative frequencies."""
    dev = pc.latent_codec.quantiles.device
    median = pc.latent_codec.quantiles[:, 0, 1].detach()
    grid = torch.arange(smin, smax + 1, device=dev, dtype=torch.float32).view(-1, 1) + median.view(1, -1)
    _, lik = pc.latent_codec(grid.contiguous(), training=False)
    return frequency_tables(lik.t().contiguous() + 1e-12)


def _content_key(anchor_q, pc, extra=None):
    """Cache key from the CONTENT of the coded anchors and the bounds they are dequantised with: two position-weighted
    int64 checksums of the 16-bit grid indices + the six bound values, one small read-back.  (Keying on data_ptr /
    _version of a temporary would let the caching allocator hand the same address to another stream's anchors.)
    `extra`: int64 device scalars the caller needs on the host as well; they ride on the same read-back and are
    returned as a list after the key."""
    q = (anchor_q.to(torch.int64) & 0xffff).reshape(-1)
    w = torch.arange(1, q.numel() + 1, device=q.device, dtype=torch.int64)
    sums = torch.stack([q.sum(), (q * (w % 65521 + 1)).sum(), (q * (w % 8191 + 7)).sum()])
    bounds = torch.cat([pc.x_bound_min.reshape(-1), pc.x_bound_max.reshape(-1)]).float().to(q.device).view(torch.int32)
    parts = [sums, bounds.to(torch.int64)] + ([] if extra is None else [torch.stack(list(extra)).to(torch.int64)])
    vals = torch.cat(parts).tolist()   # exact: integer checksums, bit patterns of the bounds
    key = (tuple(anchor_q.shape),) + tuple(vals[:9])
    return key if extra is None else (key, vals[9:])


def _plan_for(pc, anchor, key, rank, world):
    """(full level sizes, plan or shard of it), cached on the model while the coded anchors (`key`: content checksums
    + bounds, see _content_key) and the level scales are unchanged: the division and especially its dependency-root
    sharding are index gymnastics with host synchronisations that would otherwise dominate a sharded encode / decode."""
    key = (key, tuple(pc.level_scale), float(pc.voxel_size), rank, world)
    ent = getattr(pc, "_cgs_codec_plan", None)
    if ent is not None and ent[0] == key:
        return ent[1], ent[2]
    plan = build_level_plan(pc, anchor, None)
    sizes = [lv.n for lv in plan.levels]
    if world > 1:
        from .distributed import shard_level_plan
        plan = shard_level_plan(plan, rank, world)
    try:
        pc._cgs_codec_plan = (key, sizes, plan)
    except Exception:
        pass
    return sizes, plan


def _level_params(pc, lv, anchor, hyper_q, feat, scaling, offsets, masks, feat_q, scaling_q, offsets_q, sums, err, means,
                  predict_only, minmax=None, choose=None):
    """(mean, scale, Q) of every coded value of one level (and, when encoding, the level's quantised values and, in
    `minmax`, the alphabets of its three streams).  `choose` (uint8 per anchor): rows whose bits are estimated (None: all)."""
    L = _lib.lib()
    packed, in_dim = pack_grid_weights_umma(pc, lv.level)
    params = torch.empty((lv.n, PARAM_LD), dtype=torch.float32, device=anchor.device)
    p = _lib.ptr
    _lib.check(L.cgs_context_level_umma_forward_ex(
        in_dim, p(packed), p(lv.orig), p(lv.ctx_src), p(lv.level_anchor), lv.n, p(anchor), p(hyper_q), p(feat), p(scaling),
        p(offsets), p(masks), p(choose), None, means[0], means[1], means[2], p(feat_q), p(scaling_q), p(offsets_q), None,
        p(sums), p(err), p(params), int(predict_only), p(minmax), _lib.stream_ptr()), "cgs_context_level_umma_forward_ex")
    return params


@torch.no_grad()
def encode_model(pc, chunk_rows=CHUNK_ROWS, rank=0, world=1, estimate_bits=True, adaptive_chunks=True):
    """Encode every valid anchor of `pc`.  Returns a SimpleNamespace with the byte streams (CUDA uint8 tensors),
    the metadata the decoder needs, the quantised tensors that were coded (for parity checks) and the
    estimated bits of the same pass.
    chunk_rows: level rows per feat chunk (scaling / offsets: ATTR_CHUNK_MULT times as many); with adaptive_chunks small
    levels use fewer (level_chunk_rows).
    estimate_bits=False skips the entropy estimate (`estimated_bits` is then None): the reference's conduct_encoding only
    reports the sizes of the streams it wrote, and the estimate (two erf and a log per value) costs about as much as
    predicting the level.
    world > 1 (BASELINE configs[3]: anchors sharded over the GPUs): the level plan is split by
    dependency root (distributed.shard_level_plan), so rank `rank` predicts and codes only its own rows of every
    level, with no exchange; the small anchor / mask / hyper streams are produced identically on every rank (they
    are inputs of every shard).  The per-rank level streams simply sit side by side in the container."""
    L = _lib.lib()
    sel = pc.get_mask_anchor
    dev = sel.device
    tensors = (pc._anchor, pc._hyper_latent, pc._anchor_feat, pc._offset, pc.get_scaling, pc.get_mask)
    if not (pc.all_anchors_valid() if hasattr(pc, "all_anchors_valid") else bool(sel.all())):
        idx = torch.nonzero(sel)[:, 0]
        tensors = tuple(t.index_select(0, idx) for t in tensors)
    a_raw, hyper, feat, offsets, scaling, masks = (t.detach().contiguous().float() for t in tensors)
    N, K = a_raw.shape[0], pc.n_offsets
    offsets, masks = offsets.reshape(N, 3 * K).contiguous(), masks.reshape(N, K).contiguous()
    err = torch.zeros(1, dtype=torch.int32, device=dev)

    # anchors: 16-bit grid indices (saved raw, like the reference's anchor.npy)
    _, qv = Quantize_anchor.apply(a_raw, pc.x_bound_min, pc.x_bound_max)
    anchor_q = qv.to(torch.int32)
    anchor = dequantize_anchor(anchor_q, pc.x_bound_min, pc.x_bound_max).contiguous()

    # offset masks: one Bernoulli table (built on the device; P(1) reaches the host with the final read-back)
    p1_dev = masks.mean()
    # (symbols by comparison: the straight-through value of a kept offset is (1 - s) + s, not exactly 1 for every s)
    mask_pending = _table_encode((masks != 0).to(torch.int16).contiguous(), mask_table(p1_dev), chunk_rows * TABLE_CHUNK_MULT,
                                 err)

    # hyper latents under the factorised prior.  The symbol range sizes the table, so it is needed on the host: it shares
    # the read-back of the level-plan cache key
    # eval-mode quantisation of the EntropyBottleneck (round about the per-channel median) without its likelihoods
    median = pc.latent_codec.quantiles[:, 0, 1].detach().float()
    hsteps = torch.round(hyper - median.view(1, -1))
    hyper_q = hsteps + median.view(1, -1)
    hsym = hsteps.to(torch.int32)
    hlo, hhi = torch.aminmax(hsym)
    content_key, (hmin, hmax) = _content_key(anchor_q, pc, extra=(hlo, hhi))
    hyper_pending = _table_encode((hsym - hmin).to(torch.int16).contiguous(), _hyper_tables(pc, hmin, hmax),
                                  chunk_rows * TABLE_CHUNK_MULT, err)
    if getattr(pc, "disable_hyper", False):
        hyper_q = hyper_q * 0

    # level division on the DEQUANTISED anchors (what the decoder will see)
    if pc.level_scale is None:
        pc.level_scale = find_divide_scale(pc, anchor, pc.target_ratio, pc.level_num)
    n_levels_full, plan = _plan_for(pc, anchor, content_key, rank, world)

    feat_q, scaling_q, offsets_q = torch.zeros_like(feat), torch.zeros_like(scaling), torch.zeros_like(offsets)
    sums = torch.zeros(16, dtype=torch.float64, device=dev)
    terr = torch.zeros(1, dtype=torch.int32, device=dev)
    means = global_means(pc)
    no_rows = None if estimate_bits else torch.zeros(N, dtype=torch.uint8, device=dev)
    levels, pending = [], []
    for li, lv in enumerate(plan.levels):
        entry = SimpleNamespace(level=lv.level, n=lv.n, streams={}, chunk_rows=level_chunk_rows(lv.n, chunk_rows, adaptive_chunks))
        levels.append(entry)
        if lv.n == 0:
            continue
        rows3 = (ctypes.c_int * 3)(*entry.chunk_rows)
        minmax = torch.empty(6, dtype=torch.int32, device=dev)
        params = _level_params(pc, lv, anchor, hyper_q, feat, scaling, offsets, masks, feat_q, scaling_q, offsets_q,
                               sums[4 * li:4 * li + 4], terr, means, False, minmax, no_rows)
        counts = (ctypes.c_int32 * 3)()
        words = int(L.cgs_codec_gauss_level_chunks(lv.n, rows3, counts))
        counts = list(counts)
        intervals = torch.empty(lv.n * 86, dtype=torch.int32, device=dev)
        scratch = torch.empty(words, dtype=torch.int32, device=dev)
        lens = torch.zeros(sum(counts), dtype=torch.int32, device=dev)
        p = _lib.ptr
        _lib.check(L.cgs_codec_gauss_level_encode(p(lv.orig), lv.n, rows3, p(params), p(masks), p(feat_q), p(scaling_q),
                                                  p(offsets_q), p(minmax), p(intervals), p(scratch), p(lens), p(err),
                                                  _lib.stream_ptr()), "cgs_codec_gauss_level_encode")
        off = _offsets(lens)
        bounds = [counts[0], counts[0] + counts[1], counts[0] + counts[1] + counts[2]]
        pending.append((entry, lv.n, scratch, lens, off, counts, minmax, off[bounds], rows3))

    # ONE read-back for everything the host needs: error flags, stream sizes (to allocate the packed byte strings),
    # P(mask), the estimated bits
    i64 = lambda t: t.reshape(-1).to(torch.int64)
    parts = [i64(err), i64(terr), mask_pending[3][-1:], hyper_pending[3][-1:], p1_dev.double().reshape(1).view(torch.int64),
             sums.view(torch.int64)] + [pd[7] for pd in pending]
    host = torch.cat(parts).cpu()
    e, te, mask_total, hyper_total = (int(v) for v in host[:4])
    if te:
        raise _lib.CgsError("cgs_context_level_umma_forward_ex: a tensor-core completion barrier timed out")
    if e:
        raise _lib.CgsError({1: "a symbol has an empty coding interval", 2: "a chunk's alphabet exceeds 32768 symbols",
                             3: "a stream outgrew its capacity"}.get(e, f"codec error {e}"))
    p1 = float(host[4:5].view(torch.float64)[0])
    s = host[5:21].view(torch.float64).tolist()
    mask_bytes, hyper_bytes = _table_pack(mask_pending, mask_total), _table_pack(hyper_pending, hyper_total)
    for (entry, n, scratch, lens, off, counts, minmax, _, rows3), ends in zip(pending,
                                                                              host[21:].reshape(len(pending), 3).tolist()):
        packed = torch.empty(max(ends[2], 1), dtype=torch.uint8, device=dev)
        _lib.check(L.cgs_codec_gauss_level_pack(n, rows3, _lib.ptr(scratch), _lib.ptr(lens), _lib.ptr(off), _lib.ptr(packed),
                                                _lib.stream_ptr()), "cgs_codec_gauss_level_pack")
        b0 = c0 = 0
        for attr, (name, dim) in enumerate(ATTRS):
            entry.streams[name] = SimpleNamespace(bytes=packed[b0:ends[attr]], lens=lens[c0:c0 + counts[attr]],
                                                  minmax=minmax[2 * attr:2 * attr + 2])
            b0, c0 = ends[attr], c0 + counts[attr]
    meta = dict(version=STREAM_VERSION, N_total=int(pc._anchor.shape[0]), N=N, chunk_rows=chunk_rows, voxel_size=float(pc.voxel_size),
                level_scale=[float(s_) for s_ in pc.level_scale], x_bound_min=pc.x_bound_min.detach().cpu(),
                x_bound_max=pc.x_bound_max.detach().cpu(), prob_masks=p1, hyper_min=hmin, hyper_max=hmax,
                means=means, N_levels=n_levels_full, world=world)
    est = dict(hyper=None, feat=sum(s[4 * i] for i in range(3)), scaling=sum(s[4 * i + 1] for i in range(3)),
               offsets=sum(s[4 * i + 2] for i in range(3))) if estimate_bits else None
    return SimpleNamespace(meta=meta, anchor_q=anchor_q.to(torch.int16), mask_bytes=mask_bytes, mask_lens=mask_pending[2],
                           hyper_bytes=hyper_bytes, hyper_lens=hyper_pending[2], levels=levels, valid=sel,
                           quantised=dict(anchor=anchor, hyper=hyper_q, feat=feat_q, scaling=scaling_q, offsets=offsets_q,
                                          masks=masks), estimated_bits=est, plan=plan)


def encoded_bits(enc):
    """Size of every part of the encoding in bits (payload + per-chunk side information)."""
    side = lambda lens: 16 * lens.numel()    # chunk lengths are stored as 16-bit integers
    bits = dict(anchor=16 * enc.anchor_q.numel(), masks=8 * enc.mask_bytes.numel() + side(enc.mask_lens) + 32,
                hyper=8 * enc.hyper_bytes.numel() + side(enc.hyper_lens) + 32, feat=0, scaling=0, offsets=0)
    for lv in enc.levels:
        for name, st in lv.streams.items():
            bits[name] += 8 * st.bytes.numel() + 16 * st.lens.numel() + 64   # 16-bit chunk lengths + the stream's (min, max)
    bits["total"] = sum(bits.values())
    return bits


@torch.no_grad()
def decode_model(pc, meta, anchor_q, mask_bytes, mask_lens, hyper_bytes, hyper_lens, levels, rank=0, world=1):
    """Inverse of encode_model.  `pc` supplies the MLPs / entropy bottleneck (mlp.pt) and receives bounds and
    level scales from `meta`.  Returns dict(anchor, hyper, feat, offsets [N,10,3], scaling, masks [N,10,1]).
    world > 1: `levels` are the streams rank `rank` encoded; only that shard's rows of feat / scaling / offsets are
    filled (the rest stays zero), so the SUM over the ranks (one all-reduce) is the decoded model."""
    L = _lib.lib()
    if meta.get("version") != STREAM_VERSION:
        raise _lib.CgsError(f"bitstream version {meta.get('version')} cannot be decoded by this library (expects "
                            f"{STREAM_VERSION}): the coder's termination, chunking and CDF changed between versions")
    dev = pc.latent_codec.quantiles.device
    N, K, chunk_rows = meta["N"], pc.n_offsets, meta["chunk_rows"]
    pc.x_bound_min, pc.x_bound_max = meta["x_bound_min"].to(dev), meta["x_bound_max"].to(dev)
    pc.level_scale = list(meta["level_scale"])
    pc.voxel_size = meta["voxel_size"]
    anchor = dequantize_anchor(anchor_q.to(dev).to(torch.int32) & 0xffff, pc.x_bound_min, pc.x_bound_max).contiguous()

    masks = _table_decode(mask_bytes.to(dev), mask_lens.to(dev), N, K, mask_table(meta["prob_masks"]),
                          chunk_rows * TABLE_CHUNK_MULT).float().contiguous()

    hmin, hmax = meta["hyper_min"], meta["hyper_max"]
    median = pc.latent_codec.quantiles[:, 0, 1].detach()
    hsym = _table_decode(hyper_bytes.to(dev), hyper_lens.to(dev), N, median.numel(), _hyper_tables(pc, hmin, hmax),
                         chunk_rows * TABLE_CHUNK_MULT)
    hyper_q = ((hsym.to(torch.int32) + hmin).float() + median.view(1, -1)).contiguous()
    hyper_ctx = hyper_q * 0 if getattr(pc, "disable_hyper", False) else hyper_q

    sizes, plan = _plan_for(pc, anchor, _content_key(anchor_q.to(dev), pc), rank, world)
    if sizes != list(meta["N_levels"]):
        raise _lib.CgsError("decode: the level division of the decoded anchors differs from the encoder's")
    feat_q = torch.zeros((N, 50), dtype=torch.float32, device=dev)
    scaling_q = torch.zeros((N, 6), dtype=torch.float32, device=dev)
    offsets_q = torch.zeros((N, 3 * K), dtype=torch.float32, device=dev)
    sums = torch.zeros(16, dtype=torch.float64, device=dev)
    terr = torch.zeros(1, dtype=torch.int32, device=dev)
    means = tuple(meta["means"])
    for li, (lv, coded) in enumerate(zip(plan.levels, levels)):
        if lv.n == 0:
            continue
        rows3 = (ctypes.c_int * 3)(*(getattr(coded, "chunk_rows", None) or level_chunk_rows(lv.n, chunk_rows, False)))
        params = _level_params(pc, lv, anchor, hyper_ctx, None, None, None, None, feat_q, scaling_q, offsets_q,
                               sums[4 * li:4 * li + 4], terr, means, True)
        st = [coded.streams[name] for name, _ in ATTRS]
        lens = torch.cat([t.lens.to(dev).to(torch.int32) for t in st])
        minmax = torch.cat([t.minmax.to(dev).to(torch.int32) for t in st])
        data = [t.bytes.to(dev).contiguous() for t in st]   # named: the pointers must outlive the launch
        off = _offsets(lens)
        p = _lib.ptr
        _lib.check(L.cgs_codec_gauss_level_decode(p(lv.orig), lv.n, rows3, p(params), p(masks), p(data[0]), p(data[1]),
                                                  p(data[2]), p(off), p(lens), p(minmax), p(feat_q), p(scaling_q),
                                                  p(offsets_q), _lib.stream_ptr()), "cgs_codec_gauss_level_decode")
    if int(terr.item()):
        raise _lib.CgsError("cgs_context_level_umma_forward_ex: a tensor-core completion barrier timed out")
    return dict(anchor=anchor, hyper=hyper_q, feat=feat_q, offsets=offsets_q.view(N, K, 3), scaling=scaling_q,
                masks=masks.view(N, K, 1))


# ----------------------------------------------------------------------------- directory layout of the reference

def _mlp_state(pc):
    return {"opacity_mlp": pc.mlp_opacity.state_dict(), "cov_mlp": pc.mlp_cov.state_dict(),
            "color_mlp": pc.mlp_color.state_dict(), "grid_mlp": pc.mlp_grid.state_dict(),
            "latent_codec": pc.latent_codec.state_dict()}


def conduct_encoding(pc, pre_path_name, chunk_rows=CHUNK_ROWS):
    """scene/gaussian_model.py:1005-1300: writes anchor.npy, masks.b, hyper.b, {feat,scaling,offsets}{level}.b,
    meta.b, mlp.pt under `pre_path_name`; returns the reference's size summary string."""
    os.makedirs(pre_path_name, exist_ok=True)
    caps = [int(_lib.lib().cgs_codec_gauss_stream_capacity(a, chunk_rows * ATTR_CHUNK_MULT[a])) for a in range(len(ATTRS))]
    if max(caps) > 65535 or chunk_rows * TABLE_CHUNK_MULT * 12 * 2 + 16 > 65535:
        raise ValueError("chunk_rows too large for the 16-bit chunk lengths of the directory format")
    enc = encode_model(pc, chunk_rows, estimate_bits=False)
    all_lens = [enc.mask_lens, enc.hyper_lens] + [st.lens for lv in enc.levels for st in lv.streams.values()]
    if max(int(l.max()) if l.numel() else 0 for l in all_lens) > 65535:
        raise _lib.CgsError("a chunk is longer than 65535 bytes: it does not fit the 16-bit length of the directory format")
    np.save(os.path.join(pre_path_name, "anchor.npy"), enc.anchor_q.cpu().numpy().view(np.uint16))
    wr = lambda name, t: t.cpu().numpy().tofile(os.path.join(pre_path_name, name))
    wr("masks.b", enc.mask_bytes)
    wr("hyper.b", enc.hyper_bytes)
    side = dict(mask_lens=enc.mask_lens.cpu().to(torch.int16), hyper_lens=enc.hyper_lens.cpu().to(torch.int16), levels=[])
    for lv in enc.levels:
        ent = dict(level=lv.level, n=lv.n, streams={}, chunk_rows=list(lv.chunk_rows))
        for name, st in lv.streams.items():
            wr(f"{name}{lv.level}.b", st.bytes)
            ent["streams"][name] = dict(lens=st.lens.cpu().to(torch.int16), minmax=st.minmax.cpu())
        side["levels"].append(ent)
    torch.save(dict(meta=enc.meta, side=side), os.path.join(pre_path_name, "meta.b"))
    torch.save(_mlp_state(pc), os.path.join(pre_path_name, "mlp.pt"))
    bits = encoded_bits(enc)
    mb = 8 * 1024 * 1024
    return enc, "\nEncoded sizes in MB: " + ", ".join(f"{k} {round(v / mb, 4)}" for k, v in bits.items())


def conduct_decoding(pc, pre_path_name):
    """scene/gaussian_model.py:1302-1538: reads the directory written by conduct_encoding and replaces the model's
    parameters with the decoded values (`decoded_version = True`)."""
    dev = pc.latent_codec.quantiles.device
    state = torch.load(os.path.join(pre_path_name, "mlp.pt"), map_location=dev)
    pc.mlp_opacity.load_state_dict(state["opacity_mlp"]); pc.mlp_cov.load_state_dict(state["cov_mlp"])
    pc.mlp_color.load_state_dict(state["color_mlp"]); pc.mlp_grid.load_state_dict(state["grid_mlp"])
    pc.latent_codec.load_state_dict(state["latent_codec"])
    blob = torch.load(os.path.join(pre_path_name, "meta.b"), weights_only=False)
    meta, side = blob["meta"], blob["side"]
    rd = lambda name: torch.from_numpy(np.fromfile(os.path.join(pre_path_name, name), dtype=np.uint8)).to(dev)
    u16 = lambda t: t.to(torch.int32) & 0xffff
    anchor_q = torch.from_numpy(np.load(os.path.join(pre_path_name, "anchor.npy")).astype(np.int32)).to(dev)
    levels = []
    for ent in side["levels"]:
        lv = SimpleNamespace(level=ent["level"], n=ent["n"], streams={}, chunk_rows=ent.get("chunk_rows"))
        for name, st in ent["streams"].items():
            lv.streams[name] = SimpleNamespace(bytes=rd(f"{name}{ent['level']}.b"), lens=u16(st["lens"]), minmax=st["minmax"])
        levels.append(lv)
    out = decode_model(pc, meta, anchor_q, rd("masks.b"), u16(side["mask_lens"]), rd("hyper.b"), u16(side["hyper_lens"]),
                       levels)
    if hasattr(pc, "replace_with_decoded"):
        pc.replace_with_decoded(out["anchor"], out["hyper"], out["feat"], out["offsets"], out["scaling"], out["masks"])
    else:
        # bound on the reference's own GaussianModel (INTEGRATION.md hook 4): the parameter replacement of
        # scene/gaussian_model.py:1503-1533, padded back to the N_full rows the other per-anchor tensors keep
        P = torch.nn.Parameter
        n_full, n = int(meta["N_total"]), out["anchor"].shape[0]

        def full(t):
            buf = torch.zeros((n_full,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev)
            buf[:n] = t
            return P(buf)
        pc._hyper_latent, pc._anchor_feat, pc._offset = full(out["hyper"]), full(out["feat"]), full(out["offsets"])
        pc.decoded_version = True
        pc._anchor, pc._scaling, pc._mask = full(out["anchor"]), full(out["scaling"]), full(out["masks"])
    return out


# ----------------------------------------------------------------------------- anchors sharded over the GPUs

def encode_model_sharded(pc, group=None, chunk_rows=CHUNK_ROWS):
    """encode_model on this rank's shard of the level plan (one process per GPU, torch.distributed).  Returns
    (enc, total_bits): the level streams of this rank and the size of the WHOLE encoding (level streams summed over
    the ranks by one scalar all-reduce; anchors / masks / hyper counted once)."""
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if on else (0, 1)
    enc = encode_model(pc, chunk_rows, rank, world)
    bits = encoded_bits(enc)
    level_bits = torch.tensor([bits["feat"] + bits["scaling"] + bits["offsets"]], dtype=torch.float64,
                              device=enc.anchor_q.device)
    if world > 1:
        dist.all_reduce(level_bits, op=dist.ReduceOp.SUM, group=group)
    return enc, int(level_bits.item()) + bits["anchor"] + bits["masks"] + bits["hyper"]


def decode_model_sharded(pc, enc, group=None):
    """decode_model on this rank's streams, then ONE all-reduce (sum) of the three attribute arrays assembles the
    decoded model on every rank."""
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if on else (0, 1)
    out = decode_model(pc, enc.meta, enc.anchor_q, enc.mask_bytes, enc.mask_lens, enc.hyper_bytes, enc.hyper_lens,
                       enc.levels, rank, world)
    if world > 1:
        flat = torch.cat([out["feat"].reshape(-1), out["scaling"].reshape(-1), out["offsets"].reshape(-1)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        nf, ns = out["feat"].numel(), out["scaling"].numel()
        out["feat"] = flat[:nf].view_as(out["feat"])
        out["scaling"] = flat[nf:nf + ns].view_as(out["scaling"])
        out["offsets"] = flat[nf + ns:].view_as(out["offsets"])
    return out
