"""Bitstream codec parity: encode -> decode returns the encoder's quantised tensors bit for bit,
the streams are byte-identical to the CPU restatement of the range coder (table streams), their size matches the
estimated bits, and a model decoded from disk renders exactly what the in-memory decoded model renders."""
import numpy as np
import pytest
import torch

from contextgs_b200 import codec, synthetic
from contextgs_b200.context_model import multi_scale_generating
from contextgs_b200.gaussian_model import GaussianModel
from contextgs_b200.renderer import prefilter_voxel, render
from oracle import codec_ref

pytestmark = pytest.mark.gpu


def _model(N=6000, seed=11, kind="chair"):
    scene = synthetic.make_scene(kind, N, seed=seed)
    torch.manual_seed(6)
    pc = GaussianModel.from_tensors(scene, device="cuda")
    with torch.no_grad():   # break the symmetric EntropyBottleneck init, kill a few anchors entirely
        g = torch.Generator().manual_seed(9)
        for plist in (pc.latent_codec.matrices, pc.latent_codec.biases, pc.latent_codec.factors):
            for p in plist:
                p.add_((torch.randn(p.shape, generator=g) * 0.3).cuda())
        pc._mask[::17] = -8.0
    pc.eval()
    return pc


@pytest.mark.parametrize("N,chunk,adaptive", [(6000, 32, False), (1500, 64, False), (300, 1000, False), (6000, 8, True)])
def test_encode_decode_round_trip_is_bit_exact(N, chunk, adaptive):
    pc = _model(N)
    enc = codec.encode_model(pc, chunk_rows=chunk, adaptive_chunks=adaptive)
    rows = [lv.chunk_rows[0] for lv in enc.levels]
    assert rows == ([codec.MIN_CHUNK_ROWS] * 3 if adaptive else [chunk] * 3)   # small levels: short chunks
    q = enc.quantised
    # the coded values are what the scoring pass quantises (same kernels): feat_q etc. of multi_scale_generating
    sel = pc.get_mask_anchor
    assert int(sel.sum()) == q["feat"].shape[0] < pc._anchor.shape[0]
    fresh = GaussianModel(device="cuda")
    fresh.load_state_dict({k: v for k, v in pc.state_dict().items() if not k.startswith("_")}, strict=False)
    fresh.eval()
    out = codec.decode_model(fresh, enc.meta, enc.anchor_q, enc.mask_bytes, enc.mask_lens, enc.hyper_bytes, enc.hyper_lens,
                             enc.levels)
    K = pc.n_offsets
    assert torch.equal(out["anchor"], q["anchor"])
    assert torch.equal(out["masks"].view(-1, K), q["masks"])
    assert torch.equal(out["hyper"], q["hyper"])
    assert torch.equal(out["feat"], q["feat"])
    assert torch.equal(out["scaling"], q["scaling"])
    m3 = q["masks"].repeat_interleave(3, dim=1)
    assert torch.equal(out["offsets"].reshape(-1, 3 * K), q["offsets"] * m3)
    assert float(out["feat"].abs().max()) > 0 and float(out["offsets"].abs().max()) > 0


def test_decoder_rejects_streams_of_another_format_version():
    from contextgs_b200 import _lib
    pc = _model(300)
    enc = codec.encode_model(pc, estimate_bits=False)
    assert enc.meta["version"] == codec.STREAM_VERSION and enc.estimated_bits is None
    fresh = GaussianModel(device="cuda")
    fresh.load_state_dict({k: v for k, v in pc.state_dict().items() if not k.startswith("_")}, strict=False)
    with pytest.raises(_lib.CgsError, match="bitstream version"):
        codec.decode_model(fresh, dict(enc.meta, version=1), enc.anchor_q, enc.mask_bytes, enc.mask_lens, enc.hyper_bytes,
                           enc.hyper_lens, enc.levels)


def test_quantised_values_equal_the_scoring_pass_and_sizes_match_estimates():
    pc = _model(6000)
    enc = codec.encode_model(pc)
    sel = pc.get_mask_anchor
    idx = torch.nonzero(sel)[:, 0]
    t = lambda x: x.detach().index_select(0, idx)
    a = enc.quantised["anchor"]
    (res, det) = multi_scale_generating(pc, a, t(pc._hyper_latent), t(pc._anchor_feat), t(pc._offset), t(pc.get_scaling),
                                        binary_grid_masks=t(pc.get_mask), predict_bpp=True, return_sum_bits=True,
                                        return_details=True)
    assert torch.equal(det["feat_q"], enc.quantised["feat"]) and torch.equal(det["scaling_q"], enc.quantised["scaling"])
    # the encoder quantises the hyper latents inline; the EntropyBottleneck kernel must agree bit for bit
    assert torch.equal(pc.latent_codec(t(pc._hyper_latent).contiguous(), training=False)[0], enc.quantised["hyper"])
    assert torch.allclose(a, t(pc.get_anchor), atol=0, rtol=0)
    bits = codec.encoded_bits(enc)
    est = dict(hyper=res[1], feat=res[2], scaling=res[3], offsets=res[4], masks=res[5])
    chunks = dict(hyper=enc.hyper_lens.numel(), masks=enc.mask_lens.numel(),
                  **{k: sum(lv.streams[k].lens.numel() for lv in enc.levels if lv.streams) for k in ("feat", "scaling", "offsets")})
    for k in ("feat", "scaling", "offsets", "hyper", "masks"):
        # 16-bit frequencies floor every probability at 2^-16 (the estimate floors at 1e-6 ~ 2^-20), so the real
        # stream may be SHORTER than the estimate where the model is badly wrong; it must never be much longer: per chunk a
        # 16-bit length, the coder's termination byte and the rounding up to whole bytes, per stream its 64 bits of bounds
        assert bits[k] < 1.03 * est[k] + 32 * chunks[k] + 256, (k, bits[k], est[k], chunks[k])
    assert bits["anchor"] == 48 * a.shape[0]


def test_alphabets_from_the_level_kernel_equal_the_stand_alone_pass():
    """The encoder takes each stream's (min, max) symbol from the level kernel's epilogue (`symbol_minmax` of
    cgs_context_level_umma_forward_ex); cgs_codec_gauss_level_minmax recomputes them from the quantised tensors, and torch
    once more from the definition rint(value / Q) over the coded values."""
    from contextgs_b200 import _lib
    pc = _model(3000)
    enc = codec.encode_model(pc)
    q = enc.quantised
    dev = q["feat"].device
    means = codec.global_means(pc)
    sums = torch.zeros(16, dtype=torch.float64, device=dev)
    terr = torch.zeros(1, dtype=torch.int32, device=dev)
    seen = 0
    for li, (lv, coded) in enumerate(zip(enc.plan.levels, enc.levels)):
        if lv.n == 0:
            continue
        fq, sq, oq = q["feat"].clone(), q["scaling"].clone(), q["offsets"].clone()
        params = codec._level_params(pc, lv, q["anchor"], q["hyper"] * (0 if pc.disable_hyper else 1), None, None, None, None,
                                     fq, sq, oq, sums[4 * li:4 * li + 4], terr, means, True)
        mm = torch.empty(6, dtype=torch.int32, device=dev)
        p = _lib.ptr
        _lib.check(_lib.lib().cgs_codec_gauss_level_minmax(p(lv.orig), lv.n, p(params), p(q["masks"]), p(q["feat"]),
                                                           p(q["scaling"]), p(q["offsets"]), p(mm), _lib.stream_ptr()),
                   "cgs_codec_gauss_level_minmax")
        fused = torch.cat([coded.streams[name].minmax for name, _ in codec.ATTRS])
        assert torch.equal(mm, fused), (li, mm.tolist(), fused.tolist())
        o = lv.orig.long()
        for attr, (name, dim) in enumerate(codec.ATTRS):
            sym = torch.round((q["feat"], q["scaling"], q["offsets"])[attr][o] / params[:, 172 + attr:173 + attr])
            if attr == 2:
                sym = sym[q["masks"][o].repeat_interleave(3, dim=1) != 0]
            assert [int(sym.min()), int(sym.max())] == mm[2 * attr:2 * attr + 2].tolist(), (li, name)
            seen += 1
    assert seen >= 6


def test_table_streams_are_byte_identical_to_the_cpu_range_coder():
    pc = _model(700)
    enc = codec.encode_model(pc, chunk_rows=32, adaptive_chunks=False)
    q = enc.quantised
    tb = codec.mask_table(enc.meta["prob_masks"])[0].tolist()
    rows = 32 * codec.TABLE_CHUNK_MULT
    data = enc.mask_bytes.cpu().numpy().tobytes()
    lens = enc.mask_lens.tolist()
    masks = q["masks"].to(torch.int64).cpu()
    off = 0
    for c, n in enumerate(lens):
        syms = masks[c * rows:(c + 1) * rows].reshape(-1).tolist()
        chunk = data[off:off + n]
        assert codec_ref.encode(syms, lambda i: 0, [tb]) == chunk
        assert codec_ref.decode(chunk, len(syms), lambda i: 0, [tb]) == syms
        off += n
    assert off == len(data)
    # hyper stream of the first chunk: one table per channel
    median = pc.latent_codec.quantiles[:, 0, 1].detach()
    hsym = (torch.round(q["hyper"] - median.view(1, -1)).to(torch.int64) - enc.meta["hyper_min"]).cpu()
    tabs = codec._hyper_tables(pc, enc.meta["hyper_min"], enc.meta["hyper_max"]).tolist()
    n0 = enc.hyper_lens.tolist()[0]
    syms = hsym[:rows].reshape(-1).tolist()
    assert codec_ref.encode(syms, lambda i: i % 12, tabs) == enc.hyper_bytes.cpu().numpy().tobytes()[:n0]


def test_gaussian_streams_decode_with_an_independent_table_decoder():
    """Independent check of the Gaussian-CDF streams (the counterpart of utils/encodings.py:119-144, where the reference
    materialises a dense [symbols, alphabet] CDF tensor per chunk and hands it to torchac): the bytes the GPU coder wrote are
    decoded by the plain-Python range decoder of oracle/codec_ref.py from DENSE per-symbol cumulative-frequency tables
    built with torch ops (C(i) = min(rn(Phi(((smin + i) - 1/2) Q) (65536 - L)), 65536 - L) + i, the formula of
    DESIGN.md section 4, Phi = the coder's tabulated normal CDF, itself checked against math.erfc here) -- neither the
    GPU's evaluation of C nor the search of the GPU decoder is involved.  The encoder side is checked too: re-encoding the
    decoded symbols on the CPU reproduces the GPU's bytes."""
    import math
    T_np, z0, inv_h = codec.phi_table()
    want = np.array([0.5 * math.erfc(-(-4.75 + 9.5 * j / 4096) / math.sqrt(2.0)) for j in range(4097)]).astype(np.float32)
    want[0], want[-1] = 0.0, 1.0
    assert np.array_equal(T_np, want) and z0 == -4.75 and inv_h == float(np.float32(4096 / 9.5))
    pc = _model(700)
    enc = codec.encode_model(pc, chunk_rows=4, adaptive_chunks=False)
    q = enc.quantised
    dev = q["feat"].device
    means = codec.global_means(pc)
    sums = torch.zeros(16, dtype=torch.float64, device=dev)
    terr = torch.zeros(1, dtype=torch.int32, device=dev)
    checked = 0
    for li, (lv, coded) in enumerate(zip(enc.plan.levels, enc.levels)):
        if lv.n == 0:
            continue
        fq, sq, oq = q["feat"].clone(), q["scaling"].clone(), q["offsets"].clone()
        params = codec._level_params(pc, lv, q["anchor"], q["hyper"] * (0 if pc.disable_hyper else 1), None, None, None, None,
                                     fq, sq, oq, sums[4 * li:4 * li + 4], terr, means, True)
        for attr, (name, dim) in enumerate(codec.ATTRS):
            st = coded.streams[name]
            smin, smax = (int(v) for v in st.minmax.tolist())
            L = smax - smin + 1
            col0 = (0, 50, 56)[attr]
            rows = 4 * codec.ATTR_CHUNK_MULT[attr]
            values = (q["feat"], q["scaling"], q["offsets"])[attr]
            data = st.bytes.cpu().numpy().tobytes()
            off = 0
            for c, nbytes in enumerate(st.lens.tolist()[:3]):          # the first chunks of every stream
                r0, r1 = c * rows, min((c + 1) * rows, lv.n)
                o = lv.orig[r0:r1].long()
                pr = params[r0:r1]
                Q = pr[:, 172 + attr:173 + attr]
                mean, scale = pr[:, col0:col0 + dim], pr[:, 86 + col0:86 + col0 + dim]
                inv = torch.reciprocal(torch.clamp(scale, min=1e-9))
                x = values[o]
                keep = torch.ones_like(x, dtype=torch.bool) if attr != 2 else \
                    (q["masks"][o].repeat_interleave(3, dim=1) != 0)
                sym = torch.round(x / Q).to(torch.int64) - smin
                # dense tables: boundary i of every symbol position, i = 0 .. L
                i = torch.arange(L + 1, device=dev, dtype=torch.float32).view(1, 1, -1)
                z = ((smin + i) - 0.5) * Q.unsqueeze(-1)
                t = torch.clamp((((z - mean.unsqueeze(-1)) * inv.unsqueeze(-1)) - z0) * inv_h, 0.0, 4096.0)
                j = torch.clamp(t.to(torch.int64), max=4095)
                T = torch.from_numpy(T_np).to(dev)
                phi = T[j] + (t - j.to(torch.float32)) * (T[j + 1] - T[j])
                M = float(65536 - L)
                tab = (torch.clamp(torch.round(phi * M), max=M) + i).to(torch.int64)
                tabs = tab[keep].cpu().tolist()
                syms = sym[keep].cpu().tolist()
                chunk = data[off:off + nbytes]
                off += nbytes
                assert codec_ref.decode(chunk, len(syms), lambda k: k, tabs) == syms, (li, name, c)
                assert codec_ref.encode(syms, lambda k: k, tabs) == chunk, (li, name, c)
                checked += len(syms)
    assert checked > 2000 and int(terr.item()) == 0


def test_directory_round_trip_and_render(tmp_path):
    pc = _model(5000, kind="chair")
    summary = pc.conduct_encoding(str(tmp_path))
    assert "Encoded sizes in MB" in summary
    for f in ("anchor.npy", "masks.b", "hyper.b", "feat0.b", "scaling1.b", "offsets2.b", "meta.b", "mlp.pt"):
        assert (tmp_path / f).exists()
    assert np.load(tmp_path / "anchor.npy").dtype == np.uint16
    dec = GaussianModel(device="cuda")
    dec.conduct_decoding(str(tmp_path))
    dec.eval()
    assert dec.decoded_version
    # reference: the same model decoded in memory through the scoring pass (train.py:301-314 -> replace parameters)
    sel = pc.get_mask_anchor
    idx = torch.nonzero(sel)[:, 0]
    t = lambda x: x.detach().index_select(0, idx)
    a = t(pc.get_anchor)
    fq, sq, oq = multi_scale_generating(pc, a, t(pc._hyper_latent), t(pc._anchor_feat), t(pc._offset), t(pc.get_scaling),
                                        binary_grid_masks=t(pc.get_mask))
    assert torch.equal(dec._anchor_feat, fq) and torch.equal(dec._scaling, sq) and torch.equal(dec._anchor, a)
    assert torch.equal(dec._offset * dec._mask, oq * t(pc.get_mask))
    cams = synthetic.make_cameras("chair", 2, device="cuda", W=320, H=200)
    pipe = type("Pipe", (), {"debug": False})()
    bg = torch.zeros(3, device="cuda")
    ref = GaussianModel(device="cuda")
    ref.load_state_dict({k: v for k, v in pc.state_dict().items() if not k.startswith("_")}, strict=False)
    ref._rotation = torch.nn.Parameter(t(pc._rotation), requires_grad=False)
    ref.replace_with_decoded(a, t(pc._hyper_latent), fq, oq, sq, t(pc.get_mask))
    ref.eval()
    dec._rotation = torch.nn.Parameter(t(pc._rotation), requires_grad=False)
    for cam in cams:
        with torch.no_grad():
            i1 = render(cam, dec, pipe, bg, visible_mask=prefilter_voxel(cam, dec, pipe, bg))["render"]
            i2 = render(cam, ref, pipe, bg, visible_mask=prefilter_voxel(cam, ref, pipe, bg))["render"]
        assert torch.equal(i1, i2) and float(i1.max()) > 0


def test_tiny_model_and_fully_masked_offsets():
    """Ragged edges: fewer anchors than one chunk, levels with a handful of rows, and an offsets stream that is
    empty because every offset mask is 0 (decodes to zeros)."""
    for N, kill_all in ((40, False), (257, True)):
        pc = _model(N)
        if kill_all:
            with torch.no_grad():
                pc._mask.fill_(-8.0)
                pc._mask[::2, 0] = 8.0        # every second anchor keeps exactly one offset (others are pruned anchors)
        enc = codec.encode_model(pc)
        q = enc.quantised
        fresh = GaussianModel(device="cuda")
        fresh.load_state_dict({k: v for k, v in pc.state_dict().items() if not k.startswith("_")}, strict=False)
        out = codec.decode_model(fresh, enc.meta, enc.anchor_q, enc.mask_bytes, enc.mask_lens, enc.hyper_bytes,
                                 enc.hyper_lens, enc.levels)
        K = pc.n_offsets
        assert out["feat"].shape[0] == int(pc.get_mask_anchor.sum()) > 0
        assert torch.equal(out["feat"], q["feat"]) and torch.equal(out["scaling"], q["scaling"])
        assert torch.equal(out["hyper"], q["hyper"]) and torch.equal(out["masks"].view(-1, K), q["masks"])
        m3 = q["masks"].repeat_interleave(3, dim=1)
        assert torch.equal(out["offsets"].reshape(-1, 3 * K), q["offsets"] * m3)


def test_sharded_encode_decode_fake_world():
    """Anchors sharded over 4 ranks (BASELINE configs[3]) emulated on one GPU: every rank encodes / decodes only its
    dependency-closed shard; the shards' decoded rows are disjoint, their sum is the unsharded result, and the level
    streams together are about as long as the unsharded ones."""
    pc = _model(6000)
    full = codec.encode_model(pc)
    q = full.quantised
    world = 4
    K = pc.n_offsets
    acc = {k: torch.zeros_like(q[k]) for k in ("feat", "scaling", "offsets")}
    touched = torch.zeros(q["feat"].shape[0], device="cuda")
    level_bits = 0
    for rank in range(world):
        enc = codec.encode_model(pc, rank=rank, world=world)
        b = codec.encoded_bits(enc)
        level_bits += b["feat"] + b["scaling"] + b["offsets"]
        assert torch.equal(enc.hyper_bytes, full.hyper_bytes) and torch.equal(enc.mask_bytes, full.mask_bytes)
        fresh = GaussianModel(device="cuda")
        fresh.load_state_dict({k: v for k, v in pc.state_dict().items() if not k.startswith("_")}, strict=False)
        out = codec.decode_model(fresh, enc.meta, enc.anchor_q, enc.mask_bytes, enc.mask_lens, enc.hyper_bytes,
                                 enc.hyper_lens, enc.levels, rank=rank, world=world)
        touched += (out["feat"].abs().sum(dim=1) > 0).float()
        acc["feat"] += out["feat"]; acc["scaling"] += out["scaling"]; acc["offsets"] += out["offsets"].reshape(-1, 3 * K)
    assert float(touched.max()) <= 1.0
    m3 = q["masks"].repeat_interleave(3, dim=1)
    assert torch.equal(acc["feat"], q["feat"]) and torch.equal(acc["scaling"], q["scaling"])
    assert torch.equal(acc["offsets"], q["offsets"] * m3)
    fb = codec.encoded_bits(full)
    ref_bits = fb["feat"] + fb["scaling"] + fb["offsets"]
    assert abs(level_bits - ref_bits) < 0.02 * ref_bits
