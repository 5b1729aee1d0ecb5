"""Generates tests/golden/*.npz by running the REFERENCE'S OWN Python (read from /root/reference,
never copied) on small seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

What is executed from the reference, unmodified:
  * utils/entropy_models.py  (Entropy_gaussian, Low_bound)
  * utils/encodings.py       (STE_multistep, Quantize_anchor, get_binary_vxl_size)
  * utils/multi_level.py     (torch_unique_with_indices)
  * scene/gaussian_model.py  from `def multi_scale_generating` (line 1541) to EOF, exec'd because the
    module's top-level imports (compressai, plyfile, simple_knn, torch_scatter) are not installable
  * gaussian_renderer/__init__.py `generate_neural_gaussians` (lines 25-150), exec'd for the same reason
Stand-ins (not available anywhere in this container): `torchac` (stub module, never called on these
paths) and compressai's `EntropyBottleneck` (oracle.entropy_ref.EntropyBottleneckRef -- that one
piece stays "parity unpinned").
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)

from contextgs_b200 import synthetic  # noqa: E402
from oracle import entropy_ref  # noqa: E402


def load_reference():
    sys.modules.setdefault("torchac", types.ModuleType("torchac"))
    import utils.encodings as enc
    import utils.entropy_models as em
    import utils.multi_level as ml
    from einops import repeat

    class EntropyBottleneck:  # name must match the reference's isinstance assert (:1554)
        def __init__(self, impl):
            self.impl = impl

        def __call__(self, x, training=False):
            return self.impl.forward(x, training)

    ns = dict(torch=torch, repeat=repeat, EntropyBottleneck=EntropyBottleneck, STE_multistep=enc.STE_multistep,
              torch_unique_with_indices=ml.torch_unique_with_indices, get_binary_vxl_size=enc.get_binary_vxl_size)
    src = open(os.path.join(REF, "scene/gaussian_model.py")).read()
    start = src.index("def multi_scale_generating(")
    exec(compile(src[start:], "reference:scene/gaussian_model.py[1541:]", "exec"), ns)
    gsrc = open(os.path.join(REF, "gaussian_renderer/__init__.py")).read()
    g0 = gsrc.index("def generate_neural_gaussians(")
    g1 = gsrc.index("def render(")
    gns = dict(torch=torch, repeat=repeat, GaussianModel=object, multi_scale_generating=ns["multi_scale_generating"])
    exec(compile(gsrc[g0:g1], "reference:gaussian_renderer/__init__.py[25:150]", "exec"), gns)
    return ns, gns, enc, em, ml, EntropyBottleneck


def seq(w, last=None):
    """nn.Sequential(Linear, ReLU, Linear[, act]) holding the oracle's weights."""
    W1, b1, W2, b2 = w
    l1, l2 = nn.Linear(W1.shape[1], W1.shape[0]), nn.Linear(W2.shape[1], W2.shape[0])
    with torch.no_grad():
        l1.weight.copy_(W1); l1.bias.copy_(b1); l2.weight.copy_(W2); l2.bias.copy_(b2)
    mods = [l1, nn.ReLU(True), l2] + ([last] if last is not None else [])
    return nn.Sequential(*mods)


def reference_pc(pc, em, EB):
    """Duck-typed `pc` carrying what the reference functions read from GaussianModel."""
    r = types.SimpleNamespace()
    for k in ("feat_dim", "n_offsets", "voxel_size", "level_num", "target_ratio", "level_scale", "disable_hyper",
              "adaptQ_per_channel", "decoded_version", "x_bound_min", "x_bound_max", "_anchor_feat", "_offset",
              "get_scaling", "get_anchor", "get_mask", "get_mask_anchor", "_hyper_latent"):
        setattr(r, k, getattr(pc, k))
    r.latent_codec = EB(pc.latent_codec)
    r.get_grid_mlp = [seq(w) for w in pc.mlps["grid"]]
    r.entropy_gaussian = em.Entropy_gaussian(Q=1)
    r.get_opacity_mlp = seq(pc.mlps["opacity"], nn.Tanh())
    r.get_color_mlp = seq(pc.mlps["color"], nn.Sigmoid())
    r.get_cov_mlp = seq(pc.mlps["cov"])
    r.rotation_activation = torch.nn.functional.normalize
    r.update_anchor_bound = lambda: None
    return r


def npy(d):
    out = {}
    for k, v in d.items():
        if torch.is_tensor(v):
            out[k] = v.detach().cpu().numpy()
        elif isinstance(v, (list, tuple)) and v and torch.is_tensor(v[0]):
            for i, t in enumerate(v):
                out[f"{k}.{i}"] = t.detach().cpu().numpy()
        else:
            out[k] = np.asarray(v)
    return out


def model_inputs(scene, pc):
    d = dict(anchor=scene["anchor"], feat=scene["feat"], hyper=scene["hyper"], offset=scene["offset"],
             scaling=scene["scaling"], mask=scene["mask"], voxel_size=scene["voxel_size"])
    for name in ("opacity", "cov", "color"):
        d[f"mlp_{name}"] = pc.mlps[name]
    for i, w in enumerate(pc.mlps["grid"]):
        d[f"mlp_grid{i}"] = w
    eb = pc.latent_codec
    d["eb_matrices"], d["eb_biases"], d["eb_factors"], d["eb_quantiles"] = eb.matrices, eb.biases, eb.factors, eb.quantiles
    return d


def main():
    torch.set_num_threads(1)
    ns, gns, enc, em, ml, EB = load_reference()
    msg = ns["multi_scale_generating"]
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)  # the reference writes data_for_vis.pt / bit.pt into the CWD (:1681-1682)

    # ---- small pieces ------------------------------------------------------------------
    g = torch.Generator().manual_seed(123)
    x = torch.randn(257, 50, generator=g) * 3
    Q = torch.rand(257, 1, generator=g) * 1.5 + 0.05
    mean = torch.randn(257, 50, generator=g)
    scale = torch.rand(257, 50, generator=g) * 2 - 0.2  # includes values below the 1e-9 clamp
    pieces = dict(x=x, Q=Q, mean=mean, scale=scale)
    pieces["ste"] = enc.STE_multistep.apply(x, Q)
    pieces["ste_big"] = enc.STE_multistep.apply(x * 1e5, Q)
    xq = enc.STE_multistep.apply(x, Q)
    pieces["bits"] = em.Entropy_gaussian(Q=1).forward(xq, mean, scale, Q, x.mean())
    lk = torch.rand(64, 7, generator=g) * 2e-6
    gg = torch.randn(64, 7, generator=g)
    pieces["lb_x"], pieces["lb_g"] = lk, gg
    grad1 = gg.clone(); grad1[lk < 1e-6] = 0  # Low_bound.backward body (:149-156) minus its .cuda() call
    pieces["lb_out"] = grad1 * torch.Tensor(np.logical_or(lk.numpy() >= 1e-6, gg.numpy() < 0.0) + 0.0)
    anc = torch.randn(300, 3, generator=g) * 4
    mn, mx = anc.min(0, keepdim=True)[0] * 1.2, anc.max(0, keepdim=True)[0] * 1.2
    aq, qv = enc.Quantize_anchor.apply(anc, mn, mx)
    pieces.update(anc=anc, anc_min=mn, anc_max=mx, anc_q=aq, anc_qv=qv)
    bm = (torch.rand(300, 10, 1, generator=g) < 0.7).float()
    Pg, ttl_bit, _, _ = enc.get_binary_vxl_size(bm)
    pieces.update(bm=bm, bm_Pg=Pg, bm_bits=ttl_bit)
    rows = torch.round(torch.randn(500, 3, generator=g) * 2)
    u, inv, idx, cnt = ml.torch_unique_with_indices(rows)
    pieces.update(rows=rows, rows_unique=u, rows_inverse=inv, rows_first=idx)
    np.savez_compressed(os.path.join(HERE, "pieces.npz"), **npy(pieces))

    # ---- context model -----------------------------------------------------------------
    N = 1200
    scene = synthetic.make_scene("train", N, seed=3)
    pc = entropy_ref.make_model(scene)
    rpc = reference_pc(pc, em, EB)
    out = model_inputs(scene, pc)
    anchor, maskb = pc.get_anchor, pc.get_mask_anchor
    # level_scale via the reference's own find_divide_scale (cached on pc, :1559)
    rpc.level_scale = ns["find_divide_scale"](rpc, anchor[maskb], pc.target_ratio, pc.level_num)
    out["level_scale"] = np.asarray(rpc.level_scale, np.float64)
    la, inv_l, map_l, _ = ns["divide_levels"](rpc, anchor, maskb)
    out.update(div_inverse=inv_l, div_first=map_l, div_anchor=la)
    with torch.no_grad():
        # (a) non-decoded inference path, gaussian_renderer/__init__.py:93
        fq, sq, oq = msg(rpc, anchor, pc._hyper_latent, pc._anchor_feat, pc._offset, pc.get_scaling, pc.get_mask,
                         maskb, predict_bpp=False, training=False)
        out.update(eval_feat_q=fq, eval_scaling_q=sq, eval_offsets_q=oq)
        # (b) estimate_final_bits, scene/gaussian_model.py:981-992
        sel = maskb
        rpc2 = reference_pc(pc, em, EB)
        rpc2.level_scale = list(rpc.level_scale)  # cached from training in the real flow (:1559)
        sums = msg(rpc2, anchor[sel], pc._hyper_latent[sel], pc._anchor_feat[sel], pc._offset[sel],
                   pc.get_scaling[sel], binary_grid_masks=pc.get_mask[sel], predict_bpp=True, return_sum_bits=True)
        out["sum_bits"] = np.asarray(sums, np.float64)
        out["sum_level_scale"] = np.asarray(rpc2.level_scale, np.float64)
        # (c) training step > 10000, gaussian_renderer/__init__.py:73
        torch.manual_seed(7)
        res = msg(rpc, anchor, pc._hyper_latent, pc._anchor_feat, pc._offset, pc.get_scaling, pc.get_mask, maskb,
                  predict_bpp=True, training=True)
        out.update(train_feat_q=res[0], train_scaling_q=res[1], train_offsets_q=res[2],
                   train_bits=torch.stack([res[3], res[4], res[5], res[6]]))
        lb = res[7]
        out["train_level_bpp"] = np.asarray([lb[0], lb[1]] + [v for p in lb[2:] for v in p], np.float64)
    np.savez_compressed(os.path.join(HERE, "context_model.npz"), **npy(out))

    # ---- generate_neural_gaussians -------------------------------------------------------
    cam = synthetic.make_cameras("train", 3)[1]
    gen = gns["generate_neural_gaussians"]
    g = torch.Generator().manual_seed(77)
    visible = torch.rand(N, generator=g) < 0.6
    gout = dict(visible=visible, camera_center=cam.camera_center)
    rpc.get_color_mlp.train()
    with torch.no_grad():
        r = gen(cam, rpc, visible, is_training=True, step=100)
    for k, v in zip(["xyz", "color", "opacity", "scaling", "rot", "neural_opacity", "mask"], r[:7]):
        gout["train_" + k] = v
    rpc.decoded_version = True  # published-FPS path (:103-104): context model skipped
    rpc.get_color_mlp.eval()
    with torch.no_grad():
        r = gen(cam, rpc, visible, is_training=False)
    for k, v in zip(["xyz", "color", "opacity", "scaling", "rot"], r[:5]):
        gout["eval_" + k] = v
    np.savez_compressed(os.path.join(HERE, "neural_gaussians.npz"), **npy(gout))
    os.chdir(cwd)
    for f in ("pieces.npz", "context_model.npz", "neural_gaussians.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
