"""Shared test helpers: load the committed golden fixtures and rebuild the oracle model from them."""
import os

import numpy as np
import torch

from oracle import entropy_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def fixture_model(gold):
    """Oracle `pc` with exactly the inputs/weights stored in context_model.npz."""
    scene = dict(anchor=T(gold["anchor"]), feat=T(gold["feat"]), hyper=T(gold["hyper"]), offset=T(gold["offset"]),
                 scaling=T(gold["scaling"]), mask=T(gold["mask"]), voxel_size=float(gold["voxel_size"]))
    pc = entropy_ref.make_model(scene)
    for name in ("opacity", "cov", "color"):
        pc.mlps[name] = [T(gold[f"mlp_{name}.{i}"]) for i in range(4)]
    pc.mlps["grid"] = [[T(gold[f"mlp_grid{l}.{i}"]) for i in range(4)] for l in range(3)]
    eb = pc.latent_codec
    eb.matrices = [T(gold[f"eb_matrices.{i}"]) for i in range(5)]
    eb.biases = [T(gold[f"eb_biases.{i}"]) for i in range(5)]
    eb.factors = [T(gold[f"eb_factors.{i}"]) for i in range(4)]
    eb.quantiles = T(gold["eb_quantiles"])
    pc.level_scale = [float(v) for v in gold["level_scale"]]
    return scene, pc


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def cuda_model(scene, pc, **kw):
    """Product-side GaussianModel on the GPU sharing the oracle model's parameters."""
    from contextgs_b200.gaussian_model import GaussianModel
    m = GaussianModel.from_tensors(scene, pc.mlps, pc.latent_codec, **kw)
    m.level_scale = None if pc.level_scale is None else list(pc.level_scale)
    return m


def reference_noise(N, level_sizes, seed=7, K=10):
    """Noise tensors drawn in the reference's call order (EB noise on the permuted [C,1,N] tensor,
    then per level feat/scaling/offsets, then the 15% choose mask), see oracle.multi_scale_generating."""
    torch.manual_seed(seed)
    eb = torch.empty(12, 1, N).uniform_(-0.5, 0.5)
    levels = []
    for n in level_sizes:
        f = torch.empty(n, 50).uniform_(-0.5, 0.5)
        s = torch.empty(n, 6).uniform_(-0.5, 0.5)
        o = torch.empty(n, K, 3).uniform_(-0.5, 0.5)
        levels.append(torch.cat([f, s, o.reshape(n, 3 * K)], dim=1))
    choose = torch.rand(N) <= 0.15
    return dict(eb=eb.reshape(12, N).t().contiguous(), levels=levels, choose=choose)


def rel_l2_rows(a, b, row_tol=1e-3):
    """(fraction of outlier rows, rel-L2 over the other rows).  For gradients that pass through a ReLU: two valid fp32
    evaluations of the hidden layer (cuBLAS / fp32 FMA vs 3xTF32 tensor cores) can put a pre-activation that is zero to
    rounding on different sides of the kink, which changes the gradient of THAT anchor row by O(1/sqrt(units)) while
    every other row agrees to rounding.  Rows are compared one by one; a row is an outlier when its own relative error
    exceeds `row_tol`."""
    a = np.asarray(a, np.float64).reshape(a.shape[0], -1)
    b = np.asarray(b, np.float64).reshape(b.shape[0], -1)
    num = np.linalg.norm(a - b, axis=1)
    den = np.linalg.norm(b, axis=1)
    scale = np.sqrt((den ** 2).mean()) + 1e-30            # rows that are ~0 in both are judged on the global scale
    bad = num > row_tol * np.maximum(den, 1e-3 * scale)
    good = ~bad
    return float(bad.mean()), float(np.linalg.norm((a - b)[good]) / (np.linalg.norm(b[good]) + 1e-30))
