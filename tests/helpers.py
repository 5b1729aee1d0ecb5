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
