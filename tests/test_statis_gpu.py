"""cgs_training_statis (GaussianModel.training_statis) against the reference's own outputs (golden vectors) and, on a
real training iteration, against the CPU oracle."""
import os
import types

import numpy as np
import pytest
import torch

from contextgs_b200 import synthetic
from contextgs_b200.gaussian_model import GaussianModel
from contextgs_b200.renderer import prefilter_voxel, render
from oracle import statis_ref

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "statis.npz"))
KEYS = ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom")


def test_against_reference_golden():
    N, K = G["it0_vis"].shape[0], 10
    pc = GaussianModel(device="cuda")
    pc._anchor = torch.zeros(N, 3, device="cuda")
    for it in range(3):
        vsp = types.SimpleNamespace(grad=torch.from_numpy(G[f"it{it}_grad"]).cuda())
        pc.training_statis(vsp, torch.from_numpy(G[f"it{it}_opacity"]).cuda(), torch.from_numpy(G[f"it{it}_upd"]).cuda(),
                           torch.from_numpy(G[f"it{it}_keep"]).cuda(), torch.from_numpy(G[f"it{it}_vis"]).cuda())
        for k in KEYS:
            assert np.allclose(getattr(pc, k).cpu().numpy(), G[f"it{it}_{k}"], rtol=1e-6, atol=1e-6), (it, k)


def test_on_a_training_iteration_against_oracle():
    scene = synthetic.make_scene("chair", 8000, seed=2, gaussian_scale=4.0)
    torch.manual_seed(6)
    pc = GaussianModel.from_tensors(scene, device="cuda")
    pc.train()
    cam = synthetic.make_cameras("chair", 1, device="cuda", W=320, H=200)[0]
    pipe = type("Pipe", (), {"debug": False})()
    bg = torch.zeros(3, device="cuda")
    with torch.no_grad():
        vis = prefilter_voxel(cam, pc, pipe, bg)
    out = render(cam, pc, pipe, bg, visible_mask=vis, retain_grad=True, step=100)
    out["render"].mean().backward()
    vsp = out["viewspace_points"]
    assert vsp.grad is not None and vsp.grad.shape[0] == out["radii"].shape[0]
    pc.training_statis(vsp, out["neural_opacity"], out["visibility_filter"], out["selection_mask"], vis)   # train.py:243
    N, K = 8000, 10
    state = dict(opacity_accum=np.zeros((N, 1), np.float32), anchor_demon=np.zeros((N, 1), np.float32),
                 offset_gradient_accum=np.zeros((N * K, 1), np.float32), offset_denom=np.zeros((N * K, 1), np.float32))
    statis_ref.training_statis(state, K, vsp.grad.cpu().numpy(), out["neural_opacity"].detach().cpu().numpy(),
                               out["visibility_filter"].cpu().numpy(), out["selection_mask"].cpu().numpy(), vis.cpu().numpy())
    for k in KEYS:
        assert np.allclose(getattr(pc, k).cpu().numpy(), state[k], rtol=1e-5, atol=1e-9), k
    assert float(pc.offset_denom.sum()) > 0 and float(pc.anchor_demon.sum()) == float(vis.sum())
