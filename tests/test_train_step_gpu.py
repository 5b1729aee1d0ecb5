"""End-to-end training iteration through the drop-in surface (train.py:158-211 of the reference):
prefilter_voxel -> render (G1 + rasterizer, + context model after step 10000) -> loss -> backward.
The individual backward kernels are checked against autograd oracles elsewhere; this test checks the
wiring: every parameter the reference trains receives a finite, non-trivial gradient."""
import pytest
import torch

from contextgs_b200 import synthetic
from contextgs_b200.gaussian_model import GaussianModel
from contextgs_b200.renderer import prefilter_voxel, render

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("step", [100, 5000, 20000])
def test_training_iteration_reaches_every_parameter(step):
    N, W, H = 20000, 320, 200
    scene = synthetic.make_scene("chair", N, seed=4, gaussian_scale=4.0)
    torch.manual_seed(6)
    pc = GaussianModel.from_tensors(scene, device="cuda").train()
    cam = synthetic.make_cameras("chair", 3, device="cuda", W=W, H=H)[1]
    pipe = type("Pipe", (), {"debug": False})()
    bg = torch.zeros(3, device="cuda")
    gt = torch.rand(3, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    with torch.no_grad():
        vis = prefilter_voxel(cam, pc, pipe, bg)
    assert 0 < int(vis.sum()) <= N
    out = render(cam, pc, pipe, bg, visible_mask=vis, retain_grad=True, step=step)
    loss = (out["render"] - gt).abs().mean() + 0.01 * out["scaling"].prod(dim=1).mean()
    if step > 10000:
        assert out["bit_per_param"] is not None and float(out["bit_per_param"]) > 0
        loss = loss + 0.004 * out["bit_per_param"] + 5e-4 * torch.mean(torch.sigmoid(pc._mask))
    loss.backward()
    trained = {"_offset": pc._offset, "_mask": pc._mask, "_anchor_feat": pc._anchor_feat, "_scaling": pc._scaling}
    for m in (pc.mlp_opacity, pc.mlp_cov, pc.mlp_color):
        for i, p in enumerate(m.parameters()):
            trained[f"mlp{i}"] = p
    if step > 10000:
        trained["_hyper_latent"] = pc._hyper_latent
        for i, p in enumerate(pc.mlp_grid.parameters()):
            trained[f"grid{i}"] = p
        for i, p in enumerate(pc.latent_codec.parameters()):
            if p is not pc.latent_codec.quantiles:
                trained[f"eb{i}"] = p
    for name, p in trained.items():
        assert p.grad is not None, name
        assert bool(torch.isfinite(p.grad).all()), name
        assert float(p.grad.abs().max()) > 0, name
    # screen-space gradient consumed by training_statis (scene/gaussian_model.py:710)
    vsp = out["viewspace_points"]
    assert vsp.grad is not None and vsp.grad.shape[1] == 3 and float(vsp.grad[:, :2].abs().max()) > 0
    # _opacity / _rotation never receive a gradient in the reference either
    assert pc._opacity.grad is None and pc._rotation.grad is None
