"""GPU parity: fused anchor->Gaussian kernel vs golden vectors from the reference function and vs
the CPU oracle at BASELINE config-1 size."""
import numpy as np
import pytest
import torch

from contextgs_b200 import synthetic
from contextgs_b200.neural_gaussians import compact_indices, generate_neural_gaussians
from oracle import entropy_ref as er
from tests.helpers import T, cuda_model, fixture_model, load_npz, rel_l2

pytestmark = pytest.mark.gpu
REL_L2 = 1e-4


@pytest.fixture(params=["umma", "simt"], autouse=True)
def g1_impl(request, monkeypatch):
    """Every test runs against the tcgen05 kernel (default) and the fp32-FMA kernel."""
    monkeypatch.setenv("CGS_G1_IMPL", request.param)
    return request.param


def _check_against(out, ref_mask, ref, pre_opacity=None):
    xyz, color, opacity, scaling, rot, neural_opacity, mask = out[:7]
    m = mask.cpu().numpy()
    if not np.array_equal(m, ref_mask):
        # a selection can only flip where the pre-activation is within fp32 rounding of zero
        bad = np.nonzero(m != ref_mask)[0]
        assert pre_opacity is not None and np.all(np.abs(pre_opacity.reshape(-1)[bad]) < 1e-5), bad[:10]
        pytest.skip("selection flipped at a rounding-level zero crossing; positional comparison skipped")
    for name, t in zip(("xyz", "color", "opacity", "scaling", "rot"), (xyz, color, opacity, scaling, rot)):
        assert t.shape[0] == int(ref_mask.sum())
        assert rel_l2(t.cpu().numpy(), ref[name]) < REL_L2, name
    assert rel_l2(neural_opacity.cpu().numpy(), ref["neural_opacity"]) < REL_L2


@pytest.mark.parametrize("n", [0, 1, 7, 8, 2047, 2048, 2049, 100003])
def test_compact_indices_matches_nonzero(n):
    g = torch.Generator().manual_seed(n)
    mask = (torch.rand(n, generator=g) < 0.37).cuda()
    idx, cnt = compact_indices(mask)
    ref = torch.nonzero(mask)[:, 0].to(torch.int32)
    assert int(cnt.item()) == ref.numel()
    assert torch.equal(idx[: ref.numel()], ref)


def test_matches_reference_golden():
    g = load_npz("neural_gaussians.npz")
    scene, pc = fixture_model(load_npz("context_model.npz"))
    model = cuda_model(scene, pc)
    assert np.array_equal(model.get_anchor.detach().cpu().numpy(), pc.get_anchor.numpy())  # Quantize_anchor bit exact
    cam = synthetic.make_cameras("train", 3, device="cuda")[1]
    assert np.allclose(cam.camera_center.cpu().numpy(), g["camera_center"])
    vis = T(g["visible"]).cuda()
    model.train()
    out = generate_neural_gaussians(cam, model, vis, is_training=True, step=100)
    ref = {k[6:]: v for k, v in g.items() if k.startswith("train_")}
    _check_against(out, ref["mask"], ref)
    model.eval()
    model.decoded_version = False  # non-decoded eval path also runs the context model; decoded path below
    xyz = generate_neural_gaussians(cam, model, vis, is_training=True, step=0)[0]
    assert rel_l2(xyz.cpu().numpy(), g["eval_xyz"]) < REL_L2


def test_matches_oracle_config1_size():
    N = 50_000
    scene = synthetic.make_scene("chair", N, seed=1)
    pc = er.make_model(scene)
    model = cuda_model(scene, pc)
    cam_cpu = synthetic.make_cameras("chair", 4)[2]
    cam = synthetic.make_cameras("chair", 4, device="cuda")[2]
    vis = torch.rand(N, generator=torch.Generator().manual_seed(5)) < 0.5
    with torch.no_grad():
        ref = er.generate_neural_gaussians(pc, cam_cpu.camera_center, pc.get_anchor[vis], pc._anchor_feat[vis],
                                           pc._offset[vis], pc.get_scaling[vis], pc.get_mask[vis])
    model.train()
    out = generate_neural_gaussians(cam, model, vis.cuda(), is_training=True, step=0)
    refn = {k: v.numpy() for k, v in ref.items()}
    _check_against(out, refn["mask"], refn, pre_opacity=refn["pre_opacity"])
    assert 0.2 < refn["mask"].mean() < 0.6


def test_no_visible_anchor():
    scene, pc = fixture_model(load_npz("context_model.npz"))
    model = cuda_model(scene, pc)
    cam = synthetic.make_cameras("train", 3, device="cuda")[0]
    vis = torch.zeros(model._anchor.shape[0], dtype=torch.bool, device="cuda")
    model.train()
    out = generate_neural_gaussians(cam, model, vis, is_training=True, step=0)
    assert out[0].shape == (0, 3) and out[5].shape == (0, 1) and out[6].numel() == 0
