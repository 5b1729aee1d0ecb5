"""CPU: checkpoint surfaces of GaussianModel (scene/gaussian_model.py:221-286, :912-951): file keys, CompressAI key
names, capture -> restore round trip including the Adam state."""
import types

import torch

from contextgs_b200.gaussian_model import GaussianModel, _compressai_keys
from tests.helpers import load_npz
from tests.test_growing_cpu import model_from_golden


def _args():
    a = dict(percent_dense=0.01, feature_lr=0.0075, hyper_latent_lr=0.0075, opacity_lr=0.02, scaling_lr=0.007, rotation_lr=0.002)
    for n, (i, f) in dict(position=(0.0, 0.0), offset=(0.01, 1e-4), mask=(0.01, 1e-4), mlp_opacity=(0.002, 2e-5),
                          mlp_cov=(0.004, 0.004), mlp_color=(0.008, 5e-5), latent_codec=(0.005, 1e-5),
                          mlp_grid=(0.005, 1e-5)).items():
        a.update({n + "_lr_init": i, n + "_lr_final": f, n + "_lr_delay_mult": 0.01, n + "_lr_max_steps": 30000})
    return types.SimpleNamespace(**a)


def test_mlp_checkpoint_round_trip_and_reference_keys(tmp_path):
    m = model_from_golden(load_npz("growing.npz"), 1, "cpu")
    m.level_scale = [3.5, 11.0]
    path = str(tmp_path / "ck" / "checkpoint.pth")
    m.save_mlp_checkpoints(path)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"opacity_mlp", "cov_mlp", "color_mlp", "latent_codec", "grid_mlp", "bound", "level_scale"}
    # a file as the reference writes it: CompressAI 1.1.x parameter names + its buffers
    legacy = {}
    for k, v in ck["latent_codec"].items():
        for new, old in (("matrices.", "_matrix"), ("biases.", "_bias"), ("factors.", "_factor")):
            if k.startswith(new):
                k = old + k[len(new):]
        legacy[k] = v + 0.25
    legacy.update(_offset=torch.zeros(12), _quantized_cdf=torch.zeros(12, 40), _cdf_length=torch.zeros(12), target=torch.zeros(3))
    assert _compressai_keys(legacy)["matrices.2"].shape == m.latent_codec.matrices[2].shape
    ck["latent_codec"] = legacy
    torch.save(ck, path)
    fresh = GaussianModel(voxel_size=m.voxel_size, device="cpu")
    fresh.load_mlp_checkpoints(path)
    assert fresh.level_scale == [3.5, 11.0] and torch.equal(fresh.x_bound_max, m.x_bound_max)
    for a, b in zip(fresh.mlp_grid.parameters(), m.mlp_grid.parameters()):
        assert torch.equal(a, b)
    for a, b in zip(fresh.latent_codec.parameters(), m.latent_codec.parameters()):
        assert torch.equal(a, b + 0.25)


def test_capture_restore_round_trip():
    m = model_from_golden(load_npz("growing.npz"), 1, "cpu")
    m.training_setup(_args())
    m.update_learning_rate(100)
    (m._anchor_feat.sum() + m._offset.pow(2).sum() + m.mlp_cov[0].weight.sum()).backward()
    m.optimizer.step()
    m.level_scale = [2.0, 9.0]
    m.max_radii2D = torch.zeros(m._anchor.shape[0])
    state = m.capture()
    assert len(state) == 19 and state[0] is m._anchor and state[-1] == [2.0, 9.0]
    fresh = GaussianModel(voxel_size=m.voxel_size, device="cpu")
    fresh.restore(state, _args())
    assert torch.equal(fresh._offset, m._offset) and fresh.level_scale == [2.0, 9.0]
    st_a, st_b = fresh.optimizer.state[fresh._offset], m.optimizer.state[m._offset]
    assert torch.equal(st_a["exp_avg"], st_b["exp_avg"]) and torch.equal(st_a["exp_avg_sq"], st_b["exp_avg_sq"])
    for a, b in zip(fresh.mlp_cov.parameters(), m.mlp_cov.parameters()):
        assert torch.equal(a, b)
