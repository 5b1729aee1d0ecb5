"""CPU: the host side of anchor growing / pruning (contextgs_b200/densify.py: optimiser surgery, accumulator
bookkeeping, depth loop) with the device call `grow_cells` replaced by the oracle's, against the reference's golden
vectors; and the learning-rate schedule."""
import types

import numpy as np
import torch

from contextgs_b200 import densify
from contextgs_b200.gaussian_model import GaussianModel
from oracle import entropy_ref as er
from oracle import growing_ref as gr
from tests.helpers import load_npz


def model_from_golden(g, case, device):
    pre = f"c{case}_before_"
    m = GaussianModel(voxel_size=float(g[f"c{case}_voxel"]), device=device)
    groups = []
    for k in gr.NAMES:
        p = torch.nn.Parameter(torch.from_numpy(g[pre + k]).to(device))
        setattr(m, "_" + k, p)
        groups.append({"params": [p], "lr": 1e-3, "name": k})
    groups.append({"params": list(m.mlp_opacity.parameters()), "lr": 1e-3, "name": "mlp_opacity"})
    m.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    for k in gr.NAMES:          # the generator's Adam-state rule: exp_avg = p / 2, exp_avg_sq = p * p
        p = getattr(m, "_" + k)
        m.optimizer.state[p] = {"step": torch.tensor(1.0), "exp_avg": p.detach() * 0.5, "exp_avg_sq": p.detach() * p.detach()}
    m.x_bound_min = torch.from_numpy(g[f"c{case}_x_bound_min"]).to(device)
    m.x_bound_max = torch.from_numpy(g[f"c{case}_x_bound_max"]).to(device)
    for k in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom"):
        setattr(m, k, torch.from_numpy(g[pre + k]).to(device))
    return m


def check_against_golden(m, g, case):
    post = f"c{case}_after_"
    for k in gr.NAMES:
        p = getattr(m, "_" + k)
        assert isinstance(p, torch.nn.Parameter) and p.requires_grad
        assert any(gp["name"] == k and gp["params"][0] is p for gp in m.optimizer.param_groups), k
        assert np.array_equal(p.detach().cpu().numpy(), g[post + k]), k
        st = m.optimizer.state[p]
        assert np.array_equal(st["exp_avg"].cpu().numpy(), g[post + k + "_exp_avg"]), k
        if post + k + "_exp_avg_sq" in g:
            assert np.array_equal(st["exp_avg_sq"].cpu().numpy(), g[post + k + "_exp_avg_sq"]), k
    for k in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom"):
        assert np.array_equal(getattr(m, k).cpu().numpy(), g[post + k]), k
    assert len(m.optimizer.state) == len(gr.NAMES)        # no orphaned Adam state of replaced parameters
    assert m.max_radii2D.shape[0] == m._anchor.shape[0]


def oracle_grow_cells(self, candidate_mask, cur_size, n_candidates=None):
    """Stand-in for GaussianModel.grow_cells (the device call) on CPU-only hosts: the oracle's restatement."""
    n = lambda t: t.detach().numpy()
    anchor_q = er.quantize_anchor(self._anchor.detach(), self.x_bound_min, self.x_bound_max)[0]
    out = gr.grow_cells(n(anchor_q), n(self._offset), n(self.get_scaling), n(self._anchor_feat), n(self._hyper_latent),
                        candidate_mask.numpy(), cur_size)
    return tuple(torch.from_numpy(o) for o in out)


def test_host_side_of_adjust_anchor_matches_reference_golden(monkeypatch):
    g = load_npz("growing.npz")
    for case in (0, 1, 2):
        m = model_from_golden(g, case, "cpu")
        monkeypatch.setattr(GaussianModel, "grow_cells", oracle_grow_cells)
        rand = [torch.from_numpy(g[f"c{case}_rand{i}"]) for i in range(int(g[f"c{case}_n_rand"]))]
        m.adjust_anchor(check_interval=100, success_threshold=0.8, grad_threshold=0.0002, min_opacity=0.005, rand=rand)
        check_against_golden(m, g, case)


def test_training_setup_groups_and_schedule():
    a = types.SimpleNamespace(
        percent_dense=0.01, position_lr_init=0.0, position_lr_final=0.0, position_lr_delay_mult=0.01, position_lr_max_steps=30000,
        offset_lr_init=0.01, offset_lr_final=0.0001, offset_lr_delay_mult=0.01, offset_lr_max_steps=30000,
        mask_lr_init=0.01, mask_lr_final=0.0001, mask_lr_delay_mult=0.01, mask_lr_max_steps=30000,
        feature_lr=0.0075, hyper_latent_lr=0.0075, opacity_lr=0.02, scaling_lr=0.007, rotation_lr=0.002,
        mlp_opacity_lr_init=0.002, mlp_opacity_lr_final=0.00002, mlp_opacity_lr_delay_mult=0.01, mlp_opacity_lr_max_steps=30000,
        mlp_cov_lr_init=0.004, mlp_cov_lr_final=0.004, mlp_cov_lr_delay_mult=0.01, mlp_cov_lr_max_steps=30000,
        mlp_color_lr_init=0.008, mlp_color_lr_final=0.00005, mlp_color_lr_delay_mult=0.01, mlp_color_lr_max_steps=30000,
        latent_codec_lr_init=0.005, latent_codec_lr_final=0.00001, latent_codec_lr_delay_mult=0.33, latent_codec_lr_max_steps=30000,
        mlp_grid_lr_init=0.005, mlp_grid_lr_final=0.00001, mlp_grid_lr_delay_mult=0.01, mlp_grid_lr_max_steps=30000)
    g = load_npz("growing.npz")
    m = model_from_golden(g, 1, "cpu")
    m.training_setup(a)
    names = [gp["name"] for gp in m.optimizer.param_groups]
    assert names == ["anchor", "offset", "mask", "anchor_feat", "hyper_latent", "opacity", "scaling", "rotation", "mlp_opacity",
                     "mlp_cov", "mlp_color", "latent_codec", "mlp_grid"]            # scene/gaussian_model.py:455-470
    assert m.opacity_accum.shape == (m._anchor.shape[0], 1) and m.offset_denom.shape == (m._anchor.shape[0] * 10, 1)
    m.update_learning_rate(15000)
    lr = {gp["name"]: gp["lr"] for gp in m.optimizer.param_groups}
    assert lr["anchor"] == 0.0 and lr["anchor_feat"] == 0.0075                          # lr 0 disables; fixed lr untouched
    assert abs(lr["offset"] - 0.001) < 1e-12 and abs(lr["mlp_cov"] - 0.004) < 1e-12      # geometric mean at half time
    # utils/general_utils.py:67-80 by hand
    f = densify.get_expon_lr_func(0.005, 0.00001, max_steps=30000)
    t = 7000 / 30000
    assert abs(f(7000) - np.exp(np.log(0.005) * (1 - t) + np.log(0.00001) * t)) < 1e-15
    assert f(-1) == 0.0 and abs(f(10 ** 6) - 0.00001) < 1e-15
