"""The reference's own constructor calls (train.py:94-107, test.py:149-161) must work on the drop-in GaussianModel."""
import pytest

from contextgs_b200.gaussian_model import GaussianModel


def test_reference_positional_call():
    # dataset.feat_dim, n_offsets, voxel_size, update_depth, update_init_factor, update_hierachy_factor, use_feat_bank
    m = GaussianModel(50, 10, 0.001, 3, 16, 4, False, n_features_per_level=2, level_num=3, hyper_divisor=4,
                      target_ratio=0.2, disable_hyper=False, device="cpu")
    assert (m.update_depth, m.update_init_factor, m.update_hierachy_factor) == (3, 16, 4)
    assert (m.level_num, m.hyper_divisor, m.target_ratio, m.disable_hyper) == (3, 4, 0.2, False)
    m = GaussianModel(50, 10, 0.01, 2, 32, 8, False, device="cpu")
    assert (m.voxel_size, m.update_depth, m.update_init_factor, m.update_hierachy_factor) == (0.01, 2, 32, 8)


def test_unsupported_shapes_fail_loudly():
    with pytest.raises(NotImplementedError):
        GaussianModel(32, 5, 0.01, device="cpu")
    with pytest.raises(NotImplementedError):
        GaussianModel(50, 10, 0.001, 3, 16, 4, True, device="cpu")
