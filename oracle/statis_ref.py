"""TEST INFRASTRUCTURE ONLY: CPU restatement of `GaussianModel.training_statis` (scene/gaussian_model.py:696-713) as
explicit index arithmetic.  Pinned against golden vectors produced by the reference's own function
(tests/golden/make_golden_statis.py -> tests/golden/statis.npz)."""
import numpy as np


def training_statis(state, K, grad, opacity, update_filter, offset_selection_mask, anchor_visible_mask):
    """state: dict of float32 arrays opacity_accum[N,1], anchor_demon[N,1], offset_gradient_accum[N*K,1], offset_denom[N*K,1]
    (updated in place)."""
    vis_idx = np.nonzero(anchor_visible_mask)[0]
    op = np.maximum(opacity.reshape(-1, K), 0).sum(axis=1, dtype=np.float32)
    state["opacity_accum"][vis_idx, 0] += op
    state["anchor_demon"][vis_idx, 0] += 1
    kept = np.nonzero(offset_selection_mask.reshape(-1))[0]          # p-th emitted Gaussian <- p-th kept slot
    drawn = np.nonzero(update_filter)[0]
    slot = kept[drawn]
    dst = vis_idx[slot // K] * K + slot % K
    norm = np.sqrt((grad[drawn, :2].astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))
    state["offset_gradient_accum"][dst, 0] += norm
    state["offset_denom"][dst, 0] += 1
