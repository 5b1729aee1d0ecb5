"""CPU: the oracle's restatement of anchor growing / pruning (oracle/growing_ref.py) against golden vectors produced by
the reference's own methods (tests/golden/make_golden_growing.py)."""
import numpy as np
import pytest

from oracle import growing_ref as gr
from tests.helpers import load_npz


@pytest.mark.parametrize("case", [0, 1, 2])
def test_adjust_anchor_matches_reference_golden(case):
    g = load_npz("growing.npz")
    st = gr.state_from_golden(g, case)
    n0 = st["params"]["anchor"].shape[0]
    rands = [g[f"c{case}_rand{i}"] for i in range(int(g[f"c{case}_n_rand"]))]
    prune_mask = gr.adjust_anchor(st, rands, float(g[f"c{case}_voxel"]))
    post = f"c{case}_after_"
    n1 = g[post + "anchor"].shape[0]
    assert prune_mask.sum() > 0
    if case < 2:
        assert n1 + prune_mask.sum() > n0          # the case both grows and prunes
    else:
        assert n1 + prune_mask.sum() == n0         # prune only: nothing passes the gradient threshold
    for k in gr.NAMES:
        assert np.array_equal(st["params"][k], g[post + k]), k
        assert np.array_equal(st["exp_avg"][k], g[post + k + "_exp_avg"]), k
        if post + k + "_exp_avg_sq" in g:
            assert np.array_equal(st["exp_avg_sq"][k], g[post + k + "_exp_avg_sq"]), k
    for k in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom"):
        assert np.array_equal(st["stats"][k], g[post + k]), k


def test_grow_cells_rejects_occupied_cells_and_takes_the_maximum():
    anchor = np.array([[0.0, 0.0, 0.0], [0.32, 0.0, 0.0]], np.float32)
    offset = np.zeros((2, 10, 3), np.float32)
    offset[0, 0] = (0.17, 0.0, 0.0)      # -> cell (1,0,0) at size 0.16: free
    offset[0, 1] = (0.30, 0.0, 0.0)      # -> cell (2,0,0): occupied by anchor 1
    offset[1, 2] = (-0.15, 0.01, 0.0)    # -> 0.17 -> cell (1,0,0) again
    scaling = np.ones((2, 6), np.float32)
    feat = np.stack([np.arange(50), 49 - np.arange(50)]).astype(np.float32)
    hyper = np.stack([np.zeros(12), np.ones(12)]).astype(np.float32)
    cand = np.zeros(20, bool)
    cand[[0, 1, 12]] = True
    na, nf, nh = gr.grow_cells(anchor, offset, scaling, feat, hyper, cand, 0.16)
    assert na.shape == (1, 3) and np.allclose(na[0], (0.16, 0, 0))
    assert np.array_equal(nf[0], np.maximum(feat[0], feat[1])) and np.array_equal(nh[0], np.ones(12, np.float32))
