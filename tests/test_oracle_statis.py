"""oracle/statis_ref.py against golden vectors from the reference's own training_statis."""
import os

import numpy as np

from oracle import statis_ref

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "statis.npz"))
KEYS = ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom")


def test_oracle_statis_matches_reference_golden():
    N, K = G["it0_vis"].shape[0], 10
    state = dict(opacity_accum=np.zeros((N, 1), np.float32), anchor_demon=np.zeros((N, 1), np.float32),
                 offset_gradient_accum=np.zeros((N * K, 1), np.float32), offset_denom=np.zeros((N * K, 1), np.float32))
    for it in range(3):
        statis_ref.training_statis(state, K, G[f"it{it}_grad"], G[f"it{it}_opacity"], G[f"it{it}_upd"], G[f"it{it}_keep"],
                                   G[f"it{it}_vis"])
        for k in KEYS:
            assert np.allclose(state[k], G[f"it{it}_{k}"], rtol=1e-6, atol=1e-6), (it, k)
        assert state["offset_denom"].sum() > 0
