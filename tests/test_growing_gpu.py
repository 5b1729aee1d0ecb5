"""GPU parity of anchor growing / pruning: contextgs_b200.densify + csrc/anchor_growing.cu against
golden vectors produced by the reference's own methods (tests/golden/make_golden_growing.py) and against the oracle."""
import numpy as np
import pytest
import torch

from contextgs_b200 import _lib, synthetic
from oracle import entropy_ref as er
from oracle import growing_ref as gr
from tests.helpers import cuda_model, load_npz
from tests.test_growing_cpu import check_against_golden, model_from_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [0, 1])
def test_adjust_anchor_matches_reference_golden(case):
    g = load_npz("growing.npz")
    m = model_from_golden(g, case, "cuda")
    _lib.launch_counts(reset=True)
    rand = [torch.from_numpy(g[f"c{case}_rand{i}"]) for i in range(int(g[f"c{case}_n_rand"]))]
    m.adjust_anchor(check_interval=100, success_threshold=0.8, grad_threshold=0.0002, min_opacity=0.005, rand=rand)
    assert _lib.launch_counts().get("anchor_growing", 0) > 0
    check_against_golden(m, g, case)


@pytest.mark.parametrize("N,cell,frac", [(50_000, 0.016, 0.05), (400_000, 0.004, 0.3), (400_000, 0.001, 0.02)])
def test_grow_cells_matches_oracle_at_size(N, cell, frac):
    scene = synthetic.make_scene("chair", N, seed=3)
    pc = er.make_model(scene)
    m = cuda_model(scene, pc)
    g = torch.Generator().manual_seed(N)
    cand = torch.rand(N * 10, generator=g) < frac
    na, nf, nh = m.grow_cells(cand.cuda(), cell)
    ra, rf, rh = gr.grow_cells(pc.get_anchor.numpy(), pc._offset.numpy(), pc.get_scaling.numpy(), pc._anchor_feat.numpy(),
                               pc._hyper_latent.numpy(), cand.numpy(), cell)
    assert ra.shape[0] > 100
    assert np.array_equal(na.cpu().numpy(), ra)
    assert np.array_equal(nf.cpu().numpy(), rf) and np.array_equal(nh.cpu().numpy(), rh)
    # properties: cells are distinct, sorted, and none holds an existing anchor
    cells = torch.round(na / cell).int()
    assert torch.unique(cells, dim=0).shape[0] == cells.shape[0]
    occupied = torch.round(m.get_anchor.detach() / cell).int()
    both = torch.cat([occupied.unique(dim=0), cells], 0)
    assert both.unique(dim=0).shape[0] == occupied.unique(dim=0).shape[0] + cells.shape[0]


def test_grow_cells_edge_cases():
    scene = synthetic.make_scene("chair", 2000, seed=1)
    m = cuda_model(scene, er.make_model(scene))
    none = torch.zeros(20000, dtype=torch.bool, device="cuda")
    na, nf, nh = m.grow_cells(none, 0.016)
    assert na.shape == (0, 3) and nf.shape == (0, 50) and nh.shape == (0, 12)
    # every candidate lands in its own anchor's cell when offsets are zero -> everything is occupied
    with torch.no_grad():
        m._offset.zero_()
    na, _, _ = m.grow_cells(~none, 0.016)
    assert na.shape[0] == 0
    with pytest.raises(ValueError):
        m.grow_cells(none[:-1], 0.016)
    # a cell size that pushes coordinates beyond +-2^20 is reported, not wrapped
    with pytest.raises(_lib.CgsError):
        m.grow_cells(~none, 1e-7)
