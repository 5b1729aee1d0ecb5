"""GPU: distCUDA2 (csrc/knn.cu) against the oracle (oracle/knn_ref.py: all-pairs float32, bit-exact; cKDTree at size)
and `create_from_pcd` (scene/gaussian_model.py:382-423).  simple_knn is not in the reference tree: parity unpinned."""
import numpy as np
import pytest
import torch

from contextgs_b200 import _lib
from contextgs_b200.gaussian_model import GaussianModel
from contextgs_b200.knn import distCUDA2
from oracle import knn_ref

pytestmark = pytest.mark.gpu


def clouds(seed):
    g = np.random.default_rng(seed)
    uniform = g.uniform(-1, 1, (3000, 3))
    surface = np.concatenate([g.uniform(-1, 1, (2500, 2)), np.zeros((2500, 1))], 1)           # exactly planar
    voxel = np.unique(np.round(g.normal(0, 0.05, (4000, 3)) / 0.01), axis=0) * 0.01            # on a voxel grid (ties)
    sfm = np.concatenate([g.normal(0, 0.1, (2500, 3)), g.uniform(-50, 50, (40, 3)), [[1e3, 1e3, -1e3]]])  # far outliers
    dup = np.concatenate([uniform[:500], uniform[:500], uniform[:100]])                       # coincident points
    return dict(uniform=uniform, surface=surface, voxel=voxel, sfm=sfm, dup=dup)


@pytest.mark.parametrize("name", ["uniform", "surface", "voxel", "sfm", "dup"])
def test_matches_all_pairs_oracle_bit_exact(name):
    p = clouds(1)[name].astype(np.float32)
    got = distCUDA2(torch.from_numpy(p).cuda()).cpu().numpy()
    assert np.array_equal(got, knn_ref.mean_dist2_bruteforce(p))


def test_any_cell_size_gives_the_same_answer():
    p = clouds(2)["sfm"].astype(np.float32)
    ref = knn_ref.mean_dist2_bruteforce(p)
    t = torch.from_numpy(p).cuda()
    for cell in (1e-4, 0.01, 0.3, 50.0):       # far too fine (retries coarser) ... one cell holds everything
        assert np.array_equal(distCUDA2(t, cell=cell).cpu().numpy(), ref), cell


def test_matches_kdtree_at_size():
    g = np.random.default_rng(3)
    n = 400_000
    p = np.concatenate([g.normal(0, 0.5, (n * 3 // 10, 3)),
                        g.normal(0, 1, (n * 7 // 10, 3)) * g.lognormal(1.0, 0.8, (n * 7 // 10, 1))]).astype(np.float32)
    _lib.launch_counts(reset=True)
    got = distCUDA2(torch.from_numpy(p).cuda()).cpu().numpy().astype(np.float64)
    assert _lib.launch_counts().get("anchor_growing", 0) > 0
    ref = knn_ref.mean_dist2_kdtree(p)
    assert np.all(np.abs(got - ref) <= 1e-4 * ref + 1e-12)


def test_create_from_pcd():
    g = np.random.default_rng(4)
    pts = g.normal(0, 0.3, (20000, 3))
    m = GaussianModel(voxel_size=0.01)
    m.create_from_pcd(pts.copy(), spatial_lr_scale=2.5)
    vox = np.unique(np.round(pts / 0.01), axis=0) * 0.01
    assert np.array_equal(m._anchor.detach().cpu().numpy(), vox.astype(np.float32))
    d2 = np.maximum(knn_ref.mean_dist2_bruteforce(vox.astype(np.float32)), 1e-7)
    assert np.allclose(m._scaling.detach().cpu().numpy(), np.repeat(np.log(np.sqrt(d2))[:, None], 6, 1), rtol=0, atol=1e-5)
    n = vox.shape[0]
    assert m._offset.shape == (n, 10, 3) and m._mask.shape == (n, 10, 1) and m._anchor_feat.shape == (n, 50)
    assert m._hyper_latent.shape == (n, 12) and m.spatial_lr_scale == 2.5
    assert m._anchor.requires_grad and isinstance(m._rotation, torch.nn.Parameter)
    # voxel_size <= 0: the median neighbour distance of the raw cloud becomes the voxel size (:387-393)
    m2 = GaussianModel(voxel_size=0.0)
    m2.create_from_pcd(pts.copy(), spatial_lr_scale=1.0)
    med = np.sort(knn_ref.mean_dist2_kdtree(pts.astype(np.float32)))[int(20000 * 0.5) - 1]
    assert abs(m2.voxel_size - med) <= 1e-4 * med
