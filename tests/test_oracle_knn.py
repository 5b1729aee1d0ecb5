"""CPU: the two forms of the distCUDA2 oracle agree (all-pairs float32 vs scipy cKDTree float64)."""
import numpy as np

from oracle import knn_ref


def test_bruteforce_matches_kdtree():
    g = np.random.default_rng(0)
    p = g.normal(0, 1, (1500, 3)).astype(np.float32)
    a, b = knn_ref.mean_dist2_bruteforce(p), knn_ref.mean_dist2_kdtree(p)
    assert np.all(np.abs(a - b) <= 1e-5 * b)
    # hand-checkable: unit grid -> the three nearest neighbours of an inner point are at distance 1
    grid = np.stack(np.meshgrid(*[np.arange(4.0)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    d = knn_ref.mean_dist2_bruteforce(grid)
    assert d[21] == 1.0 and np.isclose(d[0], 1.0)
