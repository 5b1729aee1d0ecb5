"""TEST INFRASTRUCTURE ONLY: CPU restatement of `simple_knn._C.distCUDA2` (third-party, not in the reference tree, so
PARITY UNPINNED; call sites scene/gaussian_model.py:389,407): for every point, the mean of the three smallest squared
distances to the other points.  Two forms: an all-pairs float32 evaluation with the CUDA kernel's operation order
((dx*dx + dy*dy) + dz*dz, (b0 + b1 + b2) / 3) for bit-exact comparison at small n, and scipy's cKDTree (float64) for
large n.  The product never imports this file."""
import numpy as np


def mean_dist2_bruteforce(points):
    p = np.asarray(points, np.float32)
    n = p.shape[0]
    out = np.empty(n, np.float32)
    for i in range(n):
        d = p - p[i]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        d2[i] = np.inf
        b = np.sort(np.partition(d2, 2)[:3])
        out[i] = (b[0] + b[1] + b[2]) / np.float32(3.0)
    return out


def mean_dist2_kdtree(points):
    from scipy.spatial import cKDTree
    p = np.asarray(points, np.float64)
    d, _ = cKDTree(p).query(p, k=4)
    return (d[:, 1:] ** 2).mean(axis=1)
