"""`distCUDA2` of simple_knn (third-party, not in the reference tree; `from simple_knn._C import distCUDA2`,
scene/gaussian_model.py:22, used at :389 and :407): for every point the mean squared distance to its three nearest
neighbours.  One library call (csrc/knn.cu: uniform grid + shells, exact neighbours); no CPU path."""

import torch

from . import _lib


def _grid_cell(points, n):
    """Edge of the binning grid: the cube that holds one point on average inside the central 96 % of the cloud
    (SfM clouds carry far outliers that would blow up a bounding-box estimate), never finer than extent / 2^20."""
    lo, hi = points.min(dim=0)[0], points.max(dim=0)[0]
    if n >= 64:
        s = torch.sort(points, dim=0)[0]
        qlo, qhi = s[int(0.02 * (n - 1))], s[int(0.98 * (n - 1))]
    else:
        qlo, qhi = lo, hi
    ext = (qhi - qlo).double()
    ext = torch.clamp(ext, min=float(ext.max()) * 1e-3 + 1e-30)
    cell = float((ext.prod() / max(0.96 * n, 1.0)) ** (1.0 / 3.0))
    full = float((hi - lo).max())
    return max(cell, full / float(1 << 20) * 1.01, 1e-30), lo


@torch.no_grad()
def distCUDA2(points, cell=None):
    """points [n,3] float32 CUDA -> mean squared distance to the 3 nearest other points, [n] float32."""
    if not (points.is_cuda and points.dtype == torch.float32 and points.dim() == 2 and points.shape[1] == 3):
        raise TypeError("distCUDA2: points must be a float32 CUDA tensor of shape [n, 3] (contextgs_b200 has no CPU path)")
    L = _lib.lib()
    pts = points.contiguous()
    n = pts.shape[0]
    out = torch.empty(n, device=pts.device)
    if n == 0:
        return out
    auto, lo = _grid_cell(pts, n)
    full = float((pts.max(dim=0)[0] - lo).max())
    if full == 0.0 and n >= 4:          # every point coincides: all neighbour distances are zero
        return out.zero_()
    h = max(float(cell), full / float(1 << 20) * 1.01) if cell is not None else auto
    lo_host = lo.cpu().float().contiguous()      # host floats: part of the grid definition
    status = torch.empty(2, dtype=torch.int32, device=pts.device)
    ws = torch.empty((L.cgs_knn3_workspace_bytes(n),), dtype=torch.uint8, device=pts.device)
    for _ in range(12):
        _lib.check(L.cgs_knn3_mean_dist2(_lib.ptr(pts), n, lo_host.data_ptr(), h, _lib.ptr(out), _lib.ptr(status),
                                         _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "cgs_knn3_mean_dist2")
        if not int(status[1]):
            return out
        h *= 2.0          # too many points found no third neighbour within 3 shells: the grid was too fine
    raise _lib.CgsError("distCUDA2: could not find a grid cell size for this point cloud")
