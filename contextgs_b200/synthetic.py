"""Synthetic cameras and scenes.

Cameras reproduce the reference's matrix conventions exactly
(scene/cameras.py:48-57, utils/graphics_utils.py:38-71): `world_view_transform` and
`full_proj_transform` are the TRANSPOSED matrices (row-major storage of the transpose),
which is what `GaussianRasterizationSettings.viewmatrix/projmatrix` receive
(gaussian_renderer/__init__.py:186-187).
"""
import math
from types import SimpleNamespace

import numpy as np
import torch


def get_world2view2(R, t, translate=np.array([0.0, 0.0, 0.0]), scale=1.0):
    """utils/graphics_utils.py:38-49 (R is camera-to-world, t is world-to-camera translation)."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    cam_center = C2W[:3, 3]
    cam_center = (cam_center + translate) * scale
    C2W[:3, 3] = cam_center
    Rt = np.linalg.inv(C2W)
    return np.float32(Rt)


def get_projection_matrix(znear, zfar, fovX, fovY):
    """utils/graphics_utils.py:51-71."""
    tanHalfFovY = math.tan(fovY / 2)
    tanHalfFovX = math.tan(fovX / 2)
    top = tanHalfFovY * znear
    bottom = -top
    right = tanHalfFovX * znear
    left = -right
    P = torch.zeros(4, 4)
    z_sign = 1.0
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = z_sign
    P[2, 2] = z_sign * zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def look_at_camera(W, H, fovx, pos, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0), device="cpu", uid=0):
    """Camera-equivalent of scene/cameras.py:17-57 for a look-at pose (x right, y down, z forward)."""
    pos = np.asarray(pos, np.float64)
    f = np.asarray(target, np.float64) - pos
    f /= np.linalg.norm(f)
    down = -np.asarray(up, np.float64)
    x = np.cross(down, f)
    x /= np.linalg.norm(x)
    y = np.cross(f, x)
    Rw2c = np.stack([x, y, f], axis=0)  # rows: camera axes in world coordinates
    R = Rw2c.T  # camera-to-world, what the reference stores as `R`
    T = -Rw2c @ pos
    focal = W / (2.0 * math.tan(fovx / 2.0))
    fovy = 2.0 * math.atan(H / (2.0 * focal))
    wvt = torch.tensor(get_world2view2(R, T)).transpose(0, 1)
    proj = get_projection_matrix(znear=0.01, zfar=100.0, fovX=fovx, fovY=fovy).transpose(0, 1)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    center = wvt.inverse()[3, :3]
    return SimpleNamespace(
        uid=uid, image_width=int(W), image_height=int(H), FoVx=float(fovx), FoVy=float(fovy), znear=0.01, zfar=100.0,
        world_view_transform=wvt.contiguous().to(device), full_proj_transform=full.contiguous().to(device),
        camera_center=center.contiguous().to(device))


def ring_cameras(n, W, H, fovx, radius, height, seed=8, device="cpu"):
    g = np.random.default_rng(seed)
    phase = g.uniform(0, 2 * math.pi)
    cams = []
    for i in range(n):
        th = phase + 2 * math.pi * i / n
        cams.append(look_at_camera(W, H, fovx, (radius * math.cos(th), radius * math.sin(th), height), device=device, uid=i))
    return cams


def sphere_cameras(n, W, H, fovx, radius, seed=8, device="cpu"):
    g = np.random.default_rng(seed)
    cams = []
    for i in range(n):
        v = g.normal(size=3)
        v /= np.linalg.norm(v)
        if abs(v[2]) > 0.95:  # keep away from the up-axis singularity
            v = np.array([v[0], v[1] + 0.5, 0.5])
            v /= np.linalg.norm(v)
        cams.append(look_at_camera(W, H, fovx, radius * v, device=device, uid=i))
    return cams


def random_gaussians(P, seed=0, extent=1.0, scale_lo=0.01, scale_hi=0.08):
    """Free-standing random Gaussians for rasterizer tests (not an anchor scene)."""
    g = torch.Generator().manual_seed(seed)
    means = (torch.rand(P, 3, generator=g) * 2 - 1) * extent
    scales = torch.rand(P, 3, generator=g) * (scale_hi - scale_lo) + scale_lo
    rots = torch.randn(P, 4, generator=g)
    rots = rots / rots.norm(dim=1, keepdim=True)
    opac = torch.rand(P, 1, generator=g) * 0.9 + 0.05
    colors = torch.rand(P, 3, generator=g)
    return means, colors, opac, scales, rots


# --------------------------------------------------------------------------- anchor scenes

SCENE_KINDS = {
    # kind: (voxel_size, camera spec)
    "chair": dict(voxel=0.001, W=800, H=800, fovx=0.6911, cams="sphere", radius=4.0),
    "bicycle": dict(voxel=0.001, W=1920, H=1080, fovx=1.0, cams="ring", radius=3.0, height=0.5),
    "train": dict(voxel=0.01, W=980, H=545, fovx=1.0, cams="ring", radius=3.0, height=0.5),
}


def _anchors(kind, N, voxel, g):
    if kind == "chair":
        # union of 6 axis-aligned box surfaces inside [-1,1]^3 + jitter
        pts = []
        boxes = g.uniform(-0.8, 0.8, size=(6, 3)), g.uniform(0.1, 0.5, size=(6, 3))
        per = (N * 2) // 6 + 16
        for c, h in zip(*boxes):
            u = g.uniform(-1, 1, size=(per, 3)) * h
            face = g.integers(0, 3, size=per)
            sign = g.choice([-1.0, 1.0], size=per)
            u[np.arange(per), face] = sign * h[face]
            pts.append(c + u)
        x = np.concatenate(pts) + g.normal(0, 2 * voxel, size=(per * 6, 3))
        x = np.clip(x, -1, 1)
    elif kind == "bicycle":
        n_core = int(N * 2 * 0.3)
        n_shell = N * 2 - n_core
        core = g.normal(0, 0.5, size=(n_core, 3))
        d = g.normal(size=(n_shell, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        r = np.minimum(g.lognormal(1.0, 0.8, size=(n_shell, 1)), 60.0)
        x = np.concatenate([core, d * r])
    elif kind == "train":
        n_box = N
        box = g.uniform(-1, 1, size=(n_box, 3)) * np.array([2.0, 0.5, 0.75])
        ground = np.concatenate([g.uniform(-6, 6, size=(N, 2)), g.normal(-0.75, 0.02, size=(N, 1))], axis=1)
        x = np.concatenate([box, ground])
    else:
        raise ValueError(kind)
    g.shuffle(x)
    x = np.round(x / voxel)
    q = x.astype(np.int64) + (1 << 20)  # |coordinate / voxel| < 2^20 for every scene kind
    assert q.min() >= 0 and q.max() < (1 << 21)
    _, first = np.unique((q[:, 0] << 42) | (q[:, 1] << 21) | q[:, 2], return_index=True)
    x = x[np.sort(first)] * voxel
    if x.shape[0] < N:
        raise RuntimeError(f"synthetic scene '{kind}' produced only {x.shape[0]} unique anchors < {N}")
    return np.float32(x[:N])


def make_scene(kind, N, seed=0, feat_dim=50, n_offsets=10, hyper_divisor=4, gaussian_scale=1.0):
    """Per-anchor parameters of a synthetic ContextGS model (shapes of
    scene/gaussian_model.py:415-422).  `gaussian_scale` multiplies the base Gaussian size so
    that the 1080p workload has a realistic instances-per-Gaussian ratio."""
    spec = SCENE_KINDS[kind]
    voxel = spec["voxel"]
    anchor = torch.from_numpy(_anchors(kind, N, voxel, np.random.default_rng(seed)))
    gen = lambda s: torch.Generator().manual_seed(seed * 100 + s)
    feat = torch.round(torch.randn(N, feat_dim, generator=gen(1)) * 2.0) + torch.randn(N, feat_dim, generator=gen(11)) * 0.1
    hyper = torch.randn(N, feat_dim // hyper_divisor, generator=gen(2)) * 1.5
    offset = torch.randn(N, n_offsets, 3, generator=gen(3)) * 0.5
    u = torch.rand(N, 6, generator=gen(4))
    sc = torch.empty(N, 6)
    sc[:, :3] = torch.log(voxel * (2.0 + 18.0 * u[:, :3]))
    sc[:, 3:] = torch.log(voxel * gaussian_scale * (0.5 + 4.5 * u[:, 3:]))
    mask = (torch.rand(N, n_offsets, 1, generator=gen(5)) < 0.7).float() * 8.0 - 4.0
    return dict(kind=kind, voxel_size=voxel, anchor=anchor, feat=feat, hyper=hyper, offset=offset, scaling=sc, mask=mask)


def make_cameras(kind, n, device="cpu", W=None, H=None):
    spec = SCENE_KINDS[kind]
    W = W or spec["W"]
    H = H or spec["H"]
    if spec["cams"] == "sphere":
        return sphere_cameras(n, W, H, spec["fovx"], spec["radius"], device=device)
    return ring_cameras(n, W, H, spec["fovx"], spec["radius"], spec["height"], device=device)


def decoded_scene(scene, x_bound_min=None, x_bound_max=None):
    """The values a ContextGS model holds AFTER `conduct_decoding` (scene/gaussian_model.py:1503-1533,
    `decoded_version=True`): anchors on the 16-bit grid (utils/encodings.py:219-227), features /
    offsets on their base quantisation grids Q = 1 / 0.2 (gaussian_renderer/__init__.py:40-42),
    scaling already exponentiated (left un-rounded: the synthetic sizes are of the order of the
    base step 1e-3 and would collapse to zero), binary offset masks.
    Plain CPU torch: this is synthetic-data construction shared by both bench arms."""
    a = scene["anchor"]
    if x_bound_min is None:
        mn, mx = a.min(dim=0, keepdim=True)[0], a.max(dim=0, keepdim=True)[0]
        x_bound_min = torch.where(mn < 0, mn * 1.2, mn * 0.8)
        x_bound_max = torch.where(mx > 0, mx * 1.2, mx * 0.8)
    interval = (x_bound_max - x_bound_min) / (2 ** 16 - 1) + 1e-6
    q = torch.clamp(torch.div(a - x_bound_min, interval, rounding_mode="floor"), 0, 2 ** 16 - 1)
    anchor = q * interval + x_bound_min
    rnd = lambda x, Q: torch.round(x / Q) * Q
    mask = (torch.sigmoid(scene["mask"]) > 0.01).float()
    return dict(anchor=anchor, hyper=torch.round(scene["hyper"]), feat=rnd(scene["feat"], 1.0),
                offsets=rnd(scene["offset"], 0.2), scaling=torch.exp(scene["scaling"]), masks=mask)
