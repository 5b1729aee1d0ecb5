"""CPU tests of the rasterizer oracle (oracle/raster_ref.c).

The reference repository holds no rasterizer source, tests or golden vectors,
so the oracle is pinned the only way available: its forward against an independent float64
PyTorch restatement, and its hand-derived backward against autograd of that restatement.
"""
import math

import numpy as np
import pytest
import torch

from contextgs_b200 import synthetic
from oracle import raster_ref
from tests.torch_raster64 import render64


def _settings(cam, bg=(0.1, 0.2, 0.3), mod=1.0):
    return raster_ref.make_settings(cam.image_width, cam.image_height, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5),
                                    bg, mod, cam.world_view_transform.numpy(), cam.full_proj_transform.numpy())


def _scene(P, seed, W=48, H=32):
    cam = synthetic.look_at_camera(W, H, 0.9, (0.3, -3.0, 0.4))
    means, colors, opac, scales, rots = synthetic.random_gaussians(P, seed=seed, extent=0.7, scale_lo=0.05, scale_hi=0.3)
    return cam, means, colors, opac, scales, rots


def test_forward_matches_float64_restatement():
    cam, means, colors, opac, scales, rots = _scene(40, 1)
    st = _settings(cam)
    fwd = raster_ref.forward(st, means.numpy(), colors.numpy(), opac.numpy(), scales.numpy(), rots.numpy())
    assert fwd["R"] > 0 and (fwd["radii"] > 0).sum() > 20
    img64 = render64(st, means.double(), colors.double(), opac.double(), scales.double(), rots.double(), fwd["radii"],
                     cam.image_width, cam.image_height).numpy()
    err = np.linalg.norm(fwd["color"] - img64) / np.linalg.norm(img64)
    assert err < 1e-5, err


def test_backward_matches_autograd():
    cam, means, colors, opac, scales, rots = _scene(40, 2)
    # un-normalised quaternions: upstream does not re-normalise
    rots = rots * (0.8 + 0.4 * torch.rand(rots.shape[0], 1, generator=torch.Generator().manual_seed(5)))
    st = _settings(cam)
    fwd = raster_ref.forward(st, means.numpy(), colors.numpy(), opac.numpy(), scales.numpy(), rots.numpy())
    g = torch.Generator().manual_seed(3)
    dL = torch.randn(3, cam.image_height, cam.image_width, generator=g)
    bwd = raster_ref.backward(st, fwd, means.numpy(), colors.numpy(), scales.numpy(), rots.numpy(), dL.numpy())
    leaves = [t.double().clone().requires_grad_(True) for t in (means, colors, opac, scales, rots)]
    img = render64(st, *leaves, fwd["radii"], cam.image_width, cam.image_height)
    (img * dL.double()).sum().backward()
    for name, leaf in zip(["means3D", "colors", "opacities", "scales", "rotations"], leaves):
        ref = leaf.grad.numpy()
        got = bwd[name].reshape(ref.shape)
        err = np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30)
        assert err < 2e-4, (name, err)


def test_binning_invariants():
    cam, means, colors, opac, scales, rots = _scene(300, 4, W=100, H=70)
    st = _settings(cam)
    fwd = raster_ref.forward(st, means.numpy(), colors.numpy(), opac.numpy(), scales.numpy(), rots.numpy())
    keys, ranges = fwd["keys"], fwd["ranges"]
    assert fwd["R"] == int(fwd["tiles_touched"].sum())
    assert np.all(keys[1:] >= keys[:-1])  # sorted
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    for t in range(ranges.shape[0]):
        b, e = ranges[t]
        assert np.all(tiles[b:e] == t)
        assert e - b == int((tiles == t).sum())
    # stable: equal keys keep ascending Gaussian index
    same = keys[1:] == keys[:-1]
    assert np.all(fwd["point_list"][1:][same] > fwd["point_list"][:-1][same])


def test_filter_matches_preprocess_radii_and_empty():
    cam, means, colors, opac, scales, rots = _scene(200, 6)
    means[:20, 1] -= 10.0  # behind the camera
    st = _settings(cam)
    pre = raster_ref.preprocess(st, means.numpy(), scales.numpy(), rots.numpy(), opac.numpy())
    rad = raster_ref.preprocess(st, means.numpy(), scales.numpy(), rots.numpy(), filter_only=True)
    assert np.array_equal(rad, pre["radii"])
    assert np.all(rad[:20] == 0)
    empty = raster_ref.forward(st, np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 1)), np.zeros((0, 3)), np.zeros((0, 4)))
    assert empty["R"] == 0
    assert np.allclose(empty["color"], np.array(list(st.bg))[:, None, None])
