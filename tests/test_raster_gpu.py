"""GPU parity tests of the rasterizer: CUDA path (through the C ABI) vs the CPU oracle."""
import math

import numpy as np
import pytest
import torch

from contextgs_b200 import _lib, synthetic
from contextgs_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_forward_raw,
                                       to_c_settings)
from oracle import raster_ref

pytestmark = pytest.mark.gpu
REL_L2 = 1e-4  # north_star tolerance for floating-point outputs


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def settings_for(cam, bg=(0.1, 0.2, 0.3), mod=1.0, dev="cuda"):
    rs = GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=math.tan(cam.FoVx * 0.5),
        tanfovy=math.tan(cam.FoVy * 0.5), bg=torch.tensor(bg, dtype=torch.float32, device=dev), scale_modifier=mod,
        viewmatrix=cam.world_view_transform.to(dev), projmatrix=cam.full_proj_transform.to(dev), sh_degree=1,
        campos=cam.camera_center.to(dev), prefiltered=False, debug=False)
    st = raster_ref.make_settings(cam.image_width, cam.image_height, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5),
                                  bg, mod, cam.world_view_transform.numpy(), cam.full_proj_transform.numpy())
    return rs, st


def scene(P, seed, W, H, extent=1.2, lo=0.01, hi=0.08):
    cam = synthetic.look_at_camera(W, H, 0.9, (0.3, -3.0, 0.4))
    return (cam,) + synthetic.random_gaussians(P, seed=seed, extent=extent, scale_lo=lo, scale_hi=hi)


def run_cuda(rs, means, colors, opac, scales, rots, r_cap=None):
    cs = to_c_settings(rs)
    d = [t.cuda() for t in (means, colors, opac, scales, rots)]
    color, radii, saved = rasterize_forward_raw(cs, *d, r_cap=r_cap)
    return color, radii, saved


@pytest.mark.parametrize("n,bits", [(0, (0, 32)), (1, (0, 32)), (4095, (0, 32)), (4096, (0, 13)), (4097, (3, 19)),
                                    (300000, (0, 32)), (1 << 20, (0, 13)), (777777, (0, 8))])
def test_radix_sort_matches_stable_sort(n, bits):
    L = _lib.lib()
    g = torch.Generator().manual_seed(n + bits[1])
    cap = n + 1000
    keys = torch.randint(0, 2 ** 31 - 1, (cap,), generator=g, dtype=torch.int64)
    if n > 1000:  # heavy duplicates exercise stability
        keys[: n // 2] = keys[: n // 2] % 17
    kd = keys.to(torch.int32).cuda()
    vals = torch.randint(0, 2 ** 31 - 1, (cap,), generator=g, dtype=torch.int64).to(torch.int32).cuda()
    outs = [torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(4)]
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    ws = torch.zeros(L.cgs_sort_workspace_bytes(cap, *bits), dtype=torch.uint8, device="cuda")
    for use_vals in (True, False):
        _lib.check(L.cgs_sort_pairs_u32(_lib.ptr(kd), _lib.ptr(vals) if use_vals else None, _lib.ptr(outs[0]),
                                        _lib.ptr(outs[1]), _lib.ptr(outs[2]), _lib.ptr(outs[3]), _lib.ptr(n_dev), cap,
                                        bits[0], bits[1], _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "sort")
        torch.cuda.synchronize()
        k = keys[:n].numpy()
        digit = (k >> bits[0]) & ((1 << (bits[1] - bits[0])) - 1)
        order = np.argsort(digit, kind="stable")
        assert np.array_equal(outs[0][:n].cpu().numpy().astype(np.int64), k[order])
        exp_vals = vals[:n].cpu().numpy()[order] if use_vals else order.astype(np.int32)
        assert np.array_equal(outs[1][:n].cpu().numpy(), exp_vals)


@pytest.mark.parametrize("P,W,H", [(500, 100, 70), (20000, 320, 200), (200000, 800, 800)])
def test_forward_intermediates_bit_exact_and_image(P, W, H):
    cam, means, colors, opac, scales, rots = scene(P, seed=P, W=W, H=H)
    means[: P // 10, 1] -= 8.0  # some behind the camera
    rs, st = settings_for(cam)
    ref = raster_ref.forward(st, means.numpy(), colors.numpy(), opac.numpy(), scales.numpy(), rots.numpy())
    color, radii, saved = run_cuda(rs, means, colors, opac, scales, rots)
    geom = saved["geom"].cpu().numpy()
    # integer / index outputs: bit exact
    assert np.array_equal(radii.cpu().numpy(), ref["radii"])
    assert np.array_equal(geom[:, 10].view(np.int32), ref["radii"])
    assert np.array_equal(geom[:, 11].view(np.uint32), ref["tiles_touched"])
    assert np.array_equal(geom[:, 9].view(np.uint32), ref["depths"].view(np.uint32))
    assert np.array_equal(geom[:, 0:2].view(np.uint32), ref["xy"].view(np.uint32))
    assert saved["num_rendered"] == ref["R"]
    R = ref["R"]
    assert np.array_equal(saved["point_list"][:R].cpu().numpy().view(np.uint32), ref["point_list"])
    assert np.array_equal(saved["ranges"].cpu().numpy().view(np.uint32), ref["ranges"])
    # floats: conic is the same op sequence -> exact; image within the north_star tolerance
    vis = ref["radii"] > 0
    assert np.array_equal(geom[vis, 2:5].view(np.uint32), ref["conic_opacity"][vis, 0:3].view(np.uint32))
    assert rel_l2(color.cpu().numpy(), ref["color"]) < REL_L2
    assert rel_l2(saved["final_T"].cpu().numpy(), ref["final_T"]) < REL_L2
    nc = saved["n_contrib"].cpu().numpy().view(np.uint32)
    assert (nc != ref["n_contrib"]).mean() < 1e-4  # threshold pixels may flip with exp rounding


def test_backward_matches_oracle():
    P, W, H = 3000, 160, 120
    cam, means, colors, opac, scales, rots = scene(P, seed=11, W=W, H=H, lo=0.02, hi=0.12)
    rots = rots * (0.8 + 0.4 * torch.rand(P, 1, generator=torch.Generator().manual_seed(5)))
    rs, st = settings_for(cam)
    ref = raster_ref.forward(st, means.numpy(), colors.numpy(), opac.numpy(), scales.numpy(), rots.numpy())
    dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(3))
    refb = raster_ref.backward(st, ref, means.numpy(), colors.numpy(), scales.numpy(), rots.numpy(), dL.numpy())
    leaves = [t.cuda().requires_grad_(True) for t in (means, colors, opac, scales, rots)]
    means2D = torch.zeros(P, 3, device="cuda", requires_grad=True)
    rast = GaussianRasterizer(rs)
    img, radii = rast(means3D=leaves[0], means2D=means2D, opacities=leaves[2], colors_precomp=leaves[1],
                      scales=leaves[3], rotations=leaves[4])
    (img * dL.cuda()).sum().backward()
    assert rel_l2(img.detach().cpu().numpy(), ref["color"]) < REL_L2
    for name, leaf in zip(["means3D", "colors", "opacities", "scales", "rotations"], leaves):
        assert rel_l2(leaf.grad.cpu().numpy(), refb[name].reshape(leaf.shape)) < REL_L2, name
    assert rel_l2(means2D.grad.cpu().numpy(), refb["means2D"]) < REL_L2


def test_visible_filter_and_mark_visible():
    cam, means, colors, opac, scales, rots = scene(50000, seed=21, W=640, H=360, extent=4.0)
    rs, st = settings_for(cam)
    rast = GaussianRasterizer(rs)
    rad = rast.visible_filter(means3D=means.cuda(), scales=scales.cuda(), rotations=rots.cuda())
    ref = raster_ref.preprocess(st, means.numpy(), scales.numpy(), rots.numpy(), filter_only=True)
    assert np.array_equal(rad.cpu().numpy(), ref)
    assert 0 < (ref > 0).sum() < ref.size
    vis = rast.markVisible(means.cuda()).cpu().numpy()
    vm = cam.world_view_transform.numpy().reshape(-1)
    m = means.numpy()
    z = vm[2] * m[:, 0] + vm[6] * m[:, 1] + vm[10] * m[:, 2] + vm[14]
    assert (vis != (z > 0.2)).mean() < 1e-4


def test_edge_cases_empty_culled_overflow_and_api_errors():
    cam, means, colors, opac, scales, rots = scene(4000, seed=31, W=130, H=70)
    rs, st = settings_for(cam)
    rast = GaussianRasterizer(rs)
    e3 = torch.zeros(0, 3, device="cuda")
    img, radii = rast(means3D=e3, means2D=e3, opacities=torch.zeros(0, 1, device="cuda"), colors_precomp=e3,
                      scales=e3, rotations=torch.zeros(0, 4, device="cuda"))
    assert img.shape == (3, 70, 130) and float(img.abs().max()) == 0.0 and radii.numel() == 0
    # everything behind the camera -> background only
    far = means.clone()
    far[:, 1] -= 20.0
    color, radii, saved = run_cuda(rs, far, colors, opac, scales, rots)
    assert int(radii.max()) == 0 and saved["num_rendered"] == 0
    assert torch.allclose(color, rs.bg.view(3, 1, 1).expand_as(color))
    # capacity overflow triggers a transparent re-run with the same result
    ref = raster_ref.forward(st, means.numpy(), colors.numpy(), opac.numpy(), scales.numpy(), rots.numpy())
    color, radii, saved = run_cuda(rs, means, colors, opac, scales, rots, r_cap=64)
    assert saved["num_rendered"] == ref["R"] and saved["r_cap"] >= ref["R"]
    assert np.array_equal(saved["point_list"][: ref["R"]].cpu().numpy().view(np.uint32), ref["point_list"])
    assert rel_l2(color.cpu().numpy(), ref["color"]) < REL_L2
    with pytest.raises(Exception, match="excatly one of either SHs"):
        rast(means3D=e3, means2D=e3, opacities=e3, scales=e3, rotations=e3)
    with pytest.raises(Exception, match="scale/rotation pair"):
        rast(means3D=e3, means2D=e3, opacities=e3, colors_precomp=e3)


def test_full_size_1080p_properties():
    """BASELINE config-3 scale (1080p, millions of Gaussians): size-independent properties."""
    P, W, H = 2_000_000, 1920, 1080
    cam, means, colors, opac, scales, rots = scene(P, seed=41, W=W, H=H, extent=2.0, lo=0.002, hi=0.02)
    rs, st = settings_for(cam)
    color, radii, saved = run_cuda(rs, means, colors, opac, scales, rots)
    R = saved["num_rendered"]
    geom = saved["geom"]
    tiles_touched = geom[:, 11].view(torch.int32).long()
    assert int(tiles_touched.sum()) == R and R > P // 2
    ranges = saved["ranges"].long()
    nonempty = ranges[:, 1] > ranges[:, 0]
    r = ranges[nonempty]
    assert int((r[:, 1] - r[:, 0]).sum()) == R           # ranges partition [0, R)
    assert bool((r[1:, 0] == r[:-1, 1]).all()) and int(r[0, 0]) == 0 and int(r[-1, 1]) == R
    pl = saved["point_list"][:R].long()
    depth = geom[:, 9][pl]
    tile_of = torch.repeat_interleave(torch.nonzero(nonempty)[:, 0], (r[:, 1] - r[:, 0]))
    same_tile = tile_of[1:] == tile_of[:-1]
    assert bool((depth[1:][same_tile] >= depth[:-1][same_tile]).all())      # depth sorted inside a tile
    tie = same_tile & (depth[1:] == depth[:-1])
    assert bool((pl[1:][tie] > pl[:-1][tie]).all())                          # stable on ties
    # every instance's tile lies inside its Gaussian's rectangle
    gx = (W + 15) // 16
    tx, ty = (tile_of % gx).float(), (tile_of // gx).float()
    x, y, rad = geom[:, 0][pl], geom[:, 1][pl], geom[:, 10].view(torch.int32)[pl].float()
    assert bool(((tx * 16 <= x + rad + 15) & (tx * 16 + 16 > x - rad - 1)).all())
    assert bool(((ty * 16 <= y + rad + 15) & (ty * 16 + 16 > y - rad - 1)).all())
    assert bool(torch.isfinite(color).all())
    T = saved["final_T"]
    assert float(T.min()) >= 0 and float(T.max()) <= 1
