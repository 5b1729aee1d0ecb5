"""The inference fast path of `render` (one host read-back per frame, device-side counts) must give
exactly what the synchronous path gives, and must recover from capacity overflows."""
import pytest
import torch

from contextgs_b200 import renderer, synthetic
from contextgs_b200.gaussian_model import GaussianModel
from contextgs_b200.renderer import prefilter_voxel, render

pytestmark = pytest.mark.gpu


def _setup(N=30000, W=400, H=240, kind="chair"):
    scene = synthetic.make_scene(kind, N, seed=5, gaussian_scale=4.0)
    torch.manual_seed(6)
    pc = GaussianModel.from_tensors(scene, device="cuda")
    pc.replace_with_decoded(**{k: v.cuda() for k, v in synthetic.decoded_scene(scene).items()})
    pc.eval()
    cams = synthetic.make_cameras(kind, 4, device="cuda", W=W, H=H)
    pipe = type("Pipe", (), {"debug": False})()
    return pc, cams, pipe, torch.zeros(3, device="cuda")


def test_fast_path_equals_synchronous_path(monkeypatch):
    pc, cams, pipe, bg = _setup()
    for cam in cams:
        with torch.no_grad():
            vis = prefilter_voxel(cam, pc, pipe, bg)
            fast = render(cam, pc, pipe, bg, visible_mask=vis)
        with torch.enable_grad():                     # grad mode selects the synchronous (autograd) path
            slow = render(cam, pc, pipe, bg, visible_mask=vis)
        assert torch.equal(fast["render"], slow["render"].detach())
        assert torch.equal(fast["radii"], slow["radii"])
        assert fast["viewspace_points"].shape == slow["viewspace_points"].shape
        assert torch.equal(fast["visibility_filter"], slow["visibility_filter"])
        assert float(fast["render"].max()) > 0


def test_fast_path_recovers_from_capacity_overflow():
    pc, cams, pipe, bg = _setup()
    cam = cams[0]
    with torch.no_grad():
        vis = prefilter_voxel(cam, pc, pipe, bg)
        ref = render(cam, pc, pipe, bg, visible_mask=vis)
        dev = ref["render"].device
        renderer._p_cap_hint[dev.index] = 1000                       # far too small: Gaussian overflow
        from contextgs_b200.rasterizer import _state
        _state(dev).r_cap_hint = 0
        again = render(cam, pc, pipe, bg, visible_mask=vis)
    assert torch.equal(again["render"], ref["render"]) and torch.equal(again["radii"], ref["radii"])
    assert renderer._p_cap_hint[dev.index] >= ref["radii"].shape[0]


def test_compact_positive_i32_matches_nonzero():
    from contextgs_b200 import _lib
    L = _lib.lib()
    for n in (1, 9, 4096, 100_003):
        g = torch.Generator().manual_seed(n)
        v = torch.randint(-3, 4, (n,), generator=g, dtype=torch.int32).cuda()
        idx = torch.empty(n, dtype=torch.int32, device="cuda")
        cnt = torch.empty(1, dtype=torch.int32, device="cuda")
        ws = torch.empty(L.cgs_compact_workspace_bytes(n), dtype=torch.uint8, device="cuda")
        _lib.check(L.cgs_compact_positive_i32(_lib.ptr(v), n, _lib.ptr(idx), _lib.ptr(cnt), _lib.ptr(ws), ws.numel(),
                                              _lib.stream_ptr()), "cgs_compact_positive_i32")
        ref = torch.nonzero(v > 0)[:, 0].to(torch.int32)
        assert int(cnt.item()) == ref.numel() and torch.equal(idx[: ref.numel()], ref)


def test_prefilter_voxel_fused_equals_visible_filter():
    """cgs_prefilter_anchors (fused radius test + compaction) against the reference's own call sequence
    (gaussian_renderer/__init__.py:250-287): visible_filter(get_anchor, get_scaling[:, :3], rotation row 0
    repeated) > 0, and against the CPU oracle's radii."""
    import numpy as np
    from contextgs_b200.rasterizer import GaussianRasterizer
    from oracle import raster_ref
    pc, cams, pipe, bg = _setup(N=20011, kind="bicycle")   # cameras inside the scene: part of the anchors is culled
    partial = 0
    with torch.no_grad():
        pc._rotation[0] = torch.tensor([0.9, 0.1, -0.3, 0.2], device="cuda")   # non-trivial row 0 (normalised inside)
    for cam in cams:
        vis = prefilter_voxel(cam, pc, pipe, bg)
        rast = GaussianRasterizer(renderer._settings(cam, pipe, bg, 1.0))
        rots = pc.get_rotation[[0], :].repeat(pc.get_anchor.shape[0], 1)
        radii = rast.visible_filter(pc.get_anchor, pc.get_scaling[:, :3], rots)
        assert vis.dtype == torch.bool and torch.equal(vis, radii > 0)
        idx, cnt, ver = vis._cgs_compact
        ref_idx = torch.nonzero(vis)[:, 0].to(torch.int32)
        assert int(cnt.item()) == ref_idx.numel() and torch.equal(idx[: ref_idx.numel()], ref_idx)
        st = raster_ref.make_settings(cam.image_width, cam.image_height, np.tan(cam.FoVx * 0.5), np.tan(cam.FoVy * 0.5),
                                      (0, 0, 0), 1.0, cam.world_view_transform.cpu().numpy(),
                                      cam.full_proj_transform.cpu().numpy())
        o_radii = raster_ref.preprocess(st, pc.get_anchor.detach().cpu().numpy(),
                                        pc.get_scaling[:, :3].detach().contiguous().cpu().numpy(),
                                        rots.detach().cpu().numpy(), filter_only=True)
        assert np.array_equal(o_radii > 0, vis.cpu().numpy())
        assert int(cnt.item()) > 0
        partial += int(cnt.item()) < vis.numel()
    assert partial > 0


def test_render_accepts_plain_and_modified_masks():
    """A mask that did not come from prefilter_voxel (or was modified afterwards) is compacted again."""
    pc, cams, pipe, bg = _setup()
    cam = cams[1]
    with torch.no_grad():
        vis = prefilter_voxel(cam, pc, pipe, bg)
        ref = render(cam, pc, pipe, bg, visible_mask=vis)
        plain = render(cam, pc, pipe, bg, visible_mask=vis.clone())
        assert torch.equal(plain["render"], ref["render"])
        vis2 = prefilter_voxel(cam, pc, pipe, bg)
        vis2[::2] = False                                   # in-place edit invalidates the cached index list
        a = render(cam, pc, pipe, bg, visible_mask=vis2)
        b = render(cam, pc, pipe, bg, visible_mask=vis2.clone())
        assert torch.equal(a["render"], b["render"]) and a["radii"].shape == b["radii"].shape
        assert a["radii"].shape[0] < ref["radii"].shape[0]


def test_nothing_visible_and_everything_masked_out():
    """Edge cases of the fused frame: no visible anchor (camera looks away) and no Gaussian surviving the mask."""
    import copy
    pc, cams, pipe, bg = _setup()
    bg = torch.tensor([0.2, 0.4, 0.6], device="cuda")
    cam = copy.copy(cams[0])
    # rotate the camera by 180 degrees about its own y axis: everything is behind it
    flip = torch.diag(torch.tensor([-1.0, 1.0, -1.0, 1.0], device="cuda"))
    cam.world_view_transform = (cam.world_view_transform @ flip).contiguous()
    cam.full_proj_transform = (flip @ cam.full_proj_transform).contiguous() if False else \
        (cam.world_view_transform @ (torch.linalg.inv(cams[0].world_view_transform) @ cams[0].full_proj_transform)).contiguous()
    with torch.no_grad():
        vis = prefilter_voxel(cam, pc, pipe, bg)
        assert int(vis.sum()) == 0 and int(vis._cgs_compact[1].item()) == 0
        out = render(cam, pc, pipe, bg, visible_mask=vis)
    assert out["radii"].shape[0] == 0 and out["viewspace_points"].shape == (0, 3)
    assert float(out["render"].abs().max()) == 0.0          # upstream: zeros, not the background, when P == 0
    with torch.enable_grad():
        slow = render(cam, pc, pipe, bg, visible_mask=vis)
    assert torch.equal(out["render"], slow["render"].detach())
    # all offsets masked out: anchors visible, no Gaussian emitted
    with torch.no_grad():
        pc._mask.zero_()
        vis = prefilter_voxel(cams[0], pc, pipe, bg)
        out = render(cams[0], pc, pipe, bg, visible_mask=vis)
    assert int(vis.sum()) > 0 and out["radii"].shape[0] == 0 and float(out["render"].abs().max()) == 0.0
