"""`render` and `prefilter_voxel` with the reference's signatures and return dicts
(gaussian_renderer/__init__.py:155-229 and :232-287), wired to the contextgs_b200 kernels."""
import math

import torch

from . import _lib
from .neural_gaussians import (compact_indices, generate_neural_gaussians, g1_impl, pack_decoder_weights_umma,
                               select_attributes)
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, to_c_settings

_p_cap_hint = {}   # device index -> capacity (in Gaussians) that was enough for the recent frames


def _settings(viewpoint_camera, pipe, bg_color, scaling_modifier):
    return GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=1, campos=viewpoint_camera.camera_center,
        prefiltered=False, debug=bool(getattr(pipe, "debug", False)))


def _c_settings(viewpoint_camera, pipe, bg_color, scaling_modifier):
    """`to_c_settings(_settings(...))`, remembered on the camera object while every input is the same object at the same
    version (two calls per frame -- prefilter_voxel and render -- rebuild the same 51 numbers otherwise)."""
    wvt, prj, cc = viewpoint_camera.world_view_transform, viewpoint_camera.full_proj_transform, viewpoint_camera.camera_center
    dbg = bool(getattr(pipe, "debug", False))
    ver = lambda t: t._version if torch.is_tensor(t) else None
    sig = (int(viewpoint_camera.image_height), int(viewpoint_camera.image_width), float(viewpoint_camera.FoVx),
           float(viewpoint_camera.FoVy), float(scaling_modifier), dbg, ver(wvt), ver(prj), ver(cc), ver(bg_color))
    ent = getattr(viewpoint_camera, "_cgs_c_settings", None)
    if (ent is not None and ent[0] == sig and ent[1] is wvt and ent[2] is prj and ent[3] is cc and ent[4] is bg_color
            and torch.is_tensor(bg_color)):
        return ent[5]
    cs = to_c_settings(_settings(viewpoint_camera, pipe, bg_color, scaling_modifier))
    try:
        viewpoint_camera._cgs_c_settings = (sig, wvt, prj, cc, bg_color, cs)
    except Exception:   # cameras that do not take attributes
        pass
    return cs


class _FrameScratch:
    """Per-device buffers of the inference frame that never leave this module (Gaussian attributes, packed
    records, binning lists, sort workspace): allocated once, grown on demand and reused frame after frame
    (frames are stream-ordered and every frame ends with the status read-back, so reuse is safe)."""

    def __init__(self):
        self.p_cap = self.r_cap = self.hw = self.n = 0
        self.ws = self.g1_ws = None

    def ensure(self, L, dev, N, p_cap, r_cap, W, H):
        f32, i32 = torch.float32, torch.int32
        if p_cap > self.p_cap:
            self.p_cap = p_cap
            self.xyz, self.color = torch.empty((p_cap, 3), dtype=f32, device=dev), torch.empty((p_cap, 3), dtype=f32, device=dev)
            self.opacity, self.scaling = torch.empty((p_cap, 1), dtype=f32, device=dev), torch.empty((p_cap, 3), dtype=f32, device=dev)
            self.rot, self.geom = torch.empty((p_cap, 4), dtype=f32, device=dev), torch.empty((p_cap, _lib.GEOM_STRIDE), dtype=f32, device=dev)
            self.count = torch.empty((1,), dtype=i32, device=dev)
        if r_cap > self.r_cap:
            self.r_cap = r_cap
            self.point_list = torch.empty((r_cap,), dtype=i32, device=dev)
        if H * W != self.hw:
            self.hw = H * W
            tiles = ((W + 15) // 16) * ((H + 15) // 16)
            self.ranges = torch.empty((tiles, 2), dtype=i32, device=dev)
            self.final_T, self.n_contrib = torch.empty((H, W), dtype=f32, device=dev), torch.empty((H, W), dtype=i32, device=dev)
        if N > self.n:
            self.n = N
            self.g1_ws = torch.empty((L.cgs_neural_gaussians_umma_workspace_bytes(N),), dtype=torch.uint8, device=dev)
            self.n_all = None
        need = L.cgs_raster_workspace_bytes(self.p_cap, self.r_cap, W, H)
        if self.ws is None or self.ws.numel() < need:
            self.ws = None
            self.ws = torch.empty((need + 1024,), dtype=torch.uint8, device=dev)


_scratch = {}   # (device index, stream) -> _FrameScratch


@torch.no_grad()
def _render_inference(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, visible_mask):
    """Inference frame with ONE host read-back at the very end.  The reference synchronises three times
    inside a frame (boolean indexing of the visible anchors, `tensor[mask]` of the Gaussians,
    `num_rendered`); here the visible-anchor count, the Gaussian count and the instance count stay on the
    device -- buffers are sized by capacities learnt from earlier frames -- and a single 8-word status
    block is read after everything has been enqueued by ONE library call (cgs_render_anchors_forward).
    A capacity overflow re-runs the frame."""
    from .rasterizer import _state
    L = _lib.lib()
    anchor, feat, offsets, scaling, masks, _ = select_attributes(pc, False, 0)
    N, K, dev = anchor.shape[0], pc.n_offsets, anchor.device
    cs = _c_settings(viewpoint_camera, pipe, bg_color, scaling_modifier)
    H, W = cs.image_height, cs.image_width
    if visible_mask is None:
        vis_idx, nv_dev = None, None
    else:
        cached = getattr(visible_mask, "_cgs_compact", None)   # prefilter_voxel already compacted the mask
        if cached is not None and cached[2] == visible_mask._version:
            vis_idx, nv_dev = cached[0], cached[1]
        else:
            vis_idx, nv_dev = compact_indices(visible_mask)
    anchor, feat, scaling = anchor.contiguous(), feat.contiguous(), scaling.contiguous()
    offsets2, masks2 = offsets.reshape(N, -1).contiguous(), masks.reshape(N, -1).contiguous()
    packed = pack_decoder_weights_umma(pc)
    stream = _lib.stream_ptr()
    rst = _state(dev)
    sc = _scratch.get((dev.index, stream))
    if sc is None:
        sc = _scratch[(dev.index, stream)] = _FrameScratch()
    p_cap = _p_cap_hint.get(dev.index, min(N * K, max(4 * N, 1 << 16)))
    while True:
        r_cap = max(rst.r_cap_hint, 4 * p_cap, 1 << 16)
        sc.ensure(L, dev, N, p_cap, r_cap, W, H)
        if nv_dev is None:
            if sc.n_all is None or int(sc.n_all_n) != N:
                sc.n_all, sc.n_all_n = torch.full((1,), N, dtype=torch.int32, device=dev), N
            nv_dev = sc.n_all
        color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((sc.p_cap,), dtype=torch.int32, device=dev)
        status = torch.empty((_lib.STATUS_WORDS,), dtype=torch.int32, device=dev)
        _lib.check(L.cgs_render_anchors_forward(
            _lib.ctypes.byref(cs), _lib.ptr(packed), _lib.ptr(vis_idx), N, _lib.ptr(nv_dev), sc.p_cap, _lib.ptr(anchor),
            _lib.ptr(feat), _lib.ptr(offsets2), _lib.ptr(scaling), _lib.ptr(masks2), _lib.ptr(sc.xyz), _lib.ptr(sc.color),
            _lib.ptr(sc.opacity), _lib.ptr(sc.scaling), _lib.ptr(sc.rot), _lib.ptr(sc.count), _lib.ptr(sc.g1_ws),
            sc.g1_ws.numel(), sc.r_cap, _lib.ptr(color), _lib.ptr(radii), _lib.ptr(sc.geom), _lib.ptr(sc.point_list),
            _lib.ptr(sc.ranges), _lib.ptr(sc.final_T), _lib.ptr(sc.n_contrib), _lib.ptr(status), _lib.ptr(sc.ws),
            sc.ws.numel(), stream), "cgs_render_anchors_forward")
        st = status.tolist()  # the frame's only synchronisation
        P, R = st[_lib.STATUS_NUM_GAUSSIANS], st[_lib.STATUS_NUM_RENDERED]
        if P < 0:
            raise _lib.CgsError("cgs_neural_gaussians_umma_forward: a tensor-core completion barrier timed out")
        if st[_lib.STATUS_GAUSSIAN_OVERFLOW]:
            p_cap = min(N * K, int(P * 1.25) + 4096)
            continue
        if st[_lib.STATUS_OVERFLOW]:
            if R >= 0x7fffffff:
                raise _lib.CgsError("number of (Gaussian, tile) instances exceeds 2^31")
            rst.r_cap_hint = int(R * 1.25) + 4096
            continue
        break
    _p_cap_hint[dev.index] = max(p_cap, min(N * K, int(P * 1.25) + 4096))
    rst.r_cap_hint = max(rst.r_cap_hint, int(R * 1.25) + 4096)
    rst.last_num_rendered = R
    rst.last_num_pairs = st[_lib.STATUS_NUM_PAIRS]
    radii = radii[:P]
    if P == 0:
        color.zero_()   # upstream returns an all-zero image (not the background) when there is nothing to draw
    # by-products of the reference's return value, sized by the now known P: enqueued after the read-back, they run while
    # the host prepares the next call instead of delaying this frame's completion
    return {"render": color, "viewspace_points": torch.zeros((P, 3), dtype=torch.float32, device=dev),
            "visibility_filter": radii > 0, "radii": radii, "time_sub": 0}


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, visible_mask=None, retain_grad=False, step=0):
    is_training = pc.get_color_mlp.training
    if not is_training and not torch.is_grad_enabled() and g1_impl() == "umma":
        return _render_inference(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, visible_mask)
    out = generate_neural_gaussians(viewpoint_camera, pc, visible_mask, is_training=is_training, step=step)
    if is_training:
        (xyz, color, opacity, scaling, rot, neural_opacity, mask, bit_per_param, bit_per_anchor_param,
         bit_per_feat_param, bit_per_scaling_param, bit_per_offsets_param, bpp_per_level) = out
    else:
        xyz, color, opacity, scaling, rot, time_sub = out
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    if retain_grad:
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, pipe, bg_color, scaling_modifier))
    rendered_image, radii = rasterizer(means3D=xyz, means2D=screenspace_points, shs=None, colors_precomp=color,
                                       opacities=opacity, scales=scaling, rotations=rot, cov3D_precomp=None)
    res = {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
           "radii": radii}
    if is_training:
        res.update({"selection_mask": mask, "neural_opacity": neural_opacity, "scaling": scaling,
                    "bit_per_param": bit_per_param, "bit_per_anchor_param": bit_per_anchor_param,
                    "bit_per_feat_param": bit_per_feat_param, "bit_per_scaling_param": bit_per_scaling_param,
                    "bit_per_offsets_param": bit_per_offsets_param, "bpp_per_level": bpp_per_level})
    else:
        res["time_sub"] = time_sub
    return res


def prefilter_voxel(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None):
    """gaussian_renderer/__init__.py:232-287 -> bool[N].  One fused kernel (cgs_prefilter_anchors): radius test
    with scales = get_scaling[:, :3] and the rotation of anchor 0 for every anchor, as the reference passes them.
    The returned mask also carries the compacted index list + device-side count (`_cgs_compact`), which
    `render` picks up instead of compacting the mask again."""
    with torch.no_grad():
        L = _lib.lib()
        cs = _c_settings(viewpoint_camera, pipe, bg_color, scaling_modifier)
        means3D = pc.get_anchor.detach()
        scales = pc.get_scaling.detach()
        rot_row = pc._rotation.detach()[0]
        if means3D.dtype != torch.float32 or not means3D.is_cuda:
            raise TypeError("prefilter_voxel: anchors must be float32 CUDA tensors (contextgs_b200 has no CPU path)")
        means3D, scales, rot_row = means3D.contiguous(), scales.contiguous(), rot_row.contiguous()
        N, dev = means3D.shape[0], means3D.device
        vis = torch.empty((N,), dtype=torch.bool, device=dev)
        idx = torch.empty((max(N, 1),), dtype=torch.int32, device=dev)
        cnt = torch.empty((1,), dtype=torch.int32, device=dev)
        ws = torch.empty((L.cgs_prefilter_workspace_bytes(N),), dtype=torch.uint8, device=dev)
        _lib.check(L.cgs_prefilter_anchors(_lib.ctypes.byref(cs), N, _lib.ptr(means3D), _lib.ptr(scales),
                                           int(scales.shape[1]), _lib.ptr(rot_row), _lib.ptr(vis), _lib.ptr(idx),
                                           _lib.ptr(cnt), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                   "cgs_prefilter_anchors")
    vis._cgs_compact = (idx, cnt, vis._version)
    return vis
