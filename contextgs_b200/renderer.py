"""`render` and `prefilter_voxel` with the reference's signatures and return dicts
(gaussian_renderer/__init__.py:155-229 and :232-287), wired to the contextgs_b200 kernels."""
import math

import torch

from . import _lib
from .neural_gaussians import compact_indices, generate_neural_gaussians, generate_raw, g1_impl, select_attributes
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_forward_raw, to_c_settings)

_p_cap_hint = {}   # device index -> capacity (in Gaussians) that was enough for the recent frames


def _settings(viewpoint_camera, pipe, bg_color, scaling_modifier):
    return GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=1, campos=viewpoint_camera.camera_center,
        prefiltered=False, debug=bool(getattr(pipe, "debug", False)))


@torch.no_grad()
def _render_inference(viewpoint_camera, pc, pipe, bg_color, scaling_modifier, visible_mask):
    """Inference frame with ONE host read-back at the very end.  The reference synchronises three times
    inside a frame (boolean indexing of the visible anchors, `tensor[mask]` of the Gaussians,
    `num_rendered`); here the visible-anchor count, the Gaussian count and the instance count stay on the
    device -- buffers are sized by capacities learnt from earlier frames -- and a single 8-word status
    block is read after everything has been enqueued.  A capacity overflow re-runs the frame."""
    anchor, feat, offsets, scaling, masks, _ = select_attributes(pc, False, 0)
    N, K, dev = anchor.shape[0], pc.n_offsets, anchor.device
    cs = to_c_settings(_settings(viewpoint_camera, pipe, bg_color, scaling_modifier))
    if visible_mask is None:
        vis_idx, nv_dev = None, None
    else:
        vis_idx, nv_dev = compact_indices(visible_mask)
    offsets2, masks2 = offsets.reshape(N, -1), masks.reshape(N, -1)
    p_cap = _p_cap_hint.get(dev.index, min(N * K, max(4 * N, 1 << 16)))
    while True:
        if nv_dev is None:
            nv_dev = torch.full((1,), N, dtype=torch.int32, device=dev)
        raw = generate_raw(pc, viewpoint_camera.camera_center, anchor, feat, offsets2, scaling, masks2, vis_idx=vis_idx,
                           n_vis=N, nv_dev=nv_dev, out_cap=p_cap)
        color, radii, saved = rasterize_forward_raw(cs, raw["xyz"], raw["color"], raw["opacity"], raw["scaling"],
                                                    raw["rot"], count_dev=raw["count"])
        st = saved["status"].tolist()  # the frame's only synchronisation
        P, R = st[_lib.STATUS_NUM_GAUSSIANS], st[_lib.STATUS_NUM_RENDERED]
        if P < 0:
            raise _lib.CgsError("cgs_neural_gaussians_umma_forward: a tensor-core completion barrier timed out")
        if st[_lib.STATUS_GAUSSIAN_OVERFLOW]:
            p_cap = min(N * K, int(P * 1.25) + 4096)
            continue
        if st[_lib.STATUS_OVERFLOW]:
            from .rasterizer import _state
            _state(dev).r_cap_hint = int(R * 1.25) + 4096
            continue
        break
    _p_cap_hint[dev.index] = max(p_cap, min(N * K, int(P * 1.25) + 4096))
    from .rasterizer import _state
    _state(dev).last_num_rendered = R
    radii = radii[:P]
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
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, pipe, bg_color, scaling_modifier))
    means3D = pc.get_anchor
    scales = pc.get_scaling
    rotations = pc.get_rotation
    radii_pure = rasterizer.visible_filter(means3D=means3D, scales=scales[:, :3],
                                           rotations=rotations[[0], :].repeat(means3D.shape[0], 1),
                                           cov3D_precomp=None)
    return radii_pure > 0
