"""`render` and `prefilter_voxel` with the reference's signatures and return dicts
(gaussian_renderer/__init__.py:155-229 and :232-287), wired to the contextgs_b200 kernels."""
import math

import torch

from .neural_gaussians import generate_neural_gaussians
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def _settings(viewpoint_camera, pipe, bg_color, scaling_modifier):
    return GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=1, campos=viewpoint_camera.camera_center,
        prefiltered=False, debug=bool(getattr(pipe, "debug", False)))


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, visible_mask=None, retain_grad=False, step=0):
    is_training = pc.get_color_mlp.training
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
