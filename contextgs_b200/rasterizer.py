"""Drop-in replacement of the `diff_gaussian_rasterization` Python surface used by ContextGS.

Same names, argument meaning and error behaviour as the module the reference imports at
gaussian_renderer/__init__.py:20 and drives at :179-205 (render) and :250-285 (prefilter_voxel):
`GaussianRasterizationSettings` (12-field NamedTuple), `GaussianRasterizer(nn.Module)` with
`forward`, `visible_filter`, `markVisible`.  All work happens in libcontextgs_b200.so through
the C ABI (include/contextgs_b200.h); torch only provides memory, the stream and autograd glue.
"""
import os
import weakref
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# ------------------------------------------------------------------ host-side caches

_host_cache = {}  # id(tensor) -> (weakref, version, floats); camera matrices live on the GPU


def _host_floats(t, n):
    """Host copy of a small device tensor, cached per tensor OBJECT (weakref + version checked, so a
    recycled address or an in-place update can never return stale values)."""
    if not torch.is_tensor(t):
        vals = tuple(float(v) for v in t)
        assert len(vals) == n
        return vals
    ent = _host_cache.get(id(t))
    if ent is not None and ent[0]() is t and ent[1] == t._version:
        return ent[2]
    if len(_host_cache) > 4096:
        for k in [k for k, e in _host_cache.items() if e[0]() is None]:
            del _host_cache[k]
        if len(_host_cache) > 4096:
            _host_cache.clear()
    v = tuple(t.detach().reshape(-1).to("cpu", torch.float32).tolist())
    assert len(v) == n, f"expected {n} values, got {len(v)}"
    _host_cache[id(t)] = (weakref.ref(t), t._version, v)
    return v


def to_c_settings(rs: GaussianRasterizationSettings) -> _lib.RasterSettings:
    s = _lib.RasterSettings()
    s.image_height, s.image_width = int(rs.image_height), int(rs.image_width)
    s.tanfovx, s.tanfovy = float(rs.tanfovx), float(rs.tanfovy)
    s.bg[:] = _host_floats(rs.bg, 3)
    s.scale_modifier = float(rs.scale_modifier)
    s.viewmatrix[:] = _host_floats(rs.viewmatrix, 16)
    s.projmatrix[:] = _host_floats(rs.projmatrix, 16)
    s.sh_degree = int(rs.sh_degree)
    s.campos[:] = _host_floats(rs.campos, 3)
    s.prefiltered = int(bool(rs.prefiltered))
    s.debug = int(bool(rs.debug))
    return s


class _DeviceState:
    """Per-(device, stream) scratch: grows monotonically, reused across frames (stream ordered)."""

    def __init__(self):
        self.workspace = None
        self.r_cap_hint = 0
        self.last_num_rendered = None
        self.last_num_pairs = None


_states = {}

# CGS_RASTER_SYNC=0: never read `num_rendered` back; the caller must size R_cap generously and
# can inspect `GaussianRasterizer.last_status`.  Default (1): one read-back AFTER the whole
# forward has been enqueued, re-running only if the instance capacity overflowed.
SYNC_DEFAULT = os.environ.get("CGS_RASTER_SYNC", "1") != "0"


def _state(device):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    st = _states.get(key)
    if st is None:
        st = _states[key] = _DeviceState()
    return st


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (contextgs_b200 has no CPU path)")
    return t.contiguous()


def rasterize_forward_raw(c_settings, means3D, colors, opacities, scales, rotations, r_cap=None, sync=None,
                          count_dev=None):
    """Enqueue the full forward.  Returns (color, radii, saved-state dict).
    count_dev (optional int32 device scalar): the number of valid Gaussians; the tensors then only give
    the CAPACITY and nothing is read back here (sync is forced off; the caller inspects saved['status'])."""
    L = _lib.lib()
    dev = means3D.device
    P = int(means3D.shape[0])
    if count_dev is not None:
        sync = False
    H, W = c_settings.image_height, c_settings.image_width
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    st = _state(dev)
    sync = SYNC_DEFAULT if sync is None else sync
    if r_cap is None:
        r_cap = max(st.r_cap_hint, 4 * P, 1 << 16)
    i32, u32, f32 = torch.int32, torch.int32, torch.float32  # uint32 payloads live in int32 tensors
    color = torch.empty((3, H, W), dtype=f32, device=dev)
    radii = torch.empty((P,), dtype=i32, device=dev)
    geom = torch.empty((P, _lib.GEOM_STRIDE), dtype=f32, device=dev)
    ranges = torch.empty((tiles, 2), dtype=u32, device=dev)
    final_T = torch.empty((H, W), dtype=f32, device=dev)
    n_contrib = torch.empty((H, W), dtype=u32, device=dev)
    status = torch.empty((_lib.STATUS_WORDS,), dtype=i32, device=dev)
    stream = _lib.stream_ptr()
    while True:
        point_list = torch.empty((max(r_cap, 1),), dtype=u32, device=dev)
        need = L.cgs_raster_workspace_bytes(P, r_cap, W, H)
        if st.workspace is None or st.workspace.numel() < need:
            st.workspace = None
            st.workspace = torch.empty((int(need * 1.25) + 1024,), dtype=torch.uint8, device=dev)
        _lib.check(L.cgs_rasterize_forward_dev(
            _lib.ctypes.byref(c_settings), P, _lib.ptr(count_dev), _lib.ptr(means3D), _lib.ptr(colors), _lib.ptr(opacities),
            _lib.ptr(scales), _lib.ptr(rotations), r_cap, _lib.ptr(color), _lib.ptr(radii), _lib.ptr(geom),
            _lib.ptr(point_list), _lib.ptr(ranges), _lib.ptr(final_T), _lib.ptr(n_contrib), _lib.ptr(status),
            _lib.ptr(st.workspace), st.workspace.numel(), stream), "cgs_rasterize_forward_dev")
        if not sync:
            num_rendered = None
            break
        host = status.tolist()  # the one host read-back (upstream does it mid-pipeline)
        num_rendered = host[_lib.STATUS_NUM_RENDERED]
        if not host[_lib.STATUS_OVERFLOW]:
            break
        if num_rendered >= 0x7fffffff:
            raise _lib.CgsError("number of (Gaussian, tile) instances exceeds 2^31")
        r_cap = int(num_rendered * 1.25) + 4096
    if num_rendered is not None:
        st.last_num_rendered = num_rendered
        st.last_num_pairs = host[_lib.STATUS_NUM_PAIRS]
    st.r_cap_hint = max(st.r_cap_hint, min(r_cap, int((num_rendered or r_cap) * 1.5) + 4096))
    saved = dict(geom=geom, point_list=point_list, ranges=ranges, final_T=final_T, n_contrib=n_contrib, status=status,
                 num_rendered=num_rendered, r_cap=r_cap)
    return color, radii, saved


def rasterize_backward_raw(c_settings, means3D, scales, rotations, radii, saved, grad_color):
    L = _lib.lib()
    dev = means3D.device
    P = int(means3D.shape[0])
    f32 = torch.float32
    d_means = torch.empty((P, 3), dtype=f32, device=dev)
    d_means2D = torch.empty((P, 3), dtype=f32, device=dev)
    d_colors = torch.empty((P, 3), dtype=f32, device=dev)
    d_opac = torch.empty((P, 1), dtype=f32, device=dev)
    d_scales = torch.empty((P, 3), dtype=f32, device=dev)
    d_rots = torch.empty((P, 4), dtype=f32, device=dev)
    if P == 0:
        return d_means, d_means2D, d_colors, d_opac, d_scales, d_rots
    ws = torch.empty((L.cgs_raster_backward_workspace_bytes(P),), dtype=torch.uint8, device=dev)
    grad_color = grad_color.contiguous()
    _lib.check(L.cgs_rasterize_backward(
        _lib.ctypes.byref(c_settings), P, _lib.ptr(means3D), _lib.ptr(scales), _lib.ptr(rotations), _lib.ptr(radii),
        _lib.ptr(saved["geom"]), _lib.ptr(saved["point_list"]), _lib.ptr(saved["ranges"]), _lib.ptr(saved["final_T"]),
        _lib.ptr(saved["n_contrib"]), _lib.ptr(grad_color), _lib.ptr(d_means), _lib.ptr(d_means2D), _lib.ptr(d_colors),
        _lib.ptr(d_opac), _lib.ptr(d_scales), _lib.ptr(d_rots), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
        "cgs_rasterize_backward")
    return d_means, d_means2D, d_colors, d_opac, d_scales, d_rots


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, colors_precomp, opacities, scales, rotations, raster_settings, holder):
        cs = to_c_settings(raster_settings)
        means3D = _f32c(means3D, "means3D")
        colors_precomp = _f32c(colors_precomp, "colors_precomp")
        opacities = _f32c(opacities, "opacities")
        scales = _f32c(scales, "scales")
        rotations = _f32c(rotations, "rotations")
        P = means3D.shape[0]
        if P == 0:
            # upstream returns an all-zero image (not the background) when there is nothing to draw
            color = torch.zeros((3, cs.image_height, cs.image_width), dtype=torch.float32, device=means3D.device)
            radii = torch.zeros((0,), dtype=torch.int32, device=means3D.device)
            ctx.empty = True
            ctx.mark_non_differentiable(radii)
            return color, radii
        color, radii, saved = rasterize_forward_raw(cs, means3D, colors_precomp, opacities, scales, rotations)
        ctx.empty = False
        ctx.cs = cs
        ctx.saved_misc = saved
        ctx.save_for_backward(means3D, scales, rotations, radii)
        ctx.mark_non_differentiable(radii)
        if holder is not None:
            holder["saved"] = saved
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, _grad_radii):
        if ctx.empty:
            return (None,) * 8
        means3D, scales, rotations, radii = ctx.saved_tensors
        g = rasterize_backward_raw(ctx.cs, means3D, scales, rotations, radii, ctx.saved_misc, grad_color)
        d_means, d_means2D, d_colors, d_opac, d_scales, d_rots = g
        return d_means, d_means2D, d_colors, d_opac, d_scales, d_rots, None, None


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings
        self.last = {}  # saved state of the most recent forward (tests / diagnostics)

    def markVisible(self, positions):
        with torch.no_grad():
            cs = to_c_settings(self.raster_settings)
            positions = _f32c(positions, "positions")
            N = positions.shape[0]
            vis = torch.empty((N,), dtype=torch.uint8, device=positions.device)
            _lib.check(_lib.lib().cgs_mark_visible(_lib.ctypes.byref(cs), N, _lib.ptr(positions), _lib.ptr(vis),
                                                   _lib.stream_ptr()), "cgs_mark_visible")
        return vis.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is not None or cov3D_precomp is not None:
            raise NotImplementedError(
                "contextgs_b200 implements the ContextGS path only: colors_precomp + scales/rotations "
                "(gaussian_renderer/__init__.py:197-205 always passes shs=None, cov3D_precomp=None)")
        return _RasterizeGaussians.apply(means3D, means2D, colors_precomp, opacities, scales, rotations,
                                         self.raster_settings, self.last)

    def visible_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if cov3D_precomp is not None:
            raise NotImplementedError("cov3D_precomp is not part of the ContextGS path")
        with torch.no_grad():
            cs = to_c_settings(self.raster_settings)
            means3D = _f32c(means3D, "means3D")
            scales = _f32c(scales, "scales")
            rotations = _f32c(rotations, "rotations")
            N = means3D.shape[0]
            radii = torch.empty((N,), dtype=torch.int32, device=means3D.device)
            _lib.check(_lib.lib().cgs_visible_filter(_lib.ctypes.byref(cs), N, _lib.ptr(means3D), _lib.ptr(scales),
                                                     _lib.ptr(rotations), _lib.ptr(radii), _lib.stream_ptr()),
                       "cgs_visible_filter")
        return radii
