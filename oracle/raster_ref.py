"""ctypes/numpy front-end of oracle/raster_ref.c -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see the header of raster_ref.c): the reference repository does not
contain the rasterizer; this restates the published 3DGS algorithm that
`gaussian_renderer/__init__.py:179-205,250-285` calls.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libraster_ref.so")


class RefSettings(ctypes.Structure):
    _fields_ = [
        ("W", ctypes.c_int),
        ("H", ctypes.c_int),
        ("tanfovx", ctypes.c_float),
        ("tanfovy", ctypes.c_float),
        ("bg", ctypes.c_float * 3),
        ("scale_modifier", ctypes.c_float),
        ("view", ctypes.c_float * 16),
        ("proj", ctypes.c_float * 16),
    ]


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "raster_ref.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.ref_count_instances.restype = ctypes.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def make_settings(W, H, tanfovx, tanfovy, bg, scale_modifier, view, proj):
    s = RefSettings()
    s.W, s.H = int(W), int(H)
    s.tanfovx, s.tanfovy = float(tanfovx), float(tanfovy)
    for i in range(3):
        s.bg[i] = float(bg[i])
    s.scale_modifier = float(scale_modifier)
    v = np.asarray(view, dtype=np.float32).reshape(16)
    p = np.asarray(proj, dtype=np.float32).reshape(16)
    for i in range(16):
        s.view[i] = float(v[i])
        s.proj[i] = float(p[i])
    return s


def f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def preprocess(st, means, scales, rots, opac=None, filter_only=False):
    means, scales, rots = f32(means), f32(scales), f32(rots)
    P = means.shape[0]
    radii = np.zeros(P, np.int32)
    if filter_only:
        lib().ref_preprocess(ctypes.byref(st), P, _p(means), _p(scales), _p(rots), None, 1, _p(radii), None, None,
                             None, None, None)
        return radii
    opac = f32(opac).reshape(-1)
    xy = np.zeros((P, 2), np.float32)
    depths = np.zeros(P, np.float32)
    cov3D = np.zeros((P, 6), np.float32)
    conic_opacity = np.zeros((P, 4), np.float32)
    tiles = np.zeros(P, np.uint32)
    lib().ref_preprocess(ctypes.byref(st), P, _p(means), _p(scales), _p(rots), _p(opac), 0, _p(radii), _p(xy),
                         _p(depths), _p(cov3D), _p(conic_opacity), _p(tiles))
    return dict(radii=radii, xy=xy, depths=depths, cov3D=cov3D, conic_opacity=conic_opacity, tiles_touched=tiles)


def bin_tiles(st, pre):
    P = pre["radii"].shape[0]
    R = int(lib().ref_count_instances(P, _p(pre["tiles_touched"])))
    gx, gy = (st.W + 15) // 16, (st.H + 15) // 16
    keys = np.zeros(max(R, 1), np.uint64)
    plist = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    lib().ref_bin(ctypes.byref(st), P, _p(pre["radii"]), _p(pre["xy"]), _p(pre["depths"]), ctypes.c_int64(R),
                  _p(keys), _p(plist), _p(ranges))
    return dict(R=R, keys=keys[:R], point_list=plist[:R], ranges=ranges)


def render_forward(st, pre, binned, colors):
    colors = f32(colors)
    H, W = st.H, st.W
    out = np.zeros((3, H, W), np.float32)
    final_T = np.zeros((H, W), np.float32)
    n_contrib = np.zeros((H, W), np.uint32)
    lib().ref_render_forward(ctypes.byref(st), _p(binned["ranges"]), _p(binned["point_list"]), _p(pre["xy"]),
                             _p(colors), _p(pre["conic_opacity"]), _p(out), _p(final_T), _p(n_contrib))
    return dict(color=out, final_T=final_T, n_contrib=n_contrib)


def forward(st, means, colors, opac, scales, rots):
    """Whole forward: returns dict with every intermediate."""
    pre = preprocess(st, means, scales, rots, opac)
    binned = bin_tiles(st, pre)
    img = render_forward(st, pre, binned, colors)
    out = {}
    out.update(pre)
    out.update(binned)
    out.update(img)
    return out


def backward(st, fwd, means, colors, scales, rots, dL_dpix):
    """Backward of `forward`: gradients wrt means3D, means2D (upstream convention), colors,
    opacities, scales, rotations."""
    means, colors, scales, rots = f32(means), f32(colors), f32(scales), f32(rots)
    dL_dpix = f32(dL_dpix)
    P = means.shape[0]
    d_xy = np.zeros((P, 2), np.float64)
    d_conic = np.zeros((P, 3), np.float64)
    d_op = np.zeros(P, np.float64)
    d_col = np.zeros((P, 3), np.float64)
    lib().ref_render_backward(ctypes.byref(st), _p(fwd["ranges"]), _p(fwd["point_list"]), _p(fwd["xy"]), _p(colors),
                              _p(fwd["conic_opacity"]), _p(fwd["final_T"]), _p(fwd["n_contrib"]), _p(dL_dpix),
                              _p(d_xy), _p(d_conic), _p(d_op), _p(d_col))
    d_m2d = np.zeros((P, 3), np.float32)
    d_means = np.zeros((P, 3), np.float32)
    d_scales = np.zeros((P, 3), np.float32)
    d_rots = np.zeros((P, 4), np.float32)
    lib().ref_preprocess_backward(ctypes.byref(st), P, _p(means), _p(scales), _p(rots), _p(fwd["radii"]), _p(d_xy),
                                  _p(d_conic), _p(d_m2d), _p(d_means), _p(d_scales), _p(d_rots))
    return dict(means3D=d_means, means2D=d_m2d, colors=d_col.astype(np.float32),
                opacities=d_op.astype(np.float32).reshape(P, 1), scales=d_scales, rotations=d_rots,
                dL_dxy=d_xy, dL_dconic=d_conic)
