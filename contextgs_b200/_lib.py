"""ctypes binding of libcontextgs_b200.so (the C ABI in include/contextgs_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcontextgs_b200.so")

c_void_p, c_int, c_int64, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t


class RasterSettings(ctypes.Structure):
    """struct cgs_raster_settings"""
    _fields_ = [
        ("image_height", ctypes.c_int32),
        ("image_width", ctypes.c_int32),
        ("tanfovx", ctypes.c_float),
        ("tanfovy", ctypes.c_float),
        ("bg", ctypes.c_float * 3),
        ("scale_modifier", ctypes.c_float),
        ("viewmatrix", ctypes.c_float * 16),
        ("projmatrix", ctypes.c_float * 16),
        ("sh_degree", ctypes.c_int32),
        ("campos", ctypes.c_float * 3),
        ("prefiltered", ctypes.c_int32),
        ("debug", ctypes.c_int32),
    ]


STATUS_NUM_RENDERED, STATUS_OVERFLOW, STATUS_NUM_SORTED, STATUS_WORDS = 0, 1, 2, 8
STATUS_NUM_GAUSSIANS, STATUS_GAUSSIAN_OVERFLOW, STATUS_NUM_PAIRS = 3, 4, 5
GEOM_STRIDE = 12

# name -> (restype, argtypes); every symbol declared in include/contextgs_b200.h
_PTR = c_void_p
SIGNATURES = {
    "cgs_abi_version": (c_int, []),
    "cgs_last_error": (ctypes.c_char_p, []),
    "cgs_stage_count": (c_int, []),
    "cgs_stage_name": (ctypes.c_char_p, [c_int]),
    "cgs_stage_timing_enable": (c_int, [c_int]),
    "cgs_stage_timing_read": (c_int, [_PTR, _PTR]),
    "cgs_launch_counts": (c_int, [_PTR, c_int]),
    "cgs_umma_selftest": (c_int, [_PTR, _PTR, c_int, c_int, c_int, _PTR, _PTR, _PTR]),
    "cgs_umma_selftest_ss": (c_int, [_PTR, _PTR, c_int, c_int, c_int, _PTR, _PTR, _PTR]),
    "cgs_umma_selftest_ss_mn": (c_int, [_PTR, _PTR, c_int, c_int, c_int, _PTR, _PTR, _PTR]),
    "cgs_neural_gaussians_umma_forward_train": (c_int, [_PTR, _PTR, c_int] + [_PTR] * 19 + [_PTR, c_size_t, _PTR]),
    "cgs_neural_gaussians_bwd_umma_packed_floats": (c_int, []),
    "cgs_neural_gaussians_save_floats": (c_int, [c_int]),
    "cgs_debug_set": (c_int, [c_int, c_int]),
    "cgs_umma_mma_rate": (c_int, [c_int, c_int, c_int, c_int, _PTR, _PTR]),
    "cgs_neural_gaussians_backward_umma": (c_int, [_PTR, _PTR, c_int] + [_PTR] * 27),
    "cgs_visible_filter": (c_int, [_PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR]),
    "cgs_prefilter_workspace_bytes": (c_size_t, [c_int]),
    "cgs_prefilter_anchors": (c_int, [_PTR, c_int, _PTR, _PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, c_size_t, _PTR]),
    "cgs_render_anchors_forward": (c_int, [_PTR, _PTR, _PTR, c_int, _PTR, c_int] + [_PTR] * 12 + [c_size_t, c_int64] +
                                   [_PTR] * 9 + [c_size_t, _PTR]),
    "cgs_mark_visible": (c_int, [_PTR, c_int, _PTR, _PTR, _PTR]),
    "cgs_raster_workspace_bytes": (c_size_t, [c_int, c_int64, c_int, c_int]),
    "cgs_rasterize_forward": (c_int, [_PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, c_int64, _PTR, _PTR, _PTR, _PTR, _PTR,
                                      _PTR, _PTR, _PTR, _PTR, c_size_t, _PTR]),
    "cgs_rasterize_forward_dev": (c_int, [_PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, c_int64, _PTR, _PTR, _PTR, _PTR,
                                          _PTR, _PTR, _PTR, _PTR, _PTR, c_size_t, _PTR]),
    "cgs_raster_backward_workspace_bytes": (c_size_t, [c_int]),
    "cgs_rasterize_backward": (c_int, [_PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR,
                                       _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, c_size_t, _PTR]),
    "cgs_neural_gaussians_packed_floats": (c_int, []),
    "cgs_neural_gaussians_workspace_bytes": (c_size_t, [c_int]),
    "cgs_neural_gaussians_forward": (c_int, [_PTR, _PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR,
                                             _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, c_size_t, _PTR]),
    "cgs_neural_gaussians_umma_packed_floats": (c_int, []),
    "cgs_neural_gaussians_umma_workspace_bytes": (c_size_t, [c_int]),
    "cgs_neural_gaussians_umma_forward": (c_int, [_PTR, _PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR,
                                                  _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, c_size_t, _PTR]),
    "cgs_neural_gaussians_backward_packed_floats": (c_int, []),
    "cgs_neural_gaussians_backward_workspace_bytes": (c_size_t, [c_int]),
    "cgs_neural_gaussians_backward": (c_int, [_PTR, _PTR, _PTR, c_int] + [_PTR] * 19 + [c_size_t, _PTR]),
    "cgs_neural_gaussians_umma_forward_dev": (c_int, [_PTR, _PTR, c_int, _PTR, c_int64, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR,
                                                      _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, c_size_t,
                                                      _PTR]),
    "cgs_compact_positive_i32": (c_int, [_PTR, c_int, _PTR, _PTR, _PTR, c_size_t, _PTR]),
    "cgs_compact_workspace_bytes": (c_size_t, [c_int]),
    "cgs_compact_indices": (c_int, [_PTR, c_int, _PTR, _PTR, _PTR, c_size_t, _PTR]),
    "cgs_eb_param_floats": (c_int, []),
    "cgs_eb_forward": (c_int, [_PTR, c_int, _PTR, _PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR]),
    "cgs_context_level_packed_floats": (c_int, [c_int]),
    "cgs_context_level_forward": (c_int, [c_int, _PTR, _PTR, _PTR, _PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR,
                                          _PTR, _PTR, ctypes.c_float, ctypes.c_float, ctypes.c_float, _PTR, _PTR,
                                          _PTR, _PTR, _PTR, _PTR]),
    "cgs_context_level_umma_packed_floats": (c_int, [c_int]),
    "cgs_context_level_umma_forward": (c_int, [c_int, _PTR, _PTR, _PTR, _PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR,
                                               _PTR, _PTR, ctypes.c_float, ctypes.c_float, ctypes.c_float, _PTR, _PTR,
                                               _PTR, _PTR, _PTR, _PTR, _PTR]),
    "cgs_context_level_umma_forward_ex": (c_int, [c_int, _PTR, _PTR, _PTR, _PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR,
                                                  _PTR, _PTR, ctypes.c_float, ctypes.c_float, ctypes.c_float, _PTR, _PTR,
                                                  _PTR, _PTR, _PTR, _PTR, _PTR, c_int, _PTR, _PTR]),
    "cgs_context_level_umma_forward_train": (c_int, [c_int, _PTR, _PTR, _PTR, _PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR,
                                                     _PTR, _PTR, ctypes.c_float, ctypes.c_float, ctypes.c_float, _PTR, _PTR,
                                                     _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR]),
    "cgs_context_level_bwd_umma_packed_floats": (c_int, [c_int]),
    "cgs_context_level_backward_umma": (c_int, [c_int, _PTR, _PTR, _PTR, _PTR, c_int] + [_PTR] * 8 + [ctypes.c_float] * 3 +
                                        [_PTR, ctypes.c_float] + [_PTR] * 13 + [c_int, _PTR]),
    "cgs_codec_gauss_stream_capacity": (c_int64, [c_int, c_int]),
    "cgs_codec_phi_table": (c_int, [_PTR, c_int, _PTR, _PTR]),
    "cgs_codec_gauss_level_chunks": (c_int64, [c_int, _PTR, _PTR]),
    "cgs_codec_gauss_level_minmax": (c_int, [_PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR]),
    "cgs_codec_gauss_level_encode": (c_int, [_PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR,
                                             _PTR]),
    "cgs_codec_gauss_level_pack": (c_int, [c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR]),
    "cgs_codec_gauss_level_decode": (c_int, [_PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR,
                                             _PTR, _PTR]),
    "cgs_codec_table_encode": (c_int, [_PTR, c_int, c_int, c_int, _PTR, c_int, c_int, _PTR, c_int64, _PTR, _PTR, _PTR]),
    "cgs_codec_table_decode": (c_int, [_PTR, _PTR, _PTR, c_int, c_int, c_int, _PTR, _PTR, c_int, c_int, _PTR, _PTR]),
    "cgs_codec_pack_streams": (c_int, [_PTR, c_int64, _PTR, _PTR, c_int, _PTR, _PTR]),
    "cgs_training_statis_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cgs_training_statis": (c_int, [c_int, c_int, _PTR, _PTR, c_int, _PTR, _PTR, _PTR, c_int, _PTR, _PTR, _PTR, _PTR, _PTR,
                                    c_size_t, _PTR]),
    "cgs_anchor_growing_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "cgs_anchor_growing": (c_int, [_PTR, _PTR, _PTR, c_int, _PTR, c_int, _PTR, c_int, _PTR, c_int, c_int, ctypes.c_float,
                                   c_int, _PTR, _PTR, _PTR, c_int, _PTR, _PTR, c_size_t, _PTR]),
    "cgs_knn3_workspace_bytes": (c_size_t, [c_int]),
    "cgs_knn3_mean_dist2": (c_int, [_PTR, c_int, _PTR, ctypes.c_float, _PTR, _PTR, _PTR, c_size_t, _PTR]),
    "cgs_l1_ssim_forward": (c_int, [_PTR, _PTR, c_int, c_int, _PTR, _PTR, _PTR, _PTR, _PTR]),
    "cgs_l1_ssim_backward": (c_int, [_PTR, _PTR, c_int, c_int, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR]),
    "cgs_context_level_backward_packed_floats": (c_int, [c_int]),
    "cgs_context_level_backward": (c_int, [c_int, _PTR, _PTR, _PTR, _PTR, c_int] + [_PTR] * 8 +
                                   [ctypes.c_float] * 3 + [_PTR, ctypes.c_float] + [_PTR] * 9),
    "cgs_context_level_backward_rows": (c_int, [c_int, _PTR, _PTR, _PTR, _PTR, _PTR, c_int, c_int] + [_PTR] * 8 +
                                        [ctypes.c_float] * 3 + [_PTR, ctypes.c_float] + [_PTR] * 9),
    "cgs_eb_backward": (c_int, [_PTR, c_int, _PTR, c_int, _PTR, _PTR, ctypes.c_float, _PTR, _PTR, _PTR]),
    "cgs_gaussian_bits_forward": (c_int, [_PTR, _PTR, _PTR, _PTR, c_int, ctypes.c_float, c_int64, c_int, _PTR, _PTR]),
    "cgs_gaussian_bits_backward": (c_int, [_PTR, _PTR, _PTR, _PTR, c_int, ctypes.c_float, c_int64, c_int, _PTR, _PTR,
                                           _PTR, _PTR, _PTR, _PTR]),
    "cgs_ste_multistep": (c_int, [_PTR, _PTR, c_int64, c_int, _PTR, _PTR]),
    "cgs_quantize_anchor": (c_int, [_PTR, _PTR, _PTR, c_int64, _PTR, _PTR, _PTR]),
    "cgs_unique_voxels_workspace_bytes": (c_size_t, [c_int]),
    "cgs_unique_voxels": (c_int, [_PTR, _PTR, c_int, ctypes.c_float, ctypes.c_float, _PTR, _PTR, _PTR, _PTR, c_size_t,
                                  _PTR]),
    "cgs_sort_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "cgs_sort_pairs_u32": (c_int, [_PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, c_int64, c_int, c_int, _PTR, c_size_t,
                                   _PTR]),
}

_lib = None


class CgsError(RuntimeError):
    pass


def lib():
    """Load the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CgsError(
                f"{LIB_PATH} is missing: build it with `python -m contextgs_b200.build` "
                "(there is no CPU or PyTorch fallback for the contextgs_b200 hot path)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI and this table disagree
            fn.restype = res
            fn.argtypes = args
        if L.cgs_abi_version() != 1:
            raise CgsError("libcontextgs_b200.so ABI version mismatch")
        _lib = L
    return _lib


_deferred = []   # (int32 device flag, message): error flags of kernels whose result nobody reads back right away


def deferred_error_check(flag, message):
    """Park a device-side error flag; it is inspected at the next natural synchronisation point (`raise_deferred`,
    called by the forward passes right after their own read-back) instead of forcing one here."""
    _deferred.append((flag, message))
    if len(_deferred) > 64:
        raise_deferred()


def raise_deferred():
    while _deferred:
        flag, message = _deferred.pop()
        if int(flag.item()):
            _deferred.clear()
            raise CgsError(message)


def check(code, what=""):
    if code != 0:
        msg = lib().cgs_last_error()
        raise CgsError(f"{what} failed with code {code}: {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL).  The tensor must be contiguous."""
    if t is None:
        return None
    assert t.is_contiguous(), "contextgs_b200 kernels need contiguous tensors"
    return t.data_ptr()


def stage_names():
    L = lib()
    return [L.cgs_stage_name(i).decode() for i in range(L.cgs_stage_count())]


def launch_counts(reset=False):
    """Kernels launched per stage since the last reset (dict name -> count)."""
    L = lib()
    n = L.cgs_stage_count()
    buf = (ctypes.c_int64 * n)()
    check(L.cgs_launch_counts(buf, int(reset)), "cgs_launch_counts")
    return dict(zip(stage_names(), list(buf)))


def stage_timing(enable):
    check(lib().cgs_stage_timing_enable(int(enable)), "cgs_stage_timing_enable")


def stage_timing_read():
    """(ms_sum, scopes) dicts per stage; synchronises on the recorded events."""
    L = lib()
    n = L.cgs_stage_count()
    ms = (ctypes.c_double * n)()
    sc = (ctypes.c_int64 * n)()
    check(L.cgs_stage_timing_read(ms, sc), "cgs_stage_timing_read")
    names = stage_names()
    return dict(zip(names, list(ms))), dict(zip(names, list(sc)))


def stream_ptr():
    """Raw handle of torch's current CUDA stream on the current device (the C ABI takes it as void*)."""
    import torch
    try:   # two C calls (~1 us); torch.cuda.current_stream() builds a Stream object (~15 us, three times per frame)
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


def object_cache(obj):
    """Per-object cache dict for derived tensors (packed weights, ...).  It lives ON the object and dies with it: a
    module-level dict keyed by id(obj) would hand a later object that reuses the address -- and, through the caching
    allocator, the same data pointers at the same tensor version -- the previous object's entries."""
    d = getattr(obj, "_cgs_cache", None)
    if d is None:
        d = {}
        try:
            object.__setattr__(obj, "_cgs_cache", d)
        except Exception:
            pass          # objects without attribute storage simply do not cache
    return d
