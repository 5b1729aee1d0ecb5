"""CPU: the C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

from contextgs_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        names |= set(re.findall(r"CGS_API[^;(]*?\b(cgs_\w+)\s*\(", src))
    return names


def test_library_builds_loads_and_exports_every_declared_symbol():
    build.build_library()
    L = ctypes.CDLL(_lib.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 10
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    # the ctypes signature table covers exactly the declared ABI
    assert set(_lib.SIGNATURES) == decl
    assert _lib.lib().cgs_abi_version() == 1


def test_settings_struct_layout_matches_header():
    # 2 ints + 2 floats + 3 + 1 + 16 + 16 floats + int + 3 floats + 2 ints = 46 words
    assert ctypes.sizeof(_lib.RasterSettings) == 46 * 4


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcontextgs_b200.so")
    try:
        _lib.lib()
    except _lib.CgsError as e:
        assert "no CPU or PyTorch fallback" in str(e)
    else:
        raise AssertionError("expected CgsError")


def test_object_cache_lives_on_the_object():
    """Derived-tensor caches must not outlive their model: a module-level dict keyed by id() handed a NEW model that
    reused the address (and, via the caching allocator, the data pointers) the previous model's packed weights."""
    import gc
    import types
    from contextgs_b200 import _lib
    a = types.SimpleNamespace()
    _lib.object_cache(a)["k"] = 1
    assert _lib.object_cache(a) == {"k": 1} and a._cgs_cache is _lib.object_cache(a)
    ident = id(a)
    del a
    gc.collect()
    fresh = [types.SimpleNamespace() for _ in range(64)]         # one of them very likely reuses the address
    assert all(_lib.object_cache(o) == {} for o in fresh), ident
    import contextgs_b200.context_model as cm
    import contextgs_b200.neural_gaussians as ng
    assert not any(n.endswith("_cache") or n.endswith("_cache_umma") for m in (cm, ng) for n in vars(m)
                   if isinstance(getattr(m, n), dict))


def test_dropin_packages_resolve_the_reference_imports():
    """gaussian_renderer/__init__.py:20 and scene/gaussian_model.py:22 import these names."""
    import importlib
    import sys
    import contextgs_b200
    contextgs_b200.install()
    try:
        for mod in ("diff_gaussian_rasterization", "simple_knn", "simple_knn._C"):
            sys.modules.pop(mod, None)
        r = importlib.import_module("diff_gaussian_rasterization")
        k = importlib.import_module("simple_knn._C")
        from contextgs_b200.knn import distCUDA2
        from contextgs_b200.rasterizer import GaussianRasterizer
        assert r.GaussianRasterizer is GaussianRasterizer and k.distCUDA2 is distCUDA2
    finally:
        for mod in ("diff_gaussian_rasterization", "simple_knn", "simple_knn._C"):
            sys.modules.pop(mod, None)
