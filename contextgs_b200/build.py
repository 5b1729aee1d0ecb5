"""Builds libcontextgs_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

    python -m contextgs_b200.build [--force] [--verbose]

The shared library carries every hand-written kernel plus the C ABI declared in
include/contextgs_b200.h; Python reaches it through ctypes (contextgs_b200/_lib.py).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libcontextgs_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr"]

# per-file extra flags.  raster_preprocess.cu: no FMA contraction so that tile/sort indices are
# bit-reproducible on the CPU oracle (see the header of that file).
SOURCES = {
    "raster_api.cu": [],
    "raster_preprocess.cu": ["-fmad=false"],
    "radix_sort.cu": [],
    "raster_binning.cu": [],
    "raster_render.cu": [],
    # -fmad=false: the elementwise value-defining code (x + noise*Q, anchor + offset*scale, interval
    # arithmetic ...) rounds once per operation like the PyTorch expressions it replaces; the GEMM
    # inner loops use explicit fmaf() and are unaffected.
    "neural_gaussians.cu": ["-fmad=false"],
    "context_model.cu": ["-fmad=false"],
    "umma_selftest.cu": [],
    "neural_gaussians_umma.cu": ["-fmad=false"],
    "neural_gaussians_bwd.cu": [],
    "neural_gaussians_bwd_umma.cu": [],
    "context_model_bwd.cu": [],
    "context_model_bwd_umma.cu": [],
    "context_model_umma.cu": ["-fmad=false"],
    "level_divide.cu": ["-fmad=false"],
    "entropy_codec.cu": ["-fmad=false"],
    "loss.cu": [],
    "anchor_growing.cu": ["-fmad=false"],
    "knn.cu": ["-fmad=false"],
}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.exists(c) or c == "nvcc"):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "contextgs_b200.h"))
    headers.append(os.path.abspath(__file__))
    sources = {k: v for k, v in SOURCES.items() if os.path.exists(os.path.join(CSRC, k))}
    jobs = []
    objs = []
    for src, extra in sources.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    print(out)
    if jobs or force or _stale(LIB, objs):
        # visibility: kernels/helpers hidden, extern "C" cgs_* exported explicitly via the attribute
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs
        run(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
