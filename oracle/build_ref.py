"""TEST INFRASTRUCTURE.  Byte-compiles the reference's OWN Python modules of the hot path, from the sources where they
lie under /root/reference, into oracle/_ref/*.refbin (git-ignored, travels to the GPU box like a built .so).  No reference
source is copied into the repository: only code objects produced by this recipe, and only in this container (the GPU
box has no /root/reference and uses the prebuilt files).

    python -m oracle.build_ref

oracle/ref_loader.py turns the code objects back into modules (`gaussian_renderer`, `scene.gaussian_model`,
`utils.*`) so that tests can run the reference's unmodified `render` / `prefilter_voxel` /
`multi_scale_generating` against the drop-in rasterizer, and bench.py can time the reference's own PyTorch entropy
path on CUDA tensors as a labelled baseline."""
import os
import py_compile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")

# dotted module name -> path under /root/reference
MODULES = {
    "utils.general_utils": "utils/general_utils.py",
    "utils.graphics_utils": "utils/graphics_utils.py",
    "utils.system_utils": "utils/system_utils.py",
    "utils.encodings": "utils/encodings.py",
    "utils.entropy_models": "utils/entropy_models.py",
    "utils.multi_level": "utils/multi_level.py",
    "utils.loss_utils": "utils/loss_utils.py",
    "scene.gaussian_model": "scene/gaussian_model.py",
    "gaussian_renderer": "gaussian_renderer/__init__.py",
}


def build(verbose=False):
    """Returns the list of files written ([] when /root/reference is absent: keep whatever is there)."""
    if not os.path.isdir(REF):
        return []
    os.makedirs(OUT, exist_ok=True)
    done = []
    for name, rel in MODULES.items():
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, name + ".refbin")
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=f"reference:{rel}", doraise=True)
        done.append(dst)
        if verbose:
            print(dst)
    return done


if __name__ == "__main__":
    build(verbose=True)
