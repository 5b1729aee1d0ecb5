"""`simple_knn` drop-in (same mechanism as dropin/diff_gaussian_rasterization): with `contextgs_b200/dropin` on
PYTHONPATH the reference's `from simple_knn._C import distCUDA2` (scene/gaussian_model.py:22) resolves to
contextgs_b200.knn.distCUDA2."""
