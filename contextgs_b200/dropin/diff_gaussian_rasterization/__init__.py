"""`diff_gaussian_rasterization` drop-in: put `contextgs_b200/dropin` on PYTHONPATH (or call
`contextgs_b200.install()`) and the reference's
`from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`
(gaussian_renderer/__init__.py:20) resolves to the B200-native implementation."""
from contextgs_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]
