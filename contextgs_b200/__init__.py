"""contextgs_b200 -- B200-native (sm_100a) implementation of ContextGS's data-parallel hot path:
anchor -> neural-Gaussian generation, the tile-based differentiable rasterizer and the anchor-level
context / entropy model, behind the reference's own Python surface.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"


def install():
    """Make the reference's imports resolve to this implementation:
        from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    (gaussian_renderer/__init__.py:20) and
        from simple_knn._C import distCUDA2
    (scene/gaussian_model.py:22).  See INTEGRATION.md for the other hooks."""
    import os
    import sys
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
    if p not in sys.path:
        sys.path.insert(0, p)
