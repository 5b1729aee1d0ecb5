from contextgs_b200.knn import distCUDA2  # noqa: F401

__all__ = ["distCUDA2"]
