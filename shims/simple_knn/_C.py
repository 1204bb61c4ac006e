"""`from simple_knn._C import distCUDA2` (scene/gaussian_model.py:20) -> B200-native implementation."""
from ibgs_b200.simple_knn._C import distCUDA2  # noqa: F401
