"""ibgs_b200 -- B200-native (sm_100a) implementation of the IBGS planar Gaussian rasterizer hot path.

Public surface (mirrors the reference's two extension packages):
    ibgs_b200.diff_plane_rasterization  -> GaussianRasterizationSettings, GaussianRasterizer
    ibgs_b200.simple_knn._C             -> distCUDA2
`install_dropin()` registers them under the reference's import names so that
`gaussian_renderer/__init__.py:4-5` and `scene/gaussian_model.py:20` run unchanged.
"""
import sys

__version__ = "0.1.0"


def install_dropin():
    """Make `import diff_plane_rasterization` / `from simple_knn._C import distCUDA2` resolve to this package."""
    from . import diff_plane_rasterization as dpr
    from . import simple_knn as knn
    from .simple_knn import _C as knn_c
    sys.modules["diff_plane_rasterization"] = dpr
    sys.modules["simple_knn"] = knn
    sys.modules["simple_knn._C"] = knn_c
    return dpr, knn_c
