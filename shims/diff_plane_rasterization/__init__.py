"""Import-name shim: `import diff_plane_rasterization` (gaussian_renderer/__init__.py:5-6 of the reference) resolves
to the B200-native package when this directory is on PYTHONPATH together with the repository root."""
from ibgs_b200.diff_plane_rasterization import *  # noqa: F401,F403
from ibgs_b200.diff_plane_rasterization import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                                rasterize_gaussians)
