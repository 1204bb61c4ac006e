"""CPU: the torch expression csrc/depth_normal.cu is tested against on the GPU (ibgs_b200.gaussian_renderer.depth_normal_torch)
equals the REFERENCE's own functions -- render_normal -> normal_from_depth_image (utils/graphics_utils.py:38-83) + the
renormalisation of gaussian_renderer/__init__.py:332-335 -- in float64, up to the reference's own float32 pixel grid (depth2point_cam builds x / (W - 1) in float32 and multiplies
(W - 1) back: 1e-7 relative on the coordinates).  Together with
test_gpu_fast_glue.py::test_depth_normal_kernel_equals_torch_expressions this pins the kernel on the reference."""
import types

import pytest
import torch

import refglue as G

pytestmark = pytest.mark.skipif(not G.available(), reason="reference glue not staged (run oracle/stage_ref_py.py)")


@pytest.mark.parametrize("H,W", [(5, 7), (48, 64)])
def test_depth_normal_expression_equals_reference_functions(H, W):
    G._paths()
    from utils.graphics_utils import normal_from_depth_image           # the reference's file, staged verbatim
    from ibgs_b200.gaussian_renderer import depth_normal_torch
    g = torch.Generator().manual_seed(H * W)
    depth = (2.0 + torch.rand(H, W, generator=g, dtype=torch.float64)).requires_grad_(True)
    fx, fy, cx, cy = 61.5, 59.25, W / 2 - 0.3, H / 2 + 0.4
    K = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float64)      # Camera.get_calib_matrix_nerf
    ref = normal_from_depth_image(depth, K, torch.eye(4, dtype=torch.float64)).permute(2, 0, 1)
    ref = ref / (torch.norm(ref, dim=0, keepdim=True) + 1e-8)            # gaussian_renderer/__init__.py:333-335
    cam = types.SimpleNamespace(Fx=fx, Fy=fy, Cx=cx, Cy=cy)
    d2 = depth.detach().clone().requires_grad_(True)
    mine = depth_normal_torch(cam, d2)
    assert mine.shape == ref.shape and torch.allclose(mine, ref, rtol=0, atol=1e-6)
    cot = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    ref.backward(cot)
    mine.backward(cot)
    rel = ((d2.grad - depth.grad).norm() / depth.grad.norm()).item()
    assert rel <= 1e-5, rel          # (the float32 grid again, amplified by the normalisation of small cross products)
