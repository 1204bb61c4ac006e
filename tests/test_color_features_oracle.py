"""CPU: the numpy oracle of the colour-aggregation front half (oracle/color_features_oracle.py) against torch autograd of
the reference's own ColorFusionResidualNet module in float64 -- this pins the oracle the GPU test uses as its checker."""
import numpy as np
import pytest
import torch

import colorfeat_ref as CR
import refglue as G
from oracle import color_features_oracle as O

pytestmark = pytest.mark.skipif(not G.available(), reason="reference glue not staged (run oracle/stage_ref_py.py)")


@pytest.mark.parametrize("mode,nv", [("mean", 3), ("max", 2), ("mean", 1)])
def test_oracle_matches_reference_module_float64(mode, nv):
    G._paths()
    import color_aggregation_network as CAN
    H, W = 9, 13
    torch.manual_seed(3)
    net = CAN.ColorFusionResidualNet(height=H, width=W, feat_aggregate_mode=mode).double()
    pkg = CR.random_render_pkg(H, W, M=4, seed=11, dtype=torch.float64)
    leaves = {k: pkg[k].clone().requires_grad_(True) for k in ("render", "warped_image")}
    p2 = dict(pkg, **leaves)
    want = CR.torch_color_features(net, p2, nv)
    l1, l2 = net.per_view_mlp[0], net.per_view_mlp[2]
    N = H * W
    out, cache = O.forward(pkg["warped_image"].view(4, 3, N).numpy(), pkg["cam_feat"].view(4, 4, N).numpy(),
                           pkg["render"].view(3, N).numpy(), pkg["camera_ray"].numpy(), l1.weight.detach().numpy(),
                           l1.bias.detach().numpy(), l2.weight.detach().numpy(), l2.bias.detach().numpy(), nv, mode)
    assert np.allclose(out.T.reshape(1, 38, H, W), want.detach().numpy(), rtol=1e-12, atol=1e-12)
    g = torch.randn(want.shape, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    want.backward(g)
    d_warped, d_rendered, dw1, db1, dw2, db2 = O.backward(cache, g[0].reshape(38, N).T.numpy())
    assert np.allclose(d_rendered.reshape(3, H, W), leaves["render"].grad.numpy(), rtol=1e-9, atol=1e-12)
    gw = leaves["warped_image"].grad.view(4, 3, N).numpy()
    assert np.allclose(d_warped, gw[:nv], rtol=1e-9, atol=1e-12) and not gw[nv:].any()
    for got, p in ((dw1, l1.weight), (db1, l1.bias), (dw2, l2.weight), (db2, l2.bias)):
        assert np.allclose(got, p.grad.numpy(), rtol=1e-9, atol=1e-12)
