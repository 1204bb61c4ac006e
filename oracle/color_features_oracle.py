"""float64 numpy restatement of the colour-aggregation front half (forward AND hand-derived backward).

TEST INFRASTRUCTURE ONLY (checker for ibgs_b200/csrc/color_features.cu in tests/).  Follows the reference's torch
expressions: color_aggregation_network.py:196-206 (valid mask, residual, feature cat in fuse_color) and :121-131
(per_view_mlp, view aggregation, cat with ray / colour in ColorFusionResidualNet.forward).  Pinned in
tests/test_color_features_oracle.py against torch autograd of the reference's own module (CPU, float64).
"""
import numpy as np


def forward(warped, cam_feat, rendered, ray, w1, b1, w2, b2, n_views, mode="mean"):
    """warped (V,3,N), cam_feat (V,4,N), rendered (3,N), ray (3,N) -> (N, 38): 32 aggregated | 3 ray | 3 colour."""
    f = lambda t: np.asarray(t, dtype=np.float64)
    warped, cam_feat, rendered, ray, w1, b1, w2, b2 = map(f, (warped, cam_feat, rendered, ray, w1, b1, w2, b2))
    V = n_views
    valid = (cam_feat[:V].sum(1) > 0.0).astype(np.float64)                      # (V, N)
    res = (warped[:V] - rendered[None]) * valid[:, None]                        # (V, 3, N)
    x = np.concatenate([res, cam_feat[:V]], 1).transpose(0, 2, 1)               # (V, N, 7)
    pre1 = x @ w1.T + b1
    h1 = np.maximum(pre1, 0.0)
    pre2 = h1 @ w2.T + b2
    h2 = np.maximum(pre2, 0.0)                                                  # (V, N, 32)
    agg = h2.mean(0) if mode == "mean" else h2.max(0)
    out = np.concatenate([agg, ray.T, rendered.T], 1)
    cache = dict(valid=valid, x=x, h1=h1, h2=h2, w1=w1, w2=w2, V=V, mode=mode)
    return out, cache


def backward(cache, g):
    """g (N, 38) -> d_warped (V,3,N), d_rendered (3,N), dw1, db1, dw2, db2."""
    g = np.asarray(g, dtype=np.float64)
    V, h1, h2, x, valid, w1, w2 = cache["V"], cache["h1"], cache["h2"], cache["x"], cache["valid"], cache["w1"], cache["w2"]
    if cache["mode"] == "mean":
        g_h2 = np.broadcast_to(g[None, :, :32] / V, h2.shape)
    else:
        arg = h2.argmax(0)                                                      # first maximum
        g_h2 = (np.arange(V)[:, None, None] == arg[None]) * g[None, :, :32]
    g2 = g_h2 * (h2 > 0)
    dw2 = np.einsum("vni,vnj->ij", g2, h1)
    db2 = g2.sum((0, 1))
    g1 = (g2 @ w2) * (h1 > 0)
    dw1 = np.einsum("vni,vnk->ik", g1, x)
    db1 = g1.sum((0, 1))
    dx = g1 @ w1                                                                # (V, N, 7)
    d_res = dx[:, :, :3].transpose(0, 2, 1) * valid[:, None]                    # (V, 3, N)
    d_rendered = g[:, 35:38].T - d_res.sum(0)
    return d_res, d_rendered, dw1, db1, dw2, db2
