"""Synthetic inputs for the colour-aggregation tests + the reference's own torch expressions for the part that
ibgs_b200.color_aggregation.color_features replaces (color_aggregation_network.py:196-206 and :121-131), written with the
reference's unchanged ColorFusionResidualNet module."""
import torch


def random_render_pkg(H, W, M=4, seed=0, device="cpu", dtype=torch.float32, dead_views=0):
    """A stand-in for gaussian_renderer.render()'s dict: only the keys fuse_color reads."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64)
    cam_feat = r(M, 4, H, W) * 0.8
    hole = r(M, 1, H, W) < 0.25                       # pixels a source view does not see: all four features zero
    cam_feat = torch.where(hole, torch.zeros_like(cam_feat), cam_feat)
    warped = r(M, 3, H, W) * (~hole)
    if dead_views:
        warped[M - dead_views:] = 0
        cam_feat[M - dead_views:] = 0
    ray = torch.nn.functional.normalize(torch.randn(3, H, W, generator=g, dtype=torch.float64), dim=0)
    pkg = dict(render=r(3, H, W), warped_image=warped.reshape(M * 3, H, W), cam_feat=cam_feat.reshape(M * 4, H, W),
               min_depth_diff=r(1, H, W) * 1.2, camera_ray=ray.reshape(3, H * W),
               use_first_src_frame_mask=(r(1, H, W) < 0.7).to(torch.float64))
    return {k: v.to(device=device, dtype=dtype).contiguous() for k, v in pkg.items()}


def torch_color_features(net, pkg, n_views):
    """(1, 38, H, W) conv-decoder input exactly as fuse_color + ColorFusionResidualNet.forward build it."""
    rendered = pkg["render"]
    _, H, W = rendered.shape
    feat = pkg["cam_feat"].view(-1, 4, H, W).permute(2, 3, 0, 1)[:, :, :n_views]
    warped = pkg["warped_image"].view(-1, 3, H, W).permute(2, 3, 0, 1)[:, :, :n_views]
    valid = (torch.sum(feat, dim=-1, keepdim=True) > 0.0).to(rendered.dtype)
    residual = (warped - rendered.permute(1, 2, 0).unsqueeze(2)) * valid
    feat = torch.cat([residual, feat], dim=-1)
    x_views = feat.reshape(-1, feat.shape[2], feat.shape[3]).contiguous()
    B, M, _ = x_views.shape
    features = net.per_view_mlp(x_views.view(B * M, -1)).view(B, M, -1)
    aggregated = features.mean(dim=1) if net.feat_aggregate_mode == "mean" else features.max(dim=1).values
    ray = pkg["camera_ray"].view(3, H, W).reshape(3, -1).T
    col = rendered.permute(1, 2, 0).reshape(-1, 3)
    return torch.cat([aggregated.T.view(1, 32, H, W), ray.T.view(1, 3, H, W), col.T.view(1, 3, H, W)], dim=1)
