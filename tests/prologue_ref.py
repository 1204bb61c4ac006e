"""The reference's own torch expressions for the per-view parameter prologue (what ibgs_b200.fused replaces), written
as one function so tests can differentiate through them with autograd.  Lines follow scene/gaussian_model.py:127-147,
166-173 and gaussian_renderer/__init__.py:304-315 of the reference."""
import torch


def torch_prologue(xyz, opacity_raw, scaling_raw, rotation_raw, fdc, frest, normal_raw, offset, V, cam):
    opacity = torch.sigmoid(opacity_raw)
    scales = torch.exp(scaling_raw)
    rotations = torch.nn.functional.normalize(rotation_raw)
    shs = torch.cat((fdc, frest), dim=1)
    # get_normal(view_cam)
    offset_global = offset
    normal_global = normal_raw / torch.norm(normal_raw, dim=1, keepdim=True)
    gaussian_to_cam_global = cam - xyz
    neg_mask = (normal_global * gaussian_to_cam_global).sum(-1) < 0.0
    normal_global = torch.where(neg_mask[:, None], -normal_global, normal_global)   # in-place flip in the reference
    offset_global = offset_global * (neg_mask.to(xyz.dtype) * -2 + 1).unsqueeze(-1)
    # render(): all_map
    local_normal = normal_global @ V[:3, :3]
    global_distance = -(normal_global * xyz).sum(-1)
    global_distance = global_distance + offset_global.squeeze()
    local_distance = global_distance - torch.sum(local_normal * V[[3], :3], dim=1)
    local_distance = local_distance.abs()
    all_map = torch.zeros((xyz.shape[0], 5), device=xyz.device, dtype=xyz.dtype)
    all_map[:, :3] = local_normal
    all_map[:, 3] = 1.0
    all_map[:, 4] = local_distance
    return opacity, scales, rotations, shs, all_map


def torch_prologue_smallest_axis(xyz, opacity_raw, scaling_raw, rotation_raw, fdc, frest, V, cam):
    """learnt_normal=False: get_smallest_axis / get_normal_w_smallest_axis (scene/gaussian_model.py:149-161) with
    pytorch3d's quaternion_to_matrix, then the same all_map lines without the offset."""
    opacity = torch.sigmoid(opacity_raw)
    scales = torch.exp(scaling_raw)
    rotations = torch.nn.functional.normalize(rotation_raw)
    shs = torch.cat((fdc, frest), dim=1)
    r, i, j, k = torch.unbind(rotations, -1)
    two_s = 2.0 / (rotations * rotations).sum(-1)
    R = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1).reshape(-1, 3, 3)
    idx = scales.min(dim=-1)[1][..., None, None].expand(-1, 3, -1)
    normal_global = R.gather(2, idx).squeeze(dim=2)
    neg_mask = (normal_global * (cam - xyz)).sum(-1) < 0.0
    normal_global = torch.where(neg_mask[:, None], -normal_global, normal_global)
    local_normal = normal_global @ V[:3, :3]
    global_distance = -(normal_global * xyz).sum(-1)
    local_distance = (global_distance - torch.sum(local_normal * V[[3], :3], dim=1)).abs()
    all_map = torch.zeros((xyz.shape[0], 5), device=xyz.device, dtype=xyz.dtype)
    all_map[:, :3] = local_normal
    all_map[:, 3] = 1.0
    all_map[:, 4] = local_distance
    return opacity, scales, rotations, shs, all_map


def random_params(P, K=9, seed=0, device="cpu", dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    p = dict(xyz=r(P, 3) * 3, opacity_raw=r(P, 1) * 2, scaling_raw=r(P, 3) * 0.7 - 3, rotation_raw=r(P, 4),
             fdc=r(P, 1, 3), frest=r(P, K - 1, 3) * 0.2, normal_raw=r(P, 3) * 1.5, offset=r(P, 1) * 0.05)
    th = 0.3
    R = torch.tensor([[1, 0, 0], [0, torch.cos(torch.tensor(th)), -torch.sin(torch.tensor(th))],
                      [0, torch.sin(torch.tensor(th)), torch.cos(torch.tensor(th))]], dtype=torch.float64)
    W2C = torch.eye(4, dtype=torch.float64)
    W2C[:3, :3] = R
    W2C[:3, 3] = torch.tensor([0.2, -0.1, 4.0], dtype=torch.float64)
    V = W2C.t().contiguous()                      # world_view_transform (scene/cameras.py:102)
    cam = torch.linalg.inv(V)[3, :3].contiguous()  # camera_center (:105)
    p.update(V=V, cam=cam)
    return {k: v.to(device=device, dtype=dtype) for k, v in p.items()}
