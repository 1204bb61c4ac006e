"""Deterministic synthetic scenes for parity tests and benchmarks.

Everything is generated on the CPU with numpy from a fixed seed, so the reference extension, the CUDA
path and the CPU oracle all see bit-identical inputs.  Conventions follow the reference:
  * `viewmatrix` is torch `world_view_transform` = transpose of the true world-to-camera matrix
    (scene/cameras.py:102); `projmatrix` = viewmatrix @ getProjectionMatrix(...)^T (cameras.py:103-104,
    utils/graphics_utils.py:164-184 with znear=0.01, zfar=100);
  * `all_map` = [local normal (3), 1, |local plane distance|] built exactly like
    gaussian_renderer/__init__.py:304-315 from world normals flipped towards the camera
    (scene/gaussian_model.py:166-173), offset = 0;
  * `ref_to_src_list[i]` = world_to_src_i @ ref_to_world (true matrices, row-major),
    `src_cam_pos[i]` = translation column of src_to_world (gaussian_renderer/__init__.py:256-263).
"""
import math

import numpy as np
import torch

CONFIGS = {
    # name: (P, W, H, kind)
    "cfg1": (10_000, 256, 256, "uniform"),
    "cfg2": (500_000, 1920, 1080, "uniform"),
    "cfg3": (3_000_000, 1237, 822, "mip360"),
    "cfg3_1080p": (3_000_000, 1920, 1080, "mip360"),   # the headline "3M @1080p"
    "cfg4": (2_000_000, 977, 545, "mip360"),
    "cfg5": (6_000_000, 3840, 2160, "dense"),
    "tiny": (2_000, 96, 64, "uniform"),
}


def projection_matrix(znear, zfar, fovx, fovy):
    """utils/graphics_utils.py:164-184"""
    t = math.tan(fovy / 2) * znear
    r = math.tan(fovx / 2) * znear
    P = np.zeros((4, 4), dtype=np.float64)
    P[0, 0] = 2.0 * znear / (2 * r)
    P[1, 1] = 2.0 * znear / (2 * t)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def _rot_axis_angle(axis, ang):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def _rigid(R, t):
    M = np.eye(4)
    M[:3, :3] = R
    M[:3, 3] = t
    return M


def make_camera(w2c, W, H, fovx_deg=60.0):
    """Returns the per-view tensors of GaussianRasterizationSettings for a true world-to-camera matrix."""
    fovx = math.radians(fovx_deg)
    tanfovx = math.tan(fovx * 0.5)
    tanfovy = tanfovx * H / W  # fy == fx
    fovy = 2 * math.atan(tanfovy)
    view = np.float32(w2c).T.copy()                      # world_view_transform
    proj = np.float32(projection_matrix(0.01, 100.0, fovx, fovy)).T
    full = (view.astype(np.float32) @ proj.astype(np.float32)).astype(np.float32)
    campos = np.linalg.inv(view.astype(np.float64))[3, :3].astype(np.float32)
    return dict(viewmatrix=torch.from_numpy(view), projmatrix=torch.from_numpy(full),
                campos=torch.from_numpy(campos), tanfovx=tanfovx, tanfovy=tanfovy)


def all_map_for_view(means3D, normals_world, viewmatrix, campos):
    """gaussian_renderer/__init__.py:304-315 with learnt_normal=True, offset=0 (float32 torch ops)."""
    viewmatrix, campos = viewmatrix.to(means3D.device), campos.to(means3D.device)
    n = normals_world / torch.norm(normals_world, dim=1, keepdim=True)
    to_cam = campos.unsqueeze(0) - means3D
    neg = (n * to_cam).sum(-1) < 0.0
    n = torch.where(neg.unsqueeze(-1), -n, n)
    local_n = n @ viewmatrix[:3, :3]
    gdist = -(n * means3D).sum(-1)
    ldist = (gdist - torch.sum(local_n * viewmatrix[[3], :3], dim=1)).abs()
    am = torch.zeros((means3D.shape[0], 5), dtype=torch.float32, device=means3D.device)
    am[:, :3] = local_n
    am[:, 3] = 1.0
    am[:, 4] = ldist
    return am


def make_scene(name="cfg1", P=None, W=None, H=None, kind=None, seed=20251017, nb_src=4, sh_degree=2,
               coherent_frac=0.6, identity_pose=False):
    """Returns a dict of CPU float32 tensors + python scalars describing one reference view, its
    Gaussians and `nb_src` neighbouring source views (without source depths: those are rendered)."""
    if name in CONFIGS:
        P0, W0, H0, kind0 = CONFIGS[name]
        P, W, H, kind = P or P0, W or W0, H or H0, kind or kind0
    rng = np.random.default_rng(seed + sum(ord(c) for c in name))
    fovx = math.radians(60.0)
    tanfovx = math.tan(fovx / 2)
    tanfovy = tanfovx * H / W
    fx = W / (2 * tanfovx)

    # ---- Gaussians in the REFERENCE CAMERA frame -------------------------------------------------
    z = rng.uniform(1.5, 12.0, P)
    x = z * tanfovx * rng.uniform(-1.15, 1.15, P)
    y = z * tanfovy * rng.uniform(-1.15, 1.15, P)
    rpix = np.exp(rng.normal(math.log(1.2), 0.7, P))
    if kind == "mip360":
        blob = rng.random(P) < 0.65
        zb = rng.uniform(2.0, 5.0, P)
        rr = 0.35 * np.sqrt(rng.random(P))
        th = rng.uniform(0, 2 * math.pi, P)
        xb = zb * tanfovx * 2 * rr * np.cos(th)
        yb = zb * tanfovy * 2 * rr * np.sin(th)
        zs = rng.uniform(8.0, 40.0, P)
        xs = zs * tanfovx * rng.uniform(-1.15, 1.15, P)
        ys = zs * tanfovy * rng.uniform(-1.15, 1.15, P)
        x, y, z = np.where(blob, xb, xs), np.where(blob, yb, ys), np.where(blob, zb, zs)
        rpix = np.where(blob, rpix, 4.0 * rpix)
    elif kind == "dense":
        dense = rng.random(P) < 0.5
        cx_px, cy_px = 0.37 * W, 0.58 * H
        px = cx_px + rng.uniform(-64, 64, P)
        py = cy_px + rng.uniform(-64, 64, P)
        xd = (px - W / 2) / fx * z
        yd = (py - H / 2) / fx * z
        x, y = np.where(dense, xd, x), np.where(dense, yd, y)
        rpix = np.where(dense, np.exp(rng.normal(math.log(6.0), 0.5, P)), rpix)
    # 2 % behind / too close to the camera: exercises the near cull (auxiliary.h:156)
    near = rng.random(P) < 0.02
    z = np.where(near, rng.uniform(-1.0, 0.2, P), z)

    # coherent sheets: a fraction of the Gaussians is snapped onto a few planes so that neighbouring
    # views agree on depth (otherwise no source view would ever pass the depth-consistency test)
    normals_cam = rng.normal(size=(P, 3))
    normals_cam /= np.linalg.norm(normals_cam, axis=1, keepdims=True)
    coh = (rng.random(P) < coherent_frac) & ~near
    nsheet = 3
    sheet = rng.integers(0, nsheet, P)
    z0 = np.array([3.0, 5.5, 9.0])[sheet]
    ax = np.array([0.15, -0.10, 0.05])[sheet]
    by = np.array([-0.08, 0.12, 0.20])[sheet]
    zsafe = np.where(near, 1.0, z)
    u, v = x / zsafe, y / zsafe
    zsheet = z0 / (1.0 - ax * u - by * v)          # ray (u,v,1)*z meets the plane z = z0 + ax*x + by*y
    x = np.where(coh, u * zsheet, x)
    y = np.where(coh, v * zsheet, y)
    z = np.where(coh, zsheet, z)
    nsh = np.stack([-ax, -by, np.ones(P)], axis=1)
    nsh /= np.linalg.norm(nsh, axis=1, keepdims=True)
    normals_cam = np.where(coh[:, None], nsh, normals_cam)
    zc = np.maximum(np.abs(z), 0.3)
    long_axis = zc * rpix / fx
    log_scales = np.log(np.stack([long_axis, long_axis * rng.uniform(0.5, 1.0, P), long_axis * 0.1], axis=1))
    perm = rng.permuted(np.tile(np.arange(3), (P, 1)), axis=1)
    log_scales = np.take_along_axis(log_scales, perm, axis=1)
    quat = rng.normal(size=(P, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    opacity = 1.0 / (1.0 + np.exp(-rng.normal(0.0, 2.0, P)))
    K = (sh_degree + 1) ** 2
    Kmax = 9 if sh_degree <= 2 else 16
    shs = np.zeros((P, Kmax, 3))
    shs[:, 0, :] = rng.uniform(-1.5, 1.5, (P, 3))
    shs[:, 1:K, :] = rng.normal(0.0, 0.15, (P, K - 1, 3))

    # ---- reference pose: camera frame -> world ----------------------------------------------------
    if identity_pose:
        w2c = np.eye(4)
    else:
        Rr = _rot_axis_angle(rng.normal(size=3), math.radians(25.0))
        w2c = _rigid(Rr, rng.uniform(-0.5, 0.5, 3))
    c2w = np.linalg.inv(w2c)
    pts_cam = np.stack([x, y, z], axis=1)
    means3D = pts_cam @ c2w[:3, :3].T + c2w[:3, 3]
    normals_world = normals_cam @ c2w[:3, :3].T
    # rotate the Gaussians' own frames as well (quaternion stays a unit quaternion; any value is valid input)

    cam = make_camera(w2c, W, H)
    means3D_t = torch.from_numpy(means3D.astype(np.float32))
    normals_t = torch.from_numpy(normals_world.astype(np.float32))

    # ---- source views -------------------------------------------------------------------------------
    src_w2c, ref_to_src, src_cam_pos = [], [], []
    for _ in range(nb_src):
        D = _rigid(_rot_axis_angle(rng.normal(size=3), math.radians(rng.uniform(0.5, 3.0))),
                   rng.uniform(-0.15, 0.15, 3))
        w2s = D @ w2c
        src_w2c.append(w2s)
        ref_to_src.append(np.float32(w2s) @ np.linalg.inv(np.float32(w2c)))
        src_cam_pos.append(np.linalg.inv(w2s)[:3, 3])
    low = torch.from_numpy(rng.random((max(nb_src, 1), 3, (H + 7) // 8 + 1, (W + 7) // 8 + 1)).astype(np.float32))
    src_images = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=True)
    src_images = src_images[:nb_src].contiguous() if nb_src > 0 else src_images[:0]

    scene = dict(
        name=name, P=P, W=W, H=H, sh_degree=sh_degree, nb_src=nb_src,
        means3D=means3D_t,
        scales=torch.from_numpy(np.exp(log_scales).astype(np.float32)),
        rotations=torch.from_numpy(quat.astype(np.float32)),
        opacities=torch.from_numpy(opacity.astype(np.float32)).unsqueeze(1),
        shs=torch.from_numpy(shs.astype(np.float32)),
        normals_world=normals_t,
        all_map=all_map_for_view(means3D_t, normals_t, cam["viewmatrix"], cam["campos"]),
        bg=torch.zeros(3, dtype=torch.float32),
        w2c=torch.from_numpy(np.float32(w2c)),
        src_w2c=[torch.from_numpy(np.float32(m)) for m in src_w2c],
        ref_to_src_list=torch.from_numpy(np.stack(ref_to_src).astype(np.float32)) if nb_src else torch.zeros((0, 4, 4)),
        src_cam_pos=torch.from_numpy(np.stack(src_cam_pos).astype(np.float32)) if nb_src else torch.zeros((0, 3)),
        src_images=src_images,
        **cam,
    )
    return scene


def src_view(scene, i):
    """Camera tensors + all_map for source view i (used to render src_rendered_depths, as
    gaussian_renderer.render_depth does for each neighbour, gaussian_renderer/__init__.py:245-253)."""
    cam = make_camera(scene["src_w2c"][i].numpy(), scene["W"], scene["H"])
    cam["all_map"] = all_map_for_view(scene["means3D"], scene["normals_world"], cam["viewmatrix"], cam["campos"])
    return cam


def cotangents(scene, seed=7):
    """Fixed N(0,1) cotangents for color / normal / depth / warped image."""
    g = torch.Generator().manual_seed(seed)
    H, W = scene["H"], scene["W"]
    return dict(color=torch.randn((3, H, W), generator=g), normal=torch.randn((3, H, W), generator=g),
                depth=torch.randn((1, H, W), generator=g), warped=torch.randn((15, H, W), generator=g))
