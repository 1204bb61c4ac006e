"""float64 numpy restatement of the per-view parameter prologue (forward AND hand-derived backward).

TEST INFRASTRUCTURE ONLY (checker for ibgs_b200/csrc/prologue.cu in tests/).  Follows the reference's torch
expressions: scene/gaussian_model.py:127-147 (activations), :166-173 (get_normal), gaussian_renderer/__init__.py:304-315
(all_map).  Pinned in tests/test_prologue_oracle.py against torch autograd of those very expressions (CPU, float64).
"""
import numpy as np


def forward(xyz, opacity_raw, scaling_raw, rotation_raw, fdc, frest, normal_raw, offset, V, cam):
    f = {k: np.asarray(v, dtype=np.float64) for k, v in dict(xyz=xyz, o=opacity_raw, s=scaling_raw, r=rotation_raw,
                                                            fdc=fdc, frest=frest, n=normal_raw, off=offset, V=V,
                                                            cam=cam).items()}
    out = {}
    out["opacity"] = 1.0 / (1.0 + np.exp(-f["o"]))
    out["scales"] = np.exp(f["s"])
    out["rotations"] = f["r"] / np.maximum(np.linalg.norm(f["r"], axis=1, keepdims=True), 1e-12)
    out["shs"] = np.concatenate([f["fdc"], f["frest"]], axis=1)
    nh = f["n"] / np.linalg.norm(f["n"], axis=1, keepdims=True)
    neg = (nh * (f["cam"][None] - f["xyz"])).sum(-1) < 0.0
    sgn = np.where(neg, -1.0, 1.0)
    ng = nh * sgn[:, None]
    og = f["off"][:, 0] * sgn
    ln = ng @ f["V"][:3, :3]
    gd = -(ng * f["xyz"]).sum(-1) + og
    u = gd - (ln * f["V"][3:4, :3]).sum(1)
    out["all_map"] = np.concatenate([ln, np.ones((len(u), 1)), np.abs(u)[:, None]], axis=1)
    out["_cache"] = dict(f=f, nh=nh, sgn=sgn, ng=ng, ln=ln, u=u)
    return out


def backward(fw, g_opacity, g_scales, g_rotations, g_shs, g_all_map):
    c = fw["_cache"]
    f, nh, sgn, ng, u = c["f"], c["nh"], c["sgn"], c["ng"], c["u"]
    g = {k: np.asarray(v, dtype=np.float64) for k, v in dict(o=g_opacity, s=g_scales, r=g_rotations, sh=g_shs,
                                                            am=g_all_map).items()}
    d = {}
    o = fw["opacity"]
    d["opacity_raw"] = g["o"] * o * (1.0 - o)
    d["scaling_raw"] = g["s"] * fw["scales"]
    ln_r = np.linalg.norm(f["r"], axis=1, keepdims=True)
    y = fw["rotations"]
    d["rotation_raw"] = (g["r"] - y * (y * g["r"]).sum(1, keepdims=True)) / ln_r
    d["features_dc"] = g["sh"][:, :1]
    d["features_rest"] = g["sh"][:, 1:]
    V3, t = f["V"][:3, :3], f["V"][3, :3]
    g_u = g["am"][:, 4] * np.sign(u)
    g_ln = g["am"][:, :3] - g_u[:, None] * t[None]
    g_ng = g_ln @ V3.T - g_u[:, None] * f["xyz"]
    d["xyz"] = -g_u[:, None] * ng
    d["offset"] = (g_u * sgn)[:, None]
    g_nh = g_ng * sgn[:, None]
    d["normal_raw"] = (g_nh - nh * (nh * g_nh).sum(1, keepdims=True)) / np.linalg.norm(f["n"], axis=1, keepdims=True)
    return d


# ---- learnt_normal = False: shortest-axis plane normal (scene/gaussian_model.py:149-161) ------------------------------
# pytorch3d.transforms.quaternion_to_matrix (real part first): R = I + two_s * A(q), two_s = 2 / (q.q); column c of A
def _quat_col(y, c):
    r, i, j, k = y
    return [np.array([-(j * j + k * k), i * j + k * r, i * k - j * r]),
            np.array([i * j - k * r, -(i * i + k * k), j * k + i * r]),
            np.array([i * k + j * r, j * k - i * r, -(i * i + j * j)])][c]


def _quat_col_jac(y, c):     # J[row, m] = d col[row] / d y[m]
    r, i, j, k = y
    return [np.array([[0, 0, -2 * j, -2 * k], [k, j, i, r], [-j, k, -r, i]]),
            np.array([[-k, j, i, -r], [0, -2 * i, 0, -2 * k], [i, r, k, j]]),
            np.array([[j, k, r, i], [-i, -r, k, j], [0, -2 * i, -2 * j, 0]])][c].astype(np.float64)


def forward_smallest_axis(xyz, opacity_raw, scaling_raw, rotation_raw, fdc, frest, V, cam):
    f = {k: np.asarray(v, dtype=np.float64) for k, v in dict(xyz=xyz, o=opacity_raw, s=scaling_raw, r=rotation_raw,
                                                            fdc=fdc, frest=frest, V=V, cam=cam).items()}
    out = {}
    out["opacity"] = 1.0 / (1.0 + np.exp(-f["o"]))
    out["scales"] = np.exp(f["s"])
    y = f["r"] / np.maximum(np.linalg.norm(f["r"], axis=1, keepdims=True), 1e-12)
    out["rotations"] = y
    out["shs"] = np.concatenate([f["fdc"], f["frest"]], axis=1)
    idx = np.argmin(out["scales"], axis=1)                      # first minimum, like torch.min(dim)
    two_s = 2.0 / (y * y).sum(1)
    nh = np.stack([two_s[n] * _quat_col(y[n], idx[n]) for n in range(len(y))]) if len(y) else np.zeros((0, 3))
    nh[np.arange(len(y)), idx] += 1.0
    neg = (nh * (f["cam"][None] - f["xyz"])).sum(-1) < 0.0
    sgn = np.where(neg, -1.0, 1.0)
    ng = nh * sgn[:, None]
    ln = ng @ f["V"][:3, :3]
    gd = -(ng * f["xyz"]).sum(-1)
    u = gd - (ln * f["V"][3:4, :3]).sum(1)
    out["all_map"] = np.concatenate([ln, np.ones((len(u), 1)), np.abs(u)[:, None]], axis=1)
    out["_cache"] = dict(f=f, nh=nh, sgn=sgn, ng=ng, ln=ln, u=u, idx=idx, two_s=two_s, y=y)
    return out


def backward_smallest_axis(fw, g_opacity, g_scales, g_rotations, g_shs, g_all_map):
    c = fw["_cache"]
    f, sgn, ng, u, idx, two_s, y = c["f"], c["sgn"], c["ng"], c["u"], c["idx"], c["two_s"], c["y"]
    g = {k: np.asarray(v, dtype=np.float64) for k, v in dict(o=g_opacity, s=g_scales, r=g_rotations, sh=g_shs,
                                                            am=g_all_map).items()}
    d = {}
    o = fw["opacity"]
    d["opacity_raw"] = g["o"] * o * (1.0 - o)
    d["scaling_raw"] = g["s"] * fw["scales"]              # argmin carries no gradient
    d["features_dc"] = g["sh"][:, :1]
    d["features_rest"] = g["sh"][:, 1:]
    V3, t = f["V"][:3, :3], f["V"][3, :3]
    g_u = g["am"][:, 4] * np.sign(u)
    g_ln = g["am"][:, :3] - g_u[:, None] * t[None]
    g_ng = g_ln @ V3.T - g_u[:, None] * f["xyz"]
    d["xyz"] = -g_u[:, None] * ng
    g_nh = g_ng * sgn[:, None]
    g_y = g["r"].copy()
    for n in range(len(y)):
        col, J = _quat_col(y[n], idx[n]), _quat_col_jac(y[n], idx[n])
        g_y[n] += two_s[n] * (g_nh[n] @ J) - two_s[n] ** 2 * y[n] * (g_nh[n] @ col)
    ln_r = np.linalg.norm(f["r"], axis=1, keepdims=True)
    d["rotation_raw"] = (g_y - y * (y * g_y).sum(1, keepdims=True)) / ln_r
    return d
