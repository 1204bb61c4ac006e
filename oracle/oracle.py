"""ctypes front end of oracle/ibgs_oracle.c (the float64 CPU restatement of the hot path).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg; never by anything under ibgs_b200/.  See the header of ibgs_oracle.c for what it follows in the
reference and how it is pinned (tests/golden/ + tests/test_oracle_golden.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ibgs_oracle.c")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libibgs_oracle.so")
_lib = None
MAXS = 5


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    if (not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{r.stderr}")
    return LIB


_fp = C.c_void_p


class OIn(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("P", "D", "M", "W", "H", "nb_src", "BL", "render_geo", "depth_only")] + \
               [(n, C.c_double) for n in ("tanfovx", "tanfovy", "scale_modifier", "thr")] + \
               [(n, _fp) for n in ("bg", "means3D", "shs", "colors_precomp", "opacities", "scales", "rotations",
                                   "cov3D_precomp", "all_map", "view", "proj", "campos", "ref_to_src",
                                   "src_cam_pos", "src_images", "src_depths")]


class OGeom(C.Structure):
    _fields_ = [(n, _fp) for n in ("radii", "depths", "means2D", "conic_o", "rgb", "clamped", "cov3D", "tiles",
                                   "offsets")]


class OImg(C.Structure):
    _fields_ = [(n, _fp) for n in ("final_T", "n_contrib", "sum_w", "low", "high", "valid_idx", "valid_w")]


class OOut(C.Structure):
    _fields_ = [(n, _fp) for n in ("color", "normal", "depth", "cam_feat", "warped", "min_depth_diff",
                                   "camera_ray", "mask")]


class OGrad(C.Structure):
    _fields_ = [(n, _fp) for n in ("means3D", "means2D", "means2D_abs", "colors", "opacity", "cov3D", "sh",
                                   "scales", "rots", "all_map", "conic")]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.o_preprocess.restype = C.c_int64
        _lib.o_preprocess.argtypes = [C.POINTER(OIn), C.POINTER(OGeom)]
        _lib.o_bin.restype = None
        _lib.o_bin.argtypes = [C.POINTER(OIn), C.POINTER(OGeom), C.c_int64, _fp, _fp, _fp, _fp, _fp]
        _lib.o_render.restype = None
        _lib.o_render.argtypes = [C.POINTER(OIn), C.POINTER(OGeom), _fp, _fp, C.POINTER(OImg), C.POINTER(OOut)]
        _lib.o_render_backward.restype = None
        _lib.o_render_backward.argtypes = [C.POINTER(OIn), C.POINTER(OGeom), _fp, _fp, C.POINTER(OImg), _fp, _fp,
                                           _fp, _fp, _fp, _fp, C.POINTER(OGrad)]
        _lib.o_preprocess_backward.restype = None
        _lib.o_preprocess_backward.argtypes = [C.POINTER(OIn), C.POINTER(OGeom), C.POINTER(OGrad)]
        _lib.o_dist2.restype = None
        _lib.o_dist2.argtypes = [C.c_int, _fp, _fp]
        _lib.o_mark_visible.restype = C.c_int
        _lib.o_mark_visible.argtypes = [C.c_int, _fp, _fp, _fp]
    return _lib


def _f32(x):
    if x is None:
        return None
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data


def _fill(struct, arrays):
    for k, v in arrays.items():
        setattr(struct, k, _p(v))
    return struct


class Result(dict):
    __getattr__ = dict.__getitem__


def forward(sc, render_geo=True, render_depth_only=False, buffer_length=4, depth_error_threshold=0.01,
            colors_precomp=None, cov3D_precomp=None, cam=None):
    """sc: scene dict (ibgs_b200.synthetic.make_scene + 'src_rendered_depths' when render_geo).
    Returns a Result with the nine outputs (numpy float32, reference shapes) plus all intermediate state."""
    L = lib()
    cam = cam or sc
    P, W, H = int(sc["P"]), int(sc["W"]), int(sc["H"])
    N, T = W * H, ((W + 15) // 16) * ((H + 15) // 16)
    keep = dict(
        bg=_f32(sc["bg"]), means3D=_f32(sc["means3D"]),
        shs=None if colors_precomp is not None else _f32(sc["shs"]),
        colors_precomp=_f32(colors_precomp), opacities=_f32(sc["opacities"]).reshape(-1),
        scales=None if cov3D_precomp is not None else _f32(sc["scales"]),
        rotations=None if cov3D_precomp is not None else _f32(sc["rotations"]),
        cov3D_precomp=_f32(cov3D_precomp),
        all_map=_f32(cam["all_map"] if "all_map" in cam else sc["all_map"]) if (render_geo or render_depth_only) else None,
        view=_f32(cam["viewmatrix"]), proj=_f32(cam["projmatrix"]), campos=_f32(cam["campos"]))
    if render_geo:
        keep.update(ref_to_src=_f32(sc["ref_to_src_list"]).reshape(-1, 16), src_cam_pos=_f32(sc["src_cam_pos"]),
                    src_images=_f32(sc["src_images"]), src_depths=_f32(sc["src_rendered_depths"]))
        nb = int(sc["nb_src"])
    else:
        keep.update(ref_to_src=None, src_cam_pos=None, src_images=None, src_depths=None)
        nb = 0
    oin = OIn()
    oin.P, oin.D, oin.W, oin.H = P, int(sc["sh_degree"]), W, H
    oin.M = keep["shs"].shape[1] if keep["shs"] is not None else 0
    oin.nb_src, oin.BL = nb, int(buffer_length)
    oin.render_geo, oin.depth_only = int(render_geo), int(render_depth_only)
    oin.tanfovx, oin.tanfovy = float(np.float32(cam["tanfovx"])), float(np.float32(cam["tanfovy"]))
    oin.scale_modifier, oin.thr = 1.0, float(np.float32(depth_error_threshold))
    _fill(oin, keep)

    geom = dict(radii=np.zeros(P, np.int32), depths=np.zeros(P, np.float32), means2D=np.zeros((P, 2)),
                conic_o=np.zeros((P, 4)), rgb=np.zeros((P, 3)), clamped=np.zeros((P, 3), np.uint8),
                cov3D=np.zeros((P, 6)), tiles=np.zeros(P, np.uint32), offsets=np.zeros(P, np.uint32))
    og = _fill(OGeom(), geom)
    R = int(L.o_preprocess(C.byref(oin), C.byref(og))) if P > 0 else 0
    keys_u, vals_u = np.zeros(max(R, 1), np.uint64), np.zeros(max(R, 1), np.uint32)
    keys, plist = np.zeros(max(R, 1), np.uint64), np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((T, 2), np.uint32)
    if P > 0:
        L.o_bin(C.byref(oin), C.byref(og), R, _p(keys_u), _p(vals_u), _p(keys), _p(plist), _p(ranges))
    img = dict(final_T=np.zeros(N), n_contrib=np.zeros(N, np.uint32), sum_w=np.zeros(N), low=np.zeros(N, np.uint32),
               high=np.zeros(N, np.uint32), valid_idx=np.zeros((MAXS, N), np.int32), valid_w=np.zeros((MAXS, N)))
    out = dict(color=np.zeros((3, H, W), np.float32), normal=np.zeros((3, H, W), np.float32),
               depth=np.zeros((1, H, W), np.float32), cam_feat=np.zeros((4 * MAXS, H, W), np.float32),
               warped=np.zeros((3 * MAXS, H, W), np.float32), min_depth_diff=np.zeros((1, H, W), np.float32),
               camera_ray=np.zeros((3, H, W), np.float32), mask=np.zeros((1, H, W), np.int32))
    oi, oo = _fill(OImg(), img), _fill(OOut(), out)
    if P > 0:
        L.o_render(C.byref(oin), C.byref(og), _p(plist), _p(ranges), C.byref(oi), C.byref(oo))
    res = Result(out)
    res.update(radii=geom["radii"], num_rendered=R, geom=geom, img=img, keys_unsorted=keys_u[:R],
               vals_unsorted=vals_u[:R], keys=keys[:R], point_list=plist[:R], ranges=ranges,
               _oin=oin, _og=og, _oi=oi, _keep=keep)
    return res


def backward(sc, fw, cot, render_geo=True):
    """cot: dict color/normal/depth/warped (torch or numpy).  Returns dict of float64 numpy gradients with the
    reference's names and shapes (rasterize_points.cu:209-219)."""
    L = lib()
    P = int(sc["P"])
    oin = fw["_oin"]
    M = int(oin.M)
    g = dict(means3D=np.zeros((P, 3)), means2D=np.zeros((P, 3)), means2D_abs=np.zeros((P, 3)), colors=np.zeros((P, 3)),
             opacity=np.zeros((P, 1)), cov3D=np.zeros((P, 6)), sh=np.zeros((P, max(M, 1), 3)), scales=np.zeros((P, 3)),
             rots=np.zeros((P, 4)), all_map=np.zeros((P, 5)), conic=np.zeros((P, 4)))
    og = _fill(OGrad(), g)
    c = {k: _f32(v) for k, v in cot.items()}
    if P > 0:
        L.o_render_backward(C.byref(oin), C.byref(fw["_og"]), _p(fw["point_list"] if fw["num_rendered"] else np.zeros(1, np.uint32)),
                            _p(fw["ranges"]), C.byref(fw["_oi"]), _p(fw["depth"]), _p(fw["warped"]),
                            _p(c["color"]), _p(c.get("normal")), _p(c.get("depth")), _p(c.get("warped")), C.byref(og))
        L.o_preprocess_backward(C.byref(oin), C.byref(fw["_og"]), C.byref(og))
    return dict(means3D=g["means3D"], means2D=g["means2D"], means2D_abs=g["means2D_abs"], sh=g["sh"][:, :M],
                colors=g["colors"], opacities=g["opacity"], scales=g["scales"], rotations=g["rots"],
                cov3D=g["cov3D"], all_map=g["all_map"])


def dist2(points):
    pts = _f32(points)
    out = np.zeros(pts.shape[0], np.float32)
    if pts.shape[0]:
        lib().o_dist2(pts.shape[0], _p(pts), _p(out))
    return out


def mark_visible(means3D, viewmatrix):
    m, v = _f32(means3D), _f32(viewmatrix)
    out = np.zeros(m.shape[0], np.uint8)
    lib().o_mark_visible(m.shape[0], _p(m), _p(v), _p(out))
    return out.astype(bool)


def render_src_depths(sc, buffer_length=4):
    """src_rendered_depths via the oracle's depth-only pass from each source pose (CPU-only test scenes)."""
    import torch
    from ibgs_b200 import synthetic as S
    out = []
    for i in range(int(sc["nb_src"])):
        cam = S.src_view(sc, i)
        r = forward(sc, render_geo=False, render_depth_only=True, buffer_length=buffer_length, cam=cam)
        out.append(torch.from_numpy(r["depth"].copy()))
    return torch.stack(out, 0) if out else torch.zeros((0, 1, sc["H"], sc["W"]))
