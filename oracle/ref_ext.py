"""Loader + thin caller for the UNMODIFIED reference CUDA extensions built by oracle/build_ref.py.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py (--impl reference).
Nothing under ibgs_b200/ imports this.

The reference's Python wrapper (diff_plane_rasterization/__init__.py) is not copied; the two `_C`
entry points are called directly with the argument order the wrapper uses (__init__.py:66-98 for
forward, :182-221 for backward), and the three opaque state buffers are decoded with the layouts of
GeometryState / ImageState / BinningState::fromChunk (cuda_rasterizer/rasterizer_impl.cu:272-316).
"""
import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


def available(which="dpr"):
    name = {"dpr": "ref_dpr_C", "knn": "ref_knn_C"}[which]
    return os.path.exists(os.path.join(HERE, "_ref", which, name + ".so"))


def load(which="dpr"):
    if which in _CACHE:
        return _CACHE[which]
    name = {"dpr": "ref_dpr_C", "knn": "ref_knn_C"}[which]
    path = os.path.join(HERE, "_ref", which, name + ".so")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `python oracle/build_ref.py {which}` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _CACHE[which] = mod
    return mod


def _e(device):
    return torch.empty(0, device="cpu")  # the reference wrapper passes torch.Tensor([]) for absent inputs


def forward(sc, render_geo=True, render_depth_only=False, buffer_length=4, depth_error_threshold=0.01,
            colors_precomp=None, cov3D_precomp=None, debug=False, cam=None):
    """sc: dict of CUDA tensors (see tests/ibgs_testutil.scene_to_device).  `cam` optionally overrides the
    camera tensors + all_map (source-view depth renders).  Returns dict(outputs..., state...)."""
    C = load("dpr")
    if cam is not None:
        sc = dict(sc)
        sc.update({k: cam[k] for k in ("viewmatrix", "projmatrix", "campos", "tanfovx", "tanfovy", "all_map")})
    dev = sc["means3D"].device
    sh = sc["shs"] if colors_precomp is None else _e(dev)
    colors = colors_precomp if colors_precomp is not None else _e(dev)
    scales = sc["scales"] if cov3D_precomp is None else _e(dev)
    rots = sc["rotations"] if cov3D_precomp is None else _e(dev)
    cov = cov3D_precomp if cov3D_precomp is not None else _e(dev)
    all_map = sc["all_map"] if (render_geo or render_depth_only) else _e(dev)
    args = (sc["bg"], sc["means3D"], colors, sc["opacities"], scales, rots, 1.0, cov, all_map,
            sc["viewmatrix"], sc["projmatrix"], sc["ref_to_src_list"], sc["src_cam_pos"], sc["src_images"],
            sc["src_rendered_depths"], int(sc["nb_src"]), int(buffer_length), float(depth_error_threshold),
            float(sc["tanfovx"]), float(sc["tanfovy"]), int(sc["H"]), int(sc["W"]), sh, int(sc["sh_degree"]),
            sc["campos"], False, bool(render_geo), bool(render_depth_only), bool(debug))
    (R, color, radii, normal, depth, cam_feat, warped, min_diff, cam_ray, mask, geom, binning, img) = \
        C.rasterize_gaussians(*args)
    return dict(num_rendered=R, color=color, radii=radii, normal=normal, depth=depth, cam_feat=cam_feat,
                warped=warped, min_depth_diff=min_diff, camera_ray=cam_ray, mask=mask,
                geom=geom, binning=binning, img=img,
                _saved=dict(colors=colors, scales=scales, rots=rots, cov=cov, all_map=all_map, sh=sh))


def backward(sc, fw, cot, render_geo=True, debug=False):
    C = load("dpr")
    s = fw["_saved"]
    args = (sc["bg"], fw["normal"], fw["depth"], fw["warped"], sc["means3D"], fw["radii"], s["colors"],
            s["all_map"], s["scales"], s["rots"], 1.0, s["cov"], sc["viewmatrix"], sc["projmatrix"],
            sc["ref_to_src_list"], sc["src_cam_pos"], sc["src_images"], sc["src_rendered_depths"],
            int(sc["nb_src"]), float(sc["tanfovx"]), float(sc["tanfovy"]), cot["color"], cot["normal"],
            cot["depth"], cot["warped"], s["sh"], int(sc["sh_degree"]), sc["campos"], fw["geom"],
            int(fw["num_rendered"]), fw["binning"], fw["img"], bool(render_geo), bool(debug))
    (g_means2D, g_means2D_abs, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rots, g_all_map) = \
        C.rasterize_gaussians_backward(*args)
    return dict(means3D=g_means3D, means2D=g_means2D, means2D_abs=g_means2D_abs, sh=g_sh, colors=g_colors,
                opacities=g_opac, scales=g_scales, rotations=g_rots, cov3D=g_cov3D, all_map=g_all_map)


class RefRasterize(torch.autograd.Function):
    """Minimal autograd wrapper around the reference extension's two entry points, so that bench.py's reference arm can
    run a whole training step (loss.backward()) through the UNMODIFIED reference rasterizer.  The reference's own
    Python wrapper (diff_plane_rasterization/__init__.py:21-250) does the same bookkeeping; it is not available on the
    GPU box (no /root/reference there), and reference sources are never copied into this repo."""

    @staticmethod
    def forward(ctx, means3D, sh, opacities, scales, rotations, all_map, sc):
        sc = dict(sc)
        sc.update(means3D=means3D, shs=sh, opacities=opacities, scales=scales, rotations=rotations, all_map=all_map)
        fw = forward(sc, render_geo=True)
        ctx.sc, ctx.fw = sc, fw
        return fw["color"], fw["normal"], fw["depth"], fw["warped"]

    @staticmethod
    def backward(ctx, g_color, g_normal, g_depth, g_warped):
        z = lambda g, ref: torch.zeros_like(ref) if g is None else g.contiguous()
        fw = ctx.fw
        cot = dict(color=z(g_color, fw["color"]), normal=z(g_normal, fw["normal"]), depth=z(g_depth, fw["depth"]),
                   warped=z(g_warped, fw["warped"]))
        gr = backward(ctx.sc, fw, cot, render_geo=True)
        return (gr["means3D"], gr["sh"], gr["opacities"].view_as(ctx.sc["opacities"]), gr["scales"], gr["rotations"],
                gr["all_map"], None)


def _al(x, a=128):
    return (x + a - 1) // a * a


def _view(buf, off, count, dtype):
    nbytes = count * torch.empty(0, dtype=dtype).element_size()
    return buf[off:off + nbytes].view(dtype)


def decode_geom(geom, P):
    """GeometryState::fromChunk, rasterizer_impl.cu:272-287 (offsets relative to a >=128-aligned base)."""
    o = 0
    out = {}
    out["depths"] = _view(geom, o, P, torch.float32); o = _al(o + 4 * P)
    out["clamped"] = _view(geom, o, 3 * P, torch.uint8); o = _al(o + 3 * P)
    out["internal_radii"] = _view(geom, o, P, torch.int32); o = _al(o + 4 * P)
    out["means2D"] = _view(geom, o, 2 * P, torch.float32).view(P, 2); o = _al(o + 8 * P)
    out["cov3D"] = _view(geom, o, 6 * P, torch.float32).view(P, 6); o = _al(o + 24 * P)
    out["conic_opacity"] = _view(geom, o, 4 * P, torch.float32).view(P, 4); o = _al(o + 16 * P)
    out["rgb"] = _view(geom, o, 3 * P, torch.float32).view(P, 3); o = _al(o + 12 * P)
    out["tiles_touched"] = _view(geom, o, P, torch.int32); o = _al(o + 4 * P)
    # scanning_space (CUB-sized) sits here; point_offsets is the last array, followed by the 128-byte pad
    po = geom.numel() - 128 - 4 * P
    out["point_offsets"] = _view(geom, po, P, torch.int32)
    return out


def decode_image(img, N):
    """ImageState::fromChunk, rasterizer_impl.cu:289-301"""
    o = 0
    out = {}
    out["final_T"] = _view(img, o, N, torch.float32); o = _al(o + 4 * N)
    out["n_contrib"] = _view(img, o, N, torch.int32); o = _al(o + 4 * N)
    out["ranges"] = _view(img, o, 2 * N, torch.int32).view(N, 2); o = _al(o + 8 * N)
    out["sum_w"] = _view(img, o, N, torch.float32); o = _al(o + 4 * N)
    out["low"] = _view(img, o, N, torch.int32); o = _al(o + 4 * N)
    out["high"] = _view(img, o, N, torch.int32); o = _al(o + 4 * N)
    out["valid_idx"] = _view(img, o, 5 * N, torch.int32).view(5, N); o = _al(o + 20 * N)
    out["valid_w"] = _view(img, o, 5 * N, torch.float32).view(5, N); o = _al(o + 20 * N)
    return out


def decode_binning(binning, R):
    """BinningState::fromChunk, rasterizer_impl.cu:303-316"""
    o = 0
    out = {}
    out["point_list"] = _view(binning, o, R, torch.int32); o = _al(o + 4 * R)
    out["point_list_unsorted"] = _view(binning, o, R, torch.int32); o = _al(o + 4 * R)
    out["keys"] = _view(binning, o, R, torch.int64); o = _al(o + 8 * R)
    out["keys_unsorted"] = _view(binning, o, R, torch.int64); o = _al(o + 8 * R)
    return out


def dist2(points):
    return load("knn").distCUDA2(points)
