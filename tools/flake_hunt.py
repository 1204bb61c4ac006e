"""Dev tool: the depth-cache priming of tests/test_gpu_dropin_render.py (train.py:242-256 under both bindings), repeated,
with a report of WHICH side changes between repetitions when the caches disagree."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
# the same preamble as the test session: a few rasterizer calls on cfg1-shaped scenes first (speculation hints, allocator)
import refglue as G
from ibgs_b200 import synthetic as S
import ibgs_testutil as U
import ibgs_b200.diff_plane_rasterization as dpr

if "--preamble" in sys.argv:
    for seed in range(3):
        sc = U.scene_to_device(S.make_scene("cfg1"))
        sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
        cot = {k: v.cuda() for k, v in S.cotangents(sc).items()}
        U.ours_forward_backward(dpr, sc, cot, render_geo=True)

# record a checksum of every tensor the b200 rasterizer receives and returns, per call
import hashlib
CALLS = []
_orig = dpr.rasterize_gaussians


def _h(t):
    if t is None or not torch.is_tensor(t):
        return str(t)
    return hashlib.md5(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()[:8]


def _hooked(means3D, means2D, means2D_abs, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, all_map,
            raster_settings, *rest):
    rs = raster_settings
    rec = dict(means3D=_h(means3D), sh=_h(sh), opac=_h(opacities), scales=_h(scales), rot=_h(rotations), all_map=_h(all_map),
               view=_h(rs.viewmatrix), proj=_h(rs.projmatrix), campos=_h(rs.campos), srcd=_h(rs.src_rendered_depths),
               srci=_h(rs.src_images), r2s=_h(rs.ref_to_src_list), scp=_h(rs.src_cam_pos), bg=_h(rs.bg),
               misc=(rs.image_height, rs.image_width, rs.tanfovx, rs.tanfovy, rs.nb_src_images, rs.buffer_length,
                     rs.depth_error_threshold, rs.sh_degree, rs.render_geo, rs.render_depth_only))
    out = _orig(means3D, means2D, means2D_abs, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, all_map,
                raster_settings, *rest)
    rec["out_depth"] = _h(out[3])
    rec["out_color"] = _h(out[0])
    rec["radii"] = _h(out[1])
    CALLS.append(rec)
    return out


KEEP = []      # asynchronous variant: clones of the inputs that can influence the median depth, compared after the fact


def _keeping(means3D, means2D, means2D_abs, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, all_map,
             raster_settings, *rest):
    rs = raster_settings
    KEEP.append(dict(means3D=means3D.detach().clone(), opac=opacities.detach().clone(), scales=scales.detach().clone(),
                     rot=rotations.detach().clone(), all_map=all_map.detach().clone(), view=rs.viewmatrix.clone(),
                     proj=rs.projmatrix.clone(), campos=rs.campos.clone()))
    out = _orig(means3D, means2D, means2D_abs, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, all_map,
                raster_settings, *rest)
    KEEP[-1]["out_depth"] = out[3].detach().clone()
    KEEP[-1]["radii"] = out[1].detach().clone()
    return out


if os.environ.get("HOOK"):
    dpr.rasterize_gaussians = _hooked
elif os.environ.get("KEEP"):
    dpr.rasterize_gaussians = _keeping
glues = (G.bind("b200"), G.bind("reference"))
sc_cpu = S.make_scene("cfg1")


def prime(g):
    w = G.build_world(g, "cfg1", n_views=6, sc_cpu=sc_cpu, learnt_normal=True)
    G.prime_depth_cache(w)
    torch.cuda.synchronize()
    return w.scene.rendered_depth_list.clone()


runs = {"b200": [], "reference": []}
call_log = []
for rep in range(int(os.environ.get("REPS", "3"))):
    for g in glues:
        if g.binding == "reference" and (rep >= 2 or os.environ.get("ONLY_B200")):
            continue
        try:
            CALLS.clear()
            KEEP.clear()
            runs[g.binding].append(prime(g))
            if g.binding == "b200":
                call_log.append(list(KEEP) if os.environ.get("KEEP") else list(CALLS))
        except Exception as ex:
            print(f"{g.binding} run {rep}: EXCEPTION {ex!r}", flush=True)
            raise
ref0 = runs["reference"][0] if runs["reference"] else runs["b200"][0]
for name, lst in runs.items():
    for i, c in enumerate(lst):
        d = (c - lst[0]).abs()
        dr = (c - ref0).abs()
        per_view = [(v, int((d[v] > 1e-4).sum().item())) for v in range(d.shape[0]) if (d[v] > 1e-4).any()]
        if i > 2 and d.max().item() == 0 and dr.max().item() == 0:
            continue
        print(f"{name} run {i}: vs own run 0 max|d| {d.max().item():.3e} (views/pixels differing: {per_view}) | vs reference run 0 max|d| {dr.max().item():.3e}",
              flush=True)

# which INPUT or OUTPUT checksums of the b200 rasterizer calls differ between run 0 and run 1
if len(call_log) >= 2:
    for ci, (c0, c1) in enumerate(zip(call_log[0], call_log[1])):
        diff = [k for k in c0 if (not torch.equal(c0[k], c1[k]) if torch.is_tensor(c0[k]) else c0[k] != c1[k])]
        print(f"b200 call {ci} (view {ci}): fields differing between run 0 and run 1: {diff}", flush=True)
