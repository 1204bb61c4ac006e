"""Prints the error metrics of tests/test_gpu_dropin_render.py::test_render_unchanged_glue_vs_reference a few times
(dev tool: how much margin the gates have, run to run)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import refglue as G
import ibgs_testutil as U
from ibgs_b200 import synthetic as S
import test_gpu_dropin_render as T

glues = (G.bind("b200"), G.bind("reference"))
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    wo, wr = T._worlds(glues, learnt_normal=True)
    G.prime_depth_cache(wo); G.prime_depth_cache(wr)
    cache_err = (wo.scene.rendered_depth_list - wr.scene.rendered_depth_list).abs().max().item()
    wo.scene.rendered_depth_list.copy_(wr.scene.rendered_depth_list)
    pkgs = []
    for w in (wo, wr):
        cam = w.scene.getTrainCameras()[0]
        pkg = w.glue.render(cam, w.gaussians, w.scene, w.pipe, w.args, w.background, render_geo=True,
                            return_depth_normal=True, **G.render_kwargs(w))
        T._cotangent_loss(pkg).backward()
        pkgs.append(pkg)
    po, pr = pkgs
    outs = {k: (po[k] - pr[k]).abs().max().item() for k in T.FLOAT_KEYS}
    dn = (po["median_intersected_depth_normal"] - pr["median_intersected_depth_normal"]).abs().mean().item()
    go, gr = G.gaussian_grads(wo), G.gaussian_grads(wr)
    gerr = {n: U.rel_l2(go[n], gr[n]) for n in G.GAUSSIAN_PARAMS if gr[n] is not None}
    vs = {k: U.rel_l2(po[k].grad[:, :2], pr[k].grad[:, :2]) for k in ("viewspace_points", "viewspace_points_abs")}
    print(f"rep {rep}: cache {cache_err:.2e} | outs max " + " ".join(f"{k}={v:.1e}" for k, v in outs.items()) + f" | dnormal mean {dn:.1e}")
    print("        grads " + " ".join(f"{k}={v:.1e}" for k, v in {**gerr, **vs}.items()), flush=True)
