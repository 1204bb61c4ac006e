"""Run N forward+backward passes of one synthetic scene (dev tool; the command ncu wraps).
usage: python tools/profile_one.py cfg3_1080p [passes]
Matched kernels per pass (render_geo): preprocess_kernel, render_forward_kernel, render_backward_kernel,
preprocess_backward_kernel; the source-depth set-up adds 4 x (preprocess_kernel, render_forward_kernel) first."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from ibgs_b200 import synthetic as S
import ibgs_b200.diff_plane_rasterization as dpr
import ibgs_testutil as U
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_1080p"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = U.scene_to_device(S.make_scene(name))
sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
cot = {k: v.cuda() for k, v in S.cotangents(sc).items()}
for _ in range(n):
    U.ours_forward_backward(dpr, sc, cot, render_geo=True)
torch.cuda.synchronize()
print("done")
