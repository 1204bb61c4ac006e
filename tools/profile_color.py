"""A few forward+backward passes of the colour-aggregation fast path at cfg3's image size (dev tool; the command ncu
wraps for color_features_* and the NHWC glue kernels of nhwc_ops.cu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import colorfeat_ref as CR
from ibgs_b200 import color_aggregation as CA
from sanitize_r2_net import Net, Opts
H, W = 822, 1237
torch.manual_seed(0)
net = Net("mean").cuda()
pkg = CR.random_render_pkg(H, W, seed=1, device="cuda")
for _ in range(3):
    leaves = {k: pkg[k].clone().requires_grad_(True) for k in ("render", "warped_image")}
    out = CA.fuse_color(dict(pkg, **leaves), net, None, None, None, 20000, Opts(), precision="bf16")
    out["image_pred"].square().mean().backward()
torch.cuda.synchronize()
print("done")
