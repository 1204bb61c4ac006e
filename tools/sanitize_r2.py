"""Small instances of every kernel added in round 2 (the command compute-sanitizer wraps; dev tool): smoke() (rasterizer on
the own sort / scan primitives, colour features), the fused prologue in both normal modes forward+backward, fuse_color
forward+backward in bf16 and fp32 with the NHWC glue kernels, max mode, exposure correction."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
import __graft_entry__ as GE  # noqa: E402

GE.smoke()
import colorfeat_ref as CR  # noqa: E402
import prologue_ref as PR  # noqa: E402
from ibgs_b200 import color_aggregation as CA, fused  # noqa: E402

IN = ("xyz", "opacity_raw", "scaling_raw", "rotation_raw", "fdc", "frest", "normal_raw", "offset")
p = PR.random_params(1001, K=9, seed=3, device="cuda")
for sa in (False, True):
    leaves = {k: p[k].clone().requires_grad_(True) for k in IN}
    args = [leaves[k] for k in IN[:6]] + ([None, None] if sa else [leaves["normal_raw"], leaves["offset"]])
    outs = fused.gaussian_prologue(*args, p["V"], p["cam"], smallest_axis_normal=sa)
    sum(o.sum() for o in outs).backward()


from sanitize_r2_net import Net, Opts  # noqa: E402

H, W = 41, 57
pkg = CR.random_render_pkg(H, W, seed=2, device="cuda")
for mode, prec, expo in (("mean", "bf16", False), ("max", "fp32", True)):
    net = Net(mode).cuda()
    o = Opts()
    o.enable_exposure_correction = expo
    leaves = {k: pkg[k].clone().requires_grad_(True) for k in ("render", "warped_image")}
    out = CA.fuse_color(dict(pkg, **leaves), net, None, None, None, 20000, o, precision=prec)
    out["image_pred"].square().mean().backward()
torch.cuda.synchronize()
print("sanitize_r2 ok")
