"""Small instances of every kernel added in round 2 (the command compute-sanitizer wraps; dev tool): smoke() (rasterizer on
the own sort / scan primitives, colour features), the fused prologue in both normal modes forward+backward, fuse_color
forward+backward in bf16 and fp32 with the NHWC glue kernels, max mode, exposure correction."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
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


class ConvDecoderAE(torch.nn.Module):      # same layer shapes as color_aggregation_network.py:6-49 (38 channels)
    def __init__(self, h=38):
        super().__init__()
        c = lambda i, o, k=3: torch.nn.Sequential(torch.nn.Conv2d(i, o, k, padding=k // 2), torch.nn.ReLU())
        self.enc1, self.enc2, self.enc3 = c(h, h), c(h, h // 2), c(h // 2, h // 4)
        self.up2_conv, self.up1_conv = c(h // 4, h // 2), c(h // 2, h)
        self.dec2, self.dec1 = c(h, h // 2), c(2 * h, h)
        self.fuse_input = c(2 * h, h, 1)
        self.final = torch.nn.Conv2d(h, 3, 1)


class Net(torch.nn.Module):
    def __init__(self, mode):
        super().__init__()
        self.per_view_feat_dim, self.feat_aggregate_mode = 32, mode
        self.per_view_mlp = torch.nn.Sequential(torch.nn.Linear(7, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU())
        self.conv_decoder = ConvDecoderAE()


class Opts:
    enable_exposure_correction = False
    nb_visible_src_frames = 3
    residual_resolution_scale = 1.0


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
