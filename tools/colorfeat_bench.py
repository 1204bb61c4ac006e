"""Colour-aggregation step: reference fuse_color vs ibgs_b200.color_aggregation.fuse_color, forward+backward (dev tool).
usage: python tools/colorfeat_bench.py [H W]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import colorfeat_ref as CR  # noqa: E402
import refglue as G  # noqa: E402
from ibgs_b200 import _native as N  # noqa: E402
from ibgs_b200 import color_aggregation as CA  # noqa: E402

G._paths()
import color_aggregation_network as CAN  # noqa: E402

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (822, 1237)
torch.manual_seed(0)
net = CAN.ColorFusionResidualNet(height=H, width=W).cuda()
pkg = CR.random_render_pkg(H, W, seed=1, device="cuda")
gt = torch.rand(3, H, W, device="cuda")


class Opts:
    enable_exposure_correction = False
    nb_visible_src_frames = 3
    residual_resolution_scale = 1.0


def run(name, fn, iters=10):
    def step():
        leaves = {k: pkg[k].clone().requires_grad_(True) for k in ("render", "warped_image")}
        out = fn(dict(pkg, **leaves), color_aggregation_network=net, iter_count=None, burn_start=None, burn_end=None,
                 iteration=20000, opts=Opts())
        (out["image_pred"] - gt).abs().mean().backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) / iters:8.3f} ms fwd+bwd", flush=True)


run("reference fuse_color", CAN.fuse_color)
N.lib.ibgs_profile_reset()
N.lib.ibgs_profile_enable(1)
run("ibgs_b200 fuse_color bf16", lambda *a, **k: CA.fuse_color(*a, precision="bf16", **k))
N.lib.ibgs_profile_enable(0)
print({k: round(v[0] / v[1], 4) for k, v in N.profile_read().items() if v[1]})
run("ibgs_b200 fuse_color fp32", lambda *a, **k: CA.fuse_color(*a, precision="fp32", **k))
torch.backends.cudnn.benchmark = True
run("ibgs_b200 fuse_color bf16 + cudnn.benchmark", lambda *a, **k: CA.fuse_color(*a, precision="bf16", **k))
run("reference fuse_color + cudnn.benchmark", CAN.fuse_color)
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run("profiled", lambda *a, **k: CA.fuse_color(*a, precision="bf16", **k), iters=1)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=80))
