"""Forward+backward time of one SSIM loss term at the benchmark resolution: the reference's torch expressions
(utils/loss_utils.py:34-65, float32 transcription in tests/ssim_ref.py) vs the fused CUDA kernels (dev tool).
usage: python tools/ssim_bench.py [H W] [--iters 20]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ibgs_b200.loss_utils as LU  # noqa: E402
from ibgs_b200 import _native as N  # noqa: E402
from ssim_ref import torch_ssim_map  # noqa: E402


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def measure(H=1080, W=1920, iters=20, peak_gbs=None):
    g = torch.Generator().manual_seed(0)
    gt = torch.rand((3, H, W), generator=g).cuda()
    img = (gt + 0.05 * torch.randn((3, H, W), generator=g).cuda()).clamp(0, 1).requires_grad_(True)

    def run(fn):
        def step():
            img.grad = None
            (1.0 - fn(img, gt)).backward()
        return step

    def fwd_only(fn):
        def step():
            with torch.no_grad():
                fn(img, gt)
        return step

    out = {"shape": [3, H, W]}
    out["torch_fwd_bwd_ms"] = timed(run(lambda a, b: torch_ssim_map(a, b).mean()), iters)
    out["fused_fwd_bwd_ms"] = timed(run(LU.ssim), iters)
    out["torch_fwd_ms"] = timed(fwd_only(lambda a, b: torch_ssim_map(a, b).mean()), iters)
    out["fused_fwd_ms"] = timed(fwd_only(LU.ssim), iters)
    # kernel-only times of the two fused launches (events inside one stream, no torch ops between)
    n0 = N.lib.ibgs_launch_count()
    img.grad = None
    (1.0 - LU.ssim(img, gt)).backward()
    out["fused_launches_fwd_bwd"] = int(N.lib.ibgs_launch_count() - n0)
    N.lib.ibgs_profile_enable(1)
    N.lib.ibgs_profile_reset()
    step = run(LU.ssim)
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    pr = N.profile_read()
    N.lib.ibgs_profile_enable(0)
    for k in ("ssim_forward", "ssim_backward"):
        out[k + "_kernel_ms"] = pr[k][0] / max(pr[k][1], 1)
    px = 3 * H * W
    out["algorithmic_bytes_fwd_bwd"] = px * (8 + 4 + 12) + px * (12 + 8 + 4)   # fwd: 2 in, map + 3 partials; bwd: 3 partials + 2 images in, 1 out
    out["fused_GBps"] = out["algorithmic_bytes_fwd_bwd"] / (out["fused_fwd_bwd_ms"] * 1e-3) / 1e9
    out["kernels_GBps"] = out["algorithmic_bytes_fwd_bwd"] / ((out["ssim_forward_kernel_ms"] + out["ssim_backward_kernel_ms"]) * 1e-3) / 1e9
    if peak_gbs:
        out["kernels_frac_of_hbm_peak"] = out["kernels_GBps"] / peak_gbs
    return out


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    H, W = (int(args[0]), int(args[1])) if len(args) >= 2 else (1080, 1920)
    print(measure(H, W))
