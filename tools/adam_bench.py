"""Optimizer step over the reference's eight per-Gaussian parameter groups (scene/gaussian_model.py:227-240) at P
Gaussians: torch.optim.Adam (what the reference runs, default multi-tensor path) + zero_grad vs the one-launch
ArenaAdam (dev tool).  usage: python tools/adam_bench.py [P]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ibgs_b200.optim import ArenaAdam  # noqa: E402

GROUPS = (("xyz", (3,), 1.6e-4), ("f_dc", (1, 3), 0.0025), ("f_rest", (8, 3), 0.0025 / 20), ("opacity", (1,), 0.05),
          ("scaling", (3,), 0.005), ("rotation", (4,), 0.001), ("normal", (3,), 0.001), ("offset", (1,), 1.6e-5))


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


def measure(P=3_000_000, iters=10, peak_gbs=None):
    g = torch.Generator().manual_seed(0)
    init = {n: torch.randn((P,) + s, generator=g).cuda() for n, s, _ in GROUPS}
    grads = {n: torch.randn((P,) + s, generator=g).cuda() * 1e-3 for n, s, _ in GROUPS}
    lrs = {n: lr for n, _, lr in GROUPS}
    ps = {n: init[n].clone().requires_grad_(True) for n in init}
    ref = torch.optim.Adam([{"params": [ps[n]], "lr": lrs[n], "name": n} for n in init], lr=0.0, eps=1e-15)

    def ref_step():
        for n in ps:
            ps[n].grad = grads[n]         # stands in for autograd having produced the gradients (not timed work)
        ref.step()
        ref.zero_grad(set_to_none=True)

    opt = ArenaAdam(init, lrs)
    opt.flat_grads.copy_(torch.cat([grads[n].reshape(-1) for n in init]))

    def ours_step():
        opt.step(zero_grads=False)        # gradients stay in place for the next timed iteration

    def ours_step_zero():
        opt.step(zero_grads=True)

    out = {"P": P, "floats_per_gaussian": sum(int(torch.Size(s).numel()) for _, s, _ in GROUPS)}
    out["torch_adam_ms"] = timed(ref_step, iters)
    out["arena_adam_ms"] = timed(ours_step, iters)
    out["arena_adam_with_zero_grad_ms"] = timed(ours_step_zero, iters)
    nfl = P * out["floats_per_gaussian"]
    out["algorithmic_bytes"] = 28 * nfl
    out["arena_adam_GBps"] = 28 * nfl / (out["arena_adam_ms"] * 1e-3) / 1e9
    if peak_gbs:
        out["arena_adam_frac_of_hbm_peak"] = out["arena_adam_GBps"] / peak_gbs
    return out


if __name__ == "__main__":
    print(measure(int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000))
