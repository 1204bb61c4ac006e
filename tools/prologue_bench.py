"""Times the per-view parameter prologue forward+backward at P Gaussians: the reference's torch expressions vs
ibgs_b200.fused (dev / measurement tool).  usage: python tools/prologue_bench.py [P] [K]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import prologue_ref as PR
from ibgs_b200 import fused
P = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 9
IN = ("xyz", "opacity_raw", "scaling_raw", "rotation_raw", "fdc", "frest", "normal_raw", "offset")
p = PR.random_params(P, K=K, seed=0, device="cuda")
g = torch.Generator().manual_seed(1)
shapes = [(P, 1), (P, 3), (P, 4), (P, K, 3), (P, 5)]
cots = [torch.randn(s, generator=g).cuda() for s in shapes]
res = {}
def fused_nocat(*a):
    return fused.gaussian_prologue(*a, concat_sh=False)
for name, fn in (("torch", PR.torch_prologue), ("fused", fused.gaussian_prologue), ("fused_nocat", fused_nocat)):
    leaves = {k: p[k].clone().requires_grad_(True) for k in IN}
    def step():
        for v in leaves.values():
            v.grad = None
        outs = fn(*[leaves[k] for k in IN], p["V"], p["cam"])
        torch.autograd.backward(list(outs), cots if len(outs) == 5 else [cots[0], cots[1], cots[2], cots[4]])
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    res[name] = ts[len(ts) // 2]
# bytes the fused path must move: inputs once per direction, outputs / gradients once
words_in = 3 + 1 + 3 + 4 + 3 * K + 3 + 1
words_out = 1 + 3 + 4 + 3 * K + 5
alg = 4 * P * (2 * words_in + 2 * words_out + words_in)   # fwd: in+out, bwd: in + cotangents + grads
print(json.dumps({"P": P, "K": K, "torch_ms": res["torch"], "fused_ms": res["fused"], "fused_without_sh_concat_ms": res["fused_nocat"], "speedup": res["torch"] / res["fused"],
                  "fused_algorithmic_bytes": alg, "fused_GBps": alg / (res["fused"] * 1e-3) / 1e9}))
