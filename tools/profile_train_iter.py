"""Where the time of one training iteration (train.py:269-370 through the unchanged render()) goes (dev tool).
Sections timed with CUDA events by monkey-patching the glue's callables, then a torch.profiler table of the top GPU
kernels of one iteration.
usage: python tools/profile_train_iter.py [cfg3|cfg4] [b200|reference] [--fast]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import refglue as G  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
impl = sys.argv[2] if len(sys.argv) > 2 else "b200"
fast = "--fast" in sys.argv
glue = G.bind(impl)
w = G.build_world(glue, cfg, n_views=8, exposure=(cfg == "cfg4"))
G.prime_depth_cache(w)
G.make_data_parallel(w)

sections = {}


def timed(name, fn):
    def wrapper(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        e1.record()
        sections.setdefault(name, []).append((e0, e1, time.perf_counter() - t0))
        return r
    return wrapper


glue.render = timed("render()", glue.render)
glue.fuse_color = timed("fuse_color()", glue.fuse_color)
LU = glue.loss_utils
fns = None
if fast:
    fns = G.fast_fns()
    for k in ("ssim", "compute_photometric_ssim", "render", "fuse_color"):
        setattr(fns, k, timed(k, getattr(fns, k)))
else:
    LU.ssim = timed("ssim", LU.ssim)
    LU.compute_photometric_ssim = timed("photo_ssim", LU.compute_photometric_ssim)
_bw = torch.Tensor.backward
torch.Tensor.backward = timed("backward", _bw)

for i in range(3):
    G.dp_train_step(w, [i], 1, fns=fns)
torch.cuda.synchronize()
sections.clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
N = 4
t0 = time.perf_counter()
for i in range(N):
    G.train_iteration(w, i, fns=fns)
e1.record()
torch.cuda.synchronize()
print(f"[{impl} {cfg} fast={fast}] iteration: {e0.elapsed_time(e1) / N:.2f} ms device, {(time.perf_counter() - t0) / N * 1e3:.2f} ms host")
for k, v in sections.items():
    dev = sum(a.elapsed_time(b) for a, b, _ in v) / N
    host = sum(h for _, _, h in v) / N * 1e3
    print(f"   {k:16s} {dev:8.2f} ms device  {host:8.2f} ms host   ({len(v) // N} calls)")
w.dp.zero_grad()
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    G.train_iteration(w, 0, fns=fns)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=70, max_name_column_width=100))
