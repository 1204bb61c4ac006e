"""Timeline of one bench step with torch.profiler (dev tool): GPU busy time vs wall, top kernels, gaps."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench

class A: pass
a = A(); a.config = sys.argv[1] if len(sys.argv) > 1 else "cfg3_1080p"; a.views_per_step = 2
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
# H2D bandwidth probe
h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device=dev)
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(4): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"H2D pinned bandwidth: {4*h.numel()/dt/1e9:.1f} GB/s")
t = time.perf_counter()
for _ in range(4): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"D2H pinned bandwidth: {4*h.numel()/dt/1e9:.1f} GB/s")
del h, d
wl = bench.Workload(a, 0, 1, dev, None)
impl = sys.argv[2] if len(sys.argv) > 2 else "b200"
runner = bench.OursRunner(wl) if impl == "b200" else bench.RefRunner(wl)
e2e = len(sys.argv) > 3 and sys.argv[3] == "e2e"
stager = bench.HostStager(wl) if e2e else None
bench.run_steps(runner, wl, 3, 1, e2e, stager)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    bench.run_steps(runner, wl, 2, 1, e2e, stager)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(e.device_time for e in evs) / 1e3 if hasattr(evs[0], "device_time") else sum(e.cuda_time for e in evs) / 1e3
print(f"wall {wall:.2f} ms for 2 steps (4 views); GPU kernel+memcpy time {busy:.2f} ms; idle {wall-busy:.2f} ms")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
# gaps: sort GPU events by start, list the largest gaps and what preceded them
evs.sort(key=lambda e: e.time_range.start)
gaps = []
for p, n in zip(evs[:-1], evs[1:]):
    g = n.time_range.start - p.time_range.end
    if g > 30: gaps.append((g, p.name[:50], n.name[:50]))
gaps.sort(reverse=True)
print("largest GPU gaps (us): ")
for g in gaps[:25]: print(f"  {g[0]:8.0f}  after {g[1]:50s} before {g[2]}")
print("total gap >30us:", sum(g[0] for g in gaps)/1e3, "ms")
