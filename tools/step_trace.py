"""Per-step wall times of the bench step (dev tool): shows whether slowness is uniform or spiky."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
class A: pass
a = A(); a.config = "cfg3_1080p"; a.views_per_step = 2
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
wl = bench.Workload(a, 0, 1, dev, None)
impl = sys.argv[1] if len(sys.argv) > 1 else "b200"
runner = bench.OursRunner(wl) if impl == "b200" else bench.RefRunner(wl)
stager = bench.HostStager(wl)
for mode in ("resident", "e2e", "resident", "e2e"):
    ts = []
    for i in range(14):
        torch.cuda.synchronize(); t = time.perf_counter()
        bench.run_steps(runner, wl, 1, 1, mode == "e2e", stager)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    print(impl, mode, " ".join(f"{x:6.1f}" for x in ts), flush=True)
    print("   mem allocated GB", torch.cuda.memory_allocated() / 1e9, "reserved GB", torch.cuda.memory_reserved() / 1e9,
          "num_alloc_retries", torch.cuda.memory_stats().get("num_alloc_retries"), "segments", torch.cuda.memory_stats().get("segment.all.current"),
          "cudaMalloc calls", torch.cuda.memory_stats().get("num_device_alloc"), "cudaFree calls", torch.cuda.memory_stats().get("num_device_free"))
