"""torch.profiler view of bench.py's whole-training-step loop (dev tool): top GPU kernels per view.
usage: python tools/profile_train_step.py [b200|reference]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench

class A: pass
a = A(); a.config = "cfg3_1080p"; a.views_per_step = 8
impl = sys.argv[1] if len(sys.argv) > 1 else "b200"
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
wl = bench.Workload(a, 0, 1, dev, None)
from torch.profiler import profile, ProfilerActivity
bench.train_step_timing(wl, impl, steps=1, warmup=2)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    r = bench.train_step_timing(wl, impl, steps=1, warmup=0)
print(r)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
