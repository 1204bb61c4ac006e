"""Where does the time between kernels go (dev tool)?  Per view: device time of forward / backward from CUDA events, host
time spent inside the two calls, and the sum of the per-stage kernel times from a second pass with the stage timer on."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from ibgs_b200 import _native as N
class A: pass
a = A(); a.config = sys.argv[1] if len(sys.argv) > 1 else "cfg3_1080p"; a.views_per_step = 8
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
impl = bench.Impl("b200")
wl = bench.Workload(a, 0, 1, dev, impl)
r = bench.OursRunner(wl)
U, dpr, cot = wl.U, impl.dpr, wl.cot
N.lib.ibgs_profile_enable(0)
def one(cam, rec):
    sc = wl.scene_for(cam)
    rs = U.make_settings(dpr, sc, render_geo=True)
    am = sc["all_map"].detach().requires_grad_(True)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); t0 = time.perf_counter()
    res = dpr.GaussianRasterizer(rs)(means3D=r.leaf["means3D"], means2D=r.m2d, means2D_abs=r.m2a, opacities=r.leaf["opacities"],
                                     shs=r.leaf["shs"], scales=r.leaf["scales"], rotations=r.leaf["rotations"], all_map=am,
                                     accumulate_grads=True)
    t1 = time.perf_counter(); e[1].record(); e[2].record()
    torch.autograd.backward([res[0], res[2], res[3], res[5]], [cot["color"], cot["normal"], cot["depth"], cot["warped"]])
    t2 = time.perf_counter(); e[3].record()
    rec.append((e, t1 - t0, t2 - t1))
for rep in range(3):
    rec = []
    torch.cuda.synchronize(); tw = time.perf_counter()
    for cam in wl.views:
        one(cam, rec)
    torch.cuda.synchronize(); tw = time.perf_counter() - tw
    f = sum(x[0][0].elapsed_time(x[0][1]) for x in rec) / len(rec)
    b = sum(x[0][2].elapsed_time(x[0][3]) for x in rec) / len(rec)
    print(f"rep {rep}: wall/view {tw / len(rec) * 1e3:.3f} ms | device fwd {f:.3f} bwd {b:.3f} | host in fwd call {sum(x[1] for x in rec) / len(rec) * 1e3:.3f} in bwd call {sum(x[2] for x in rec) / len(rec) * 1e3:.3f} ms", flush=True)
N.lib.ibgs_profile_reset(); N.lib.ibgs_profile_enable(1)
rec = []
for cam in wl.views:
    one(cam, rec)
torch.cuda.synchronize()
st = N.profile_read()
fw = ("preprocess", "depth_order_sort", "scan", "duplicate_with_keys", "radix_sort", "identify_tile_ranges", "texture_fill", "render_forward")
print("stage sums: fwd", round(sum(st[k][0] / st[k][1] for k in fw if st[k][1]), 3), "bwd", round(sum(st[k][0] / st[k][1] for k in ("render_backward", "preprocess_backward") if st[k][1]), 3))
print({k: round(v[0] / v[1], 3) for k, v in st.items() if v[1]})
