"""Per-stage device times of one view (forward + backward) on a synthetic scene, from the library's built-in event
timer, plus gradient parity against the reference extension (dev tool for kernel experiments).
usage: [IBGS_B200_LIB=<variant .so>] python tools/stage_times.py cfg3_1080p [--iters 10] [--no-ref] [--bwd-variant N]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ibgs_b200 import synthetic as S  # noqa: E402
from ibgs_b200 import _native as N  # noqa: E402
import ibgs_b200.diff_plane_rasterization as dpr  # noqa: E402
import ibgs_testutil as U  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--bwd-variant", type=int, default=0)
    ap.add_argument("--fwd-variant", type=int, default=0)
    ap.add_argument("--geo", type=int, default=1)
    a = ap.parse_args()
    if a.bwd_variant:
        N.check(N.lib.ibgs_set_backward_variant(a.bwd_variant), "bwd variant")
    if a.fwd_variant:
        N.check(N.lib.ibgs_set_forward_variant(a.fwd_variant), "fwd variant")
    sc = U.scene_to_device(S.make_scene(a.name))
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    cot = {k: v.cuda() for k, v in S.cotangents(sc).items()}
    geo = bool(a.geo)
    for _ in range(3):
        U.ours_forward_backward(dpr, sc, cot, render_geo=geo, keep_state=False)
    torch.cuda.synchronize()
    N.lib.ibgs_profile_reset()
    N.lib.ibgs_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        outs, grads, _ = U.ours_forward_backward(dpr, sc, cot, render_geo=geo, keep_state=False)
    e1.record()
    torch.cuda.synchronize()
    N.lib.ibgs_profile_enable(0)
    st = N.profile_read()
    tag = os.environ.get("IBGS_B200_LIB", "default")
    line = " ".join(f"{k}={v[0] / v[1]:.3f}" for k, v in st.items() if v[1])
    tot = sum(v[0] / v[1] for v in st.values() if v[1])
    print(f"[{os.path.basename(os.path.dirname(tag))}] {a.name} geo={geo} kernels_sum={tot:.3f} ms  wall/iter={e0.elapsed_time(e1) / a.iters:.3f}  {line}", flush=True)
    if not a.no_ref:
        from oracle import ref_ext
        fw = ref_ext.forward(sc, render_geo=geo)
        rg = ref_ext.backward(sc, fw, cot, render_geo=geo)
        errs = {k: U.rel_l2(grads[k], rg[k].view_as(grads[k])) for k in U.GRAD_NAMES if grads.get(k) is not None}
        mx = {k: float((outs[k] - fw[k]).abs().max()) for k in ("color", "normal", "depth", "warped")}
        print("   grad rel-L2 vs reference:", " ".join(f"{k}={v:.2e}" for k, v in errs.items()),
              "| max-abs:", " ".join(f"{k}={v:.1e}" for k, v in mx.items()), flush=True)


if __name__ == "__main__":
    main()
