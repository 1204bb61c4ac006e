"""Quick A/B timing of this implementation vs the reference extension on one synthetic scene (dev tool).
usage: python tools/quick_ab.py cfg2 [--iters 10] [--no-ref]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ibgs_b200 import synthetic as S  # noqa: E402
import ibgs_b200.diff_plane_rasterization as dpr  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ibgs_testutil as U  # noqa: E402


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--geo", type=int, default=1)
    ap.add_argument("--depth-only", action="store_true", help="time the depth-only forward (render.py's source-depth renders)")
    ap.add_argument("--depth-batch", type=int, default=0, metavar="V",
                    help="time V source-view depth renders: V single calls vs one ibgs_forward_depth_batch call")
    ap.add_argument("--P", type=int, default=0, help="override the number of Gaussians of the scene")
    ap.add_argument("--bwd-variant", type=int, default=0, help="force the backward variant (1 / 2 pixels per lane; 0 = automatic)")
    a = ap.parse_args()
    if a.bwd_variant:
        from ibgs_b200 import _native as _N
        _N.check(_N.lib.ibgs_set_backward_variant(a.bwd_variant), "ibgs_set_backward_variant")
    t0 = time.time()
    sc = U.scene_to_device(S.make_scene(a.name, P=a.P or None))
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    cot = {k: v.cuda() for k, v in S.cotangents(sc).items()}
    print(f"scene {a.name}: P={sc['P']} {sc['W']}x{sc['H']} built in {time.time()-t0:.1f}s", flush=True)
    if a.depth_batch:
        import ibgs_b200.depth_batch as DB
        V = a.depth_batch
        cams = []
        for i in range(V):
            cam = S.src_view(sc, i % sc["nb_src"])
            cams.append({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in cam.items()})
        st = DB.DepthBatchSettings(sc["H"], sc["W"], sc["tanfovx"], sc["tanfovy"], 1.0,
                                   torch.stack([c["viewmatrix"] for c in cams]), torch.stack([c["projmatrix"] for c in cams]), 4)
        centers = torch.stack([c["campos"] for c in cams])
        all_maps = torch.stack([c["all_map"] for c in cams]).contiguous()
        rss = [U.make_settings(dpr, sc, render_geo=False, render_depth_only=True, cam=c) for c in cams]
        z = torch.zeros_like(sc["means3D"])

        def torch_all_map(cam):   # render_depth's per-view plane parameters (gaussian_renderer/__init__.py:119-132)
            return S.all_map_for_view(sc["means3D"], sc["normals_world"], cam["viewmatrix"], cam["campos"])

        def singles(with_maps):
            with torch.no_grad():
                outs = []
                for c, rs in zip(cams, rss):
                    am = torch_all_map(c) if with_maps else c["all_map"]
                    outs.append(dpr.GaussianRasterizer(rs)(means3D=sc["means3D"], means2D=z, means2D_abs=z,
                                                           opacities=sc["opacities"], shs=sc["shs"], scales=sc["scales"],
                                                           rotations=sc["rotations"], all_map=am)[3])
                return torch.stack(outs)

        def batch_maps():
            return DB.render_depth_batch(st, sc["means3D"], sc["opacities"], scales=sc["scales"],
                                         rotations=sc["rotations"], all_maps=all_maps)

        def batch_fused():
            return DB.render_depth_batch(st, sc["means3D"], sc["opacities"], scales=sc["scales"],
                                         rotations=sc["rotations"], normals=sc["normals_world"], camera_centers=centers)

        same = torch.equal(singles(False), batch_maps())
        err = (batch_fused() - batch_maps()).abs().max().item()
        t1 = timed(lambda: singles(False), a.iters)
        t1m = timed(lambda: singles(True), a.iters)
        t2 = timed(batch_maps, a.iters)
        t3 = timed(batch_fused, a.iters)
        print(f"OURS  {a.name} {V} source-depth renders: {V} calls {t1:.3f} ms ({t1m:.3f} ms with the torch all_map ops)  "
              f"one batch call {t2:.3f} ms (given all_maps) / {t3:.3f} ms (plane terms in-kernel)   "
              f"bit-identical={same} fused-vs-maps max|d|={err:.2e}", flush=True)
        from ibgs_b200 import _native as N
        N.lib.ibgs_profile_enable(1)
        N.lib.ibgs_profile_reset()
        for _ in range(5):
            batch_fused()
        torch.cuda.synchronize()
        print("   batch stages (ms per call): " + "  ".join(f"{k} {ms / 5:.3f}" for k, (ms, n) in N.profile_read().items() if n),
              f"  pixels differing > 1e-4 fused-vs-maps: {((batch_fused() - batch_maps()).abs() > 1e-4).float().mean().item():.2e}", flush=True)
        N.lib.ibgs_profile_enable(0)
        if not a.no_ref:
            from oracle import ref_ext

            def ref_singles():
                for c in cams:
                    am = torch_all_map(c)
                    ref_ext.forward(sc, render_geo=False, render_depth_only=True, cam=dict(c, all_map=am))
            tr = timed(ref_singles, a.iters)
            print(f"REF   {a.name} {V} source-depth renders (torch all_map + rasterizer per view): {tr:.3f} ms   "
                  f"speedup x{tr / t3:.2f}", flush=True)
        return
    if a.depth_only:
        rs = U.make_settings(dpr, sc, render_geo=False, render_depth_only=True)
        z = torch.zeros_like(sc["means3D"])

        def ours_depth():
            with torch.no_grad():
                dpr.GaussianRasterizer(rs)(means3D=sc["means3D"], means2D=z, means2D_abs=z, opacities=sc["opacities"],
                                           shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"], all_map=sc["all_map"])
        t = timed(ours_depth, a.iters)
        line = f"OURS  {a.name} depth-only forward {t:.3f} ms"
        if not a.no_ref:
            from oracle import ref_ext
            tr = timed(lambda: ref_ext.forward(sc, render_geo=False, render_depth_only=True), a.iters)
            line += f"   REF {tr:.3f} ms   speedup x{tr / t:.2f}"
        print(line, flush=True)
        return
    geo = bool(a.geo)
    rs = U.make_settings(dpr, sc, render_geo=geo)
    leaf = {k: sc[k].detach().clone().requires_grad_(True)
            for k in ("means3D", "shs", "opacities", "scales", "rotations", "all_map")}
    m2d = torch.zeros_like(sc["means3D"], requires_grad=True)
    m2a = torch.zeros_like(sc["means3D"], requires_grad=True)
    rast = dpr.GaussianRasterizer(rs)
    state = {}

    def ours_fwd():
        state["res"] = rast(means3D=leaf["means3D"], means2D=m2d, means2D_abs=m2a, opacities=leaf["opacities"],
                            shs=leaf["shs"], scales=leaf["scales"], rotations=leaf["rotations"],
                            all_map=leaf["all_map"] if geo else None)

    def ours_bwd():
        r = state["res"]
        outs = [r[0]] + ([r[2], r[3], r[5]] if geo else [])
        gs = [cot["color"]] + ([cot["normal"], cot["depth"], cot["warped"]] if geo else [])
        torch.autograd.backward(outs, gs)

    def ours_both():
        ours_fwd()
        ours_bwd()

    f = timed(ours_fwd, a.iters)
    ours_fwd()
    torch.cuda.synchronize()
    # backward alone: re-run forward untimed each iteration
    tb = []
    for _ in range(a.iters):
        ours_fwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ours_bwd(); e1.record(); torch.cuda.synchronize()
        tb.append(e0.elapsed_time(e1))
    tb.sort()
    both = timed(ours_both, a.iters)
    U.ours_forward_backward  # noqa
    dpr.KEEP_STATE = True
    ours_fwd()
    dpr.KEEP_STATE = False
    R = dpr.LAST_STATE["num_rendered"]
    print(f"OURS  {a.name} geo={geo}: R={R} fwd {f:.3f} ms  bwd {tb[len(tb)//2]:.3f} ms  fwd+bwd {both:.3f} ms", flush=True)

    if not a.no_ref:
        from oracle import ref_ext
        sc1 = sc
        if not geo:
            H, W = sc["H"], sc["W"]
            sc1 = dict(sc)
            sc1.update(nb_src=1, ref_to_src_list=torch.zeros((1, 16), device="cuda"),
                       src_images=torch.zeros((1, 3, H * W), device="cuda"),
                       src_rendered_depths=torch.zeros((1, 1, H * W), device="cuda"),
                       src_cam_pos=torch.zeros((1, 3), device="cuda"))
        st = {}

        def ref_fwd():
            st["fw"] = ref_ext.forward(sc1, render_geo=geo)

        def ref_bwd():
            ref_ext.backward(sc1, st["fw"], cot, render_geo=geo)

        def ref_both():
            ref_fwd(); ref_bwd()

        rf = timed(ref_fwd, a.iters)
        ref_fwd(); torch.cuda.synchronize()
        rb = timed(ref_bwd, a.iters)
        rboth = timed(ref_both, a.iters)
        print(f"REF   {a.name} geo={geo}: R={st['fw']['num_rendered']} fwd {rf:.3f} ms  bwd {rb:.3f} ms  fwd+bwd {rboth:.3f} ms"
              f"   speedup fwd+bwd x{rboth/both:.2f}", flush=True)


if __name__ == "__main__":
    main()
