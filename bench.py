#!/usr/bin/env python
"""bench.py -- headline benchmark of the IBGS planar Gaussian rasterizer hot path.

Metric (BASELINE.json): forward+backward rasterization of 3M Gaussians at 1920x1080 (render_geo mode: colour +
normal + median plane depth + 4 warped source views), reported as whole-job views/s (ms/view = 1000 * n_gpus *
views_per_step / (value)).  A "step" = every rank renders `views_per_step` different camera views
(forward + backward, gradients accumulated into one flat per-Gaussian arena) and, for N > 1, one NCCL
all-reduce of that arena.  Weak scaling: per-GPU work is fixed.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config cfg3_1080p]

  value   inputs resident in HBM when the timed region starts
  e2e     same step through the public API, but every view's source images / depths / camera block start in
          PINNED HOST memory and are copied H2D inside the timed region (prefetched on a copy stream), and a
          scalar loss is read back D2H every step
`--impl reference` times the UNMODIFIED reference CUDA extension (oracle/_ref, built from /root/reference by
oracle/build_ref.py) on the same scene, metric and step; under torchrun only rank 0 runs it (the reference is
single-GPU: train.py:277-292).  If the extension is not available the CPU oracle port is timed instead.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "fwd+bwd rasterize views/s @1080p, 3M Gaussians (render_geo, 4 src views)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3_1080p")
    # the data-parallel step is a batch of 8 views per GPU (gradient accumulation) followed
    # by one all-reduce of the per-Gaussian gradient arena
    ap.add_argument("--views-per-step", type=int, default=8)
    ap.add_argument("--streams", type=int, default=1,
                    help="CUDA streams the views of a step are spread over (front-end kernels of one view overlap the "
                         "tile renderers of another)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the extra whole-training-step timing (N = 1)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------
def start_clock_sampler(dev_index):
    path = tempfile.mktemp(suffix=".csv")
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        f = open(path, "w")
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                              "-i", str(dev_index)], stdout=f, stderr=subprocess.DEVNULL)
        return p, path
    except Exception:
        return None, path


def stop_clock_sampler(p, path):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    if p is None:
        return out
    p.terminate()
    try:
        p.wait(timeout=5)
    except Exception:
        p.kill()
    try:
        rows = [r.strip().split(",") for r in open(path) if r.strip()]
        sm = [float(r[1]) for r in rows if len(r) >= 9]
        if sm:
            out["samples"] = len(sm)
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(rows[0][2])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for i, n in enumerate(names):
                if any(r[5 + i].strip().lower().startswith("active") for r in rows if len(r) >= 9):
                    out["reasons"].append(n)
        os.unlink(path)
    except Exception:
        pass
    return out


# ----------------------------------------------------------------------------------------------------------
class Workload:
    """Scene + per-rank view batch.  Everything derived (all_map per view, source depths) is prepared
    untimed, as the training loop would have it resident before calling the rasterizer."""

    def __init__(self, args, rank, world, device, impl_mod):
        from ibgs_b200 import synthetic as S
        import ibgs_testutil as U
        self.S, self.U = S, U
        self.device = device
        sc_cpu = S.make_scene(args.config)
        self.P, self.W, self.H = sc_cpu["P"], sc_cpu["W"], sc_cpu["H"]
        self.sc = U.scene_to_device(sc_cpu, device)
        import ibgs_b200.diff_plane_rasterization as dpr
        self.dpr = dpr
        self.sc["src_rendered_depths"] = U.render_src_depths(dpr, self.sc)
        g = torch.Generator().manual_seed(7)
        self.cot = {k: v.to(device) for k, v in S.cotangents(sc_cpu).items()}
        # view batch of this rank: the reference view perturbed by a small rigid motion (seeded per global view id)
        self.views = []
        V = args.views_per_step
        w2c = sc_cpu["w2c"].double().numpy()
        for i in range(V):
            gid = rank * V + i
            rng = np.random.default_rng(1000 + gid)
            D = S._rigid(S._rot_axis_angle(rng.normal(size=3), np.radians(rng.uniform(0.0, 2.0))),
                         rng.uniform(-0.1, 0.1, 3)) if gid > 0 else np.eye(4)
            w2v = D @ w2c
            cam = S.make_camera(w2v, self.W, self.H)
            cam = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in cam.items()}
            cam["all_map"] = S.all_map_for_view(self.sc["means3D"], self.sc["normals_world"], cam["viewmatrix"],
                                                cam["campos"]).contiguous()
            r2s = torch.stack([torch.from_numpy(np.float32(m.numpy().astype(np.float64) @ np.linalg.inv(w2v)))
                               for m in sc_cpu["src_w2c"]]).to(device).contiguous()
            cam["ref_to_src_list"] = r2s
            self.views.append(cam)
        del g

    def scene_for(self, cam, src_images=None, src_depths=None):
        sc = dict(self.sc)
        sc.update({k: cam[k] for k in ("viewmatrix", "projmatrix", "campos", "tanfovx", "tanfovy", "all_map",
                                       "ref_to_src_list")})
        if src_images is not None:
            sc["src_images"], sc["src_rendered_depths"] = src_images, src_depths
        return sc


class OursRunner:
    name = "b200"

    def __init__(self, wl):
        from ibgs_b200 import parallel as PL
        from ibgs_b200 import _native as N
        self.wl, self.N = wl, N
        sc = wl.sc
        self.leaf = {k: sc[k].detach().clone().requires_grad_(True)
                     for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        shapes = {k: tuple(v.shape) for k, v in self.leaf.items()}
        shapes.update(means2D=(wl.P, 3), means2D_abs=(wl.P, 3))
        self.arena = PL.GradArena(shapes, device=wl.device)
        self.m2d = torch.zeros((wl.P, 3), device=wl.device, requires_grad=True)
        self.m2a = torch.zeros((wl.P, 3), device=wl.device, requires_grad=True)
        for k, v in self.leaf.items():
            v.grad = self.arena.views[k]
        self.m2d.grad = self.arena.views["means2D"]
        self.m2a.grad = self.arena.views["means2D_abs"]
        self.last_R = 0

    def view_fwd_bwd(self, sc):
        dpr, U = self.wl.dpr, self.wl.U
        rs = U.make_settings(dpr, sc, render_geo=True)
        am = sc["all_map"].detach().requires_grad_(True)
        res = dpr.GaussianRasterizer(rs)(means3D=self.leaf["means3D"], means2D=self.m2d, means2D_abs=self.m2a,
                                         opacities=self.leaf["opacities"], shs=self.leaf["shs"],
                                         scales=self.leaf["scales"], rotations=self.leaf["rotations"], all_map=am,
                                         accumulate_grads=True)   # gradients add into the arena-backed .grad in-kernel
        cot = self.wl.cot
        torch.autograd.backward([res[0], res[2], res[3], res[5]],
                                [cot["color"], cot["normal"], cot["depth"], cot["warped"]])
        return res[0]

    def launches(self):
        return int(self.N.lib.ibgs_launch_count())


class RefRunner:
    name = "reference"

    def __init__(self, wl):
        from ibgs_b200 import parallel as PL
        from oracle import ref_ext
        self.wl, self.ref = wl, ref_ext
        ref_ext.load("dpr")
        sc = wl.sc
        shapes = {k: tuple(sc[k].shape) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        shapes.update(means2D=(wl.P, 3), means2D_abs=(wl.P, 3))
        self.arena = PL.GradArena(shapes, device=wl.device)

    def view_fwd_bwd(self, sc):
        fw = self.ref.forward(sc, render_geo=True)
        gr = self.ref.backward(sc, fw, self.wl.cot, render_geo=True)
        self.arena.accumulate({"means3D": gr["means3D"], "shs": gr["sh"], "opacities": gr["opacities"],
                               "scales": gr["scales"], "rotations": gr["rotations"], "means2D": gr["means2D"],
                               "means2D_abs": gr["means2D_abs"]})
        return fw["color"]

    def launches(self):
        return 0


STREAMS = []


def run_steps(runner, wl, steps, world, e2e=False, stager=None):
    """K steps; returns the last colour image's checksum tensor (device)."""
    loss = None
    main = torch.cuda.current_stream(wl.device)
    for _ in range(steps):
        runner.arena.zero_()
        for st in STREAMS:
            st.wait_stream(main)
        for vi, cam in enumerate(wl.views):
            if e2e:
                simg, sdep, cam_d = stager.fetch(vi)
                cam2 = dict(cam)
                cam2.update(cam_d)
                sc = wl.scene_for(cam2, simg, sdep)
            else:
                sc = wl.scene_for(cam)
            if STREAMS and not e2e:
                with torch.cuda.stream(STREAMS[vi % len(STREAMS)]):
                    color = runner.view_fwd_bwd(sc)
            else:
                color = runner.view_fwd_bwd(sc)
            if e2e:
                stager.release(vi)
        for st in STREAMS:
            main.wait_stream(st)
        if world > 1:
            runner.arena.all_reduce()
        if e2e:
            loss = float((color * wl.cot["color"]).sum().item())   # D2H read of the step's scalar result
    return loss


class HostStager:
    """Pinned-host copies of every view's per-step inputs + double-buffered H2D prefetch on a copy stream."""

    def __init__(self, wl):
        self.wl = wl
        dev = wl.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.host = []
        for cam in wl.views:
            h = dict(src_images=wl.sc["src_images"].cpu().pin_memory(),
                     src_depths=wl.sc["src_rendered_depths"].cpu().pin_memory(),
                     viewmatrix=cam["viewmatrix"].cpu().pin_memory(), projmatrix=cam["projmatrix"].cpu().pin_memory(),
                     campos=cam["campos"].cpu().pin_memory(), ref_to_src_list=cam["ref_to_src_list"].cpu().pin_memory(),
                     src_cam_pos=wl.sc["src_cam_pos"].cpu().pin_memory())
            self.host.append(h)
        self.nbuf = 2
        self.dev = [{k: torch.empty_like(v, device=dev) for k, v in self.host[0].items()} for _ in range(self.nbuf)]
        self.ready = [torch.cuda.Event() for _ in range(self.nbuf)]
        self.free = [torch.cuda.Event() for _ in range(self.nbuf)]
        self.bytes_per_view = sum(v.numel() * v.element_size() for v in self.host[0].values())
        self.counter = 0
        self.issued = -1
        for e in self.free:
            e.record(torch.cuda.current_stream(dev))

    def _issue(self, n):
        b = n % self.nbuf
        vi = n % len(self.host)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[b])
            for k, v in self.host[vi].items():
                self.dev[b][k].copy_(v, non_blocking=True)
            self.ready[b].record(self.copy_stream)
        self.issued = n

    def fetch(self, vi):
        n = self.counter
        if self.issued < n:
            self._issue(n)
        if self.issued < n + 1:
            self._issue(n + 1)          # prefetch the next view while this one computes
        b = n % self.nbuf
        torch.cuda.current_stream(self.wl.device).wait_event(self.ready[b])
        d = self.dev[b]
        cam = {k: d[k] for k in ("viewmatrix", "projmatrix", "campos", "ref_to_src_list")}
        return d["src_images"], d["src_depths"], cam

    def release(self, vi):
        b = self.counter % self.nbuf
        self.free[b].record(torch.cuda.current_stream(self.wl.device))
        self.counter += 1


PREWARM_STEPS = 2   # untimed set-up passes: let torch's caching allocator reach its steady state (the per-view
                    # state buffers are tens of MB to GB; a cold cache means cudaMalloc + implicit syncs)


def timed(runner, wl, steps, warmup, world, e2e=False, stager=None):
    dev = wl.device
    run_steps(runner, wl, PREWARM_STEPS, world, e2e, stager)
    torch.cuda.synchronize(dev)
    import gc
    gc.collect()
    run_steps(runner, wl, warmup, world, e2e, stager)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(runner, wl, steps, world, e2e, stager)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def cpu_baseline(args):
    """Float64 CPU oracle (oracle/ibgs_oracle.c, OpenMP) on a bounded sample of the same workload."""
    from ibgs_b200 import synthetic as S
    from oracle import oracle as O
    P0, W, H, _ = S.CONFIGS[args.config]
    Ps = min(P0, 3_000_000)   # ~15 s of CPU work on 16 cores for the headline scene (full workload, no scaling)
    sc = S.make_scene(args.config, P=Ps)
    sc["src_rendered_depths"] = torch.full((4, 1, H, W), 4.0)
    cot = S.cotangents(sc)
    t = time.perf_counter()
    fw = O.forward(sc)
    O.backward(sc, fw, cot)
    dt = time.perf_counter() - t
    scale = P0 / Ps
    return {"value": 1.0 / (dt * scale), "unit": "views/s", "cores": os.cpu_count(), "kind": "port",
            "sample": (f"the full {P0}-Gaussian scene" if Ps == P0 else
                       f"first-{Ps}-Gaussian subsample of the {P0}-Gaussian scene (scaled x{scale:.1f}, linear in P)") +
                      f" at {W}x{H}, 1 view fwd+bwd in {dt:.2f}s; float64 C oracle with OpenMP"}


def prologue_timing(P, device, K=9, iters=10):
    """SURVEY.md section 8f rank 1 (next row, reported beside the headline, not part of it): forward+backward of the
    per-view parameter prologue at the workload's P -- the reference's torch expressions vs ibgs_b200.fused."""
    import prologue_ref as PR
    from ibgs_b200 import fused
    names = ("xyz", "opacity_raw", "scaling_raw", "rotation_raw", "fdc", "frest", "normal_raw", "offset")
    p = PR.random_params(P, K=K, seed=0, device=device)
    g = torch.Generator().manual_seed(1)
    cots = [torch.randn(s, generator=g).to(device) for s in ((P, 1), (P, 3), (P, 4), (P, K, 3), (P, 5))]
    out = {}
    def fused_nocat(*a):
        return fused.gaussian_prologue(*a, concat_sh=False)

    for name, fn in (("torch", PR.torch_prologue), ("fused", fused.gaussian_prologue),
                     ("fused_without_sh_concat", fused_nocat)):
        leaves = {k: p[k].clone().requires_grad_(True) for k in names}

        def step():
            for v in leaves.values():
                v.grad = None
            outs = list(fn(*[leaves[k] for k in names], p["V"], p["cam"]))
            torch.autograd.backward(outs, cots if len(outs) == 5 else [cots[0], cots[1], cots[2], cots[4]])
        for _ in range(3):
            step()
        torch.cuda.synchronize(device)
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); torch.cuda.synchronize(device)
            ts.append(e0.elapsed_time(e1))
        out[name + "_ms_per_view"] = sorted(ts)[len(ts) // 2]
        del leaves
    words_in, words_out = 15 + 3 * K, 13 + 3 * K
    alg = 4 * P * (3 * words_in + 2 * words_out)
    out["fused_GBps"] = alg / (out["fused_ms_per_view"] * 1e-3) / 1e9
    return out


def depth_batch_timing(wl, iters=10):
    """SURVEY.md section 8f rank 2 (next row, reported beside the headline, not part of it): the source-view depth
    renders of one test view (gaussian_renderer/__init__.py:245-253) -- V rasterizer calls, each behind the torch
    all_map construction, vs ONE ibgs_forward_depth_batch call that derives the plane terms in-kernel."""
    import ibgs_b200.depth_batch as DB
    S, U, dpr, sc, dev = wl.S, wl.U, wl.dpr, wl.sc, wl.device
    cams = []
    for i in range(sc["nb_src"]):
        cam = S.src_view(sc, i)
        cams.append({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in cam.items()})
    st = DB.DepthBatchSettings(sc["H"], sc["W"], sc["tanfovx"], sc["tanfovy"], 1.0,
                               torch.stack([c["viewmatrix"] for c in cams]),
                               torch.stack([c["projmatrix"] for c in cams]), 4)
    centers = torch.stack([c["campos"] for c in cams])
    rss = [U.make_settings(dpr, sc, render_geo=False, render_depth_only=True, cam=c) for c in cams]
    z = torch.zeros_like(sc["means3D"])

    def singles():
        with torch.no_grad():
            for c, rs in zip(cams, rss):
                am = S.all_map_for_view(sc["means3D"], sc["normals_world"], c["viewmatrix"], c["campos"])
                dpr.GaussianRasterizer(rs)(means3D=sc["means3D"], means2D=z, means2D_abs=z, opacities=sc["opacities"],
                                           shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"], all_map=am)

    def batch():
        DB.render_depth_batch(st, sc["means3D"], sc["opacities"], scales=sc["scales"], rotations=sc["rotations"],
                              normals=sc["normals_world"], camera_centers=centers)

    out = {"views": len(cams)}
    for name, fn in (("per_view_calls_ms", singles), ("one_batch_call_ms", batch)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        out[name] = sorted(ts)[len(ts) // 2]
    return out


def train_step_timing(wl, impl, steps=4, warmup=3):
    """One whole training step around the rasterizer, per view (extra key, N = 1): raw GaussianModel parameters ->
    per-view prologue -> render_geo rasterizer -> image loss (L1 + SSIM, train.py:302-305) + multi-view photometric
    loss over the 4 warped source images (L1 + SSIM map, train.py:318-336) + a normal / depth term -> backward;
    every 8 views one Adam step over the eight parameter groups + zero_grad (train.py:422-424).
    impl "b200": every stage through this repo (fused prologue without the SH concat, split-SH rasterizer, fused SSIM,
    one-launch ArenaAdam).  impl "reference": the reference's way (its torch expressions for prologue and SSIM, its
    unmodified CUDA rasterizer, torch.optim.Adam).  The colour-aggregation network is not part of either."""
    import prologue_ref as PR
    S, U, sc, dev = wl.S, wl.U, wl.sc, wl.device
    P = wl.P
    op = sc["opacities"].clamp(1e-4, 1 - 1e-4)
    raw = {"xyz": sc["means3D"].clone(), "f_dc": sc["shs"][:, :1, :].contiguous(), "f_rest": sc["shs"][:, 1:, :].contiguous(),
           "opacity": torch.log(op / (1 - op)), "scaling": torch.log(sc["scales"]), "rotation": sc["rotations"].clone(),
           "normal": sc["normals_world"].clone(), "offset": torch.zeros((P, 1), device=dev)}
    lrs = {"xyz": 1.6e-4, "f_dc": 0.0025, "f_rest": 0.0025 / 20, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001,
           "normal": 0.001, "offset": 1.6e-5}
    g = torch.Generator().manual_seed(5)
    gt = torch.rand((3, wl.H, wl.W), generator=g).to(dev)
    nref = torch.nn.functional.normalize(torch.randn((3, wl.H, wl.W), generator=g), dim=0).to(dev)
    Vn = len(wl.views)
    if impl == "b200":
        from ibgs_b200.fused import gaussian_prologue
        from ibgs_b200.loss_utils import ssim, compute_photometric_ssim
        from ibgs_b200.optim import ArenaAdam
        opt = ArenaAdam(raw, lrs)
        pr = opt.params
        z = torch.zeros((P, 3), device=dev)

        def render(cam, scv):
            opacity, scales, rotations, all_map = gaussian_prologue(
                pr["xyz"], pr["opacity"], pr["scaling"], pr["rotation"], pr["f_dc"], pr["f_rest"], pr["normal"],
                pr["offset"], cam["viewmatrix"], cam["campos"], concat_sh=False)
            rs = U.make_settings(wl.dpr, scv, render_geo=True)
            res = wl.dpr.GaussianRasterizer(rs)(means3D=pr["xyz"], means2D=z, means2D_abs=z, opacities=opacity,
                                                shs=pr["f_dc"], shs_rest=pr["f_rest"], scales=scales,
                                                rotations=rotations, all_map=all_map, accumulate_grads=True)
            return res[0], res[2], res[3], res[5]

        def finish():
            opt.step(grad_scale=1.0 / Vn, zero_grads=True)
    else:
        from oracle import ref_ext
        from ssim_ref import torch_ssim_map
        pr = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
        opt = torch.optim.Adam([{"params": [pr[k]], "lr": lrs[k], "name": k} for k in pr], lr=0.0, eps=1e-15)

        def ssim(a, b):
            return torch_ssim_map(a, b).mean()

        def compute_photometric_ssim(a, b, size_average=True):
            m = torch_ssim_map(a, b)
            return m.mean() if size_average else m

        def render(cam, scv):
            opacity, scales, rotations, shs, all_map = PR.torch_prologue(
                pr["xyz"], pr["opacity"], pr["scaling"], pr["rotation"], pr["f_dc"], pr["f_rest"], pr["normal"],
                pr["offset"], cam["viewmatrix"], cam["campos"])
            return ref_ext.RefRasterize.apply(pr["xyz"], shs, opacity, scales, rotations, all_map, scv)

        def finish():
            opt.step()
            opt.zero_grad(set_to_none=True)

    def step():
        for cam in wl.views:
            scv = wl.scene_for(cam)
            image, normal, depth, warped = render(cam, scv)
            loss = 0.8 * (image - gt).abs().mean() + 0.2 * (1.0 - ssim(image, gt))
            w = warped.view(5, 3, wl.H, wl.W)[:4]
            ph_ssim = 1 - torch.stack([compute_photometric_ssim(gt, w[i], size_average=False).mean(0) for i in range(4)])
            ph_l1 = (gt[None] - w).abs().mean(1)
            loss = loss + 0.15 * (0.15 * ph_l1 + 0.85 * ph_ssim).mean()
            loss = loss + 0.015 * (1 - (normal * nref).sum(0)).mean() + 0.01 * depth.mean()
            (loss / 1.0).backward()
        finish()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / (steps * Vn)
    return {"ms_per_view": ms, "views_per_s": 1000.0 / ms, "views_per_step": Vn, "steps": steps,
            "stages": "prologue + rasterizer fwd/bwd + L1/SSIM image loss + 4-view photometric L1/SSIM + Adam every "
                      f"{Vn} views; no colour-aggregation network"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return 0                                     # the reference is single-GPU; only rank 0 runs it
    if not torch.cuda.is_available():
        print(json.dumps({"impl": args.impl, "error": "no CUDA device: the rasterizer has no CPU path"}))
        return 1
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    eff_world = 1 if args.impl == "reference" else world
    if eff_world > 1:
        dist.init_process_group("nccl", device_id=device)

    from ibgs_b200 import synthetic as S
    from ibgs_b200 import _native as N
    wl = Workload(args, rank if eff_world > 1 else 0, eff_world, device, None)
    kind = "reference"
    if args.impl == "reference":
        from oracle import ref_ext
        if not ref_ext.available("dpr"):
            cb = cpu_baseline(args)
            line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "views/s", "n_gpus": args.gpus,
                    "steps": 1, "warmup": 0, "ms_per_step": 1000.0 / cb["value"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": args.config}, "cpu_baseline": cb,
                    "e2e": {"value": cb["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            print(json.dumps(line))
            return 0
        runner = RefRunner(wl)
    else:
        runner = OursRunner(wl)
        N.lib.ibgs_profile_enable(0 if os.environ.get("IBGS_BENCH_NOPROF") else 1)

    if args.streams > 1 and args.impl == "b200":
        STREAMS.extend(torch.cuda.Stream(device=device) for _ in range(args.streams))
    V = args.views_per_step
    # ---- device-resident arm ---------------------------------------------------------------------------
    sampler, spath = start_clock_sampler(local_rank) if (rank == 0 and not os.environ.get("IBGS_BENCH_NOCLOCK")) else (None, "")
    N.lib.ibgs_profile_reset()
    l0 = runner.launches()
    ms = timed(runner, wl, args.steps, args.warmup, eff_world)
    launches = ((runner.launches() - l0) * args.steps // (args.steps + args.warmup + PREWARM_STEPS)
                if args.impl == "b200" else None)
    stages = N.profile_read() if args.impl == "b200" else {}
    clocks = stop_clock_sampler(sampler, spath) if rank == 0 else {}
    views = eff_world * V * args.steps
    value = views / (ms / 1000.0)

    # ---- end-to-end arm (host buffers) -------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        N.lib.ibgs_profile_enable(0)
        stager = HostStager(wl)
        ms_e = timed(runner, wl, args.steps, args.warmup, eff_world, e2e=True, stager=stager)
        e2e = {"value": views / (ms_e / 1000.0), "unit": "views/s",
               "h2d_bytes_per_step": int(stager.bytes_per_view * V), "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e / args.steps}

    if rank != 0:
        if eff_world > 1:
            dist.destroy_process_group()
        return 0

    # ---- bookkeeping ---------------------------------------------------------------------------------------
    wl.dpr.KEEP_STATE = True
    with torch.no_grad():
        sc0 = wl.scene_for(wl.views[0])
        rs = wl.U.make_settings(wl.dpr, sc0, render_geo=True)
        z = torch.zeros_like(sc0["means3D"])
        wl.dpr.GaussianRasterizer(rs)(means3D=sc0["means3D"], means2D=z, means2D_abs=z, opacities=sc0["opacities"],
                                      shs=sc0["shs"], scales=sc0["scales"], rotations=sc0["rotations"],
                                      all_map=sc0["all_map"])
    wl.dpr.KEEP_STATE = False
    st = wl.dpr.LAST_STATE
    R = int(st["num_rendered"])
    dec = wl.U.decode_ours(st)
    pairs = int(dec["n_contrib"].long().sum().item())          # upper bound on blended pairs (last contributor idx)
    P_vis = int((dec["tiles_touched"] > 0).sum().item())
    Npix = wl.W * wl.H
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    roofline = None
    if args.impl == "b200" and stages.get("render_backward", (0, 0))[1] > 0:
        bw_ms, bw_n = stages["render_backward"]
        per_launch_ms = bw_ms / bw_n
        # algorithmic bytes of the backward tile renderer: point list + one 64 B record
        # read and one 64 B accumulator write per visible Gaussian + 212 B per pixel of cotangents / saved state
        # + the per-pixel median-pair lists (12 B written and read back per entry, <= buffer_length entries)
        alg = 4 * R + 128 * P_vis + 212 * Npix + 24 * 4 * Npix
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json"))).get("bytes_per_launch")
        except Exception:
            pass
        ach = alg / (per_launch_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "render_backward_pairs_kernel<geo>", "achieved": ach, "peak": peak_gbs,
                    "unit": "GB/s", "frac": ach / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg, "ms_per_launch": per_launch_ms,
                    "note": "the pair loop is issue/latency bound, not HBM bound (DESIGN.md): see pairs_per_s"}
    bytes_view = 730 * wl.P + 28 * R + 460 * Npix              # SURVEY.md section 8d whole-view figure
    ms_view = ms / (V * args.steps)
    line = {
        "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": args.gpus if args.impl == "b200" else 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {wl.P} Gaussians, {wl.W}x{wl.H}, render_geo, 4 src views, sh_degree 2, "
                               f"buffer_length 4", "views_per_step_per_gpu": V, "num_rendered_R": R,
                   "visible_gaussians": P_vis, "sum_n_contrib": pairs,
                   "l2": "per-view working set (192 MB records + 456 MB inputs + lists) exceeds the 126 MB L2; no flush",
                   "parallelism": f"dp{eff_world} over views, replicated Gaussians, 1 all-reduce of "
                                  f"{runner.arena.nbytes / 1e6:.0f} MB per step" if eff_world > 1 else "single GPU"},
        "ms_per_view": ms_view,
        "view_roofline": {"algorithmic_bytes_per_view": bytes_view, "achieved_GBps": bytes_view / (ms_view * 1e-3) / 1e9,
                          "frac_of_hbm_peak": bytes_view / (ms_view * 1e-3) / 1e9 / peak_gbs},
        "pairs_per_s": pairs / (ms_view * 1e-3),
        "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roofline,
        "stages_ms_per_launch": {k: (v[0] / v[1] if v[1] else None) for k, v in stages.items()},
    }
    if args.impl == "reference":
        line["impl"] = "reference"
        line["gpu_launches"] = None
        line["cpu_baseline"] = {"kind": kind, "cores": 0, "value": value, "unit": "views/s",
                                "sample": "the reference has no CPU path: this arm is its unmodified CUDA extension "
                                          "(oracle/_ref) on the same GPU, full workload"}
    elif args.gpus == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    if args.gpus == 1 and not args.no_train_step:
        try:
            if args.impl == "b200":
                del runner
            torch.cuda.empty_cache()
            line["train_step"] = train_step_timing(wl, args.impl)
        except Exception as ex:  # extra information only: never lose the headline line over it
            line["train_step"] = {"error": repr(ex)}
        runner = None
    if args.impl == "b200" and args.gpus == 1:
        try:
            line["source_depth_batch"] = depth_batch_timing(wl)
        except Exception as ex:  # extra information only: never lose the headline line over it
            line["source_depth_batch"] = {"error": repr(ex)}
        try:  # one SSIM loss term (forward+backward) at the workload's resolution
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import ssim_bench
            line["loss_ssim"] = ssim_bench.measure(wl.H, wl.W, iters=10, peak_gbs=peak_gbs)
        except Exception as ex:
            line["loss_ssim"] = {"error": repr(ex)}
        try:  # the optimizer step over the eight per-Gaussian parameter groups
            import adam_bench
            line["optimizer_step"] = adam_bench.measure(wl.P, iters=10, peak_gbs=peak_gbs)
        except Exception as ex:
            line["optimizer_step"] = {"error": repr(ex)}
        try:
            del runner
            torch.cuda.empty_cache()
            line["parameter_prologue"] = prologue_timing(wl.P, device)
        except Exception as ex:  # extra information only: never lose the headline line over it
            line["parameter_prologue"] = {"error": repr(ex)}
    print(json.dumps(line))
    if eff_world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
