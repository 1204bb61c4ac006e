#!/usr/bin/env python
"""bench.py -- headline benchmark of the IBGS planar Gaussian rasterizer hot path.

Metric (BASELINE.json): forward+backward rasterization of 3M Gaussians at 1920x1080 (render_geo mode: colour +
normal + median plane depth + 4 warped source views), reported as whole-job views/s (ms/view = 1000 * n_gpus *
views_per_step / (value * steps) ...).  A "step" = every rank renders `views_per_step` different camera views
(forward + backward, gradients accumulated into one flat per-Gaussian arena) and, for N > 1, one NCCL all-reduce of
the parameter-gradient part of that arena.  Weak scaling: per-GPU work is fixed.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config cfg3_1080p]

  value        inputs resident in HBM when the timed region starts
  e2e          same step through the public API with the step's INPUT DATA starting in pinned host memory: every
               view's four source images (uint8, as an image file decodes; converted on the device) and its camera
               block are copied H2D inside the timed region (prefetched on a copy stream) and a scalar loss is read
               back D2H every step.  The source-view DEPTHS are not host data: they are the rasterizer's own cached
               renders (train.py:299 keeps scene.rendered_depth_list on data_device = cuda) and stay resident.
  train_step   BASELINE metric (ii): whole training iterations through the reference's UNCHANGED glue
               (gaussian_renderer.render -> utils.loss_utils -> color_aggregation_network.fuse_color -> backward ->
               optimizers; tests/refglue.py restates train.py:269-430) on configs 3 and 4, data-parallel over views
               at N > 1 (ibgs_b200.parallel).
`--impl reference` times the UNMODIFIED reference CUDA extension (oracle/_ref, built from /root/reference by
oracle/build_ref.py) on the same scene, metric and step, and the same train steps under the reference's own autograd
wrapper; that process never loads libibgs_b200.so.  Under torchrun only rank 0 runs it (the reference is single-GPU:
train.py:277-292).  If the extension is not available the CPU oracle port is timed instead.
"""
import argparse
import gc
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
PARAM_KEYS = ("means3D", "shs", "opacities", "scales", "rotations")


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
    ap.add_argument("--no-train-step", action="store_true", help="skip the training-step timings through render()")
    ap.add_argument("--no-extras", action="store_true", help="skip the section-8f side measurements and the parity block")
    ap.add_argument("--train-configs", default="cfg3,cfg4", help="configs of the train_step measurement (N = 1); "
                    "N > 1 runs the last one data-parallel")
    ap.add_argument("--train-views", type=int, default=8, help="views per optimisation step and rank in train_step")
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
class Impl:
    """What differs between the two arms: which rasterizer renders the (untimed) source depths, runs the views and
    reports its state.  The reference arm imports nothing of ibgs_b200 except the pure-Python scene generator."""

    def __init__(self, name):
        self.name = name
        if name == "b200":
            import ibgs_b200.diff_plane_rasterization as dpr
            from ibgs_b200 import _native as N
            self.dpr, self.N = dpr, N
        else:
            from oracle import ref_ext
            ref_ext.load("dpr")
            self.ref = ref_ext

    def depth_only(self, sc, cam):
        """Plane depth of one source pose (gaussian_renderer.render_depth), by this arm's own rasterizer."""
        import ibgs_testutil as U
        if self.name == "b200":
            rs = U.make_settings(self.dpr, sc, render_geo=False, render_depth_only=True, cam=cam)
            z = torch.zeros_like(sc["means3D"])
            with torch.no_grad():
                return self.dpr.GaussianRasterizer(rs)(means3D=sc["means3D"], means2D=z, means2D_abs=z,
                                                       opacities=sc["opacities"], shs=sc["shs"], scales=sc["scales"],
                                                       rotations=sc["rotations"], all_map=cam["all_map"])[3]
        return self.ref.forward(sc, render_geo=False, render_depth_only=True, cam=cam)["depth"]


class Workload:
    """Scene + per-rank view batch.  Everything derived (all_map per view, source depths) is prepared untimed, as
    the training loop would have it resident before calling the rasterizer."""

    def __init__(self, args, rank, world, device, impl, config=None, views=None):
        from ibgs_b200 import synthetic as S
        import ibgs_testutil as U
        self.S, self.U, self.impl = S, U, impl
        self.device = device
        sc_cpu = S.make_scene(config or args.config)
        self.sc_cpu = sc_cpu
        self.P, self.W, self.H = sc_cpu["P"], sc_cpu["W"], sc_cpu["H"]
        self.sc = U.scene_to_device(sc_cpu, device)
        nb = self.sc["nb_src"]
        self.sc["src_rendered_depths"] = torch.zeros((nb, 1, self.H, self.W), device=device)
        deps = []
        for i in range(nb):
            cam = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in S.src_view(self.sc, i).items()}
            deps.append(impl.depth_only(self.sc, cam))
        self.sc["src_rendered_depths"] = torch.stack(deps, 0).contiguous()
        self.cot = {k: v.to(device) for k, v in S.cotangents(sc_cpu).items()}
        # view batch of this rank: the reference view perturbed by a small rigid motion (seeded per global view id)
        self.views = []
        V = views or args.views_per_step
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

    def scene_for(self, cam, src_images=None):
        sc = dict(self.sc)
        sc.update({k: cam[k] for k in ("viewmatrix", "projmatrix", "campos", "tanfovx", "tanfovy", "all_map",
                                       "ref_to_src_list")})
        if src_images is not None:
            sc["src_images"] = src_images
        return sc


class _Timer:
    """CUDA-event brackets around the rasterizer calls only (kernel_ms_per_view): what the two implementations spend
    on the device for one view, without the Python / allocator time between calls."""

    def __init__(self):
        self.pairs = []
        self.on = False

    def begin(self):
        if not self.on:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, e0):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.pairs.append((e0, e1))

    def total_ms(self):
        return sum(a.elapsed_time(b) for a, b in self.pairs)


class OursRunner:
    """The headline arm: this repo's rasterizer through its public Python API (-> ctypes -> C ABI -> sm_100a).
    accumulate=True is the batch-of-views mode (gradients added into the arena-backed .grad by the backward kernel);
    accumulate=False is the stock reference API, autograd's AccumulateGrad doing the adds."""
    name = "b200"

    def __init__(self, wl, accumulate=True):
        from ibgs_b200 import parallel as PL
        self.wl, self.N, self.accumulate = wl, wl.impl.N, accumulate
        sc = wl.sc
        self.leaf = {k: sc[k].detach().clone().requires_grad_(True) for k in PARAM_KEYS}
        shapes = {k: tuple(v.shape) for k, v in self.leaf.items()}
        shapes.update(means2D=(wl.P, 3), means2D_abs=(wl.P, 3))     # last two: local densification statistics only
        self.arena = PL.GradArena(shapes, device=wl.device)
        self.m2d = torch.zeros((wl.P, 3), device=wl.device, requires_grad=True)
        self.m2a = torch.zeros((wl.P, 3), device=wl.device, requires_grad=True)
        for k, v in self.leaf.items():
            v.grad = self.arena.views[k]
        self.m2d.grad = self.arena.views["means2D"]
        self.m2a.grad = self.arena.views["means2D_abs"]
        self.timer = _Timer()

    def view_fwd_bwd(self, sc):
        dpr, U = self.wl.impl.dpr, self.wl.U
        rs = U.make_settings(dpr, sc, render_geo=True)
        am = sc["all_map"].detach().requires_grad_(True)
        kw = dict(accumulate_grads=True) if self.accumulate else {}
        t = self.timer.begin()
        res = dpr.GaussianRasterizer(rs)(means3D=self.leaf["means3D"], means2D=self.m2d, means2D_abs=self.m2a,
                                         opacities=self.leaf["opacities"], shs=self.leaf["shs"],
                                         scales=self.leaf["scales"], rotations=self.leaf["rotations"], all_map=am, **kw)
        self.timer.end(t)
        cot = self.wl.cot
        t = self.timer.begin()
        torch.autograd.backward([res[0], res[2], res[3], res[5]],
                                [cot["color"], cot["normal"], cot["depth"], cot["warped"]])
        self.timer.end(t)
        return res[0]

    def all_reduce(self):
        self.arena.all_reduce(upto="means2D")

    def launches(self):
        return int(self.N.lib.ibgs_launch_count())


class RefRunner:
    name = "reference"

    def __init__(self, wl):
        from ibgs_b200 import parallel as PL          # pure torch (no native library)
        self.wl, self.ref = wl, wl.impl.ref
        sc = wl.sc
        shapes = {k: tuple(sc[k].shape) for k in PARAM_KEYS}
        shapes.update(means2D=(wl.P, 3), means2D_abs=(wl.P, 3))
        self.arena = PL.GradArena(shapes, device=wl.device)
        self.timer = _Timer()

    def view_fwd_bwd(self, sc):
        t = self.timer.begin()
        fw = self.ref.forward(sc, render_geo=True)
        self.timer.end(t)
        t = self.timer.begin()
        gr = self.ref.backward(sc, fw, self.wl.cot, render_geo=True)
        self.timer.end(t)
        self.arena.accumulate({"means3D": gr["means3D"], "shs": gr["sh"], "opacities": gr["opacities"],
                               "scales": gr["scales"], "rotations": gr["rotations"], "means2D": gr["means2D"],
                               "means2D_abs": gr["means2D_abs"]})
        return fw["color"]

    def all_reduce(self):
        pass

    def launches(self):
        return 0


STREAMS = []


def run_steps(runner, wl, steps, world, e2e=False, stager=None):
    """K steps; returns the last step's scalar (read back D2H in e2e mode)."""
    loss = None
    main = torch.cuda.current_stream(wl.device)
    for _ in range(steps):
        runner.arena.zero_()
        for st in STREAMS:
            st.wait_stream(main)
        for vi, cam in enumerate(wl.views):
            if e2e:
                simg, cam_d = stager.fetch(vi)
                cam2 = dict(cam)
                cam2.update(cam_d)
                sc = wl.scene_for(cam2, simg)
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
            runner.all_reduce()
        if e2e:
            loss = float((color * wl.cot["color"]).sum().item())   # D2H read of the step's scalar result
    return loss


class HostStager:
    """Pinned-host copies of every view's per-step input data + double-buffered H2D prefetch on a copy stream.
    Source images travel as uint8 (what the dataset's image files decode to; scene/cameras.py:23-40 converts to float
    once at load) and are converted to the float32 planes the rasterizer reads on the device, on the copy stream."""

    def __init__(self, wl):
        self.wl = wl
        dev = wl.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.host = []
        img_u8 = (wl.sc["src_images"].clamp(0, 1) * 255.0).round().to(torch.uint8).cpu()
        for cam in wl.views:
            h = dict(src_images_u8=img_u8.clone().pin_memory(),
                     viewmatrix=cam["viewmatrix"].cpu().pin_memory(), projmatrix=cam["projmatrix"].cpu().pin_memory(),
                     campos=cam["campos"].cpu().pin_memory(), ref_to_src_list=cam["ref_to_src_list"].cpu().pin_memory(),
                     src_cam_pos=wl.sc["src_cam_pos"].cpu().pin_memory())
            self.host.append(h)
        self.nbuf = 2
        self.dev = [{k: torch.empty_like(v, device=dev) for k, v in self.host[0].items()} for _ in range(self.nbuf)]
        self.img_f32 = [torch.empty_like(wl.sc["src_images"]) for _ in range(self.nbuf)]
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
            torch.mul(self.dev[b]["src_images_u8"], 1.0 / 255.0, out=self.img_f32[b])
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
        return self.img_f32[b], cam

    def release(self, vi):
        b = self.counter % self.nbuf
        self.free[b].record(torch.cuda.current_stream(self.wl.device))
        self.counter += 1


PREWARM_STEPS = 2   # untimed set-up passes: let torch's caching allocator reach its steady state (the per-view
                    # state buffers are tens of MB to GB; a cold cache means cudaMalloc + implicit syncs)


def timed(runner, wl, steps, warmup, world, e2e=False, stager=None, prewarm=PREWARM_STEPS):
    dev = wl.device
    run_steps(runner, wl, prewarm, world, e2e, stager)
    torch.cuda.synchronize(dev)
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


# ---- BASELINE metric (ii): training iterations through the reference's unchanged glue ----------------------------------
def train_step_through_render(impl_name, config, device, rank, world, views_per_rank, steps=3, warmup=2, fused_losses=False,
                              fast=False):
    """views/s of whole optimisation steps -- `views_per_rank` x train.py:269-370 (render -> L1/SSIM -> normal loss ->
    multi-view photometric loss -> fuse_color + ColorFusionResidualNet -> backward; + AppModel affine and the exposure
    lstsq on cfg4) followed by the optimizer steps of train.py:421-430 -- through the UNCHANGED gaussian_renderer.render
    with `diff_plane_rasterization` bound to this arm's rasterizer.  At world > 1 the views of a step are sharded over the
    ranks (ibgs_b200.parallel.GaussianDataParallel: one gradient all-reduce, depth-cache exchange).  iter_time is what
    the reference itself logs: CUDA events around render..backward of one view (train.py:269,370)."""
    import refglue as G
    glue = G.bind("b200" if impl_name == "b200" else "reference")
    n_cams = 16
    w = G.build_world(glue, config, n_views=n_cams, device=str(device), exposure=(config == "cfg4"))
    G.prime_depth_cache(w)
    G.make_data_parallel(w)
    fns = None
    if fast:          # every section-8f fast path: fused losses, fused-prologue render(), fused colour aggregation (bf16)
        fns = G.fast_fns(precision="bf16")
    elif fused_losses:
        import ibgs_b200.loss_utils as FL
        import types
        fns = types.SimpleNamespace(ssim=FL.ssim, compute_photometric_ssim=FL.compute_photometric_ssim)
    V = views_per_rank
    total = V * world
    mine = [(rank * V + i) % n_cams for i in range(V)]
    dev = torch.device(device)

    def step():
        G.dp_train_step(w, mine, total, fns=fns)

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    # iter_time as the reference logs it: one view, render..backward
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = G.train_iteration(w, mine[0], fns=fns)
    b.record()
    torch.cuda.synchronize(dev)
    w.dp.zero_grad()
    res = {"config": f"{config}: {w.P} Gaussians, {w.W}x{w.H}" + (", exposure compensation + correction" if config == "cfg4" else ""),
           "views_per_s": total * steps / (ms / 1000.0), "ms_per_view_per_gpu": ms / (steps * V),
           "iter_time_ms": a.elapsed_time(b), "views_per_step": total, "steps": steps, "world": world,
           "loss": float(out["loss"].item()),
           "stages": "unchanged gaussian_renderer.render() + utils.loss_utils L1/SSIM + normal loss + 3-view photometric "
                     "L1/SSIM + fuse_color/ColorFusionResidualNet + backward; torch.optim.Adam steps every "
                     f"{total} views" + ("; SSIM terms through ibgs_b200.loss_utils" if fused_losses and not fast else "")
                     + ("; FAST PATHS: ibgs_b200.gaussian_renderer.render (fused prologue), ibgs_b200.loss_utils (fused SSIM), "
                        "ibgs_b200.color_aggregation.fuse_color (fused feature/MLP kernel + channel-padded NHWC bf16 conv "
                        "decoder) instead of the reference's functions" if fast else "")}
    del w
    gc.collect()
    torch.cuda.empty_cache()
    return res


def dp_check(device, rank, world):
    """N > 1 only: correctness of the sharded step on real devices.  Every rank renders its 2 views of a 2*world-view
    batch on cfg1 and the arenas are all-reduced; rank 0 also renders ALL views alone and compares (rel-L2 <= 1e-3:
    float atomics do not sum in a fixed order)."""
    from ibgs_b200 import synthetic as S
    import ibgs_testutil as U
    import types
    args = types.SimpleNamespace(config="cfg1", views_per_step=2)
    impl = Impl("b200")
    wl = Workload(args, rank, world, device, impl)
    r = OursRunner(wl)
    run_steps(r, wl, 1, world)
    torch.cuda.synchronize(device)
    out = None
    if rank == 0:
        single = OursRunner(wl)
        single.arena.zero_()
        for rr in range(world):
            wl_r = wl if rr == 0 else Workload(args, rr, world, device, impl)
            for cam in wl_r.views:
                single.view_fwd_bwd(wl_r.scene_for(cam))
        off = r.arena.offsets["means2D"][0]
        a, b = r.arena.flat[:off].double(), single.arena.flat[:off].double()
        out = {"config": "cfg1, 2 views per rank", "rel_l2": float(((a - b).norm() / b.norm()).item()), "gate": 1e-3}
        out["ok"] = out["rel_l2"] <= out["gate"]
    return out


def parity_block(wl, runner_cls):
    """Untimed: this repo's rasterizer against the UNMODIFIED reference extension on the benchmark's own scene and view
    (both are in the process here; the reference is the checker, never the thing measured in this arm)."""
    from oracle import ref_ext
    if not ref_ext.available("dpr"):
        return {"unavailable": "oracle/_ref/dpr/ref_dpr_C.so not built"}
    U = wl.U
    sc = wl.scene_for(wl.views[0])
    outs, grads, state = U.ours_forward_backward(wl.impl.dpr, sc, wl.cot, render_geo=True, keep_state=False)
    fw = ref_ext.forward(sc, render_geo=True)
    rg = ref_ext.backward(sc, fw, wl.cot, render_geo=True)
    res = {"scene": "view 0 of the timed batch", "num_rendered_equal": int(state.get("num_rendered", -1)) in (-1, int(fw["num_rendered"])),
           "radii_equal": bool(torch.equal(outs["radii"], fw["radii"])), "mask_equal": bool(torch.equal(outs["mask"], fw["mask"])),
           "max_abs": {k: float((outs[k] - fw[k]).abs().max().item())
                       for k in ("color", "normal", "depth", "cam_feat", "warped", "min_depth_diff", "camera_ray")},
           "grad_rel_l2": {k: U.rel_l2(grads[k], rg[k].view_as(grads[k])) for k in U.GRAD_NAMES},
           "gates": {"max_abs": 1e-4, "grad_rel_l2": 1e-3}}
    res["ok"] = (res["radii_equal"] and res["mask_equal"] and max(res["max_abs"].values()) <= 1e-4
                 and max(res["grad_rel_l2"].values()) <= 1e-3)
    return res


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
    S, U, dpr, sc, dev = wl.S, wl.U, wl.impl.dpr, wl.sc, wl.device
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


def fast_path_step_timing(wl, steps=4, warmup=3):
    """Every section-8f fast path of this repo together (extra key, N = 1, b200 arm only; NOT the reference's glue --
    that is `train_step`): raw parameters -> fused prologue without the SH concat -> split-SH rasterizer with in-kernel
    gradient accumulation -> L1 + fused SSIM image loss + 4-view photometric L1 / fused SSIM map + normal / depth terms
    -> backward; one ArenaAdam launch + zero_grad every 8 views.  No colour-aggregation network."""
    from ibgs_b200.fused import gaussian_prologue
    from ibgs_b200.loss_utils import ssim, compute_photometric_ssim
    from ibgs_b200.optim import ArenaAdam
    U, sc, dev = wl.U, wl.sc, wl.device
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
    opt = ArenaAdam(raw, lrs)
    pr = opt.params
    z = torch.zeros((P, 3), device=dev)
    dpr = wl.impl.dpr

    def step():
        for cam in wl.views:
            scv = wl.scene_for(cam)
            opacity, scales, rotations, all_map = gaussian_prologue(
                pr["xyz"], pr["opacity"], pr["scaling"], pr["rotation"], pr["f_dc"], pr["f_rest"], pr["normal"],
                pr["offset"], cam["viewmatrix"], cam["campos"], concat_sh=False)
            rs = U.make_settings(dpr, scv, render_geo=True)
            res = dpr.GaussianRasterizer(rs)(means3D=pr["xyz"], means2D=z, means2D_abs=z, opacities=opacity,
                                             shs=pr["f_dc"], shs_rest=pr["f_rest"], scales=scales,
                                             rotations=rotations, all_map=all_map, accumulate_grads=True)
            image, normal, depth, warped = res[0], res[2], res[3], res[5]
            loss = 0.8 * (image - gt).abs().mean() + 0.2 * (1.0 - ssim(image, gt))
            w = warped.view(5, 3, wl.H, wl.W)[:4]
            ph_ssim = 1 - torch.stack([compute_photometric_ssim(gt, w[i], size_average=False).mean(0) for i in range(4)])
            ph_l1 = (gt[None] - w).abs().mean(1)
            loss = loss + 0.15 * (0.15 * ph_l1 + 0.85 * ph_ssim).mean()
            loss = loss + 0.015 * (1 - (normal * nref).sum(0)).mean() + 0.01 * depth.mean()
            loss.backward()
        opt.step(grad_scale=1.0 / Vn, zero_grads=True)

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
            "stages": "fused prologue + rasterizer fwd/bwd (in-kernel grad accumulation) + L1/fused-SSIM image loss + "
                      f"4-view photometric L1/fused-SSIM + ArenaAdam every {Vn} views; no colour-aggregation network"}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def load_profile_facts():
    """Per-launch facts read off the committed ncu captures (profiles/dominant_kernel_traffic.json): DRAM bytes and
    executed warp instructions of the two pair-loop kernels on the headline scene."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
    except Exception:
        return {}


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
    impl = Impl(args.impl)
    wl = Workload(args, rank if eff_world > 1 else 0, eff_world, device, impl)
    runner = RefRunner(wl) if args.impl == "reference" else OursRunner(wl)
    N = impl.N if args.impl == "b200" else None
    if N is not None:
        N.lib.ibgs_profile_enable(0)              # the stage timer stays OFF in every timed region

    if args.streams > 1 and args.impl == "b200":
        STREAMS.extend(torch.cuda.Stream(device=device) for _ in range(args.streams))
    V = args.views_per_step
    # ---- device-resident arm ---------------------------------------------------------------------------
    sampler, spath = start_clock_sampler(local_rank) if (rank == 0 and not os.environ.get("IBGS_BENCH_NOCLOCK")) else (None, "")
    l0 = runner.launches()
    ms = timed(runner, wl, args.steps, args.warmup, eff_world)
    launches = ((runner.launches() - l0) * args.steps // (args.steps + args.warmup + PREWARM_STEPS)
                if args.impl == "b200" else None)
    clocks = stop_clock_sampler(sampler, spath) if rank == 0 else {}
    views = eff_world * V * args.steps
    value = views / (ms / 1000.0)

    # ---- end-to-end arm (host buffers) -------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        stager = HostStager(wl)
        ms_e = timed(runner, wl, args.steps, args.warmup, eff_world, e2e=True, stager=stager, prewarm=1)
        e2e = {"value": views / (ms_e / 1000.0), "unit": "views/s",
               "h2d_bytes_per_step": int(stager.bytes_per_view * V), "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e / args.steps,
               "inputs": "per view: 4 source images as uint8 + camera block from pinned host memory (converted to "
                         "float32 on the device); source-view depths are the rasterizer's own cached renders and stay "
                         "resident (train.py:299)"}
        del stager

    # ---- untimed passes: device time inside the rasterizer calls, per-stage times (b200) -----------------
    runner.timer.on = True
    run_steps(runner, wl, 1, 1)
    torch.cuda.synchronize(device)
    kernel_ms_per_view = runner.timer.total_ms() / V
    runner.timer.on = False
    runner.timer.pairs = []
    stages = {}
    if N is not None:
        N.lib.ibgs_profile_reset()
        N.lib.ibgs_profile_enable(1)
        run_steps(runner, wl, 2, 1)
        torch.cuda.synchronize(device)
        stages = N.profile_read()
        N.lib.ibgs_profile_enable(0)

    dpc = None
    if eff_world > 1 and not args.no_extras:
        try:
            dpc = dp_check(device, rank, eff_world)
        except Exception as ex:
            dpc = {"error": repr(ex)}

    train = {}
    if not args.no_train_step:
        cfgs = [c for c in args.train_configs.split(",") if c]
        if eff_world > 1:
            cfgs = cfgs[-1:]
        del runner
        runner = None
        gc.collect()
        torch.cuda.empty_cache()
        for c in cfgs:
            try:
                train[c] = train_step_through_render(args.impl, c, device, rank, eff_world, args.train_views)
            except Exception as ex:  # extra information only: never lose the headline line over it
                train[c] = {"error": repr(ex)}
        if args.impl == "b200" and cfgs and not args.no_extras:
            # the same step with this repo's section-8f fast paths switched in for the reference's Python glue
            for c in cfgs:
                try:
                    train[c + "_fast"] = train_step_through_render(args.impl, c, device, rank, eff_world, args.train_views,
                                                                   fast=True)
                except Exception as ex:
                    train[c + "_fast"] = {"error": repr(ex)}

    if rank != 0:
        if eff_world > 1:
            dist.destroy_process_group()
        return 0

    # ---- bookkeeping ---------------------------------------------------------------------------------------
    sc0 = wl.scene_for(wl.views[0])
    Npix = wl.W * wl.H
    if args.impl == "b200":
        dpr = impl.dpr
        dpr.KEEP_STATE = True
        with torch.no_grad():
            rs = wl.U.make_settings(dpr, sc0, render_geo=True)
            z = torch.zeros_like(sc0["means3D"])
            dpr.GaussianRasterizer(rs)(means3D=sc0["means3D"], means2D=z, means2D_abs=z, opacities=sc0["opacities"],
                                       shs=sc0["shs"], scales=sc0["scales"], rotations=sc0["rotations"],
                                       all_map=sc0["all_map"])
        dpr.KEEP_STATE = False
        st = dpr.LAST_STATE
        R = int(st["num_rendered"])
        dec = wl.U.decode_ours(st)
        pairs = int(dec["n_contrib"].long().sum().item())          # upper bound on blended pairs (last contributor idx)
        P_vis = int((dec["tiles_touched"] > 0).sum().item())
        del dec, st
        dpr.LAST_STATE.clear()
    else:
        fw = impl.ref.forward(sc0, render_geo=True)
        R = int(fw["num_rendered"])
        pairs = int(impl.ref.decode_image(fw["img"], Npix)["n_contrib"].long().sum().item())
        P_vis = int((fw["radii"] > 0).sum().item())
        del fw
    peaks = load_peaks()
    facts = load_profile_facts()
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    stage_ms = {k: (v[0] / v[1] if v[1] else None) for k, v in stages.items()}
    roofline = None
    roofline_issue = None
    hbm_stages = None
    if args.impl == "b200" and stage_ms.get("render_backward"):
        per_launch_ms = stage_ms["render_backward"]
        # algorithmic bytes of the backward tile renderer: point list + one 64 B record
        # read and one 64 B accumulator write per visible Gaussian + 212 B per pixel of cotangents / saved state
        # + the per-pixel median-pair lists (12 B written and read back per entry, <= buffer_length entries)
        alg = 4 * R + 128 * P_vis + 212 * Npix + 24 * 4 * Npix
        ach = alg / (per_launch_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "render_backward (dominant kernel of the step)", "achieved": ach,
                    "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                    "traffic": (facts.get("render_backward") or {}).get("dram_bytes_per_launch", facts.get("bytes_per_launch")),
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "ms_per_launch": per_launch_ms,
                    "note": "contract block: HBM is NOT what bounds this kernel -- it is bound by instruction issue "
                            "(see roofline_issue); the HBM-bound stages are listed in hbm_stages"}
        # the two pair loops against the issue ceiling: warp instructions per launch (committed ncu capture of the same
        # scene) / measured launch time / (148 SMs x 4 schedulers x SM clock)
        sm_mhz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
        ceiling = 148 * 4 * sm_mhz * 1e6
        roofline_issue = {"ceiling_warp_instr_per_s": ceiling, "sm_mhz": sm_mhz}
        for kname in ("render_backward", "render_forward"):
            wi = (facts.get(kname) or {}).get("warp_instructions_per_launch")
            if wi and stage_ms.get(kname):
                rate = wi / (stage_ms[kname] * 1e-3)
                roofline_issue[kname] = {"warp_instructions_per_launch": wi, "ms_per_launch": stage_ms[kname],
                                         "achieved_warp_instr_per_s": rate, "frac": rate / ceiling,
                                         "source": (facts.get(kname) or {}).get("source")}
        K = 9
        hb = {"preprocess": wl.P * (44 + 12 * K + 20) + P_vis * 73 + (wl.P - P_vis) * 8,
              "preprocess_backward": P_vis * (64 + 172 + 248 + 200), "duplicate_with_keys": 8 * R + 16 * wl.P,
              "radix_sort": 2 * 16 * R, "depth_order_sort": 4 * 16 * wl.P, "identify_tile_ranges": 4 * R}
        hbm_stages = {k: {"algorithmic_bytes": b, "ms": stage_ms[k], "GBps": b / (stage_ms[k] * 1e-3) / 1e9,
                          "frac_of_hbm_peak": b / (stage_ms[k] * 1e-3) / 1e9 / peak_gbs}
                      for k, b in hb.items() if stage_ms.get(k)}
    bytes_view = 730 * wl.P + 28 * R + 460 * Npix              # SURVEY.md section 8d whole-view figure
    ms_view = ms / (V * args.steps)
    arena_mb = 4 * sum(int(np.prod(wl.sc[k].shape)) for k in PARAM_KEYS) / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": args.gpus if args.impl == "b200" else 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {wl.P} Gaussians, {wl.W}x{wl.H}, render_geo, 4 src views, sh_degree 2, "
                               f"buffer_length 4", "views_per_step_per_gpu": V, "num_rendered_R": R,
                   "visible_gaussians": P_vis, "sum_n_contrib": pairs,
                   "l2": "per-view working set (192 MB records + 456 MB inputs + lists) exceeds the 126 MB L2; no flush",
                   "parallelism": f"dp{eff_world} over views, replicated Gaussians, 1 all-reduce of "
                                  f"{arena_mb:.0f} MB per step" if eff_world > 1 else "single GPU"},
        "ms_per_view": ms_view, "kernel_ms_per_view": kernel_ms_per_view,
        "view_roofline": {"algorithmic_bytes_per_view": bytes_view, "achieved_GBps": bytes_view / (ms_view * 1e-3) / 1e9,
                          "frac_of_hbm_peak": bytes_view / (ms_view * 1e-3) / 1e9 / peak_gbs},
        "pairs_per_s": pairs / (ms_view * 1e-3),
        "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roofline,
        "roofline_issue": roofline_issue, "hbm_stages": hbm_stages, "stages_ms_per_launch": stage_ms,
        "train_step": train or None,
    }
    if dpc is not None:
        line["dp_check"] = dpc
    if args.impl == "reference":
        line["impl"] = "reference"
        line["gpu_launches"] = None
        line["cpu_baseline"] = {"kind": "reference", "cores": 0, "value": value, "unit": "views/s",
                                "sample": "the reference has no CPU path: this arm is its unmodified CUDA extension "
                                          "(oracle/_ref) on the same GPU, full workload"}
        line["native_libraries"] = "oracle/_ref/dpr/ref_dpr_C.so only (libibgs_b200.so is never loaded in this arm)"
        assert "ibgs_b200._native" not in sys.modules, "the reference arm must not load libibgs_b200.so"
    elif args.gpus == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    if args.impl == "b200" and args.gpus == 1 and not args.no_extras:
        try:   # the stock reference API (no accumulate_grads): what the unchanged gaussian_renderer would call
            stock = OursRunner(wl, accumulate=False)
            ms_s = timed(stock, wl, max(2, args.steps // 3), 1, 1, prewarm=1)
            line["stock_api"] = {"ms_per_view": ms_s / (V * max(2, args.steps // 3)),
                                 "note": "same step without accumulate_grads: gradients returned to autograd, "
                                         "AccumulateGrad adds them into .grad"}
            del stock
        except Exception as ex:
            line["stock_api"] = {"error": repr(ex)}
        try:
            line["parity"] = parity_block(wl, OursRunner)
        except Exception as ex:
            line["parity"] = {"error": repr(ex)}
        torch.cuda.empty_cache()
        try:
            line["fast_path_step"] = fast_path_step_timing(wl)
        except Exception as ex:  # extra information only: never lose the headline line over it
            line["fast_path_step"] = {"error": repr(ex)}
        try:
            line["source_depth_batch"] = depth_batch_timing(wl)
        except Exception as ex:
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
            torch.cuda.empty_cache()
            line["parameter_prologue"] = prologue_timing(wl.P, device)
        except Exception as ex:
            line["parameter_prologue"] = {"error": repr(ex)}
    print(json.dumps(line))
    if eff_world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
