"""Harness that runs the reference's UNCHANGED Python glue on synthetic scenes (test / bench infrastructure).

The files staged verbatim by oracle/stage_ref_py.py (gaussian_renderer.render / render_depth, scene.GaussianModel,
scene.cameras.Camera, scene.Scene's train-buffer code, scene.app_model.AppModel, utils.loss_utils,
color_aggregation_network.fuse_color / ColorFusionResidualNet, arguments.*Params) are imported as they are; only the
two extension import names are chosen:

    bind("b200")       `diff_plane_rasterization`, `simple_knn._C`  -> this repo (shims/ -> ibgs_b200 -> C ABI -> sm_100a)
    bind("reference")  the reference's own autograd wrapper + oracle/_ref/*.so (its unmodified CUDA code); this binding
                       never imports ibgs_b200._native, so libibgs_b200.so is not mapped in a reference-arm process

Both bindings can live in one process (each gets its own `gaussian_renderer` module object; `scene`, `utils`, ... are
shared, they do not depend on the binding except for `distCUDA2`, which is bound by whichever comes first).

`build_world` replaces what needs a COLMAP dataset on disk (SURVEY.md section 8d "Harness duck-typing"): Gaussians of
ibgs_b200.synthetic.make_scene loaded into a real GaussianModel, real Camera objects on perturbed poses with synthetic
images, and a Scene whose buffers / nearest_id lists are produced by the reference's own Scene methods.
`train_iteration` restates the body of train.py:269-370 (render -> losses -> fuse_color -> backward) and
`optimizer_step` train.py:421-430, calling the reference's functions; the same code drives both bindings.
"""
import importlib
import importlib.util
import math
import os
import sys
import tempfile
import types
from argparse import ArgumentParser

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGE = os.path.join(ROOT, "baseline", "_ref")
PY_ROOT = os.path.join(STAGE, "py")
REF_WRAPPER = os.path.join(STAGE, "py_ref_dpr", "diff_plane_rasterization", "__init__.py")
STUBS = os.path.join(HERE, "stubs")

_GLUE = {}


def available():
    return os.path.exists(os.path.join(PY_ROOT, "gaussian_renderer", "__init__.py"))


def reference_available():
    return (available() and os.path.exists(REF_WRAPPER)
            and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "dpr", "ref_dpr_C.so")))


def _paths():
    for p in (STUBS, PY_ROOT, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)


def _load_reference_binding():
    """The reference's own diff_plane_rasterization/__init__.py (staged verbatim) over its own extension.  Its
    `from . import _C` is satisfied by registering the prebuilt module under `<package>._C` first."""
    from oracle import ref_ext
    C = ref_ext.load("dpr")
    name = "ref_diff_plane_rasterization"
    if name in sys.modules:
        return sys.modules[name]
    sys.modules[name + "._C"] = C
    spec = importlib.util.spec_from_file_location(name, REF_WRAPPER,
                                                  submodule_search_locations=[os.path.dirname(REF_WRAPPER)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _reference_knn():
    from oracle import ref_ext
    pkg = types.ModuleType("simple_knn")
    pkg.__path__ = []
    c = types.ModuleType("simple_knn._C")
    if ref_ext.available("knn"):
        c.distCUDA2 = ref_ext.load("knn").distCUDA2
    else:
        def distCUDA2(points):
            raise RuntimeError("oracle/_ref/knn/ref_knn_C.so is missing")
        c.distCUDA2 = distCUDA2
    pkg._C = c
    return pkg, c


def bind(binding):
    """Import the staged glue with the extension names bound to `binding`; returns a namespace of the reference's
    own callables (render, render_depth, GaussianModel, Camera, Scene, AppModel, fuse_color, ...)."""
    if binding in _GLUE:
        return _GLUE[binding]
    if not available():
        raise FileNotFoundError(f"{PY_ROOT} is not staged: run `python oracle/stage_ref_py.py` where /root/reference exists")
    _paths()
    if binding == "b200":
        shims = os.path.join(ROOT, "shims")
        saved = {k: sys.modules.pop(k, None) for k in ("diff_plane_rasterization", "simple_knn", "simple_knn._C")}
        sys.path.insert(0, shims)
        try:
            dpr = importlib.import_module("diff_plane_rasterization")      # shims/ -> ibgs_b200
            knn = importlib.import_module("simple_knn")
            knn_c = importlib.import_module("simple_knn._C")
        finally:
            sys.path.remove(shims)
        del saved
    elif binding == "reference":
        dpr = _load_reference_binding()
        knn, knn_c = _reference_knn()
    else:
        raise ValueError(binding)
    sys.modules["diff_plane_rasterization"] = dpr
    if "scene.gaussian_model" not in sys.modules:          # distCUDA2 is bound at the first import of scene.*
        sys.modules["simple_knn"], sys.modules["simple_knn._C"] = knn, knn_c
    # a private copy of gaussian_renderer per binding: it binds the rasterizer names at import time (:5-6)
    name = f"gaussian_renderer__{binding}"
    path = os.path.join(PY_ROOT, "gaussian_renderer", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path)
    gr = importlib.util.module_from_spec(spec)
    sys.modules[name] = gr
    spec.loader.exec_module(gr)
    sys.modules.pop("diff_plane_rasterization", None)

    import scene as scene_mod
    from scene.gaussian_model import GaussianModel
    from scene.cameras import Camera
    from scene.app_model import AppModel
    import utils.loss_utils as loss_utils
    import utils.general_utils as general_utils
    import color_aggregation_network as can
    import arguments as arguments_mod
    g = types.SimpleNamespace(
        binding=binding, dpr=dpr, distCUDA2=knn_c.distCUDA2, gaussian_renderer=gr, render=gr.render,
        render_depth=gr.render_depth, Scene=scene_mod.Scene, GaussianModel=GaussianModel, Camera=Camera,
        AppModel=AppModel, loss_utils=loss_utils, general_utils=general_utils, fuse_color=can.fuse_color,
        ColorFusionResidualNet=can.ColorFusionResidualNet, arguments=arguments_mod)
    _GLUE[binding] = g
    return g


# ----------------------------------------------------------------------------------------------------------------------
def default_args(glue, **overrides):
    """The reference's own defaults (arguments/__init__.py:57-138) through its own ParamGroup -> argparse reflection."""
    A = glue.arguments
    parser = ArgumentParser()
    lp, op, pp = A.ModelParams(parser), A.OptimizationParams(parser), A.PipelineParams(parser)
    args = parser.parse_args([])
    for k, v in overrides.items():
        if not hasattr(args, k):
            raise AttributeError(k)
        setattr(args, k, v)
    args.model_path = tempfile.mkdtemp(prefix="ibgs_refglue_")
    return args, lp.extract(args), op.extract(args), pp.extract(args)


def _smooth_image(rng, H, W):
    low = torch.from_numpy(rng.random((1, 3, (H + 7) // 8 + 1, (W + 7) // 8 + 1)).astype(np.float32))
    return torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=True)[0].contiguous()


def view_poses(sc_cpu, n_views, seed=1000, first=0):
    """World-to-camera matrices of the train views: view 0 is the scene's reference pose, the others are small rigid
    perturbations of it (<= 2 degrees, <= 0.1 units: every view sees the same Gaussians and passes the reference's
    neighbour filter, scene/__init__.py:215-222)."""
    from ibgs_b200 import synthetic as S
    w2c = sc_cpu["w2c"].double().numpy()
    out = []
    for gid in range(first, first + n_views):
        rng = np.random.default_rng(seed + gid)
        D = S._rigid(S._rot_axis_angle(rng.normal(size=3), np.radians(rng.uniform(0.3, 2.0))),
                     rng.uniform(-0.1, 0.1, 3)) if gid > 0 else np.eye(4)
        out.append(D @ w2c)
    return out


class World:
    pass


def build_world(glue, config="cfg1", n_views=6, device="cuda", P=None, seed=1000, first_view=0, sc_cpu=None,
                exposure=False, learnt_normal=True, **arg_overrides):
    """Synthetic stand-in for `Scene(dataset, gaussians, args)` + `gaussians.training_setup(opt)` + AppModel +
    ColorFusionResidualNet (train.py:181-231), built with the reference's own classes."""
    from ibgs_b200 import synthetic as S
    if exposure:   # BASELINE config 4: exposure compensation (AppModel affine) + exposure correction (lstsq in fuse_color)
        arg_overrides.setdefault("exposure_compensation", True)
        arg_overrides.setdefault("enable_exposure_correction", True)
    arg_overrides.setdefault("learnt_normal", learnt_normal)
    args, dataset, opt, pipe = default_args(glue, **arg_overrides)
    sc_cpu = sc_cpu if sc_cpu is not None else S.make_scene(config, P=P)
    H, W, Pn = sc_cpu["H"], sc_cpu["W"], sc_cpu["P"]
    w = World()
    w.glue, w.args, w.dataset, w.opt, w.pipe, w.sc_cpu = glue, args, dataset, opt, pipe, sc_cpu
    w.H, w.W, w.P, w.device = H, W, Pn, device

    # ---- cameras: the reference's Camera class (scene/cameras.py:51-118), images attached instead of loaded -----
    fovx = 2 * math.atan(sc_cpu["tanfovx"])
    fovy = 2 * math.atan(sc_cpu["tanfovy"])
    cams = []
    rng = np.random.default_rng(seed + 77)
    for i, w2c in enumerate(view_poses(sc_cpu, n_views, seed, first_view)):
        R = np.ascontiguousarray(w2c[:3, :3].T)          # Camera stores the transposed rotation (getWorld2View2)
        T = np.ascontiguousarray(w2c[:3, 3])
        cam = glue.Camera(colmap_id=i, R=R, T=T, FoVx=fovx, FoVy=fovy, image_width=W, image_height=H,
                          image_path="", image_name=f"synthetic_{first_view + i:04d}", uid=i, preload_img=False,
                          data_device=args.data_device)
        cam.original_image = _smooth_image(rng, H, W).to(device)
        cams.append(cam)

    # ---- scene: the reference's own buffer / neighbour code on a Scene that skips the dataset readers ------------
    class SynthScene(glue.Scene):
        def __init__(self, cams, args):
            self.model_path = args.model_path
            self.loaded_iter = None
            self.gaussians = None
            self.cameras_extent = 5.0
            self.multi_view_num = args.multi_view_num
            self.train_cameras = {1.0: cams}
            self.test_cameras = {1.0: []}
            self._initialize_train_buffers(cams, args)                       # scene/__init__.py:113-141
            self._write_train_multiview(cams, self._compute_train_metrics(), args, args)   # :143-176 -> nearest_id

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        w.scene = SynthScene(cams, args)
    w.cams = cams

    # ---- Gaussians: a real GaussianModel holding the synthetic scene (what load_ply does, gaussian_model.py:312-360)
    gm = glue.GaussianModel(dataset.sh_degree)
    dev = torch.device(device)
    K = (dataset.sh_degree + 1) ** 2
    op = sc_cpu["opacities"].clamp(1e-4, 1 - 1e-4)
    par = lambda t: torch.nn.Parameter(t.to(dev).contiguous().requires_grad_(True))
    gm._xyz = par(sc_cpu["means3D"])
    gm._features_dc = par(sc_cpu["shs"][:, :1, :])
    gm._features_rest = par(sc_cpu["shs"][:, 1:K, :])
    gm._opacity = par(torch.log(op / (1 - op)))
    gm._scaling = par(torch.log(sc_cpu["scales"]))
    gm._rotation = par(sc_cpu["rotations"])
    gm._normal = par(sc_cpu["normals_world"])
    gm._offset = par(torch.zeros((Pn, 1)))
    gm.max_radii2D = torch.zeros((Pn,), device=dev)
    gm.max_weight = torch.zeros((Pn,), device=dev)
    gm.active_sh_degree = dataset.sh_degree
    gm.spatial_lr_scale = w.scene.cameras_extent
    gm.training_setup(opt)
    w.gaussians = gm
    w.scene.gaussians = gm

    w.app_model = glue.AppModel()
    w.app_model.train()
    w.app_model.cuda()
    w.background = torch.tensor([0, 0, 0], dtype=torch.float32, device=dev)
    w.color_net = None
    w.color_opt = None
    if opt.use_color_aggregation:
        g = torch.Generator().manual_seed(seed + 5)
        state = torch.random.get_rng_state()
        torch.manual_seed(seed + 5)
        w.color_net = glue.ColorFusionResidualNet(height=int(H * opt.residual_resolution_scale),
                                                  width=int(W * opt.residual_resolution_scale),
                                                  feat_aggregate_mode=opt.feat_aggregate_mode).cuda()
        torch.random.set_rng_state(state)
        del g
        w.color_opt = torch.optim.Adam(w.color_net.parameters(), lr=0.001)
    # steady state of the schedule: render_geo on, normal + multi-view losses on, colour aggregation on and burnt in
    w.iteration = 20000
    w.color_iter_count = w.iteration
    w.color_burn_start = w.iteration - 2 * opt.color_aggregate_burnin_steps
    w.color_burn_end = w.color_burn_start + opt.color_aggregate_burnin_steps
    return w


def render_kwargs(w):
    o = w.opt
    return dict(learnt_normal=o.learnt_normal, nb_src_frames=o.number_src_frames, buffer_length=o.buffer_length,
                depth_error_threshold=o.depth_error_threshold, app_model=w.app_model)


def prime_depth_cache(w):
    """train.py:242-256 -- fill scene.rendered_depth_list with every train view's median depth before training."""
    g = w.glue
    if w.iteration > 1000 and w.opt.exposure_compensation:
        w.gaussians.use_app = True
    with torch.no_grad():
        for idx, cam in enumerate(w.scene.getTrainCameras()):
            pkg = g.render(cam, w.gaussians, w.scene, w.pipe, w.args, w.background, render_geo=True,
                           return_depth_normal=True, **render_kwargs(w))
            w.scene.rendered_depth_list[idx] = pkg["median_intersected_depth"].detach().to(device=w.args.data_device)


def train_iteration(w, cam_idx, loss_scale=1.0, fns=None):
    """Body of the training loop the reference times as `iter_time` (train.py:269-370): render -> image loss ->
    single-view normal loss -> multi-view photometric loss -> colour-aggregation prediction + loss -> backward.
    `fns` may override ssim / compute_photometric_ssim / l1_loss (the fused-SSIM fast path) and render / fuse_color
    (ibgs_b200.gaussian_renderer / ibgs_b200.color_aggregation); default = the reference's."""
    g, opt = w.glue, w.opt
    LU = g.loss_utils
    ssim = getattr(fns, "ssim", LU.ssim)
    l1_loss = getattr(fns, "l1_loss", LU.l1_loss)
    photo_ssim = getattr(fns, "compute_photometric_ssim", LU.compute_photometric_ssim)
    render = getattr(fns, "render", g.render)
    fuse_color = getattr(fns, "fuse_color", g.fuse_color)
    iteration = w.iteration
    gaussians = w.gaussians
    cams = w.scene.getTrainCameras()
    viewpoint_cam = cams[cam_idx]
    gt_image = viewpoint_cam.original_image.cuda()
    if iteration > 1000 and opt.exposure_compensation:
        gaussians.use_app = True
    geo = iteration > opt.single_view_weight_from_iter - len(cams) * 2
    render_pkg = render(viewpoint_cam, gaussians, w.scene, w.pipe, w.args, w.background, render_geo=geo,
                        return_depth_normal=geo, **render_kwargs(w))
    image = render_pkg["render"]
    if geo:
        w.scene.rendered_depth_list[cam_idx] = render_pkg["median_intersected_depth"].detach().to(device=w.args.data_device)

    ssim_loss = (1.0 - ssim(image, gt_image))
    if render_pkg["app_image"] is not None and ssim_loss < 0.5:
        Ll1 = l1_loss(render_pkg["app_image"], gt_image)
    else:
        Ll1 = l1_loss(image, gt_image)
    image_loss = (1.0 - opt.lambda_dssim) * Ll1 + opt.lambda_dssim * ssim_loss

    normal_loss = torch.tensor(0.0).float().cuda()
    if iteration > opt.single_view_weight_from_iter:
        weight = opt.single_view_weight
        normal = render_pkg["rendered_normal"]
        depth_normal = render_pkg["median_intersected_depth_normal"]
        normal_loss1 = weight * (((depth_normal - normal)).abs().sum(0)).mean()
        normal_loss2 = weight * (1 - ((depth_normal * normal).sum(0))).mean()
        normal_loss = (0.4 * normal_loss1 + 0.6 * normal_loss2)

    photometric_loss = torch.tensor(0.0).float().cuda()
    if iteration > opt.multi_view_weight_from_iter:
        Hc, Wc = viewpoint_cam.image_height, viewpoint_cam.image_width
        warped_image = render_pkg["warped_image"].view(-1, 3, Hc, Wc)[:opt.nb_visible_src_frames]
        cam_feat = render_pkg["cam_feat"].view(-1, 4, Hc, Wc)[:opt.nb_visible_src_frames]
        valid_mask = torch.sum(cam_feat, dim=1, keepdim=True) > 0
        ref_image = gt_image.unsqueeze(0)
        masked_warped_image = valid_mask.float() * warped_image + (1 - valid_mask.float()) * ref_image
        if torch.sum(valid_mask) > 0:
            p_ssim = 1 - torch.stack([photo_ssim(ref_image[0], masked_warped_image[i], size_average=False).mean(0)
                                      for i in range(len(masked_warped_image))])
            p_ssim = torch.sum(p_ssim * valid_mask[:, 0]) / torch.sum(valid_mask[:, 0])
            p_l1 = torch.abs(ref_image - masked_warped_image).mean(1)
            p_l1 = torch.sum(p_l1 * valid_mask[:, 0]) / torch.sum(valid_mask[:, 0])
            photometric_loss = ((1 - opt.photo_ssim_weight) * p_l1 + opt.photo_ssim_weight * p_ssim)
            photometric_loss = (photometric_loss * opt.photo_weight).mean()

    aggregate_image_loss = torch.tensor(0.0).float().cuda()
    fusion = None
    if opt.use_color_aggregation and iteration > opt.start_color_aggregation_iter:
        fusion = fuse_color(render_pkg, color_aggregation_network=w.color_net, iter_count=w.color_iter_count,
                            burn_start=w.color_burn_start, burn_end=w.color_burn_end, iteration=iteration, opts=opt)
        if fusion is not None:
            image_pred = fusion["image_pred"]
            aggregate_image_loss = ((1.0 - opt.lambda_dssim) * l1_loss(image_pred, gt_image)
                                    + opt.lambda_dssim * (1.0 - ssim(image_pred, gt_image)))
    loss = normal_loss + photometric_loss
    if fusion is None:
        loss = loss + image_loss
    else:
        loss = loss + (image_loss + aggregate_image_loss) / 2
    if loss_scale != 1.0:
        loss = loss * loss_scale
    loss.backward()
    return dict(loss=loss.detach(), render_pkg=render_pkg, fusion=fusion, image_loss=image_loss.detach(),
                normal_loss=normal_loss.detach(), photometric_loss=photometric_loss.detach())


def densification_stats(w, out):
    """train.py:399-405 (per view, under no_grad)."""
    with torch.no_grad():
        pkg = out["render_pkg"]
        mask, radii = pkg["visibility_filter"], pkg["radii"]
        gm = w.gaussians
        gm.max_radii2D[mask] = torch.max(gm.max_radii2D[mask], radii[mask])
        gm.add_densification_stats(pkg["viewspace_points"], pkg["viewspace_points_abs"], mask)


def optimizer_step(w):
    """train.py:421-430."""
    w.gaussians.optimizer.step()
    w.app_model.optimizer.step()
    w.gaussians.optimizer.zero_grad(set_to_none=True)
    w.app_model.optimizer.zero_grad(set_to_none=True)
    if w.opt.use_color_aggregation and w.iteration > w.opt.start_color_aggregation_iter:
        w.color_opt.step()
        w.color_opt.zero_grad()
        w.color_iter_count += 1


def fast_fns(precision="bf16", losses=True, renderer=True, color=True):
    """The section-8f fast paths of ibgs_b200 as `fns` for train_iteration: fused SSIM losses, the fused-prologue render()
    and the fused colour-aggregation step."""
    import functools
    f = types.SimpleNamespace()
    if losses:
        import ibgs_b200.loss_utils as FL
        f.ssim, f.compute_photometric_ssim = FL.ssim, FL.compute_photometric_ssim
    if renderer:
        import ibgs_b200.gaussian_renderer as FR
        f.render = FR.render
    if renderer:
        import ibgs_b200.densify as FD
        f.add_densification_stats = FD.add_densification_stats
    if color:
        import ibgs_b200.color_aggregation as CA
        f.fuse_color = functools.partial(CA.fuse_color, precision=precision)
    return f


GAUSSIAN_PARAMS = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_normal", "_offset")


def gaussian_grads(w):
    return {n: getattr(w.gaussians, n).grad for n in GAUSSIAN_PARAMS}


# ----------------------------------------------------------------------------------------------------------------------
def make_data_parallel(w):
    """View-sharded data parallelism around the unchanged GaussianModel (ibgs_b200.parallel)."""
    from ibgs_b200.parallel import GaussianDataParallel
    w.dp = GaussianDataParallel(w.gaussians, extra_modules=(w.app_model, w.color_net), average=True)
    return w.dp


def dp_train_step(w, cam_indices, views_total, fns=None, sync_stats=False):
    """One data-parallel optimisation step: this rank runs train_iteration (train.py:269-370) for its views of the batch,
    gradients accumulate in the flat arenas, ONE all-reduce (+ one small bucket) averages them over the `views_total`
    views of all ranks, the depth-cache entries the views produced are exchanged, then every rank takes the identical
    optimizer step (train.py:421-430 with the arena-preserving zero_grad)."""
    dp = w.dp
    outs = []
    for ci in cam_indices:
        out = train_iteration(w, ci, fns=fns)
        if w.iteration < w.opt.densify_until_iter or sync_stats:     # train.py:399: only while densification is on
            if hasattr(fns, "add_densification_stats"):
                with torch.no_grad():
                    fns.add_densification_stats(w.gaussians, out["render_pkg"])
            else:
                densification_stats(w, out)
        outs.append(out)
    dp.all_reduce_grads(views_total=views_total)
    dp.sync_depth_cache(w.scene.rendered_depth_list, cam_indices)
    if sync_stats:
        dp.sync_densification_stats()
    w.gaussians.optimizer.step()
    w.app_model.optimizer.step()
    if w.opt.use_color_aggregation and w.iteration > w.opt.start_color_aggregation_iter:
        w.color_opt.step()
        w.color_iter_count += 1
    dp.zero_grad()
    return outs
