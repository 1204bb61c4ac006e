"""Optional fast path for the per-view glue around the rasterizer.

    from ibgs_b200.gaussian_renderer import render, render_depth      # instead of gaussian_renderer.render / render_depth

Same arguments, same result dict as the reference's gaussian_renderer/__init__.py:143-365 (`render`) and :41-140
(`render_depth`), on the reference's own objects (scene.cameras.Camera, scene.GaussianModel, scene.Scene, AppModel).
The unchanged reference functions keep working on top of `diff_plane_rasterization` (that is the drop-in boundary);
this module is what a caller switches to when it also wants the Python prologue gone:

  * activations + plane normal / distance (`all_map`) come from ONE fused launch per direction (ibgs_b200.fused: sigmoid /
    exp / normalize / normal flip / plane distance, learnt normal or shortest axis) instead of ~30 torch kernels, and the
    SH coefficients are read in place (`shs` = _features_dc, `shs_rest` = _features_rest) instead of through torch.cat;
    leaves that already hold a `.grad` (a view batch accumulating into an arena) get their gradient ADDED by the
    backward kernel itself (`accumulate_grads`) instead of through a temporary + autograd's add pass;
  * with do_render_src_depth the source-view depths are rendered by ONE batched depth-only pass
    (ibgs_b200.depth_batch.render_depth_views) instead of one rasterizer call per source view (:245-252);
  * the two torch.inverse calls per view on camera matrices (each a device-wide host synchronisation) are cached on the
    scene / camera objects for as long as the matrices are the same unmodified tensors;
  * the depth-to-normal map (utils/graphics_utils.py:38-75 through render_normal, :15-26) is one CUDA kernel per direction
    (csrc/depth_normal.cu) instead of ~12 + ~25 elementwise torch kernels.

Python-side SH / covariance evaluation (pipe.convert_SHs_python, pipe.compute_cov3D_python) is not covered:
NotImplementedError (those flags exist to bypass the CUDA path).  There is no CPU path.
"""
import math
import random
from typing import Optional

import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import fused
from .depth_batch import render_depth_views
from .diff_plane_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def _check_pipe(pipe):
    if pipe.convert_SHs_python or pipe.compute_cov3D_python:
        raise NotImplementedError("ibgs_b200.gaussian_renderer covers the CUDA SH / covariance path only "
                                  "(pipe.convert_SHs_python / pipe.compute_cov3D_python are False)")


def _threshold(args, depth_error_threshold):
    if depth_error_threshold is None:
        depth_error_threshold = getattr(args, "depth_error_threshold", 0.01)
    return float(depth_error_threshold)


def _prologue(pc, cam, learnt_normal, with_map):
    """(opacity, scales, rotations, all_map) of gaussian_renderer/__init__.py:168-195,304-315 from the raw parameters."""
    if with_map:
        nrm, off = (pc._normal, pc._offset) if learnt_normal else (None, None)
        return fused.gaussian_prologue(pc._xyz, pc._opacity, pc._scaling, pc._rotation, pc._features_dc, pc._features_rest,
                                       nrm, off, cam.world_view_transform, cam.camera_center, concat_sh=False,
                                       smallest_axis_normal=not learnt_normal)
    opacity, scales, rotations = fused.gaussian_prologue(pc._xyz, pc._opacity, pc._scaling, pc._rotation,
                                                         pc._features_dc, pc._features_rest, concat_sh=False)
    return opacity, scales, rotations, None


def _cached_inverse(owner, attr, t, transposed=False):
    """torch.inverse(t) (or of t.T), cached on `owner` for as long as `t` is the same unmodified tensor.  Camera poses are constants of
    a training run; torch.inverse on CUDA tensors synchronises the device with the host (its error check reads `info`
    back), and the reference evaluates two of them per render() call (gaussian_renderer/__init__.py:258-260)."""
    hit = getattr(owner, attr, None)
    if hit is not None and hit[0] is t and hit[1] == t._version:      # the very same tensor object, not written since
        return hit[2]
    inv = torch.inverse(t.T if transposed else t)
    try:
        setattr(owner, attr, (t, t._version, inv))     # holds `t`: its storage cannot be recycled under the cache
    except Exception:      # objects that refuse new attributes: no cache
        pass
    return inv


def _empty_sources(cam):
    n = int(cam.image_height) * int(cam.image_width)
    z = lambda *s: torch.zeros(s, device="cuda")
    return 1, z(1, 16), z(1, 3, n), z(1, 1, n), z(1, 3)


def render_depth(viewpoint_camera, pc, scene, pipe, args, bg_color, learnt_normal: bool, nb_src_frames: int,
                 buffer_length: int, depth_error_threshold: Optional[float] = None, scaling_modifier=1.0,
                 override_color=None):
    """Plane depth of one view, (1, H, W): gaussian_renderer/__init__.py:41-140."""
    _check_pipe(pipe)
    return render_depth_views([viewpoint_camera], pc, scene, pipe, args, bg_color, learnt_normal, nb_src_frames,
                              buffer_length, depth_error_threshold, scaling_modifier, override_color)[0]


def _closest_frames(viewpoint_camera, scene, args):
    """Source-view choice at test time, gaussian_renderer/__init__.py:198-226: sort the train views by (distance, angle)
    to this camera, keep those inside the angle / distance window, at most multi_view_num; with exposure correction the
    view with the most similar pose goes first."""
    camera_center = viewpoint_camera.camera_center
    R = torch.as_tensor(viewpoint_camera.R).float()
    center_ray = torch.tensor([0.0, 0.0, 1.0], device="cuda") @ R.to("cuda").transpose(-1, -2)
    dist = torch.norm(camera_center.unsqueeze(0) - scene.camera_centers, dim=-1).detach().cpu().numpy()
    ang = (torch.arccos(torch.sum(center_ray.unsqueeze(0) * scene.center_rays, dim=-1)) * 180 / torch.pi).detach().cpu().numpy()
    order = np.lexsort((ang, dist))
    keep = (ang[order] < args.multi_view_max_angle) & (dist[order] > args.multi_view_min_dis) & (dist[order] < args.multi_view_max_dis)
    order = order[keep][:min(args.multi_view_num, int(keep.sum()))].tolist()
    if args.enable_exposure_correction:
        rel = torch.matmul(viewpoint_camera.world_view_transform.T.unsqueeze(0), torch.inverse(scene.world_view_transforms))
        diff = torch.mean(torch.abs(rel - torch.eye(4, device="cuda").unsqueeze(0)), dim=[1, 2]).detach().cpu().numpy()
        best = order[int(np.argmin(diff[order]))]
        order.remove(best)
        order = [best] + order
    return np.array(order)


class _DepthNormal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, fx, fy, cx, cy):
        if not depth.is_cuda:
            raise RuntimeError("ibgs_b200.gaussian_renderer: tensors must be CUDA tensors (there is no CPU path)")
        d = depth.detach().float().contiguous()
        H, W = d.shape
        out = torch.empty((3, H, W), dtype=torch.float32, device=d.device)
        with torch.cuda.device(d.device):
            N.check(N.lib.ibgs_depth_normal_forward(d.data_ptr(), out.data_ptr(), H, W, fx, fy, cx, cy,
                                                    C.c_void_p(torch.cuda.current_stream(d.device).cuda_stream)),
                    "ibgs_depth_normal_forward")
        ctx.save_for_backward(d)
        ctx.k = (fx, fy, cx, cy)
        return out

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        H, W = d.shape
        g = g.detach().float().contiguous()
        gd = torch.empty_like(d)
        with torch.cuda.device(d.device):
            N.check(N.lib.ibgs_depth_normal_backward(d.data_ptr(), g.data_ptr(), gd.data_ptr(), H, W, *ctx.k,
                                                     C.c_void_p(torch.cuda.current_stream(d.device).cuda_stream)),
                    "ibgs_depth_normal_backward")
        return gd, None, None, None, None


def depth_normal(viewpoint_cam, depth):
    """Unit normal map (3, H, W) of a depth image (H, W): render_normal (gaussian_renderer/__init__.py:15-26) followed by
    the renormalisation of :332-335, one CUDA kernel per direction (csrc/depth_normal.cu).  Points are back-projected with
    the pinhole intrinsics of Camera.get_calib_matrix_nerf (scene/cameras.py:115-118); the normal is the cross product of
    the horizontal and vertical central differences (utils/graphics_utils.py:65-72), zero on the 1-pixel border."""
    return _DepthNormal.apply(depth, float(viewpoint_cam.Fx), float(viewpoint_cam.Fy), float(viewpoint_cam.Cx),
                              float(viewpoint_cam.Cy))


def depth_normal_torch(viewpoint_cam, depth):
    """The same map written with torch ops (what the kernel is tested against; not used by render())."""
    H, W = depth.shape
    fx, fy, cx, cy = float(viewpoint_cam.Fx), float(viewpoint_cam.Fy), float(viewpoint_cam.Cx), float(viewpoint_cam.Cy)
    xs = (torch.arange(W, device=depth.device, dtype=depth.dtype) - cx) / fx
    ys = (torch.arange(H, device=depth.device, dtype=depth.dtype) - cy) / fy
    pts = torch.stack((depth * xs[None, :], depth * ys[:, None], depth), 0)                                  # (3, H, W)
    l2r = pts[:, 1:H - 1, 2:W] - pts[:, 1:H - 1, 0:W - 2]
    b2t = pts[:, 0:H - 2, 1:W - 1] - pts[:, 2:H, 1:W - 1]
    n = torch.nn.functional.normalize(torch.cross(l2r, b2t, dim=0), p=2, dim=0)
    n = torch.nn.functional.pad(n, (1, 1, 1, 1), mode="constant")
    return n / (torch.norm(n, dim=0, keepdim=True) + 1e-8)


def render(viewpoint_camera, pc, scene, pipe, args, bg_color, learnt_normal: bool, nb_src_frames: int, buffer_length: int,
           depth_error_threshold: Optional[float] = None, scaling_modifier=1.0, override_color=None, app_model=None,
           render_geo=True, return_depth_normal=True, do_find_closest_frame=False, do_render_src_depth=False,
           render_depth_only=False):
    """Render one view: gaussian_renderer/__init__.py:143-365 (same result dict)."""
    _check_pipe(pipe)
    xyz = pc._xyz
    # the two dummy leaves whose .grad the densification reads (:153-159, train.py:401-404)
    screenspace_points = torch.zeros_like(xyz, requires_grad=True) + 0
    screenspace_points_abs = torch.zeros_like(xyz, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
        screenspace_points_abs.retain_grad()
    except Exception:
        pass
    depth_error_threshold = _threshold(args, depth_error_threshold)
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)

    with_map = bool(render_geo or render_depth_only)
    opacity, scales, rotations, input_all_map = _prologue(pc, viewpoint_camera, learnt_normal, with_map)

    if render_geo:
        nearest = _closest_frames(viewpoint_camera, scene, args) if do_find_closest_frame else viewpoint_camera.nearest_id
        if len(nearest) == 0:
            nb_src_frames, ref_to_src_list, src_images, src_rendered_depths, src_cam_pos = _empty_sources(viewpoint_camera)
        else:
            nb_src_frames = min(nb_src_frames, len(nearest))
            if args.shuffle_source_frame:
                selected = random.sample(list(nearest), nb_src_frames)
            else:
                selected = nearest[:nb_src_frames]
            src_images = scene.original_image_list[selected]
            if do_render_src_depth:      # :245-252, one batched depth-only pass instead of one render per source view
                src_views = [scene.getTrainCameras()[i] for i in selected]
                src_rendered_depths = render_depth_views(src_views, pc, scene, pipe, args, bg_color, learnt_normal,
                                                         nb_src_frames, buffer_length, depth_error_threshold,
                                                         scaling_modifier, override_color)
            else:
                src_rendered_depths = scene.rendered_depth_list[selected]
            world_to_src = scene.world_view_transforms[selected]
            src_to_world = _cached_inverse(scene, "_ibgs_b200_view_to_world", scene.world_view_transforms)[selected]
            ref_to_world = _cached_inverse(viewpoint_camera, "_ibgs_b200_cam_to_world", viewpoint_camera.world_view_transform,
                                           transposed=True)
            ref_to_src_list = (world_to_src @ ref_to_world.unsqueeze(0)).cuda()
            src_cam_pos = src_to_world[:, :3, 3].cuda()
            src_rendered_depths = src_rendered_depths.cuda()
            src_images = src_images.cuda()
    else:
        nb_src_frames, ref_to_src_list, src_images, src_rendered_depths, src_cam_pos = _empty_sources(viewpoint_camera)

    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx,
        tanfovy=tanfovy,
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        ref_to_src_list=ref_to_src_list,
        src_cam_pos=src_cam_pos,
        src_images=src_images,
        src_rendered_depths=src_rendered_depths,
        nb_src_images=nb_src_frames,
        buffer_length=buffer_length,
        depth_error_threshold=depth_error_threshold,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        render_geo=render_geo,
        render_depth_only=render_depth_only,
        debug=pipe.debug)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    if override_color is None:
        sh_kw = dict(shs=pc._features_dc, shs_rest=pc._features_rest, colors_precomp=None)
    else:
        sh_kw = dict(shs=None, colors_precomp=override_color)
    (rendered_image, radii, out_normal_map, out_median_intersected_depth, out_cam_feat, out_warped_image,
     out_min_depth_diff, out_camera_ray, use_first_src_frame_mask) = rasterizer(
        means3D=xyz, means2D=screenspace_points, means2D_abs=screenspace_points_abs, opacities=opacity, scales=scales,
        rotations=rotations, all_map=input_all_map, cov3D_precomp=None, accumulate_grads=True, **sh_kw)

    rendered_normal = out_normal_map[0:3] if render_geo else None
    if return_depth_normal:
        median_intersected_depth_normal = depth_normal(viewpoint_camera, out_median_intersected_depth.squeeze())
    else:
        median_intersected_depth_normal = None

    if app_model is not None and pc.use_app:
        appear_ab = app_model.appear_ab[torch.tensor(viewpoint_camera.uid, device="cuda")]
        app_image = torch.exp(appear_ab[0]) * rendered_image + appear_ab[1]
    else:
        app_image = None

    return {"render": rendered_image,
            "app_image": app_image,
            "viewspace_points": screenspace_points,
            "viewspace_points_abs": screenspace_points_abs,
            "visibility_filter": radii > 0,
            "radii": radii,
            "rendered_normal": rendered_normal,
            "median_intersected_depth": out_median_intersected_depth,
            "median_intersected_depth_normal": median_intersected_depth_normal,
            "cam_feat": out_cam_feat,
            "warped_image": out_warped_image,
            "min_depth_diff": out_min_depth_diff,
            "camera_ray": out_camera_ray,
            "use_first_src_frame_mask": use_first_src_frame_mask}
