"""Drop-in replacement for the reference package `diff_plane_rasterization`
(submodules/diff-plane-rasterization/diff_plane_rasterization/__init__.py).

Same public names, argument order, return order and error behaviour:
  GaussianRasterizationSettings  (reference __init__.py:252-276)
  GaussianRasterizer             (reference __init__.py:278-331)
  rasterize_gaussians / _RasterizeGaussians  (reference __init__.py:21-250)

The reference calls a pybind11 module `_C` that owns torch tensors; here Python owns every tensor and
hands raw device pointers to the C ABI in include/ibgs_b200.h through ctypes.  There is no CPU or
PyTorch fallback: without libibgs_b200.so the import fails.
"""
from typing import NamedTuple

import ctypes as C
import os
import torch
import torch.nn as nn

from .. import _native as N

# tests flip this to keep the forward scratch (unsorted / sorted keys) alive for bit-exact comparison
# IBGS_POISON_OUTPUTS=1 (debugging aid): every tensor the kernels are trusted to write COMPLETELY -- outputs, gradients, state
# and scratch buffers, all allocated with torch.empty -- is pre-filled with NaN / 0xFF bytes, so a word a kernel fails to
# write shows up in any comparison instead of depending on what the caching allocator handed out.
POISON = os.environ.get("IBGS_POISON_OUTPUTS", "") not in ("", "0")


def _poisoned_empty(shape, **kw):
    t = torch.empty(shape, **kw)
    if POISON and t.numel():
        t.view(torch.uint8).fill_(0xFF) if t.dtype != torch.float32 else t.fill_(float("nan"))
    return t


KEEP_STATE = False
LAST_STATE = {}

_M = N.MAX_SRC


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def rasterize_gaussians(means3D, means2D, means2D_abs, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, all_map, raster_settings, sh_rest=None, accumulate_grads=False):
    return _RasterizeGaussians.apply(means3D, means2D, means2D_abs, sh, colors_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, all_map, raster_settings, sh_rest, accumulate_grads)


def _accumulable(t):
    """A leaf whose .grad already exists as a dense float32 tensor of its own shape: autograd would run
    `t.grad += returned gradient` as a separate pass; with accumulate_grads the kernel adds into t.grad itself."""
    if not (isinstance(t, torch.Tensor) and t.is_leaf and t.requires_grad):
        return False            # (.grad of a non-leaf must not even be looked at: torch warns)
    g = t.grad
    return (g is not None and g.is_cuda and g.dtype == torch.float32 and g.is_contiguous()
            and tuple(g.shape) == tuple(t.shape) and g.data_ptr() % 16 == 0)   # (float4 stores into quaternion rows)


def _ptr(t):
    """Device address of a tensor, or NULL for an absent (empty) one -- the reference's convention
    (rasterizer_impl.cu:470,595,643; forward.cu:244,280)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t, device):
    if t is None:
        return None
    if t.numel() == 0:
        return t
    if t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        # a view at an odd offset of a larger buffer (e.g. one group of a packed parameter arena): the kernels read
        # quaternion rows as float4, so hand them a fresh (512-byte aligned) copy
        t = t.clone()
    return t


class _Allocator:
    """The C side asks for its state / scratch buffers through this callback -- the ctypes spelling of
    the reference's resizeFunctional lambdas (rasterize_points.cu:29-35)."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        self.scratch = []
        self.error = None
        self.fn = N.ALLOC_FN(self._alloc)

    def _alloc(self, user, which, nbytes):
        try:
            t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=self.device)
            if POISON:
                t.fill_(0xFF)
            if which == N.IBGS_BUF_SCRATCH:
                self.scratch.append(t)
            else:
                self.bufs[which] = t
            return t.data_ptr()
        except Exception as ex:  # surfaced as IBGS_EALLOC by the library
            self.error = ex
            return None

    def release(self):
        """Break the allocator <-> ctypes-callback reference cycle so the scratch tensors return to torch's
        caching allocator immediately (not at the next cyclic GC pass)."""
        self.fn = None
        self.scratch = []
        self.bufs = {}


def _fill_view(view, rs, device, sh_coeffs, keep):
    H, W = int(rs.image_height), int(rs.image_width)
    view.image_height, view.image_width = H, W
    view.tanfovx, view.tanfovy = float(rs.tanfovx), float(rs.tanfovy)
    view.scale_modifier = float(rs.scale_modifier)
    view.sh_degree = int(rs.sh_degree)
    view.sh_coeffs = int(sh_coeffs)
    view.nb_src_images = int(rs.nb_src_images)
    view.buffer_length = int(rs.buffer_length)
    view.depth_error_threshold = float(rs.depth_error_threshold)
    view.prefiltered = int(bool(rs.prefiltered))
    view.render_geo = int(bool(rs.render_geo))
    view.render_depth_only = int(bool(rs.render_depth_only))
    view.debug = int(bool(rs.debug))
    for name in ("bg", "viewmatrix", "projmatrix", "campos", "ref_to_src_list", "src_cam_pos", "src_images",
                 "src_rendered_depths"):
        t = _f32c(getattr(rs, name), device)
        keep.append(t)
        setattr(view, name, _ptr(t))
    return keep[-2], keep[-1]  # src_images, src_rendered_depths (contiguous versions)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, means2D_abs, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, all_maps, raster_settings, sh_rest=None, accumulate_grads=False):
        rs = raster_settings
        if means3D.ndimension() != 2 or means3D.size(1) != 3:
            # rasterize_points.cu:69-71
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        if not means3D.is_cuda:
            raise RuntimeError("ibgs_b200: means3D must be a CUDA tensor (there is no CPU path)")
        device = means3D.device
        P = means3D.size(0)
        H, W = int(rs.image_height), int(rs.image_width)

        means3D_c = _f32c(means3D, device)
        sh_c = _f32c(sh, device)
        # extension over the reference API: `sh` may hold only the DC coefficient [P,1,3] with coefficients 1..M-1 in
        # their own tensor `sh_rest` [P,M-1,3] (GaussianModel stores them apart, scene/gaussian_model.py:139-143)
        split_sh = sh_rest is not None and sh_rest.numel() != 0
        sh_rest_c = _f32c(sh_rest, device) if split_sh else torch.empty(0, device=device)
        if split_sh and (sh_c is None or sh_c.numel() == 0 or sh_c.size(1) != 1):
            raise RuntimeError("shs_rest needs shs to be the [P,1,3] DC coefficients")
        colors_c = _f32c(colors_precomp, device)
        opac_c = _f32c(opacities, device)
        scales_c = _f32c(scales, device)
        rot_c = _f32c(rotations, device)
        cov_c = _f32c(cov3Ds_precomp, device)
        allmap_c = _f32c(all_maps, device)

        # The reference binding zero-fills all nine outputs (torch::full, rasterize_points.cu:80-90) and its kernels
        # leave unused slots untouched.  In render_geo mode our tile renderer writes every word of every output
        # (zeros included) and preprocess writes every radius, so torch.empty is enough; in the other modes the
        # untouched outputs must read as zeros.  (Separate tensors on purpose: callers keep single outputs alive.)
        make = _poisoned_empty if (rs.render_geo and P > 0) else torch.zeros
        fopt = dict(dtype=torch.float32, device=device)
        iopt = dict(dtype=torch.int32, device=device)
        color = make((3, H, W), **fopt)
        radii = make((P,), **iopt)
        out_normal_map = make((3, H, W), **fopt)
        out_median_intersected_depth = make((1, H, W), **fopt)
        out_cam_feat = make((4 * _M, H, W), **fopt)
        out_warped_image = make((3 * _M, H, W), **fopt)
        out_min_depth_diff = make((1, H, W), **fopt)
        out_camera_ray = make((3, H, W), **fopt)
        out_use_first_src_frame = make((1, H, W), **iopt)

        alloc = _Allocator(device)
        keep = []
        a = N.IbgsForwardArgs()
        a.P = P
        M_sh = sh_c.size(1) if (sh_c is not None and sh_c.numel() != 0) else 0
        if split_sh:
            M_sh = 1 + sh_rest_c.size(1)
        src_images_c, src_depths_c = _fill_view(a.view, rs, device, M_sh, keep)
        a.means3D = _ptr(means3D_c)
        a.shs = _ptr(sh_c)
        a.shs_rest = _ptr(sh_rest_c) if split_sh else None
        a.colors_precomp = _ptr(colors_c)
        a.opacities = _ptr(opac_c)
        a.scales = _ptr(scales_c)
        a.rotations = _ptr(rot_c)
        a.cov3D_precomp = _ptr(cov_c)
        a.all_map = _ptr(allmap_c)
        a.out_color = color.data_ptr()
        a.radii = radii.data_ptr() if P > 0 else None
        a.out_normal_map = out_normal_map.data_ptr()
        a.out_median_intersected_depth = out_median_intersected_depth.data_ptr()
        a.out_cam_feat = out_cam_feat.data_ptr()
        a.out_warped_image = out_warped_image.data_ptr()
        a.out_min_depth_diff = out_min_depth_diff.data_ptr()
        a.out_camera_ray = out_camera_ray.data_ptr()
        a.out_use_first_src_frame = out_use_first_src_frame.data_ptr()
        a.alloc = alloc.fn
        a.alloc_user = None

        def run():
            with torch.cuda.device(device):
                stream = torch.cuda.current_stream(device).cuda_stream
                rc = N.lib.ibgs_forward(C.byref(a), C.c_void_p(stream))
            if rc < 0:
                if alloc.error is not None:
                    raise alloc.error
                raise RuntimeError(f"ibgs_forward failed ({rc}): {N.last_error()}")
            return int(rc)

        if rs.debug:
            # same crash-repro aid as the reference (__init__.py:101-114)
            cpu_args = cpu_deep_copy_tuple((rs.bg, means3D, colors_precomp, opacities, scales, rotations,
                                            rs.scale_modifier, cov3Ds_precomp, all_maps, rs.viewmatrix,
                                            rs.projmatrix, rs.ref_to_src_list, rs.src_cam_pos, rs.src_images,
                                            rs.src_rendered_depths, rs.nb_src_images, rs.buffer_length,
                                            rs.depth_error_threshold, rs.tanfovx, rs.tanfovy, rs.image_height,
                                            rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered,
                                            rs.render_geo, rs.render_depth_only, rs.debug))
            try:
                num_rendered = run()
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered = run()

        empty = torch.empty(0, dtype=torch.uint8, device=device)
        geomBuffer = alloc.bufs.get(N.IBGS_BUF_GEOM, empty)
        binningBuffer = alloc.bufs.get(N.IBGS_BUF_BINNING, empty)
        imgBuffer = alloc.bufs.get(N.IBGS_BUF_IMAGE, empty)
        if KEEP_STATE:
            LAST_STATE.clear()
            LAST_STATE.update(geom=geomBuffer, binning=binningBuffer, image=imgBuffer,
                              scratch=alloc.scratch[-1] if alloc.scratch else empty,
                              num_rendered=num_rendered, P=P, H=H, W=W,
                              scratch_capacity=int(a.scratch_capacity_out))
        a.alloc = N.ALLOC_FN()
        alloc.release()

        ctx.raster_settings = rs
        # accumulate_grads (extension over the reference API): the original input tensors, so that backward can add into
        # the .grad of those that are leaves with an allocated gradient (gradient accumulation over a view batch)
        ctx.accum_inputs = dict(means3D=means3D, means2D=means2D, means2D_abs=means2D_abs, sh=sh, sh_rest=sh_rest,
                                opacities=opacities, scales=scales, rotations=rotations,
                                all_map=all_maps) if accumulate_grads else None
        ctx.num_rendered = num_rendered
        ctx.opacity_shape = tuple(opacities.shape)  # reference returns [P,1] (rasterize_points.cu:215)
        ctx.tex_token = (int(a.tex_generation_out), src_images_c, src_depths_c,
                         getattr(src_images_c, "_version", 0) if src_images_c is not None else 0,
                         getattr(src_depths_c, "_version", 0) if src_depths_c is not None else 0)
        ctx.view_keep = keep
        ctx.save_for_backward(out_normal_map, out_median_intersected_depth, out_warped_image, colors_c,
                              allmap_c, means3D_c, scales_c, rot_c, cov_c, radii, sh_c, geomBuffer,
                              binningBuffer, imgBuffer, sh_rest_c)
        ctx.mark_non_differentiable(radii, out_use_first_src_frame)
        return (color, radii, out_normal_map, out_median_intersected_depth, out_cam_feat, out_warped_image,
                out_min_depth_diff, out_camera_ray, out_use_first_src_frame)

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_out_normal_map, grad_out_median_intersected_depth,
                 grad_out_cam_feat, grad_out_warped_image, grad_out_min_depth_diff, grad_out_camera_ray,
                 grad_out_use_first_src_frame):
        rs = ctx.raster_settings
        (normal_map_pixels, median_intersected_depth_pixels, warped_image_pixels, colors_precomp, all_maps,
         means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer, sh_rest) = ctx.saved_tensors
        split_sh = sh_rest.numel() != 0
        device = means3D.device
        P = means3D.size(0)
        H, W = int(rs.image_height), int(rs.image_width)
        M_sh = sh.size(1) if sh.numel() != 0 else 0
        if split_sh:
            M_sh = 1 + sh_rest.size(1)
        fopt = dict(dtype=torch.float32, device=device)

        def cot(g, shape):
            if g is None:
                return torch.zeros(shape, **fopt)
            return _f32c(g, device)

        g_color = cot(grad_out_color, (3, H, W))
        if rs.render_geo:
            g_normal = cot(grad_out_normal_map, (3, H, W))
            g_depth = cot(grad_out_median_intersected_depth, (1, H, W))
            g_warped = cot(grad_out_warped_image, (3 * _M, H, W))
        else:  # not read by the kernel (backward.cu:587-593 only loads them under render_geo)
            g_normal = g_depth = g_warped = None

        # every row is written by the kernel (zeros for culled Gaussians) -> no memset needed
        E = _poisoned_empty
        dL_dmeans3D = E((P, 3), **fopt)
        dL_dmeans2D = E((P, 3), **fopt)
        dL_dmeans2D_abs = E((P, 3), **fopt)
        dL_dcolors = E((P, 3), **fopt)
        dL_dall_map = E((P, 5), **fopt)
        dL_dopacity = E((P, 1), **fopt)
        need_cov = cov3Ds_precomp.numel() != 0
        dL_dcov3D = E((P, 6), **fopt) if need_cov else torch.zeros((0,), **fopt)
        dL_dsh = E((P, 1 if split_sh else M_sh, 3), **fopt)
        dL_dsh_rest = E((P, M_sh - 1, 3), **fopt) if split_sh else None
        dL_dscales = E((P, 3), **fopt)
        dL_drotations = E((P, 4), **fopt)

        accumulated = set()   # inputs whose gradient the kernel added into .grad: autograd gets None for them
        if P != 0:
            alloc = _Allocator(device)
            keep = []
            a = N.IbgsBackwardArgs()
            a.P = P
            a.R = int(ctx.num_rendered)
            _fill_view(a.view, rs, device, M_sh, keep)
            a.means3D = _ptr(means3D)
            a.shs = _ptr(sh)
            a.shs_rest = _ptr(sh_rest) if split_sh else None
            a.colors_precomp = _ptr(colors_precomp)
            a.scales = _ptr(scales)
            a.rotations = _ptr(rotations)
            a.cov3D_precomp = _ptr(cov3Ds_precomp)
            a.all_map = _ptr(all_maps)
            a.radii = radii.data_ptr()
            a.out_median_intersected_depth = median_intersected_depth_pixels.data_ptr()
            a.out_warped_image = warped_image_pixels.data_ptr()
            a.geom_buffer = geomBuffer.data_ptr()
            a.binning_buffer = binningBuffer.data_ptr()
            a.image_buffer = imgBuffer.data_ptr()
            gen, simg, sdep, v_img, v_dep = ctx.tex_token
            same = (simg is not None and sdep is not None and keep[-2] is not None and keep[-1] is not None
                    and simg.numel() and keep[-2].data_ptr() == simg.data_ptr()
                    and keep[-1].data_ptr() == sdep.data_ptr()
                    and simg._version == v_img and sdep._version == v_dep)
            a.tex_generation = gen if same else 0
            a.dL_dout_color = g_color.data_ptr()
            a.dL_dout_normal_map = _ptr(g_normal)
            a.dL_dout_median_intersected_depth = _ptr(g_depth)
            a.dL_dout_warped_image = _ptr(g_warped)
            a.dL_dmeans3D = dL_dmeans3D.data_ptr()
            a.dL_dmeans2D = dL_dmeans2D.data_ptr()
            a.dL_dmeans2D_abs = dL_dmeans2D_abs.data_ptr()
            a.dL_dcolors = dL_dcolors.data_ptr()
            a.dL_dopacity = dL_dopacity.data_ptr()
            a.dL_dcov3D = dL_dcov3D.data_ptr() if need_cov else None
            a.dL_dsh = dL_dsh.data_ptr() if M_sh else None
            a.dL_dsh_rest = dL_dsh_rest.data_ptr() if split_sh else None
            a.dL_dscales = dL_dscales.data_ptr()
            a.dL_drotations = dL_drotations.data_ptr()
            a.dL_dall_map = dL_dall_map.data_ptr()
            mask = 0
            if ctx.accum_inputs is not None:
                for name, field, bit in (("means3D", "dL_dmeans3D", N.ACC_MEANS3D), ("means2D", "dL_dmeans2D", N.ACC_MEANS2D),
                                         ("means2D_abs", "dL_dmeans2D_abs", N.ACC_MEANS2D_ABS),
                                         ("opacities", "dL_dopacity", N.ACC_OPACITY), ("sh", "dL_dsh", N.ACC_SH),
                                         ("sh_rest", "dL_dsh_rest", N.ACC_SH_REST), ("scales", "dL_dscales", N.ACC_SCALES),
                                         ("rotations", "dL_drotations", N.ACC_ROTATIONS),
                                         ("all_map", "dL_dall_map", N.ACC_ALL_MAP)):
                    t = ctx.accum_inputs.get(name)
                    if t is not None and t.numel() and _accumulable(t) and getattr(a, field):
                        setattr(a, field, t.grad.data_ptr())
                        mask |= bit
                        accumulated.add(name)
            a.accumulate_mask = mask
            a.alloc = alloc.fn
            a.alloc_user = None

            def run():
                with torch.cuda.device(device):
                    stream = torch.cuda.current_stream(device).cuda_stream
                    rc = N.lib.ibgs_backward(C.byref(a), C.c_void_p(stream))
                if rc < 0:
                    if alloc.error is not None:
                        raise alloc.error
                    raise RuntimeError(f"ibgs_backward failed ({rc}): {N.last_error()}")

            if rs.debug:
                try:
                    run()
                except Exception as ex:
                    torch.save(cpu_deep_copy_tuple((rs.bg, normal_map_pixels, median_intersected_depth_pixels,
                                                    warped_image_pixels, means3D, radii, colors_precomp, all_maps,
                                                    scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                                                    rs.viewmatrix, rs.projmatrix, grad_out_color, sh,
                                                    rs.sh_degree, rs.campos)), "snapshot_bw.dump")
                    print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                    raise ex
            else:
                run()
            a.alloc = N.ALLOC_FN()
            alloc.release()

        need = ctx.needs_input_grad
        acc = accumulated
        grads = (
            dL_dmeans3D if (need[0] and "means3D" not in acc) else None,
            dL_dmeans2D if (need[1] and "means2D" not in acc) else None,
            dL_dmeans2D_abs if (need[2] and "means2D_abs" not in acc) else None,
            dL_dsh if (need[3] and M_sh and "sh" not in acc) else None,
            dL_dcolors if (need[4] and colors_precomp.numel()) else None,
            dL_dopacity.view(ctx.opacity_shape) if (need[5] and "opacities" not in acc) else None,
            dL_dscales if (need[6] and scales.numel() and "scales" not in acc) else None,
            dL_drotations if (need[7] and rotations.numel() and "rotations" not in acc) else None,
            dL_dcov3D if (need[8] and need_cov) else None,
            dL_dall_map if (need[9] and all_maps.numel() and "all_map" not in acc) else None,
            None,
            dL_dsh_rest if (split_sh and len(need) > 11 and need[11] and "sh_rest" not in acc) else None,
            None,
        )
        return grads


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    ref_to_src_list: torch.Tensor
    src_cam_pos: torch.Tensor
    src_images: torch.Tensor
    src_rendered_depths: torch.Tensor
    nb_src_images: int
    buffer_length: int
    depth_error_threshold: float
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    render_geo: bool
    render_depth_only: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # reference __init__.py:283-292 -> rasterize_points.cu:273-292
        with torch.no_grad():
            rs = self.raster_settings
            device = positions.device
            pos = _f32c(positions, device)
            P = pos.size(0)
            visible = torch.zeros((P,), dtype=torch.bool, device=device)
            if P != 0:
                view = _f32c(rs.viewmatrix, device)
                proj = _f32c(rs.projmatrix, device)
                with torch.cuda.device(device):
                    stream = torch.cuda.current_stream(device).cuda_stream
                    N.check(N.lib.ibgs_mark_visible(P, pos.data_ptr(), view.data_ptr(), proj.data_ptr(),
                                                    visible.data_ptr(), C.c_void_p(stream)), "ibgs_mark_visible")
        return visible

    def forward(self, means3D, means2D, means2D_abs, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None, all_map=None, shs_rest=None, accumulate_grads=False):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        if all_map is None:
            all_map = torch.Tensor([])

        # shs_rest (not in the reference API): shs = _features_dc, shs_rest = _features_rest, read in place.
        # accumulate_grads (not in the reference API): inputs that are leaves with an allocated float32 .grad get their
        # gradient ADDED into it by the backward kernel itself (what autograd's AccumulateGrad would do in a separate
        # pass per tensor); every other input receives its gradient through autograd as usual.
        return rasterize_gaussians(means3D, means2D, means2D_abs, shs, colors_precomp, opacities, scales,
                                   rotations, cov3D_precomp, all_map, raster_settings, shs_rest, accumulate_grads)
