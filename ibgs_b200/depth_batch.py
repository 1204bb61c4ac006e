"""Batched source-view depth renders -- optional fast path in front of the C ABI's
`ibgs_forward_depth_batch`.

The reference renders the plane depth of every source view with its own rasterizer call
(gaussian_renderer/__init__.py:245-253 -> render_depth, :33-145): per view ~10 torch kernels for the all_map, then
preprocess + sort + depth-only blend over the SAME Gaussians.  `render_depth_batch` renders V views of the same
image size in one pass (one preprocess launch, one binning pass over the V*P (view, Gaussian) items, one tile-renderer
launch); `render_depth_views` has render_depth's argument list with a LIST of cameras and returns the stacked
[V,1,H,W] tensor the caller builds at :252.

No gradient: the reference calls render_depth for source depths only and never back-propagates through it (the
stacked depths go into the rasterizer settings, not into the graph).  There is no CPU or PyTorch fallback.
"""
import ctypes as C
import math
from typing import NamedTuple, Optional

import torch

from . import _native as N
from .diff_plane_rasterization import _Allocator, _f32c, _poisoned_empty, _ptr


class DepthBatchSettings(NamedTuple):
    """The view-independent fields of GaussianRasterizationSettings (reference __init__.py:252-276) plus the stacked
    per-view matrices."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    scale_modifier: float
    viewmatrices: torch.Tensor      # [V,4,4] world_view_transform of every view
    projmatrices: torch.Tensor      # [V,4,4] full_proj_transform of every view
    buffer_length: int
    prefiltered: bool = False
    debug: bool = False


def render_depth_batch(settings: DepthBatchSettings, means3D, opacities, scales=None, rotations=None,
                       cov3D_precomp=None, all_maps=None, normals=None, offsets=None, camera_centers=None,
                       return_radii=False, return_counts=False):
    """Plane-intersection depth of V views: [V,1,H,W] float32 (what V depth-only GaussianRasterizer calls return as
    out_median_intersected_depth).  Plane parameters: `all_maps` [V,P,5] as render_depth builds them per view, or
    `normals` [P,3] (world space: GaussianModel._normal, or the shortest axis) + optional `offsets` [P] / [P,1] +
    `camera_centers` [V,3] and the kernel derives them per view."""
    st = settings
    if means3D.ndimension() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("ibgs_b200: means3D must be a CUDA tensor (there is no CPU path)")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
    if (all_maps is None) == (normals is None):
        raise Exception('Please provide exactly one of either all_maps or normals (+ camera_centers)!')
    device = means3D.device
    P = means3D.size(0)
    H, W = int(st.image_height), int(st.image_width)
    vm = _f32c(st.viewmatrices, device).reshape(-1, 16)
    pm = _f32c(st.projmatrices, device).reshape(-1, 16)
    V = vm.size(0)
    if pm.size(0) != V or not (1 <= V <= N.MAX_DEPTH_BATCH):
        raise RuntimeError(f"need 1..{N.MAX_DEPTH_BATCH} views with one projmatrix each, got {V} / {pm.size(0)}")
    if all_maps is not None and tuple(all_maps.shape) != (V, P, 5):
        raise RuntimeError(f"all_maps must be [V,P,5] = {(V, P, 5)}, got {tuple(all_maps.shape)}")
    if normals is not None:
        if tuple(normals.shape) != (P, 3):
            raise RuntimeError(f"normals must be [P,3], got {tuple(normals.shape)}")
        if camera_centers is None or camera_centers.numel() != 3 * V:
            raise RuntimeError("normals need camera_centers [V,3]")
        if offsets is not None and offsets.numel() != P:
            raise RuntimeError(f"offsets must hold P = {P} values")

    keep = [_f32c(t, device) for t in (means3D, opacities, scales, rotations, cov3D_precomp, all_maps, normals,
                                       offsets, camera_centers)]
    # zeros: with P == 0 (or an empty view) the untouched planes must read as 0 like the reference's outputs
    depths = (torch.zeros if P == 0 else _poisoned_empty)((V, 1, H, W), dtype=torch.float32, device=device)
    radii = torch.zeros((V, P), dtype=torch.int32, device=device) if return_radii else None
    counts = (C.c_int64 * V)()
    alloc = _Allocator(device)
    a = N.IbgsDepthBatchArgs()
    a.P, a.V = P, V
    a.image_height, a.image_width = H, W
    a.tanfovx, a.tanfovy, a.scale_modifier = float(st.tanfovx), float(st.tanfovy), float(st.scale_modifier)
    a.buffer_length = int(st.buffer_length)
    a.prefiltered, a.debug = int(bool(st.prefiltered)), int(bool(st.debug))
    a.viewmatrices, a.projmatrices = vm.data_ptr(), pm.data_ptr()
    (a.means3D, a.opacities, a.scales, a.rotations, a.cov3D_precomp, a.all_maps, a.normals, a.offsets,
     a.camera_centers) = [_ptr(t) for t in keep]
    a.out_depths = depths.data_ptr()
    a.radii = _ptr(radii)
    a.num_rendered = counts
    a.alloc = alloc.fn
    a.alloc_user = None
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = N.lib.ibgs_forward_depth_batch(C.byref(a), C.c_void_p(stream))
    a.alloc = N.ALLOC_FN()
    err = alloc.error
    alloc.release()
    if rc < 0:
        if err is not None:
            raise err
        raise RuntimeError(f"ibgs_forward_depth_batch failed ({rc}): {N.last_error()}")
    out = (depths,)
    if return_radii:
        out += (radii,)
    if return_counts:
        out += ([int(c) for c in counts],)
    return out[0] if len(out) == 1 else out


def render_depth_views(viewpoint_cameras, pc, scene, pipe, args, bg_color, learnt_normal: bool, nb_src_frames: int,
                       buffer_length: int, depth_error_threshold: Optional[float] = None, scaling_modifier=1.0,
                       override_color=None):
    """render_depth (gaussian_renderer/__init__.py:33-145) for a list of cameras of one image size: returns the stacked
    [V,1,H,W] depths, i.e. `torch.stack([render_depth(v, ...) for v in viewpoint_cameras])` (:245-252).  Arguments the
    depth-only render does not use (scene, args, bg_color, nb_src_frames, depth_error_threshold, override_color) are
    accepted for signature compatibility."""
    cams = list(viewpoint_cameras)
    c0 = cams[0]
    for c in cams[1:]:
        if (c.image_height, c.image_width, c.FoVx, c.FoVy) != (c0.image_height, c0.image_width, c0.FoVx, c0.FoVy):
            raise RuntimeError("render_depth_views needs cameras of one image size and field of view")
    st = DepthBatchSettings(
        image_height=int(c0.image_height), image_width=int(c0.image_width),
        tanfovx=math.tan(c0.FoVx * 0.5), tanfovy=math.tan(c0.FoVy * 0.5), scale_modifier=scaling_modifier,
        viewmatrices=torch.stack([c.world_view_transform for c in cams]),
        projmatrices=torch.stack([c.full_proj_transform for c in cams]),
        buffer_length=buffer_length, prefiltered=False, debug=bool(pipe.debug))
    kw = {}
    if pipe.compute_cov3D_python:
        kw["cov3D_precomp"] = pc.get_covariance(scaling_modifier)
    else:
        kw["scales"], kw["rotations"] = pc.get_scaling, pc.get_rotation
    if learnt_normal:
        # get_normal (scene/gaussian_model.py:166-173) normalises, flips towards each camera and flips the offset with
        # it: done per view inside the kernel
        kw["normals"], kw["offsets"] = pc._normal, pc.get_offset()
    else:
        kw["normals"] = pc.get_smallest_axis()
    kw["camera_centers"] = torch.stack([c.camera_center for c in cams])
    with torch.no_grad():
        return render_depth_batch(st, pc.get_xyz, pc.get_opacity, **kw)
