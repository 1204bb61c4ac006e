"""Optional fast path for the per-view densification statistics.

    from ibgs_b200.densify import add_densification_stats
    add_densification_stats(gaussians, render_pkg)          # instead of train.py:399-405

does what the training loop does after every view while densification is on --

    mask = visibility_filter
    gaussians.max_radii2D[mask] = torch.max(gaussians.max_radii2D[mask], radii[mask])
    gaussians.add_densification_stats(viewspace_point_tensor, viewspace_point_tensor_abs, visibility_filter)

(scene/gaussian_model.py:600-604) -- in ONE launch over the Gaussians and without the five host synchronisations of the
boolean-mask indexing.  The statistics tensors of the (unchanged) GaussianModel are updated in place.  No CPU path.
"""
import ctypes as C

import torch

from . import _native as N


def add_densification_stats(gaussians, render_pkg):
    radii = render_pkg["radii"]
    g = render_pkg["viewspace_points"].grad
    ga = render_pkg["viewspace_points_abs"].grad
    if g is None or ga is None:
        raise RuntimeError("viewspace_points(.abs).grad is missing: call this after loss.backward()")
    if not radii.is_cuda:
        raise RuntimeError("ibgs_b200.densify: tensors must be CUDA tensors (there is no CPU path)")
    P = radii.numel()
    stats = [gaussians.max_radii2D, gaussians.xyz_gradient_accum, gaussians.xyz_gradient_accum_abs, gaussians.denom,
             gaussians.denom_abs]
    for t in stats:
        if t.numel() != P or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("densification statistics must be contiguous float32 tensors with one value per Gaussian")
    radii_c = radii.contiguous() if radii.dtype == torch.int32 else radii.to(torch.int32)
    g_c, ga_c = g.detach().float().contiguous(), ga.detach().float().contiguous()
    with torch.cuda.device(radii.device), torch.no_grad():
        N.check(N.lib.ibgs_densification_stats(P, radii_c.data_ptr(), g_c.data_ptr(), ga_c.data_ptr(),
                                               *[t.data_ptr() for t in stats],
                                               C.c_void_p(torch.cuda.current_stream(radii.device).cuda_stream)),
                "ibgs_densification_stats")
