"""Optional fast path in front of the rasterizer: the per-view parameter prologue of
`gaussian_renderer.render()` as two CUDA launches per direction instead of ~30 PyTorch kernels.

    opacity, scales, rotations, shs, all_map = gaussian_prologue(pc_xyz, pc_opacity, pc_scaling, pc_rotation,
                                                                 pc_features_dc, pc_features_rest, pc_normal, pc_offset,
                                                                 camera.world_view_transform, camera.camera_center)

is a drop-in for these lines of the reference (same values, same gradients):
    opacity   = pc.get_opacity                    scene/gaussian_model.py:145-147   torch.sigmoid(_opacity)
    scales    = pc.get_scaling                    :127-129                          torch.exp(_scaling)
    rotations = pc.get_rotation                   :131-133                          F.normalize(_rotation)
    shs       = pc.get_features                   :139-143                          cat(_features_dc, _features_rest)
    all_map   = [normal_cam, 1, |plane distance|] gaussian_renderer/__init__.py:304-315 with pc.get_normal(camera)
                                                  (scene/gaussian_model.py:166-173, learnt_normal=True)
                                                  or, with smallest_axis_normal=True (learnt_normal=False),
                                                  pc.get_normal_w_smallest_axis(camera) (:149-161): the column
                                                  argmin(scales) of quaternion_to_matrix(get_rotation), no offset
The outputs feed `GaussianRasterizer` unchanged.  Like the rasterizer there is no CPU / PyTorch fallback.
"""
import ctypes as C

import torch

from . import _native as N


def _c(t):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:   # a slice at an odd offset of a packed arena: quaternion rows are accessed as float4
        t = t.clone()
    return t


class _GaussianPrologue(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, opacity_raw, scaling_raw, rotation_raw, features_dc, features_rest, normal_raw, offset,
                world_view_transform, camera_center, concat_sh=True, smallest_axis_normal=False):
        if not xyz.is_cuda:
            raise RuntimeError("ibgs_b200.fused: parameters must be CUDA tensors (there is no CPU path)")
        device = xyz.device
        P = xyz.size(0)
        K1 = features_rest.size(1) if features_rest.numel() else 0
        learnt = normal_raw is not None
        with_map = learnt or smallest_axis_normal
        ins = [_c(t) for t in (xyz, opacity_raw, scaling_raw, rotation_raw, features_dc, features_rest)]
        nrm, off = (_c(normal_raw), _c(offset)) if learnt else (None, None)
        view, cam = _c(world_view_transform).to(device), _c(camera_center).to(device)
        fopt = dict(dtype=torch.float32, device=device)
        opacity = torch.empty((P, 1), **fopt)
        scales = torch.empty((P, 3), **fopt)
        rotations = torch.empty((P, 4), **fopt)
        shs = torch.empty((P, K1 + 1, 3), **fopt) if concat_sh else None
        all_map = torch.empty((P, 5), **fopt) if with_map else None
        a = N.IbgsPrologueArgs()
        a.P, a.sh_rest = P, K1
        for name, t in zip(("xyz", "opacity_raw", "scaling_raw", "rotation_raw", "features_dc", "features_rest"), ins):
            setattr(a, name, t.data_ptr() if t.numel() else None)
        if learnt:
            a.normal_raw, a.offset = nrm.data_ptr(), off.data_ptr()
        a.smallest_axis_normal = int(bool(smallest_axis_normal))
        a.world_view_transform, a.camera_center = view.data_ptr(), cam.data_ptr()
        a.opacity, a.scales, a.rotations = opacity.data_ptr(), scales.data_ptr(), rotations.data_ptr()
        a.shs = shs.data_ptr() if concat_sh else None
        a.all_map = all_map.data_ptr() if with_map else None
        if P:
            with torch.cuda.device(device):
                N.check(N.lib.ibgs_prologue_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
                        "ibgs_prologue_forward")
        ctx.with_map = with_map
        ctx.learnt = learnt
        ctx.concat_sh = concat_sh
        ctx.shapes = [tuple(t.shape) for t in (xyz, opacity_raw, scaling_raw, rotation_raw, features_dc, features_rest)]
        ctx.save_for_backward(*ins, *((nrm, off) if learnt else ()), view, cam)
        outs = [opacity, scales, rotations]
        if concat_sh:
            outs.append(shs)
        if with_map:
            outs.append(all_map)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_opacity, g_scales, g_rotations, *rest):
        rest = list(rest)
        g_shs = rest.pop(0) if ctx.concat_sh else None
        g_all_map = rest.pop(0) if ctx.with_map else None
        saved = ctx.saved_tensors
        xyz, opacity_raw, scaling_raw, rotation_raw, fdc, frest = saved[:6]
        with_map, learnt = ctx.with_map, ctx.learnt
        nrm, off = (saved[6], saved[7]) if learnt else (None, None)
        view, cam = saved[-2], saved[-1]
        device = xyz.device
        P = xyz.size(0)
        K1 = frest.size(1) if frest.numel() else 0
        keep = [None if g is None else _c(g) for g in (g_opacity, g_scales, g_rotations, g_shs, g_all_map)]
        d = {n: torch.empty_like(t) for n, t in (("opacity_raw", opacity_raw), ("scaling_raw", scaling_raw),
                                                ("rotation_raw", rotation_raw))}
        if ctx.concat_sh:
            d.update(features_dc=torch.empty_like(fdc), features_rest=torch.empty_like(frest))
        if with_map:
            d.update(xyz=torch.empty_like(xyz))
        if learnt:
            d.update(normal_raw=torch.empty_like(nrm), offset=torch.empty_like(off))
        a = N.IbgsPrologueArgs()
        a.P, a.sh_rest = P, K1
        for name, t in (("xyz", xyz), ("opacity_raw", opacity_raw), ("scaling_raw", scaling_raw),
                        ("rotation_raw", rotation_raw), ("features_dc", fdc), ("features_rest", frest)):
            setattr(a, name, t.data_ptr() if t.numel() else None)
        if learnt:
            a.normal_raw, a.offset = nrm.data_ptr(), off.data_ptr()
        a.smallest_axis_normal = int(with_map and not learnt)
        a.world_view_transform, a.camera_center = view.data_ptr(), cam.data_ptr()
        for name, g in zip(("g_opacity", "g_scales", "g_rotations", "g_shs", "g_all_map"), keep):
            setattr(a, name, None if g is None else g.data_ptr())
        for name, t in d.items():
            setattr(a, "d_" + name, t.data_ptr() if t.numel() else None)
        if P:
            with torch.cuda.device(device):
                N.check(N.lib.ibgs_prologue_backward(C.byref(a), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
                        "ibgs_prologue_backward")
        need = ctx.needs_input_grad
        sh = ctx.shapes

        def out(i, name):
            return d[name].view(sh[i]) if (need[i] and name in d) else None

        return (out(0, "xyz"), out(1, "opacity_raw"), out(2, "scaling_raw"), out(3, "rotation_raw"),
                out(4, "features_dc"), out(5, "features_rest"),
                d["normal_raw"] if (learnt and need[6]) else None, d["offset"] if (learnt and need[7]) else None,
                None, None, None, None)


def gaussian_prologue(xyz, opacity_raw, scaling_raw, rotation_raw, features_dc, features_rest, normal_raw=None,
                      offset=None, world_view_transform=None, camera_center=None, concat_sh=True,
                      smallest_axis_normal=False):
    """Activated rasterizer inputs from the raw GaussianModel parameters; see the module docstring.
    With normal_raw/offset = None no all_map is produced (4 outputs instead of 5).  With concat_sh=False no `shs` is
    produced either (the tuple is opacity, scales, rotations[, all_map]): pass features_dc / features_rest to the
    rasterizer as `shs=` / `shs_rest=` and it reads them in place, which saves the torch.cat round trip per view.
    smallest_axis_normal=True is render(..., learnt_normal=False): the plane normal is the Gaussian's shortest axis
    (normal_raw / offset must be None); its gradient reaches rotation_raw and xyz."""
    if (normal_raw is None) != (offset is None):
        raise ValueError("normal_raw and offset must be given together")
    if smallest_axis_normal and normal_raw is not None:
        raise ValueError("smallest_axis_normal=True excludes normal_raw / offset")
    if (normal_raw is not None or smallest_axis_normal) and (world_view_transform is None or camera_center is None):
        raise ValueError("all_map needs world_view_transform and camera_center")
    if world_view_transform is None:
        world_view_transform = torch.eye(4, device=xyz.device)
        camera_center = torch.zeros(3, device=xyz.device)
    return _GaussianPrologue.apply(xyz, opacity_raw, scaling_raw, rotation_raw, features_dc, features_rest, normal_raw,
                                   offset, world_view_transform, camera_center, concat_sh, smallest_axis_normal)
