"""GPU: Gaussian counts that are not a multiple of 4 through the packed-arena fast paths (round-1 advisor finding).

ArenaAdam / GradArena pack the per-Gaussian groups into one buffer; the rasterizer, its backward (accumulate mode) and
the fused prologue access quaternion rows as float4.  With P % 4 != 0 -- any P after a densification -- a group packed
back to back would start at a 4- or 8-byte offset and the kernels would fault ("misaligned address").  Groups now start
on 256-byte boundaries, the Python layer re-aligns foreign tensors, and the C ABI rejects a misaligned quaternion pointer
with IBGS_EINVAL instead of faulting."""
import ctypes as C

import pytest
import torch

from ibgs_b200 import synthetic as S
import ibgs_testutil as U

pytestmark = pytest.mark.gpu


def _raw_params(sc):
    op = sc["opacities"].clamp(1e-4, 1 - 1e-4)
    P = sc["means3D"].shape[0]
    return {"xyz": sc["means3D"].clone(), "f_dc": sc["shs"][:, :1, :].contiguous(), "f_rest": sc["shs"][:, 1:, :].contiguous(),
            "opacity": torch.log(op / (1 - op)), "scaling": torch.log(sc["scales"]), "rotation": sc["rotations"].clone(),
            "normal": sc["normals_world"].clone(), "offset": torch.zeros((P, 1), device="cuda")}


def _step(dpr, opt, sc, cot):
    from ibgs_b200.fused import gaussian_prologue
    pr = opt.params
    opacity, scales, rotations, all_map = gaussian_prologue(
        pr["xyz"], pr["opacity"], pr["scaling"], pr["rotation"], pr["f_dc"], pr["f_rest"], pr["normal"], pr["offset"],
        sc["viewmatrix"], sc["campos"], concat_sh=False)
    rs = U.make_settings(dpr, sc, render_geo=True, depth_error_threshold=0.05)
    z = torch.zeros_like(pr["xyz"])
    res = dpr.GaussianRasterizer(rs)(means3D=pr["xyz"], means2D=z, means2D_abs=z, opacities=opacity, shs=pr["f_dc"],
                                     shs_rest=pr["f_rest"], scales=scales, rotations=rotations, all_map=all_map,
                                     accumulate_grads=True)
    torch.autograd.backward([res[0], res[2], res[3], res[5]], [cot["color"], cot["normal"], cot["depth"], cot["warped"]])


@pytest.mark.parametrize("P", [1001, 1002, 1003])
def test_odd_gaussian_counts_through_arena_prologue_rasterizer(P):
    import ibgs_b200.diff_plane_rasterization as dpr
    from ibgs_b200.optim import ArenaAdam
    lrs = {"xyz": 1.6e-4, "f_dc": 0.0025, "f_rest": 0.0025 / 20, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001,
           "normal": 0.001, "offset": 1.6e-5}
    sc = U.scene_to_device(S.make_scene("tiny", P=P))
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    cot = {k: v.cuda() for k, v in S.cotangents(sc).items()}
    opt = ArenaAdam(_raw_params(sc), lrs)
    for name, p in opt.params.items():
        assert p.data_ptr() % 256 == 0 and p.grad.data_ptr() % 256 == 0, name
    _step(dpr, opt, sc, cot)
    torch.cuda.synchronize()
    # same step with ordinary (separately allocated) leaves: the arena-backed gradients must equal autograd's
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in opt.params.items()}
    from ibgs_b200.fused import gaussian_prologue
    opacity, scales, rotations, all_map = gaussian_prologue(
        leaves["xyz"], leaves["opacity"], leaves["scaling"], leaves["rotation"], leaves["f_dc"], leaves["f_rest"],
        leaves["normal"], leaves["offset"], sc["viewmatrix"], sc["campos"], concat_sh=False)
    rs = U.make_settings(dpr, sc, render_geo=True, depth_error_threshold=0.05)
    z = torch.zeros_like(leaves["xyz"])
    res = dpr.GaussianRasterizer(rs)(means3D=leaves["xyz"], means2D=z, means2D_abs=z, opacities=opacity,
                                     shs=leaves["f_dc"], shs_rest=leaves["f_rest"], scales=scales, rotations=rotations,
                                     all_map=all_map)
    torch.autograd.backward([res[0], res[2], res[3], res[5]], [cot["color"], cot["normal"], cot["depth"], cot["warped"]])
    for k in leaves:
        assert U.rel_l2(opt.params[k].grad, leaves[k].grad) <= 1e-4, k
    opt.step(zero_grads=True)
    # densification surgery changes P again: prune to an odd count, extend by an odd count, keep stepping
    keep = torch.ones(P, dtype=torch.bool, device="cuda")
    keep[::7] = False
    opt2 = opt.prune(keep)
    P2 = int(keep.sum())
    sc2 = dict(sc, P=P2)
    for k in ("means3D", "shs", "opacities", "scales", "rotations", "normals_world", "all_map"):
        sc2[k] = sc[k][keep].contiguous()
    _step(dpr, opt2, sc2, cot)
    opt2.step(zero_grads=True)
    ext = {k: v.detach()[:5].clone() for k, v in opt2.params.items()}
    opt3 = opt2.extend(ext)
    sc3 = dict(sc2, P=P2 + 5)
    _step(dpr, opt3, sc3, cot)
    opt3.step(zero_grads=True)
    torch.cuda.synchronize()
    for p in opt3.params.values():
        assert torch.isfinite(p).all()


def test_misaligned_views_are_realigned_by_python_and_rejected_by_the_c_abi():
    import ibgs_b200.diff_plane_rasterization as dpr
    from ibgs_b200 import _native as N
    sc = U.scene_to_device(S.make_scene("tiny", P=1001))
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    base, _, _ = U.ours_forward_backward(dpr, sc, None, depth_error_threshold=0.05)
    # a rotation tensor that starts 4 bytes into its storage
    flat = torch.empty(sc["rotations"].numel() + 1, device="cuda")
    rot = flat[1:].view(-1, 4)
    rot.copy_(sc["rotations"])
    assert rot.data_ptr() % 16 == 4
    out, _, _ = U.ours_forward_backward(dpr, dict(sc, rotations=rot), None, depth_error_threshold=0.05)   # clones inputs
    z = torch.zeros_like(sc["means3D"])
    rs = U.make_settings(dpr, sc, depth_error_threshold=0.05)
    res = dpr.GaussianRasterizer(rs)(means3D=sc["means3D"], means2D=z, means2D_abs=z, opacities=sc["opacities"],
                                     shs=sc["shs"], scales=sc["scales"], rotations=rot, all_map=sc["all_map"])
    assert torch.equal(res[0], base["color"]) and torch.equal(res[1], base["radii"])
    # the raw C ABI with the misaligned pointer: an error code, not a sticky CUDA fault
    a = N.IbgsPrologueArgs()
    a.P = 1001
    a.rotation_raw = rot.data_ptr()
    dummy = torch.zeros((1001, 16), device="cuda")
    for f in ("xyz", "opacity_raw", "scaling_raw", "features_dc", "world_view_transform", "camera_center", "opacity",
              "scales", "rotations"):
        setattr(a, f, dummy.data_ptr())
    rc = N.lib.ibgs_prologue_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc < 0 and "aligned" in N.last_error()
    torch.cuda.synchronize()
