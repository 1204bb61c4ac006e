"""GPU: the fused parameter prologue (ibgs_b200.fused -> C ABI -> prologue.cu) against (a) the numpy float64 oracle
and (b) the reference's torch expressions in float32 with autograd, values and gradients; plus the end-to-end property
that rasterizing from raw parameters through the fused prologue equals rasterizing through the torch prologue."""
import numpy as np
import pytest
import torch

import prologue_ref as PR

pytestmark = pytest.mark.gpu
NAMES = ("opacity", "scales", "rotations", "shs", "all_map")
IN = ("xyz", "opacity_raw", "scaling_raw", "rotation_raw", "fdc", "frest", "normal_raw", "offset")


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("P,K", [(1, 9), (1000, 1), (4097, 9), (100_003, 16)])
def test_fused_prologue_matches_oracle_and_torch(P, K):
    from ibgs_b200 import fused
    from oracle import prologue_oracle as O
    p = PR.random_params(P, K=K, seed=P, device="cuda")
    leaves = {k: p[k].clone().requires_grad_(True) for k in IN}
    outs = fused.gaussian_prologue(*[leaves[k] for k in IN], p["V"], p["cam"])
    tl = {k: p[k].clone().requires_grad_(True) for k in IN}
    touts = PR.torch_prologue(*[tl[k] for k in IN], p["V"], p["cam"])
    fw = O.forward(*[p[k].cpu().numpy() for k in IN], p["V"].cpu().numpy(), p["cam"].cpu().numpy())
    for n, o, t in zip(NAMES, outs, touts):
        assert _rel(o, t) <= 2e-6, (n, _rel(o, t))
        assert _rel(o, torch.from_numpy(fw[n]).cuda().view_as(o)) <= 2e-6, n
    # the flip decision must agree with torch wherever it is not a rounding-level tie
    assert torch.equal(outs[3], touts[3])
    g = torch.Generator().manual_seed(3)
    cots = [torch.randn(o.shape, generator=g).cuda() for o in outs]
    torch.autograd.backward(list(outs), cots)
    torch.autograd.backward(list(touts), cots)
    d = O.backward(fw, *[c.cpu().numpy() for c in cots])
    name = dict(fdc="features_dc", frest="features_rest")
    for k in IN:
        if leaves[k].numel() == 0:
            continue
        assert _rel(leaves[k].grad, tl[k].grad) <= 1e-5, (k, _rel(leaves[k].grad, tl[k].grad))
        want = torch.from_numpy(d[name.get(k, k)]).cuda().view_as(leaves[k].grad)
        assert _rel(leaves[k].grad, want) <= 1e-5, k


def test_without_plane_parameters_and_argument_errors():
    from ibgs_b200 import fused
    p = PR.random_params(257, device="cuda")
    outs = fused.gaussian_prologue(p["xyz"], p["opacity_raw"], p["scaling_raw"], p["rotation_raw"], p["fdc"], p["frest"])
    assert len(outs) == 4 and torch.allclose(outs[0], torch.sigmoid(p["opacity_raw"]), rtol=1e-6, atol=1e-7)
    with pytest.raises(ValueError):
        fused.gaussian_prologue(p["xyz"], p["opacity_raw"], p["scaling_raw"], p["rotation_raw"], p["fdc"], p["frest"],
                                normal_raw=p["normal_raw"])
    with pytest.raises(RuntimeError):
        fused.gaussian_prologue(*[p[k].cpu() for k in IN], p["V"].cpu(), p["cam"].cpu())


def test_rasterizing_through_fused_prologue_equals_torch_prologue():
    """End to end: raw parameters -> prologue -> rasterizer -> loss.backward(), fused vs the reference's torch ops."""
    from ibgs_b200 import fused, synthetic as S
    import ibgs_b200.diff_plane_rasterization as dpr
    import ibgs_testutil as U
    sc = U.scene_to_device(S.make_scene("cfg1"))
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    cot = {k: v.cuda() for k, v in S.cotangents(sc).items()}
    P = sc["P"]
    raw = dict(xyz=sc["means3D"], opacity_raw=torch.logit(sc["opacities"].clamp(1e-4, 1 - 1e-4)).view(P, 1),
               scaling_raw=sc["scales"].log(), rotation_raw=sc["rotations"] * 1.7,
               fdc=sc["shs"][:, :1].contiguous(), frest=sc["shs"][:, 1:].contiguous(),
               normal_raw=sc["normals_world"] * 0.6, offset=torch.zeros((P, 1), device="cuda"))
    V, cam = sc["viewmatrix"], sc["campos"]
    rs = U.make_settings(dpr, sc, render_geo=True)
    grads = []
    for pro in (fused.gaussian_prologue, PR.torch_prologue):
        leaves = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
        opacity, scales, rotations, shs, all_map = pro(*[leaves[k] for k in IN], V, cam)
        z = torch.zeros_like(sc["means3D"])
        res = dpr.GaussianRasterizer(rs)(means3D=leaves["xyz"], means2D=z, means2D_abs=z, opacities=opacity, shs=shs,
                                         scales=scales, rotations=rotations, all_map=all_map)
        torch.autograd.backward([res[0], res[2], res[3], res[5]], [cot["color"], cot["normal"], cot["depth"], cot["warped"]])
        grads.append({k: v.grad for k, v in leaves.items()})
    for k in IN:
        assert _rel(grads[0][k], grads[1][k]) <= 1e-3, (k, _rel(grads[0][k], grads[1][k]))


@pytest.mark.parametrize("deg,P", [(2, None), (3, 2011), (1, 777)])
def test_split_sh_tensors_equal_concatenated_path(deg, P):
    """`shs=_features_dc, shs_rest=_features_rest` (read in place) must give bit-identical outputs, and the same SH
    gradients, as the reference API's single concatenated tensor."""
    from ibgs_b200 import synthetic as S
    import ibgs_b200.diff_plane_rasterization as dpr
    import ibgs_testutil as U
    sc = U.scene_to_device(S.make_scene("cfg1" if P is None else "tiny", P=P, sh_degree=deg))
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    cot = {k: v.cuda() for k, v in S.cotangents(sc).items()}
    rs = U.make_settings(dpr, sc, render_geo=True, depth_error_threshold=0.05)
    z = torch.zeros_like(sc["means3D"])
    common = dict(means3D=sc["means3D"], means2D=z, means2D_abs=z, opacities=sc["opacities"], scales=sc["scales"],
                  rotations=sc["rotations"], all_map=sc["all_map"])

    def run(**sh_kw):
        leaves = {k: v.clone().requires_grad_(True) for k, v in sh_kw.items()}
        res = dpr.GaussianRasterizer(rs)(**common, **leaves)
        torch.autograd.backward([res[0], res[2], res[3], res[5]], [cot["color"], cot["normal"], cot["depth"], cot["warped"]])
        return res, {k: v.grad for k, v in leaves.items()}

    r1, g1 = run(shs=sc["shs"])
    r2, g2 = run(shs=sc["shs"][:, :1].contiguous(), shs_rest=sc["shs"][:, 1:].contiguous())
    for a, b in zip(r1, r2):
        assert torch.equal(a, b)
    # the SH gradient is a per-Gaussian function of dL/dcolour, which the tile renderer sums with float atomics:
    # two runs agree to rounding, not bit for bit
    assert _rel(g2["shs"], g1["shs"][:, :1]) <= 1e-4 and _rel(g2["shs_rest"], g1["shs"][:, 1:]) <= 1e-4
    with pytest.raises(RuntimeError, match="DC"):
        dpr.GaussianRasterizer(rs)(**common, shs=sc["shs"], shs_rest=sc["shs"][:, 1:].contiguous())


def test_fused_prologue_without_sh_concat_end_to_end():
    from ibgs_b200 import fused, synthetic as S
    import ibgs_b200.diff_plane_rasterization as dpr
    import ibgs_testutil as U
    sc = U.scene_to_device(S.make_scene("cfg1"))
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    cot = {k: v.cuda() for k, v in S.cotangents(sc).items()}
    P = sc["P"]
    raw = dict(xyz=sc["means3D"], opacity_raw=torch.logit(sc["opacities"].clamp(1e-4, 1 - 1e-4)).view(P, 1),
               scaling_raw=sc["scales"].log(), rotation_raw=sc["rotations"] * 0.5,
               fdc=sc["shs"][:, :1].contiguous(), frest=sc["shs"][:, 1:].contiguous(),
               normal_raw=sc["normals_world"] * 2.0, offset=torch.zeros((P, 1), device="cuda"))
    V, cam = sc["viewmatrix"], sc["campos"]
    rs = U.make_settings(dpr, sc, render_geo=True)
    z = torch.zeros_like(sc["means3D"])
    out = []
    for concat in (True, False):
        leaves = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
        pro = fused.gaussian_prologue(*[leaves[k] for k in IN], V, cam, concat_sh=concat)
        if concat:
            opacity, scales, rotations, shs, all_map = pro
            sh_kw = dict(shs=shs)
        else:
            opacity, scales, rotations, all_map = pro
            sh_kw = dict(shs=leaves["fdc"], shs_rest=leaves["frest"])
        res = dpr.GaussianRasterizer(rs)(means3D=leaves["xyz"], means2D=z, means2D_abs=z, opacities=opacity,
                                         scales=scales, rotations=rotations, all_map=all_map, **sh_kw)
        torch.autograd.backward([res[0], res[2], res[3], res[5]], [cot["color"], cot["normal"], cot["depth"], cot["warped"]])
        out.append((res, {k: v.grad for k, v in leaves.items()}))
    for a, b in zip(out[0][0], out[1][0]):
        assert torch.equal(a, b)
    for k in IN:
        assert _rel(out[0][1][k], out[1][1][k]) <= 1e-4, k   # float atomics: run-to-run summation order


@pytest.mark.parametrize("P,K", [(1, 1), (1001, 9), (65_537, 16)])
def test_fused_prologue_smallest_axis_normal(P, K):
    """learnt_normal=False (render(..., learnt_normal=False), scene/gaussian_model.py:149-161): plane normal = shortest
    axis of the Gaussian.  Values and gradients against the reference's torch expressions (float32 autograd) and the
    float64 oracle; the chosen axis and the flip decision are discrete and must agree exactly."""
    from ibgs_b200 import fused
    from oracle import prologue_oracle as O
    IN6 = IN[:6]
    p = PR.random_params(P, K=K, seed=P + 1, device="cuda")
    leaves = {k: p[k].clone().requires_grad_(True) for k in IN6}
    outs = fused.gaussian_prologue(*[leaves[k] for k in IN6], None, None, p["V"], p["cam"], smallest_axis_normal=True)
    assert len(outs) == 5
    tl = {k: p[k].clone().requires_grad_(True) for k in IN6}
    touts = PR.torch_prologue_smallest_axis(*[tl[k] for k in IN6], p["V"], p["cam"])
    fw = O.forward_smallest_axis(*[p[k].cpu().numpy() for k in IN6], p["V"].cpu().numpy(), p["cam"].cpu().numpy())
    for n, o, t in zip(NAMES, outs, touts):
        assert _rel(o, t) <= 2e-6, (n, _rel(o, t))
        assert _rel(o, torch.from_numpy(fw[n]).cuda().view_as(o)) <= 2e-6, n
    g = torch.Generator().manual_seed(4)
    cots = [torch.randn(o.shape, generator=g).cuda() for o in outs]
    torch.autograd.backward(list(outs), cots)
    torch.autograd.backward(list(touts), cots)
    d = O.backward_smallest_axis(fw, *[c.cpu().numpy() for c in cots])
    name = dict(fdc="features_dc", frest="features_rest")
    for k in IN6:
        if leaves[k].numel() == 0:
            continue
        assert _rel(leaves[k].grad, tl[k].grad) <= 1e-5, (k, _rel(leaves[k].grad, tl[k].grad))
        want = torch.from_numpy(d[name.get(k, k)]).cuda().view_as(leaves[k].grad)
        assert _rel(leaves[k].grad, want) <= 1e-5, k
    with pytest.raises(ValueError):
        fused.gaussian_prologue(*[p[k] for k in IN], p["V"], p["cam"], smallest_axis_normal=True)
