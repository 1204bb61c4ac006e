"""GPU: the section-8f fast paths as a whole -- ibgs_b200.gaussian_renderer.render / render_depth and
ibgs_b200.color_aggregation.fuse_color with the fused losses -- against the reference's UNCHANGED glue
(gaussian_renderer.render, color_aggregation_network.fuse_color, utils.loss_utils) on the same synthetic world, both
running on this repo's rasterizer.  Output dict, losses and every parameter gradient of a training iteration."""
import pytest
import torch

import refglue as G
import ibgs_testutil as U
from ibgs_b200 import synthetic as S

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not G.available(), reason="reference glue not staged")]

FLOAT_KEYS = ("render", "rendered_normal", "median_intersected_depth", "cam_feat", "warped_image", "min_depth_diff",
              "camera_ray")


def _world(config="cfg1", **kw):
    g = G.bind("b200")
    w = G.build_world(g, config, n_views=6, sc_cpu=S.make_scene(config), **kw)
    G.prime_depth_cache(w)
    return w


@pytest.mark.parametrize("learnt_normal", [True, False])
def test_fast_render_equals_unchanged_render(learnt_normal):
    import ibgs_b200.gaussian_renderer as FR
    w = _world(learnt_normal=learnt_normal)
    cam = w.scene.getTrainCameras()[0]
    pkgs, grads = [], []
    for fn in (w.glue.render, FR.render):
        w.gaussians.optimizer.zero_grad(set_to_none=True)
        pkg = fn(cam, w.gaussians, w.scene, w.pipe, w.args, w.background, render_geo=True, return_depth_normal=True,
                 **G.render_kwargs(w))
        g = torch.Generator().manual_seed(11)
        loss = 0.0
        for k in ("render", "rendered_normal", "median_intersected_depth", "warped_image", "median_intersected_depth_normal"):
            loss = loss + (pkg[k] * torch.randn(pkg[k].shape, generator=g).to(pkg[k].device)).sum()
        loss.backward()
        pkgs.append(pkg)
        grads.append({n: (None if v is None else v.clone()) for n, v in G.gaussian_grads(w).items()})
    po, pr = pkgs[1], pkgs[0]
    assert torch.equal(po["radii"], pr["radii"]) and torch.equal(po["visibility_filter"], pr["visibility_filter"])
    # same rasterizer, inputs equal to float rounding: every map agrees to rounding except at isolated pixels where the
    # median-depth selection (a discontinuous function of the inputs: which Gaussian crosses T = 0.5) lands on a neighbour
    for k in FLOAT_KEYS:
        d = (po[k] - pr[k]).abs()
        assert d.median().item() <= 1e-6 and (d > 1e-4).float().mean().item() <= 1e-3, (k, d.max().item())
    d = (po["median_intersected_depth_normal"] - pr["median_intersected_depth_normal"]).abs()
    assert d.mean().item() <= 1e-5 and (d > 1e-3).float().mean().item() <= 1e-3
    for n in G.GAUSSIAN_PARAMS:
        if grads[0][n] is None:
            assert grads[1][n] is None or grads[1][n].abs().max().item() == 0.0, n
            continue
        assert U.rel_l2(grads[1][n], grads[0][n]) <= 5e-3, (n, U.rel_l2(grads[1][n], grads[0][n]))
    for k in ("viewspace_points", "viewspace_points_abs"):
        assert U.rel_l2(po[k].grad[:, :2], pr[k].grad[:, :2]) <= 5e-3, k


def test_fast_render_modes_and_test_time_path():
    import ibgs_b200.gaussian_renderer as FR
    w = _world()
    cam = w.scene.getTrainCameras()[2]
    o = w.opt
    with torch.no_grad():
        for mode in (dict(render_geo=False, return_depth_normal=False),
                     dict(render_geo=False, return_depth_normal=False, render_depth_only=True),
                     dict(render_geo=True, return_depth_normal=True, do_find_closest_frame=True, do_render_src_depth=True)):
            a = w.glue.render(cam, w.gaussians, w.scene, w.pipe, w.args, w.background, **mode, **G.render_kwargs(w))
            b = FR.render(cam, w.gaussians, w.scene, w.pipe, w.args, w.background, **mode, **G.render_kwargs(w))
            assert torch.equal(a["radii"], b["radii"]), mode
            for k in ("render", "median_intersected_depth"):
                d = (a[k] - b[k]).abs()
                assert d.median().item() <= 1e-6 and (d > 1e-4).float().mean().item() <= 1e-3, (mode, k, d.max().item())
            if mode.get("do_render_src_depth"):
                for k in ("warped_image", "cam_feat", "min_depth_diff"):
                    bad = ((a[k] - b[k]).abs() > 1e-4).float().mean().item()
                    assert bad <= 2e-3, (k, bad)
        d0 = w.glue.render_depth(cam, w.gaussians, w.scene, w.pipe, w.args, w.background, o.learnt_normal,
                                 o.number_src_frames, o.buffer_length, o.depth_error_threshold)
        d1 = FR.render_depth(cam, w.gaussians, w.scene, w.pipe, w.args, w.background, o.learnt_normal,
                             o.number_src_frames, o.buffer_length, o.depth_error_threshold)
        assert d0.shape == d1.shape and ((d0 - d1).abs() > 1e-4).float().mean().item() <= 1e-3
    w.pipe.convert_SHs_python = True
    with pytest.raises(NotImplementedError):
        FR.render(cam, w.gaussians, w.scene, w.pipe, w.args, w.background, **G.render_kwargs(w))


@pytest.mark.parametrize("exposure,precision", [(False, "fp32"), (True, "fp32"), (False, "bf16")])
def test_fast_train_iteration_equals_unchanged_glue(exposure, precision):
    """train.py:269-370 with every fast path on vs the unchanged glue: losses and gradients."""
    w = _world(exposure=exposure)
    res = []
    for fns in (None, G.fast_fns(precision=precision)):
        w.dp = None
        w.gaussians.optimizer.zero_grad(set_to_none=True)
        w.color_net.zero_grad(set_to_none=True)
        w.app_model.optimizer.zero_grad(set_to_none=True)
        out = G.train_iteration(w, 0, fns=fns)
        assert out["fusion"] is not None
        res.append((out, {n: v.clone() for n, v in G.gaussian_grads(w).items()},
                    [p.grad.clone() for p in w.color_net.parameters()]))
    (o0, g0, n0), (o1, g1, n1) = res
    ltol = 2e-4 if precision == "fp32" else 3e-3
    for k in ("loss", "image_loss", "normal_loss", "photometric_loss"):
        a, b = o1[k].item(), o0[k].item()
        assert abs(a - b) <= ltol * max(1.0, abs(b)), (k, a, b)
    gtol = 1e-2 if precision == "fp32" else 8e-2      # the reference's own convolutions run in tf32 (cuDNN default)
    for n in G.GAUSSIAN_PARAMS:
        assert U.rel_l2(g1[n], g0[n]) <= gtol, (n, U.rel_l2(g1[n], g0[n]))
    for a, b in zip(n1, n0):
        assert U.rel_l2(a, b) <= gtol, U.rel_l2(a, b)


@pytest.mark.parametrize("H,W", [(3, 3), (37, 53), (240, 321)])
def test_depth_normal_kernel_equals_torch_expressions(H, W):
    """csrc/depth_normal.cu against the same map written with torch ops (and through them against the reference's
    render_normal, compared in test_fast_render_equals_unchanged_render): values and the depth gradient."""
    import types
    import ibgs_b200.gaussian_renderer as FR
    cam = types.SimpleNamespace(Fx=310.5, Fy=305.25, Cx=W / 2 - 0.3, Cy=H / 2 + 0.4)
    g = torch.Generator().manual_seed(H + W)
    d0 = (2.0 + torch.rand(H, W, generator=g) + 0.2 * torch.randn(H, W, generator=g).cumsum(1) / W).cuda()
    d0[H // 2:H // 2 + 2, W // 2:W // 2 + 2] = 3.0          # a flat patch: zero cross products, the eps branches
    da, db = d0.clone().requires_grad_(True), d0.clone().requires_grad_(True)
    na, nb = FR.depth_normal(cam, da), FR.depth_normal_torch(cam, db)
    assert na.shape == (3, H, W) and (na - nb).abs().max().item() <= 3e-5   # differences of nearby depths cancel
    assert not na[:, 0].any() and not na[:, -1].any() and not na[:, :, 0].any() and not na[:, :, -1].any()
    cot = torch.randn(3, H, W, generator=g).cuda()
    na.backward(cot)
    nb.backward(cot)
    assert U.rel_l2(da.grad, db.grad) <= 1e-4, U.rel_l2(da.grad, db.grad)


def test_densification_stats_kernel_equals_reference_statements():
    """ibgs_b200.densify.add_densification_stats (one launch) against train.py:399-405 executed with the reference's own
    GaussianModel.add_densification_stats, over three views (statistics accumulate)."""
    from ibgs_b200.densify import add_densification_stats
    w = _world()
    gm = w.gaussians
    names = ("max_radii2D", "xyz_gradient_accum", "xyz_gradient_accum_abs", "denom", "denom_abs")
    gm.max_radii2D = torch.zeros((w.P,), device="cuda")       # what densification leaves behind (gaussian_model.py:212)
    base = {n: getattr(gm, n).clone() for n in names}
    outs = []
    for i in range(3):
        out = G.train_iteration(w, i)
        outs.append(out["render_pkg"])
        w.gaussians.optimizer.zero_grad(set_to_none=True)
    for pkg in outs:                                                     # the reference's statements
        G.densification_stats(w, dict(render_pkg=pkg))
    want = {n: getattr(gm, n).clone() for n in names}
    for n in names:
        getattr(gm, n).copy_(base[n])
    for pkg in outs:
        add_densification_stats(gm, pkg)
    for n in names:
        a, b = getattr(gm, n), want[n]
        assert torch.allclose(a, b, rtol=1e-6, atol=0), (n, (a - b).abs().max().item())
    assert torch.equal(gm.denom, want["denom"]) and torch.equal(gm.max_radii2D, want["max_radii2D"])
    assert gm.denom.max().item() == 3.0


def test_fast_train_iteration_at_full_size():
    """BASELINE config 2's size (500 k Gaussians at 1920x1080): the fast iteration against the unchanged glue -- losses to
    bf16 accuracy, every gradient in direction and magnitude."""
    w = _world("cfg2")
    res = []
    for fns in (None, G.fast_fns(precision="bf16")):
        w.gaussians.optimizer.zero_grad(set_to_none=True)
        w.color_net.zero_grad(set_to_none=True)
        out = G.train_iteration(w, 1, fns=fns)
        assert out["fusion"] is not None
        res.append((out, {n: v.clone() for n, v in G.gaussian_grads(w).items()}))
    (o0, g0), (o1, g1) = res
    for k in ("loss", "image_loss", "normal_loss", "photometric_loss"):
        a, b = o1[k].item(), o0[k].item()
        assert abs(a - b) <= 3e-3 * max(1.0, abs(b)), (k, a, b)
    for n in G.GAUSSIAN_PARAMS:
        assert U.rel_l2(g1[n], g0[n]) <= 8e-2, (n, U.rel_l2(g1[n], g0[n]))
