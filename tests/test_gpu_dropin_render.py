"""The drop-in claim, tested: the reference's UNCHANGED `gaussian_renderer.render()` / `render_depth()`
(gaussian_renderer/__init__.py:143-365, :41-140), `scene.GaussianModel`, `Camera`, `Scene` buffers, losses and
`color_aggregation_network.fuse_color` run twice on the same synthetic world -- once with `diff_plane_rasterization`
bound to this repo (shims/ -> ibgs_b200 -> C ABI), once bound to the reference's own wrapper + unmodified CUDA
extension (oracle/_ref) -- and the whole output dict and the GaussianModel parameter gradients are compared.

Gates: integer outputs exact; float maps <= 1e-4 max-abs; gradients <= 1e-3 relative L2.
The staged reference files come from oracle/stage_ref_py.py (baseline/_ref/py, git-ignored, shipped by gpurun).
"""
import pytest
import torch

import refglue as G
import ibgs_testutil as U
from ibgs_b200 import synthetic as S

pytestmark = pytest.mark.gpu

FLOAT_KEYS = ("render", "rendered_normal", "median_intersected_depth", "cam_feat", "warped_image", "min_depth_diff",
              "camera_ray")


@pytest.fixture(scope="module")
def glues():
    if not G.reference_available():
        pytest.skip("baseline/_ref/py or oracle/_ref/dpr not staged (run __graft_entry__.build() where /root/reference exists)")
    gl = (G.bind("b200"), G.bind("reference"))
    # One throw-away priming pass per binding before anything is compared.  The comparisons below are exact (the outputs
    # of the two rasterizers are bit-identical on this scene), so they rely on the torch glue in front of them --
    # get_normal, the `global_normal @ world_view_transform` product, torch.inverse ... -- producing bit-identical inputs in
    # both worlds.  On a fresh process the FIRST execution of that glue occasionally differs in the last bit from every
    # later one (library first-call algorithm selection; seen on ~1 fresh box in 10: a handful of median-depth selections
    # flip, tools/flake_hunt.py), which is a property of neither rasterizer.
    for g in gl:
        w = G.build_world(g, "cfg1", n_views=6, sc_cpu=S.make_scene("cfg1"))
        G.prime_depth_cache(w)
        del w
    torch.cuda.synchronize()
    return gl


def _worlds(glues, config="cfg1", n_views=6, **kw):
    sc_cpu = S.make_scene(config)
    return [G.build_world(g, config, n_views=n_views, sc_cpu=sc_cpu, **kw) for g in glues]


def _share_depth_cache(wo, wr):
    """Each binding primes its own cache (train.py:242-256); the caches are compared, then the reference's is handed
    to both worlds so that every depth-consistency decision downstream sees identical inputs."""
    G.prime_depth_cache(wo)
    G.prime_depth_cache(wr)
    err = (wo.scene.rendered_depth_list - wr.scene.rendered_depth_list).abs().max().item()
    # (if this ever fails: were the two worlds' camera matrices -- torch bmm / inverse at Camera construction -- bit-equal?)
    cams = [all(torch.equal(getattr(a, k), getattr(b, k)) for k in ("world_view_transform", "full_proj_transform", "camera_center"))
            for a, b in zip(wo.scene.getTrainCameras(), wr.scene.getTrainCameras())]
    assert err <= 1e-4, f"rendered_depth_list: {err}; camera matrices bit-equal per view: {cams}"
    wo.scene.rendered_depth_list.copy_(wr.scene.rendered_depth_list)


def _describe(po, pr):
    """One line per output for a failing comparison: which side holds non-finite values, how many entries differ."""
    lines = []
    for k in ("radii", "visibility_filter", "use_first_src_frame_mask") + FLOAT_KEYS + ("median_intersected_depth_normal",):
        a, b = po.get(k), pr.get(k)
        if a is None or b is None:
            lines.append(f"{k}: ours {None if a is None else tuple(a.shape)} / reference {None if b is None else tuple(b.shape)}")
            continue
        af, bf = a.float(), b.float()
        d = (af - bf).abs()
        lines.append(f"{k}: max|d| {d.max().item():.3e}  differing entries {(d > 1e-4).float().mean().item():.2e}  "
                     f"non-finite ours {(~torch.isfinite(af)).sum().item()} / reference {(~torch.isfinite(bf)).sum().item()}")
    return "\n".join(lines)


def _compare_pkg(po, pr, tol=1e-4):
    try:
        _compare_pkg_checks(po, pr, tol)
    except AssertionError as ex:
        raise AssertionError(f"{ex}\n--- all outputs, this repo vs reference ---\n{_describe(po, pr)}") from None


def _compare_pkg_checks(po, pr, tol=1e-4):
    assert torch.equal(po["radii"], pr["radii"])
    assert torch.equal(po["visibility_filter"], pr["visibility_filter"])
    assert torch.equal(po["use_first_src_frame_mask"], pr["use_first_src_frame_mask"])
    for k in FLOAT_KEYS:
        if pr[k] is None:
            assert po[k] is None, k
            continue
        assert po[k].shape == pr[k].shape and po[k].dtype == pr[k].dtype, k
        err = (po[k] - pr[k]).abs().max().item()
        assert err <= tol, f"{k}: max-abs {err}"
    if pr["median_intersected_depth_normal"] is not None:
        # derived by torch from the depth map (normalised cross products): compare where the depth is smooth enough
        # that 1e-4 depth noise cannot flip a normal -- mean error instead of max
        d = (po["median_intersected_depth_normal"] - pr["median_intersected_depth_normal"]).abs()
        assert d.mean().item() <= 1e-4, d.mean().item()
    if pr["app_image"] is not None:
        assert (po["app_image"] - pr["app_image"]).abs().max().item() <= tol


def _cotangent_loss(pkg, seed=11):
    g = torch.Generator().manual_seed(seed)
    loss = 0.0
    for k in ("render", "rendered_normal", "median_intersected_depth", "warped_image"):
        c = torch.randn(pkg[k].shape, generator=g).to(pkg[k].device)
        loss = loss + (pkg[k] * c).sum()
    return loss


@pytest.mark.parametrize("learnt_normal", [True, False])
def test_render_unchanged_glue_vs_reference(glues, learnt_normal):
    wo, wr = _worlds(glues, learnt_normal=learnt_normal)
    _share_depth_cache(wo, wr)
    pkgs = []
    for w in (wo, wr):
        cam = w.scene.getTrainCameras()[0]
        assert len(cam.nearest_id) >= 4, cam.nearest_id          # the reference's own neighbour selection ran
        pkg = w.glue.render(cam, w.gaussians, w.scene, w.pipe, w.args, w.background, render_geo=True,
                            return_depth_normal=True, **G.render_kwargs(w))
        _cotangent_loss(pkg).backward()
        pkgs.append(pkg)
    _compare_pkg(*pkgs)
    assert pkgs[0]["warped_image"].abs().sum().item() > 0, "no source view ever passed the depth-consistency test"
    go, gr = G.gaussian_grads(wo), G.gaussian_grads(wr)
    for n in G.GAUSSIAN_PARAMS:
        if gr[n] is None:
            assert go[n] is None or go[n].abs().max().item() == 0.0, n
            continue
        e = U.rel_l2(go[n], gr[n])
        assert e <= 1e-3, f"grad {n}: rel-L2 {e}"
    for k in ("viewspace_points", "viewspace_points_abs"):      # densification statistics read these (.grad[:, :2])
        e = U.rel_l2(pkgs[0][k].grad[:, :2], pkgs[1][k].grad[:, :2])
        assert e <= 1e-3, f"{k}.grad: rel-L2 {e}"


def test_render_colour_only_and_depth_only_modes(glues):
    """render_geo=False (the first 7000 - 2*Nv iterations, train.py:291) and render_depth_only=True."""
    wo, wr = _worlds(glues)
    for mode in (dict(render_geo=False, return_depth_normal=False),
                 dict(render_geo=False, return_depth_normal=False, render_depth_only=True)):
        pkgs = []
        for w in (wo, wr):
            cam = w.scene.getTrainCameras()[1]
            with torch.no_grad():
                pkgs.append(w.glue.render(cam, w.gaussians, w.scene, w.pipe, w.args, w.background,
                                          **mode, **G.render_kwargs(w)))
        assert torch.equal(pkgs[0]["radii"], pkgs[1]["radii"])
        for k in ("render", "median_intersected_depth"):
            err = (pkgs[0][k] - pkgs[1][k]).abs().max().item()
            assert err <= 1e-4, f"{mode} {k}: {err}"


def test_render_depth_and_test_time_path(glues):
    """render_depth() directly, then the test-time call of render.py:132 -- do_find_closest_frame + do_render_src_depth:
    neighbour search at call time and four nested depth-only renders per view."""
    wo, wr = _worlds(glues)
    outs = []
    for w in (wo, wr):
        cam = w.scene.getTrainCameras()[2]
        o = w.opt
        with torch.no_grad():
            d = w.glue.render_depth(cam, w.gaussians, w.scene, w.pipe, w.args, w.background, o.learnt_normal,
                                    o.number_src_frames, o.buffer_length, o.depth_error_threshold)
            pkg = w.glue.render(cam, w.gaussians, w.scene, w.pipe, w.args, w.background, render_geo=True,
                                return_depth_normal=True, do_find_closest_frame=True, do_render_src_depth=True,
                                **G.render_kwargs(w))
        outs.append((d, pkg))
    assert (outs[0][0] - outs[1][0]).abs().max().item() <= 1e-4
    # the nested source depths differ by <= 1e-4 between the bindings, so a few pixels sit on the other side of the
    # depth-consistency threshold: masks / warps are compared by mismatch fraction here, exactly elsewhere
    po, pr = outs[0][1], outs[1][1]
    assert torch.equal(po["radii"], pr["radii"])
    for k in ("render", "rendered_normal", "median_intersected_depth", "camera_ray"):
        assert (po[k] - pr[k]).abs().max().item() <= 1e-4, k
    for k in ("warped_image", "cam_feat", "min_depth_diff"):
        bad = ((po[k] - pr[k]).abs() > 1e-4).float().mean().item()
        assert bad <= 2e-3, f"{k}: {bad} of the values differ"


@pytest.mark.parametrize("exposure", [False, True])
def test_train_iteration_unchanged_glue_vs_reference(glues, exposure):
    """One whole training iteration (train.py:269-370) incl. fuse_color + ColorFusionResidualNet (+ AppModel affine and
    the exposure lstsq with exposure=True, BASELINE config 4) under both bindings: losses and gradients agree."""
    wo, wr = _worlds(glues, exposure=exposure)
    _share_depth_cache(wo, wr)
    res = []
    for w in (wo, wr):
        out = G.train_iteration(w, 0)
        assert out["fusion"] is not None, "colour aggregation did not run"
        G.densification_stats(w, out)
        res.append(out)
    for k in ("loss", "image_loss", "normal_loss", "photometric_loss"):
        a, b = res[0][k].item(), res[1][k].item()
        assert abs(a - b) <= 2e-5 + 1e-4 * abs(b), f"{k}: {a} vs {b}"
    go, gr = G.gaussian_grads(wo), G.gaussian_grads(wr)
    for n in G.GAUSSIAN_PARAMS:
        e = U.rel_l2(go[n], gr[n])
        assert e <= 2e-3, f"grad {n}: rel-L2 {e}"       # cuDNN convolutions in the loss are not run-to-run exact
    for (no, po), (_, pr) in zip(wo.color_net.named_parameters(), wr.color_net.named_parameters()):
        e = U.rel_l2(po.grad, pr.grad)
        assert e <= 2e-3, f"colour net grad {no}: rel-L2 {e}"
    assert U.rel_l2(wo.gaussians.xyz_gradient_accum, wr.gaussians.xyz_gradient_accum) <= 1e-3
    assert torch.equal(wo.gaussians.max_radii2D, wr.gaussians.max_radii2D)
    # and the optimiser step of train.py:421-430 leaves both worlds with the same parameters
    for w in (wo, wr):
        G.optimizer_step(w)
    for n in G.GAUSSIAN_PARAMS:
        e = U.rel_l2(getattr(wo.gaussians, n).data, getattr(wr.gaussians, n).data)
        assert e <= 1e-4, f"param {n} after the step: rel-L2 {e}"
