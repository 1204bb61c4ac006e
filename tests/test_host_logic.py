"""CPU: host-side mirror of the reference's Python API (names, field order, error behaviour)."""
import pytest
import torch

import ibgs_b200
import ibgs_b200.diff_plane_rasterization as dpr

# reference: submodules/diff-plane-rasterization/diff_plane_rasterization/__init__.py:252-276
REFERENCE_SETTINGS_FIELDS = (
    "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
    "ref_to_src_list", "src_cam_pos", "src_images", "src_rendered_depths", "nb_src_images", "buffer_length",
    "depth_error_threshold", "sh_degree", "campos", "prefiltered", "render_geo", "render_depth_only", "debug")


def _settings(**kw):
    d = dict(image_height=8, image_width=8, tanfovx=0.5, tanfovy=0.5, bg=torch.zeros(3), scale_modifier=1.0,
             viewmatrix=torch.eye(4), projmatrix=torch.eye(4), ref_to_src_list=torch.zeros(1, 16),
             src_cam_pos=torch.zeros(1, 3), src_images=torch.zeros(1, 3, 64), src_rendered_depths=torch.zeros(1, 1, 64),
             nb_src_images=1, buffer_length=4, depth_error_threshold=0.01, sh_degree=2, campos=torch.zeros(3),
             prefiltered=False, render_geo=False, render_depth_only=False, debug=False)
    d.update(kw)
    return dpr.GaussianRasterizationSettings(**d)


def test_settings_fields_match_reference_order():
    assert dpr.GaussianRasterizationSettings._fields == REFERENCE_SETTINGS_FIELDS
    assert isinstance(dpr.GaussianRasterizer(_settings()), torch.nn.Module)


def test_forward_argument_validation_messages_match_reference():
    r = dpr.GaussianRasterizer(_settings())
    z = torch.zeros(4, 3)
    # reference __init__.py:298-302
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(z, z, z, torch.zeros(4, 1), scales=z, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="Please provide excatly one of either SHs or precomputed colors!"):
        r(z, z, z, torch.zeros(4, 1), shs=torch.zeros(4, 9, 3), colors_precomp=z, scales=z, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(z, z, z, torch.zeros(4, 1), shs=torch.zeros(4, 9, 3))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(z, z, z, torch.zeros(4, 1), shs=torch.zeros(4, 9, 3), scales=z, rotations=torch.zeros(4, 4),
          cov3D_precomp=torch.zeros(4, 6))


def test_bad_shape_and_cpu_tensors_raise():
    r = dpr.GaussianRasterizer(_settings())
    with pytest.raises(RuntimeError, match=r"means3D must have dimensions \(num_points, 3\)"):   # rasterize_points.cu:69-71
        r(torch.zeros(4, 2), torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(4, 1), shs=torch.zeros(4, 9, 3),
          scales=torch.zeros(4, 3), rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        z = torch.zeros(4, 3)
        r(z, z, z, torch.zeros(4, 1), shs=torch.zeros(4, 9, 3), scales=z, rotations=torch.zeros(4, 4))
    from ibgs_b200.simple_knn._C import distCUDA2
    with pytest.raises(RuntimeError, match="no CPU path"):
        distCUDA2(torch.zeros(10, 3))


def test_install_dropin_registers_reference_module_names():
    import sys
    saved = {k: sys.modules.get(k) for k in ("diff_plane_rasterization", "simple_knn", "simple_knn._C")}
    try:
        ibgs_b200.install_dropin()
        from diff_plane_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401
        from simple_knn._C import distCUDA2  # noqa: F401  (scene/gaussian_model.py:20)
        assert GaussianRasterizer is dpr.GaussianRasterizer
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_synthetic_scene_is_deterministic_and_consistent():
    from ibgs_b200 import synthetic as S
    a, b = S.make_scene("tiny"), S.make_scene("tiny")
    for k in ("means3D", "scales", "rotations", "opacities", "shs", "all_map", "src_images", "ref_to_src_list"):
        assert torch.equal(a[k], b[k]), k
    # viewmatrix is the transpose of W2C; projecting a point in front of the camera lands inside NDC
    w2c = a["w2c"]
    assert torch.allclose(a["viewmatrix"], w2c.T)
    p = torch.linalg.inv(w2c) @ torch.tensor([0.1, -0.2, 4.0, 1.0])
    hom = p @ a["projmatrix"]
    ndc = hom[:2] / hom[3]
    assert ndc.abs().max() < 1.0
    # all_map: unit normals facing the camera, positive distance column
    am = a["all_map"]
    assert torch.allclose(am[:, :3].norm(dim=1), torch.ones(am.shape[0]), atol=1e-5)
    assert (am[:, 3] == 1).all() and (am[:, 4] >= 0).all()
    # ref_to_src maps the reference camera centre to -R t of the relative pose: consistent with src_cam_pos
    c_ref_world = torch.linalg.inv(w2c)[:3, 3]
    for i in range(a["nb_src"]):
        w2s = a["src_w2c"][i]
        assert torch.allclose(a["ref_to_src_list"][i], w2s @ torch.linalg.inv(w2c), atol=1e-5)
        assert torch.allclose(torch.linalg.inv(w2s)[:3, 3], a["src_cam_pos"][i], atol=1e-5)
    assert c_ref_world.shape == (3,)


def test_shim_directory_provides_reference_import_names():
    """with shims/ on PYTHONPATH the reference's import lines work unchanged."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "shims"), root]))
    code = ("from diff_plane_rasterization import GaussianRasterizationSettings, GaussianRasterizer\n"
            "from simple_knn._C import distCUDA2\n"
            "import ibgs_b200.diff_plane_rasterization as d\n"
            "assert GaussianRasterizer is d.GaussianRasterizer\n"
            "assert GaussianRasterizationSettings._fields[:4] == ('image_height', 'image_width', 'tanfovx', 'tanfovy')\n"
            "print('ok')")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_python_layer_helpers_exist():
    """The wrapper's helpers that only GPU calls exercise must at least exist (guards against an edit removing them)."""
    import ibgs_b200.diff_plane_rasterization as d
    import ibgs_b200.depth_batch as db
    import ibgs_b200.loss_utils as lu
    import ibgs_b200.optim as op
    import ibgs_b200.fused as fu
    for mod, names in ((d, ("_ptr", "_f32c", "_Allocator", "_accumulable", "_fill_view", "rasterize_gaussians")),
                       (db, ("render_depth_batch", "render_depth_views", "DepthBatchSettings")),
                       (lu, ("ssim", "compute_photometric_ssim", "ssim2", "ssim_map")),
                       (op, ("ArenaAdam",)), (fu, ("gaussian_prologue",))):
        for n in names:
            assert hasattr(mod, n), (mod.__name__, n)
    import torch
    assert d._ptr(None) is None and d._ptr(torch.empty(0)) is None
    assert d._accumulable(torch.zeros(3)) is False


def test_fast_paths_fail_loudly_without_cuda_tensors():
    """No CPU / PyTorch fallback anywhere: every optional fast path raises on CPU tensors instead of computing."""
    import pytest
    import torch
    import ibgs_b200.depth_batch as db
    import ibgs_b200.loss_utils as lu
    import ibgs_b200.fused as fu
    from ibgs_b200.optim import ArenaAdam
    x = torch.rand((3, 16, 16))
    with pytest.raises(RuntimeError):
        lu.ssim(x, x)
    with pytest.raises(NotImplementedError):
        lu.ssim(x, x, window_size=7)
    with pytest.raises(RuntimeError):
        ArenaAdam({"w": torch.zeros(4)}, {"w": 0.1})
    with pytest.raises(ValueError):
        ArenaAdam({}, {})
    P = 5
    st = db.DepthBatchSettings(8, 8, 0.5, 0.5, 1.0, torch.eye(4)[None], torch.eye(4)[None], 4)
    with pytest.raises(RuntimeError):
        db.render_depth_batch(st, torch.zeros((P, 3)), torch.zeros((P, 1)), scales=torch.zeros((P, 3)),
                              rotations=torch.zeros((P, 4)), all_maps=torch.zeros((1, P, 5)))
    with pytest.raises(Exception):   # neither all_maps nor normals
        db.render_depth_batch(st, torch.zeros((P, 3)), torch.zeros((P, 1)), scales=torch.zeros((P, 3)),
                              rotations=torch.zeros((P, 4)))
    with pytest.raises(RuntimeError):
        fu.gaussian_prologue(torch.zeros((P, 3)), torch.zeros((P, 1)), torch.zeros((P, 3)), torch.zeros((P, 4)),
                             torch.zeros((P, 1, 3)), torch.zeros((P, 8, 3)))


def test_kernel_variant_knobs_validate_their_argument():
    from ibgs_b200 import _native as N
    for fn in (N.lib.ibgs_set_backward_variant, N.lib.ibgs_set_forward_variant):
        assert fn(7) < 0 and "pixels_per_lane" in N.last_error()
        assert fn(2) == 0 and fn(0) == 0


def test_round2_fast_paths_fail_loudly_without_cuda_tensors_and_validate_arguments():
    """The colour-aggregation / render / densification fast paths: no CPU fallback, argument errors before any launch."""
    import types
    import pytest
    import torch
    import ibgs_b200.color_aggregation as ca
    import ibgs_b200.densify as dn
    import ibgs_b200.gaussian_renderer as gr
    from ibgs_b200 import _native as N
    H, W = 8, 8
    mlp = torch.nn.Sequential(torch.nn.Linear(7, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU())
    with pytest.raises(RuntimeError, match="no CPU path"):
        ca.color_features(torch.zeros(9, H, W), torch.zeros(12, H, W), torch.zeros(3, H, W), torch.zeros(3, H, W), mlp, 3)
    net = types.SimpleNamespace(per_view_feat_dim=32, feat_aggregate_mode="mean", per_view_mlp=mlp, conv_decoder=None)
    opts = types.SimpleNamespace(residual_resolution_scale=1.0, enable_exposure_correction=False, nb_visible_src_frames=3)
    pkg = {}
    with pytest.raises(ValueError):
        ca.fuse_color(pkg, net, None, None, None, 1, opts, precision="fp8")
    opts.residual_resolution_scale = 0.5
    with pytest.raises(NotImplementedError):
        ca.fuse_color(pkg, net, None, None, None, 1, opts)
    opts.residual_resolution_scale = 1.0
    net.per_view_feat_dim = 16
    with pytest.raises(NotImplementedError):
        ca.fuse_color(pkg, net, None, None, None, 1, opts)
    assert ca.fuse_color(pkg, None, None, None, None, 1, opts) is None
    for fn in (ca.max_pool2, lambda t: ca.upsample_nearest(t, (4, 4))):
        with pytest.raises(RuntimeError):
            fn(torch.zeros(1, 8, 4, 4))                     # CPU tensor
    with pytest.raises(RuntimeError, match="no CPU path"):
        gr.depth_normal(types.SimpleNamespace(Fx=1.0, Fy=1.0, Cx=0.0, Cy=0.0), torch.ones(4, 4))
    with pytest.raises(NotImplementedError):
        gr.render(None, None, None, types.SimpleNamespace(convert_SHs_python=True, compute_cov3D_python=False), None, None,
                  True, 4, 4)
    gm = types.SimpleNamespace()
    leaf = torch.zeros(4, 3, requires_grad=True)
    with pytest.raises(RuntimeError, match="after loss.backward"):
        dn.add_densification_stats(gm, dict(radii=torch.zeros(4, dtype=torch.int32), viewspace_points=leaf,
                                            viewspace_points_abs=leaf))
    # C ABI argument checks of the new entry points (no launch happens)
    a = N.IbgsColorFeatArgs()
    a.height, a.width, a.n_views, a.channel_pitch = 4, 4, 9, 40
    assert N.lib.ibgs_color_features_forward(a, None) < 0 and "n_views" in N.last_error()
    a.n_views, a.channel_pitch = 3, 38
    assert N.lib.ibgs_color_features_forward(a, None) < 0 and "channel_pitch" in N.last_error()
    assert N.lib.ibgs_nhwc_maxpool2_forward(None, None, None, 4, 4, 7, 1, None) < 0
    assert N.lib.ibgs_depth_normal_forward(None, None, 4, 4, 1.0, 1.0, 0.0, 0.0, None) < 0
    assert N.lib.ibgs_densification_stats(-1, None, None, None, None, None, None, None, None, None) < 0
    assert N.lib.ibgs_forward_backward_h(None, None) < 0
