"""GPU: the host-buffer entry points of the C ABI (`ibgs_forward_h`, `ibgs_dist2_h`, include/ibgs_b200.h) -- what a
caller without a device allocator binds -- called through ctypes on plain numpy arrays and compared bit for bit with
the device entry points the Python package uses."""
import ctypes as C

import numpy as np
import pytest
import torch

from ibgs_b200 import synthetic as S
import ibgs_testutil as U

pytestmark = pytest.mark.gpu


def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy())


@pytest.mark.parametrize("geo", [True, False])
def test_forward_h_matches_device_path(geo):
    import ibgs_b200.diff_plane_rasterization as dpr
    from ibgs_b200 import _native as N
    sc = U.scene_to_device(S.make_scene("tiny"))
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    want, _, state = U.ours_forward_backward(dpr, sc, None, render_geo=geo, depth_error_threshold=0.05)
    P, H, W, nb = sc["P"], sc["H"], sc["W"], sc["nb_src"]
    host = {k: _np(sc[k]) for k in ("means3D", "shs", "opacities", "scales", "rotations", "all_map", "bg", "viewmatrix",
                                    "projmatrix", "campos", "ref_to_src_list", "src_cam_pos", "src_images",
                                    "src_rendered_depths")}
    outs = dict(color=np.zeros((3, H, W), np.float32), radii=np.zeros((P,), np.int32),
                normal=np.zeros((3, H, W), np.float32), depth=np.zeros((1, H, W), np.float32),
                cam_feat=np.zeros((20, H, W), np.float32), warped=np.zeros((15, H, W), np.float32),
                min_depth_diff=np.zeros((1, H, W), np.float32), camera_ray=np.zeros((3, H, W), np.float32),
                mask=np.zeros((1, H, W), np.int32))
    a = N.IbgsForwardArgs()
    a.P = P
    v = a.view
    v.image_height, v.image_width = H, W
    v.tanfovx, v.tanfovy, v.scale_modifier = sc["tanfovx"], sc["tanfovy"], 1.0
    v.sh_degree, v.sh_coeffs = sc["sh_degree"], host["shs"].shape[1]
    v.nb_src_images, v.buffer_length, v.depth_error_threshold = (nb if geo else 0), 4, 0.05
    v.render_geo, v.render_depth_only, v.prefiltered, v.debug = int(geo), 0, 0, 0
    ptr = lambda x: x.ctypes.data_as(C.c_void_p)
    for k in ("bg", "viewmatrix", "projmatrix", "campos"):
        setattr(v, k, ptr(host[k]))
    if geo:
        for k in ("ref_to_src_list", "src_cam_pos", "src_images", "src_rendered_depths"):
            setattr(v, k, ptr(host[k]))
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        setattr(a, k, ptr(host[k]))
    a.all_map = ptr(host["all_map"]) if geo else None
    a.out_color, a.radii, a.out_normal_map = ptr(outs["color"]), ptr(outs["radii"]), ptr(outs["normal"])
    a.out_median_intersected_depth, a.out_cam_feat = ptr(outs["depth"]), ptr(outs["cam_feat"])
    a.out_warped_image, a.out_min_depth_diff = ptr(outs["warped"]), ptr(outs["min_depth_diff"])
    a.out_camera_ray, a.out_use_first_src_frame = ptr(outs["camera_ray"]), ptr(outs["mask"])
    R = N.lib.ibgs_forward_h(C.byref(a))
    assert R == state["num_rendered"], (R, N.last_error())
    for k, arr in outs.items():
        assert np.array_equal(arr, _np(want[k])), k


def test_dist2_h_matches_device_path():
    from ibgs_b200 import _native as N
    from ibgs_b200.simple_knn._C import distCUDA2
    pts = torch.randn((20_011, 3), generator=torch.Generator().manual_seed(9)) * 2.0
    want = distCUDA2(pts.cuda()).cpu().numpy()
    host = np.ascontiguousarray(pts.numpy())
    out = np.zeros((pts.shape[0],), np.float32)
    rc = N.lib.ibgs_dist2_h(pts.shape[0], host.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rc == 0, N.last_error()
    assert np.array_equal(out, want)
    # error path: NULL output
    assert N.lib.ibgs_dist2_h(5, host.ctypes.data_as(C.c_void_p), None) < 0


def _host_forward_args(N, sc, host, outs, geo):
    P, H, W, nb = sc["P"], sc["H"], sc["W"], sc["nb_src"]
    a = N.IbgsForwardArgs()
    a.P = P
    v = a.view
    v.image_height, v.image_width = H, W
    v.tanfovx, v.tanfovy, v.scale_modifier = sc["tanfovx"], sc["tanfovy"], 1.0
    v.sh_degree, v.sh_coeffs = sc["sh_degree"], host["shs"].shape[1]
    v.nb_src_images, v.buffer_length, v.depth_error_threshold = (nb if geo else 0), 4, 0.05
    v.render_geo, v.render_depth_only, v.prefiltered, v.debug = int(geo), 0, 0, 0
    ptr = lambda x: x.ctypes.data_as(C.c_void_p)
    for k in ("bg", "viewmatrix", "projmatrix", "campos"):
        setattr(v, k, ptr(host[k]))
    if geo:
        for k in ("ref_to_src_list", "src_cam_pos", "src_images", "src_rendered_depths"):
            setattr(v, k, ptr(host[k]))
    for k in ("means3D", "shs", "opacities", "scales", "rotations"):
        setattr(a, k, ptr(host[k]))
    a.all_map = ptr(host["all_map"]) if geo else None
    a.out_color, a.radii, a.out_normal_map = ptr(outs["color"]), ptr(outs["radii"]), ptr(outs["normal"])
    a.out_median_intersected_depth, a.out_cam_feat = ptr(outs["depth"]), ptr(outs["cam_feat"])
    a.out_warped_image, a.out_min_depth_diff = ptr(outs["warped"]), ptr(outs["min_depth_diff"])
    a.out_camera_ray, a.out_use_first_src_frame = ptr(outs["camera_ray"]), ptr(outs["mask"])
    return a


@pytest.mark.parametrize("geo", [True, False])
def test_forward_backward_h_matches_device_path(geo):
    """One training view through host buffers only (`ibgs_forward_backward_h`): outputs bit-equal to the device path,
    gradients equal up to the order of the float atomics."""
    import ibgs_b200.diff_plane_rasterization as dpr
    from ibgs_b200 import _native as N
    sc = U.scene_to_device(S.make_scene("tiny"))
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    cot = {k: v.cuda() for k, v in S.cotangents(sc).items()}
    want, wgrads, state = U.ours_forward_backward(dpr, sc, cot, render_geo=geo, depth_error_threshold=0.05)
    P, H, W = sc["P"], sc["H"], sc["W"]
    host = {k: _np(sc[k]) for k in ("means3D", "shs", "opacities", "scales", "rotations", "all_map", "bg", "viewmatrix",
                                    "projmatrix", "campos", "ref_to_src_list", "src_cam_pos", "src_images",
                                    "src_rendered_depths")}
    outs = dict(color=np.zeros((3, H, W), np.float32), radii=np.zeros((P,), np.int32),
                normal=np.zeros((3, H, W), np.float32), depth=np.zeros((1, H, W), np.float32),
                cam_feat=np.zeros((20, H, W), np.float32), warped=np.zeros((15, H, W), np.float32),
                min_depth_diff=np.zeros((1, H, W), np.float32), camera_ray=np.zeros((3, H, W), np.float32),
                mask=np.zeros((1, H, W), np.int32))
    a = _host_forward_args(N, sc, host, outs, geo)
    K = host["shs"].shape[1]
    hc = {k: _np(cot[k]) for k in ("color", "normal", "depth", "warped")}
    g = dict(means3D=np.zeros((P, 3), np.float32), means2D=np.zeros((P, 3), np.float32), means2D_abs=np.zeros((P, 3), np.float32),
             colors=np.zeros((P, 3), np.float32), opacity=np.zeros((P, 1), np.float32), sh=np.zeros((P, K, 3), np.float32),
             scales=np.zeros((P, 3), np.float32), rotations=np.zeros((P, 4), np.float32), all_map=np.zeros((P, 5), np.float32))
    b = N.IbgsBackwardArgs()
    ptr = lambda x: x.ctypes.data_as(C.c_void_p)
    b.dL_dout_color = ptr(hc["color"])
    if geo:
        b.dL_dout_normal_map, b.dL_dout_median_intersected_depth = ptr(hc["normal"]), ptr(hc["depth"])
        b.dL_dout_warped_image = ptr(hc["warped"])
    b.dL_dmeans3D, b.dL_dmeans2D, b.dL_dmeans2D_abs = ptr(g["means3D"]), ptr(g["means2D"]), ptr(g["means2D_abs"])
    b.dL_dcolors, b.dL_dopacity, b.dL_dsh = ptr(g["colors"]), ptr(g["opacity"]), ptr(g["sh"])
    b.dL_dscales, b.dL_drotations, b.dL_dall_map = ptr(g["scales"]), ptr(g["rotations"]), ptr(g["all_map"])
    R = N.lib.ibgs_forward_backward_h(C.byref(a), C.byref(b))
    assert R == state["num_rendered"], (R, N.last_error())
    for k, arr in outs.items():
        assert np.array_equal(arr, _np(want[k])), k
    pairs = dict(means3D="means3D", means2D="means2D", means2D_abs="means2D_abs", sh="sh", opacity="opacities", scales="scales",
                 rotations="rotations")
    if geo:
        pairs["all_map"] = "all_map"
    for hk, wk in pairs.items():
        w = wgrads[wk]
        assert w is not None, wk
        got = torch.from_numpy(g[hk]).cuda().view_as(w)
        assert U.rel_l2(got, w) <= 1e-5, (hk, U.rel_l2(got, w))
    # error path: no colour cotangent
    b.dL_dout_color = None
    assert N.lib.ibgs_forward_backward_h(C.byref(a), C.byref(b)) < 0
