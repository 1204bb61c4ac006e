"""ctypes binding of libibgs_b200.so (the C ABI in include/ibgs_b200.h).

There is NO fallback: if the shared library is missing or does not load, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IBGS_B200_LIB selects another build of the SAME library (kernel experiments, tools/build_variant.py); there is still no fallback
LIB_PATH = os.environ.get("IBGS_B200_LIB") or os.path.join(_HERE, "_lib", "libibgs_b200.so")

IBGS_BUF_GEOM, IBGS_BUF_BINNING, IBGS_BUF_IMAGE, IBGS_BUF_SCRATCH = 0, 1, 2, 3
MAX_SRC = 5
MAX_BUFFER_LENGTH = 8

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_size_t)

_fp = C.c_void_p  # all data pointers travel as raw addresses


class IbgsView(C.Structure):
    _fields_ = [
        ("image_height", C.c_int32), ("image_width", C.c_int32),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
        ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32),
        ("nb_src_images", C.c_int32), ("buffer_length", C.c_int32),
        ("depth_error_threshold", C.c_float),
        ("prefiltered", C.c_int32), ("render_geo", C.c_int32), ("render_depth_only", C.c_int32),
        ("debug", C.c_int32),
        ("bg", _fp), ("viewmatrix", _fp), ("projmatrix", _fp), ("campos", _fp),
        ("ref_to_src_list", _fp), ("src_cam_pos", _fp), ("src_images", _fp), ("src_rendered_depths", _fp),
    ]


class IbgsForwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("view", IbgsView),
        ("means3D", _fp), ("shs", _fp), ("shs_rest", _fp), ("colors_precomp", _fp), ("opacities", _fp), ("scales", _fp),
        ("rotations", _fp), ("cov3D_precomp", _fp), ("all_map", _fp),
        ("out_color", _fp), ("radii", _fp), ("out_normal_map", _fp), ("out_median_intersected_depth", _fp),
        ("out_cam_feat", _fp), ("out_warped_image", _fp), ("out_min_depth_diff", _fp),
        ("out_camera_ray", _fp), ("out_use_first_src_frame", _fp),
        ("alloc", ALLOC_FN), ("alloc_user", C.c_void_p),
        ("tex_generation_out", C.c_int64), ("scratch_capacity_out", C.c_int64),
    ]


class IbgsBackwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("R", C.c_int64), ("view", IbgsView),
        ("means3D", _fp), ("shs", _fp), ("shs_rest", _fp), ("colors_precomp", _fp), ("scales", _fp), ("rotations", _fp),
        ("cov3D_precomp", _fp), ("all_map", _fp), ("radii", _fp),
        ("out_median_intersected_depth", _fp), ("out_warped_image", _fp),
        ("geom_buffer", _fp), ("binning_buffer", _fp), ("image_buffer", _fp),
        ("tex_generation", C.c_int64),
        ("dL_dout_color", _fp), ("dL_dout_normal_map", _fp), ("dL_dout_median_intersected_depth", _fp),
        ("dL_dout_warped_image", _fp),
        ("dL_dmeans3D", _fp), ("dL_dmeans2D", _fp), ("dL_dmeans2D_abs", _fp), ("dL_dcolors", _fp),
        ("dL_dopacity", _fp), ("dL_dcov3D", _fp), ("dL_dsh", _fp), ("dL_dsh_rest", _fp), ("dL_dscales", _fp),
        ("dL_drotations", _fp), ("dL_dall_map", _fp),
        ("alloc", ALLOC_FN), ("alloc_user", C.c_void_p), ("accumulate_mask", C.c_uint32),
    ]


ACC_MEANS3D, ACC_MEANS2D, ACC_MEANS2D_ABS, ACC_OPACITY, ACC_SH, ACC_SH_REST, ACC_SCALES, ACC_ROTATIONS, ACC_ALL_MAP = \
    (1 << i for i in range(9))


class IbgsPrologueArgs(C.Structure):
    _fields_ = [("P", C.c_int32), ("sh_rest", C.c_int32)] + [(n, _fp) for n in (
        "xyz", "opacity_raw", "scaling_raw", "rotation_raw", "features_dc", "features_rest", "normal_raw", "offset",
        "world_view_transform", "camera_center",
        "opacity", "scales", "rotations", "shs", "all_map",
        "g_opacity", "g_scales", "g_rotations", "g_shs", "g_all_map",
        "d_xyz", "d_opacity_raw", "d_scaling_raw", "d_rotation_raw", "d_features_dc", "d_features_rest",
        "d_normal_raw", "d_offset")] + [("smallest_axis_normal", C.c_int32)]


class IbgsDepthBatchArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("V", C.c_int32), ("image_height", C.c_int32), ("image_width", C.c_int32),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
        ("buffer_length", C.c_int32), ("prefiltered", C.c_int32), ("debug", C.c_int32),
    ] + [(n, _fp) for n in (
        "viewmatrices", "projmatrices", "means3D", "opacities", "scales", "rotations", "cov3D_precomp", "all_maps",
        "normals", "offsets", "camera_centers", "out_depths", "radii")] + [
        ("num_rendered", C.POINTER(C.c_int64)), ("alloc", ALLOC_FN), ("alloc_user", C.c_void_p)]


class IbgsSsimArgs(C.Structure):
    _fields_ = [("planes", C.c_int32), ("height", C.c_int32), ("width", C.c_int32)] + [(k, _fp) for k in (
        "img1", "img2", "ssim_map", "dm_dmu1", "dm_de11", "dm_de12", "dm_dmu2", "dL_dmap")] + [
        ("dL_dmap_is_scalar", C.c_int32), ("dL_dmap_scale", C.c_float), ("dL_dimg1", _fp), ("dL_dimg2", _fp)]


class IbgsColorFeatArgs(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("height", "width", "n_views", "mode", "channel_pitch", "bf16")] + [
        (k, _fp) for k in ("warped", "cam_feat", "rendered", "camera_ray", "w1", "b1", "w2", "b2", "cnn_input",
                           "g_cnn_input", "d_warped", "d_rendered", "d_w1", "d_b1", "d_w2", "d_b2")]


ADAM_MAX_GROUPS = 16


class IbgsAdamGroup(C.Structure):
    _fields_ = [("offset", C.c_int64), ("count", C.c_int64), ("lr", C.c_float)]


class IbgsAdamArgs(C.Structure):
    _fields_ = [("params", _fp), ("grads", _fp), ("exp_avg", _fp), ("exp_avg_sq", _fp), ("num_groups", C.c_int32),
                ("groups", IbgsAdamGroup * ADAM_MAX_GROUPS), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("step", C.c_int64), ("grad_scale", C.c_float), ("zero_grads", C.c_int32)]


MAX_DEPTH_BATCH = 16

EXPORTS = [
    "ibgs_forward", "ibgs_backward", "ibgs_mark_visible", "ibgs_dist2_scratch_bytes", "ibgs_dist2",
    "ibgs_forward_h", "ibgs_forward_backward_h", "ibgs_dist2_h", "ibgs_state_layout", "ibgs_sort_bits", "ibgs_last_error",
    "ibgs_abi_version", "ibgs_launch_count", "ibgs_release_cached", "ibgs_profile_enable", "ibgs_profile_reset",
    "ibgs_profile_read", "ibgs_profile_name", "ibgs_profile_stages", "ibgs_prologue_forward",
    "ibgs_prologue_backward", "ibgs_forward_depth_batch", "ibgs_ssim_forward", "ibgs_ssim_backward", "ibgs_adam_step", "ibgs_set_backward_variant", "ibgs_set_forward_variant",
    "ibgs_sort_temp_bytes", "ibgs_sort_pairs", "ibgs_scan_temp_bytes", "ibgs_scan_gather",
    "ibgs_color_features_forward", "ibgs_color_features_backward",
    "ibgs_nhwc_maxpool2_forward", "ibgs_nhwc_maxpool2_backward", "ibgs_nhwc_upsample_cat_forward",
    "ibgs_nhwc_upsample_backward", "ibgs_nhwc_relu_bias_backward",
    "ibgs_depth_normal_forward", "ibgs_depth_normal_backward", "ibgs_densification_stats",
]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ibgs_b200.build` "
            "(there is deliberately no CPU / PyTorch fallback for the rasterizer)")
    lib = C.CDLL(LIB_PATH)
    lib.ibgs_forward.restype = C.c_int64
    lib.ibgs_forward.argtypes = [C.POINTER(IbgsForwardArgs), C.c_void_p]
    lib.ibgs_backward.restype = C.c_int
    lib.ibgs_backward.argtypes = [C.POINTER(IbgsBackwardArgs), C.c_void_p]
    lib.ibgs_mark_visible.restype = C.c_int
    lib.ibgs_mark_visible.argtypes = [C.c_int32, _fp, _fp, _fp, _fp, C.c_void_p]
    lib.ibgs_dist2_scratch_bytes.restype = C.c_size_t
    lib.ibgs_dist2_scratch_bytes.argtypes = [C.c_int32]
    lib.ibgs_dist2.restype = C.c_int
    lib.ibgs_dist2.argtypes = [C.c_int32, _fp, _fp, _fp, C.c_size_t, C.c_void_p]
    lib.ibgs_forward_h.restype = C.c_int64
    lib.ibgs_forward_h.argtypes = [C.POINTER(IbgsForwardArgs)]
    lib.ibgs_forward_backward_h.restype = C.c_int64
    lib.ibgs_forward_backward_h.argtypes = [C.POINTER(IbgsForwardArgs), C.POINTER(IbgsBackwardArgs)]
    lib.ibgs_dist2_h.restype = C.c_int
    lib.ibgs_dist2_h.argtypes = [C.c_int32, _fp, _fp]
    lib.ibgs_state_layout.restype = C.c_int
    lib.ibgs_state_layout.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t), C.c_int,
                                      C.POINTER(C.c_size_t)]
    lib.ibgs_sort_bits.restype = C.c_int
    lib.ibgs_sort_bits.argtypes = [C.c_int32]
    lib.ibgs_last_error.restype = C.c_char_p
    lib.ibgs_last_error.argtypes = []
    lib.ibgs_abi_version.restype = C.c_int
    lib.ibgs_launch_count.restype = C.c_int64
    lib.ibgs_release_cached.restype = None
    lib.ibgs_profile_enable.restype = None
    lib.ibgs_profile_enable.argtypes = [C.c_int]
    lib.ibgs_profile_reset.restype = None
    lib.ibgs_profile_read.restype = C.c_int
    lib.ibgs_profile_read.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.ibgs_profile_name.restype = C.c_char_p
    lib.ibgs_profile_name.argtypes = [C.c_int]
    lib.ibgs_profile_stages.restype = C.c_int
    for fn in (lib.ibgs_prologue_forward, lib.ibgs_prologue_backward):
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(IbgsPrologueArgs), C.c_void_p]
    for fn in (lib.ibgs_ssim_forward, lib.ibgs_ssim_backward):
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(IbgsSsimArgs), C.c_void_p]
    for fn in (lib.ibgs_color_features_forward, lib.ibgs_color_features_backward):
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(IbgsColorFeatArgs), C.c_void_p]
    i32 = C.c_int32
    lib.ibgs_nhwc_maxpool2_forward.restype = C.c_int
    lib.ibgs_nhwc_maxpool2_forward.argtypes = [_fp, _fp, _fp, i32, i32, i32, i32, C.c_void_p]
    lib.ibgs_nhwc_maxpool2_backward.restype = C.c_int
    lib.ibgs_nhwc_maxpool2_backward.argtypes = [_fp, _fp, _fp, i32, i32, i32, i32, C.c_void_p]
    lib.ibgs_nhwc_upsample_cat_forward.restype = C.c_int
    lib.ibgs_nhwc_upsample_cat_forward.argtypes = [_fp, _fp, _fp, i32, i32, i32, i32, i32, i32, i32, C.c_void_p]
    lib.ibgs_nhwc_upsample_backward.restype = C.c_int
    lib.ibgs_nhwc_upsample_backward.argtypes = [_fp, _fp, i32, i32, i32, i32, i32, i32, i32, C.c_void_p]
    f32 = C.c_float
    lib.ibgs_depth_normal_forward.restype = C.c_int
    lib.ibgs_depth_normal_forward.argtypes = [_fp, _fp, i32, i32, f32, f32, f32, f32, C.c_void_p]
    lib.ibgs_depth_normal_backward.restype = C.c_int
    lib.ibgs_depth_normal_backward.argtypes = [_fp, _fp, _fp, i32, i32, f32, f32, f32, f32, C.c_void_p]
    lib.ibgs_densification_stats.restype = C.c_int
    lib.ibgs_densification_stats.argtypes = [i32] + [_fp] * 8 + [C.c_void_p]
    lib.ibgs_nhwc_relu_bias_backward.restype = C.c_int
    lib.ibgs_nhwc_relu_bias_backward.argtypes = [_fp, i32, _fp, _fp, _fp, C.c_int64, i32, i32, C.c_void_p]
    for fn in (lib.ibgs_set_backward_variant, lib.ibgs_set_forward_variant):
        fn.restype = C.c_int
        fn.argtypes = [C.c_int]
    lib.ibgs_adam_step.restype = C.c_int
    lib.ibgs_adam_step.argtypes = [C.POINTER(IbgsAdamArgs), C.c_void_p]
    lib.ibgs_sort_temp_bytes.restype = C.c_size_t
    lib.ibgs_sort_temp_bytes.argtypes = [C.c_int64, C.c_int, C.c_int]
    lib.ibgs_sort_pairs.restype = C.c_int
    lib.ibgs_sort_pairs.argtypes = [_fp, _fp, _fp, _fp, C.c_int64, C.c_int, C.c_int, _fp, C.c_size_t, C.c_void_p]
    lib.ibgs_scan_temp_bytes.restype = C.c_size_t
    lib.ibgs_scan_temp_bytes.argtypes = [C.c_int64]
    lib.ibgs_scan_gather.restype = C.c_int
    lib.ibgs_scan_gather.argtypes = [C.c_int64, _fp, _fp, _fp, _fp, C.c_size_t, C.c_void_p]
    lib.ibgs_forward_depth_batch.restype = C.c_int64
    lib.ibgs_forward_depth_batch.argtypes = [C.POINTER(IbgsDepthBatchArgs), C.c_void_p]
    return lib


lib = _load()


def last_error():
    return lib.ibgs_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc < 0:
        raise RuntimeError(f"{what} failed ({rc}): {last_error()}")
    return rc


def profile_read():
    """{stage name: (total ms, launches)} from the library's built-in event timer."""
    out = {}
    for i in range(lib.ibgs_profile_stages()):
        ms, n = C.c_double(0), C.c_int64(0)
        lib.ibgs_profile_read(i, C.byref(ms), C.byref(n))
        out[lib.ibgs_profile_name(i).decode()] = (ms.value, n.value)
    return out


def state_layout(which, count, aux=0):
    offs = (C.c_size_t * 16)()
    total = C.c_size_t(0)
    n = check(lib.ibgs_state_layout(which, count, aux, offs, 16, C.byref(total)), "ibgs_state_layout")
    return [int(offs[i]) for i in range(n)], int(total.value)
