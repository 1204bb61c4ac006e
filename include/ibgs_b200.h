/*
 * ibgs_b200.h -- C ABI of the B200-native IBGS planar Gaussian rasterizer.
 *
 * This is the drop-in boundary for ONE path of HoangChuongNguyen/ibgs: what the reference
 * binds through pybind11 in submodules/diff-plane-rasterization/ext.cpp:15-19
 * (rasterize_gaussians / rasterize_gaussians_backward / mark_visible) and
 * submodules/simple-knn/ext.cpp:15-17 (distCUDA2).  Plain pointers and sizes only; no torch
 * types.  All data pointers are DEVICE pointers on the current CUDA device unless a `_h`
 * entry point says otherwise; `stream` is a cudaStream_t passed as void* (NULL = legacy
 * default stream, which is what the reference always uses: rasterizer_impl.cu launches
 * `<<<grid,block>>>` without a stream).
 *
 * Every entry point returns >= 0 on success and a negative IBGS_E* code on failure;
 * ibgs_last_error() returns a human-readable message for the calling thread.
 */
#ifndef IBGS_B200_H_INCLUDED
#define IBGS_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IBGS_ABI_VERSION 5   /* 5: IbgsPrologueArgs.smallest_axis_normal; colour-feature, NHWC-glue and depth-normal entry points */

/* Compile-time constants of the reference (cuda_rasterizer/config.h:15-19, auxiliary.h:21-23). */
#define IBGS_NUM_CHANNELS 3
#define IBGS_NUM_PLANE_PARAMS 5
#define IBGS_TILE 16
#define IBGS_MAX_BUFFER_LENGTH 8
#define IBGS_MAX_SRC 5

enum {
  IBGS_OK = 0,
  IBGS_EINVAL = -1,  /* bad argument (shape / null / range) -- reference: rasterize_points.cu:69-71 */
  IBGS_ECUDA = -2,   /* CUDA runtime error (message has cudaGetErrorString) */
  IBGS_EALLOC = -3,  /* allocator callback returned NULL */
  IBGS_ELIMIT = -4   /* num_rendered would exceed the reference's int32 limit (rasterizer_impl.cu:429) */
};

/* Scratch/state buffers.  The reference hands Rasterizer::forward three std::function<char*(size_t)>
 * resizers (rasterize_points.cu:29-35,94-99; rasterizer_impl.cu:321-323); this is the C spelling:
 * the library calls `alloc(user, which, bytes)` and gets device memory owned by the caller.
 * GEOM / BINNING / IMAGE must stay alive until the matching ibgs_backward; SCRATCH may be released
 * (stream-ordered) as soon as ibgs_forward returns. */
enum { IBGS_BUF_GEOM = 0, IBGS_BUF_BINNING = 1, IBGS_BUF_IMAGE = 2, IBGS_BUF_SCRATCH = 3 };
typedef void* (*ibgs_alloc_fn)(void* user, int which, size_t bytes);

/* View / mode parameters shared by forward and backward
 * (reference: GaussianRasterizationSettings, diff_plane_rasterization/__init__.py:252-276). */
typedef struct IbgsView {
  int32_t image_height, image_width;
  float tanfovx, tanfovy;
  float scale_modifier;
  int32_t sh_degree;          /* active degree D */
  int32_t sh_coeffs;          /* M = coefficients per channel in `shs` (0 when shs absent) */
  int32_t nb_src_images;      /* <= IBGS_MAX_SRC */
  int32_t buffer_length;      /* 1..IBGS_MAX_BUFFER_LENGTH */
  float depth_error_threshold;
  int32_t prefiltered, render_geo, render_depth_only, debug;
  const float* bg;            /* [3] */
  const float* viewmatrix;    /* [16], torch world_view_transform memory (= column-major W2C) */
  const float* projmatrix;    /* [16] */
  const float* campos;        /* [3] */
  const float* ref_to_src_list;     /* [nb_src,16] row-major */
  const float* src_cam_pos;         /* [nb_src,3] */
  const float* src_images;          /* [nb_src,3,H,W] planar */
  const float* src_rendered_depths; /* [nb_src,1,H,W] */
} IbgsView;

/* Forward.  Replaces CudaRasterizer::Rasterizer::forward (rasterizer_impl.cu:320-515) as called by
 * RasterizeGaussiansCUDA (rasterize_points.cu:37-160).  NULL pointer == "absent", as in the
 * reference (forward.cu:244,280).  Outputs must be pre-zeroed by the caller exactly like the
 * binding's torch::full(...,0) (rasterize_points.cu:80-90). */
typedef struct IbgsForwardArgs {
  int32_t P;
  IbgsView view;
  const float* means3D;        /* [P,3] */
  const float* shs;            /* [P,M,3] or NULL; with shs_rest given: the DC coefficient only, [P,1,3] */
  const float* shs_rest;       /* NULL, or coefficients 1..M-1 as their own tensor [P,M-1,3] (GaussianModel keeps
                                  _features_dc / _features_rest apart, scene/gaussian_model.py:139-143; reading them
                                  in place saves the torch.cat the reference does per view) */
  const float* colors_precomp; /* [P,3] or NULL */
  const float* opacities;      /* [P] */
  const float* scales;         /* [P,3] or NULL */
  const float* rotations;      /* [P,4] or NULL */
  const float* cov3D_precomp;  /* [P,6] or NULL */
  const float* all_map;        /* [P,5] or NULL (required for render_geo / render_depth_only) */
  /* outputs (CHW planar, float32 unless noted) */
  float* out_color;                     /* [3,H,W] */
  int32_t* radii;                       /* [P] */
  float* out_normal_map;                /* [3,H,W] */
  float* out_median_intersected_depth;  /* [1,H,W] */
  float* out_cam_feat;                  /* [4*5,H,W] */
  float* out_warped_image;              /* [3*5,H,W] */
  float* out_min_depth_diff;            /* [1,H,W] */
  float* out_camera_ray;                /* [3,H,W] */
  int32_t* out_use_first_src_frame;     /* [1,H,W] */
  ibgs_alloc_fn alloc;
  void* alloc_user;
  int64_t tex_generation_out;  /* written by the library: id of the texture fill, see backward */
  int64_t scratch_capacity_out; /* written by the library: instance capacity the LAST scratch request was carved for
                                 * (>= num_rendered; pass it as `count` to ibgs_state_layout(IBGS_BUF_SCRATCH, ...)) */
} IbgsForwardArgs;

/* returns num_rendered (tile instances R) */
int64_t ibgs_forward(IbgsForwardArgs* args, void* stream);

enum {
  IBGS_ACC_MEANS3D = 1u << 0, IBGS_ACC_MEANS2D = 1u << 1, IBGS_ACC_MEANS2D_ABS = 1u << 2, IBGS_ACC_OPACITY = 1u << 3,
  IBGS_ACC_SH = 1u << 4, IBGS_ACC_SH_REST = 1u << 5, IBGS_ACC_SCALES = 1u << 6, IBGS_ACC_ROTATIONS = 1u << 7,
  IBGS_ACC_ALL_MAP = 1u << 8
};

/* Backward.  Replaces CudaRasterizer::Rasterizer::backward (rasterizer_impl.cu:519-666) as called
 * by RasterizeGaussiansBackwardCUDA (rasterize_points.cu:162-271).  Every dL_* output is fully
 * written (no pre-zeroing needed), except that rows of Gaussians with radii==0 are written as 0. */
typedef struct IbgsBackwardArgs {
  int32_t P;
  int64_t R;
  IbgsView view;
  const float* means3D;
  const float* shs;            /* as in the forward: [P,M,3], or the DC part when shs_rest is given */
  const float* shs_rest;
  const float* colors_precomp;
  const float* scales;
  const float* rotations;
  const float* cov3D_precomp;
  const float* all_map;
  const int32_t* radii;
  /* saved forward outputs */
  const float* out_median_intersected_depth;
  const float* out_warped_image;
  /* state buffers from the forward allocator */
  const void* geom_buffer;
  const void* binning_buffer;
  const void* image_buffer;
  int64_t tex_generation;      /* from forward; textures are refilled if it is stale */
  /* cotangents; only these four are consumed (rasterize_points.cu:205-211) */
  const float* dL_dout_color;                     /* [3,H,W] */
  const float* dL_dout_normal_map;                /* [3,H,W] */
  const float* dL_dout_median_intersected_depth;  /* [1,H,W] */
  const float* dL_dout_warped_image;              /* [15,H,W] */
  /* gradients */
  float* dL_dmeans3D;     /* [P,3] */
  float* dL_dmeans2D;     /* [P,3] (z column = 0) */
  float* dL_dmeans2D_abs; /* [P,3] */
  float* dL_dcolors;      /* [P,3] */
  float* dL_dopacity;     /* [P,1] */
  float* dL_dcov3D;       /* [P,6] */
  float* dL_dsh;          /* [P,M,3] or NULL; with shs_rest given: [P,1,3] */
  float* dL_dsh_rest;     /* [P,M-1,3] when shs_rest is given, else NULL */
  float* dL_dscales;      /* [P,3] */
  float* dL_drotations;   /* [P,4] */
  float* dL_dall_map;     /* [P,5] */
  ibgs_alloc_fn alloc;    /* SCRATCH only (per-Gaussian accumulation arena, 64 B/Gaussian) */
  void* alloc_user;
  uint32_t accumulate_mask; /* IBGS_ACC_* bits: these gradient outputs are accumulated into (dst += grad; rows of culled
                               Gaussians untouched) instead of written -- gradient accumulation over a batch of views
                               without autograd's separate add pass.  0 = the reference behaviour (every output written). */
} IbgsBackwardArgs;

int ibgs_backward(IbgsBackwardArgs* args, void* stream);

/* markVisible (rasterizer_impl.cu:258-270, kernel :171-183): present[i] = view-space z > 0.2 */
int ibgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                      const float* projmatrix, uint8_t* present, void* stream);

/* simple-knn distCUDA2 (simple_knn.cu:185-221, spatial.cu:15-26): mean squared distance to the three
 * nearest neighbours.  `scratch` is device memory of at least ibgs_dist2_scratch_bytes(P). */
size_t ibgs_dist2_scratch_bytes(int32_t P);
int ibgs_dist2(int32_t P, const float* points, float* mean_dists, void* scratch, size_t scratch_bytes,
               void* stream);

/* Fused per-view parameter prologue (optional fast path, the reference has no such
 * entry point).  Replaces the ~30 PyTorch kernels gaussian_renderer.render() runs over all Gaussians before every
 * rasterizer call: the activations of the GaussianModel getters (scene/gaussian_model.py:127-147), the learnt plane
 * normal (get_normal, :166-173) and the all_map construction (gaussian_renderer/__init__.py:304-315) -- and their
 * autograd backward.  normal_raw/offset may both be NULL (no all_map, e.g. colour-only renders).  shs may be NULL
 * (no concatenation: the caller passes features_dc / features_rest to ibgs_forward as shs / shs_rest), and likewise
 * d_features_dc in the backward call.  In the backward
 * call every g_* may be NULL (= zero cotangent) and d_xyz holds only the all_map path's share of dL/dxyz. */
typedef struct IbgsPrologueArgs {
  int32_t P;
  int32_t sh_rest;               /* coefficients per channel in features_rest (K-1) */
  const float* xyz;              /* [P,3]  _xyz */
  const float* opacity_raw;      /* [P,1]  _opacity (pre-sigmoid) */
  const float* scaling_raw;      /* [P,3]  _scaling (log) */
  const float* rotation_raw;     /* [P,4]  _rotation (un-normalised) */
  const float* features_dc;      /* [P,1,3] */
  const float* features_rest;    /* [P,K-1,3] */
  const float* normal_raw;       /* [P,3]  _normal or NULL */
  const float* offset;           /* [P,1]  _offset or NULL */
  const float* world_view_transform; /* [16] */
  const float* camera_center;    /* [3] */
  /* forward outputs */
  float* opacity;                /* [P,1] */
  float* scales;                 /* [P,3] */
  float* rotations;              /* [P,4] */
  float* shs;                    /* [P,K,3] */
  float* all_map;                /* [P,5] */
  /* backward inputs (cotangents of the forward outputs) */
  const float* g_opacity; const float* g_scales; const float* g_rotations; const float* g_shs; const float* g_all_map;
  /* backward outputs */
  float* d_xyz; float* d_opacity_raw; float* d_scaling_raw; float* d_rotation_raw;
  float* d_features_dc; float* d_features_rest; float* d_normal_raw; float* d_offset;
  /* learnt_normal = False (gaussian_renderer/__init__.py:305-306, scene/gaussian_model.py:149-161): the plane normal is
   * the Gaussian's SHORTEST AXIS -- column argmin(scales) of the rotation matrix of the normalised quaternion -- flipped
   * towards the camera; no offset.  Set it with normal_raw = offset = NULL and all_map != NULL; the normal's gradient
   * flows into d_rotation_raw (and d_xyz). */
  int32_t smallest_axis_normal;
} IbgsPrologueArgs;
int ibgs_prologue_forward(const IbgsPrologueArgs* args, void* stream);
int ibgs_prologue_backward(const IbgsPrologueArgs* args, void* stream);

/* Batched source-view depth renders (optional fast path, the reference has no such
 * entry point).  At test time and with do_render_src_depth the reference renders the plane depth of every source
 * view with its own rasterizer call (gaussian_renderer/__init__.py:245-253 -> render_depth :33-145, i.e. V x
 * [per-view all_map torch ops + preprocess + sort + depth-only render] over the SAME Gaussians).  This entry point
 * renders V views (same image size and intrinsics) in ONE pass: one preprocess launch reads each Gaussian once and
 * writes V records, the V*P (view, Gaussian) items go through ONE depth-order sort / scan / emission / tile sort
 * with the view folded into the tile id, and one depth-only tile-renderer launch with gridDim.z = V.  Results per
 * view are those of ibgs_forward with render_depth_only=1 (radii, tile lists and the blend order are bit-identical).
 * Plane parameters: either all_maps [V,P,5] (what render_depth builds per view, :123-132), or -- all_maps NULL --
 * the world-space normals [P,3] (un-normalised _normal, or the unit shortest axis) and optional offsets [P] plus the
 * V camera centres, from which the kernel derives each view's (local normal, 1, |local distance|) itself
 * (scene/gaussian_model.py:158-173 + gaussian_renderer/__init__.py:123-129). */
#define IBGS_MAX_DEPTH_BATCH 16
typedef struct IbgsDepthBatchArgs {
  int32_t P, V;
  int32_t image_height, image_width;
  float tanfovx, tanfovy, scale_modifier;
  int32_t buffer_length, prefiltered, debug;
  const float* viewmatrices;   /* [V,16] */
  const float* projmatrices;   /* [V,16] */
  const float* means3D;        /* [P,3] */
  const float* opacities;      /* [P] */
  const float* scales;         /* [P,3] or NULL */
  const float* rotations;      /* [P,4] or NULL */
  const float* cov3D_precomp;  /* [P,6] or NULL */
  const float* all_maps;       /* [V,P,5] or NULL */
  const float* normals;        /* [P,3] world space; used when all_maps is NULL */
  const float* offsets;        /* [P] or NULL */
  const float* camera_centers; /* [V,3]; used when all_maps is NULL */
  float* out_depths;           /* [V,1,H,W] */
  int32_t* radii;              /* [V,P] or NULL */
  int64_t* num_rendered;       /* HOST [V] or NULL: tile instances per view */
  ibgs_alloc_fn alloc;         /* SCRATCH only */
  void* alloc_user;
} IbgsDepthBatchArgs;
/* returns the total number of tile instances over the V views */
int64_t ibgs_forward_depth_batch(IbgsDepthBatchArgs* args, void* stream);

/* Fused SSIM map + backward (the loss step next to the rasterizer; optional fast path).
 * Replaces utils/loss_utils.py:34-65 (ssim / _ssim), :67-90 (compute_photometric_ssim) and :92-117 (ssim2): the five
 * depthwise 11x11 Gaussian-window convolutions (sigma 1.5, zero padding 5) and the SSIM formula with C1 = 0.01^2,
 * C2 = 0.03^2, for `planes` independent H x W planes (batch and channels folded; the window is depthwise).
 * The forward optionally stores the partial derivatives of the map with respect to the convolution outputs
 * (mu1, e11 = W*img1^2, e12 = W*img1*img2; and mu2 for a differentiable img2 -- dm/de22 equals dm/de11); the backward turns them and the
 * map's cotangent into dL_dimg1
 * (and dL_dimg2 when given). */
typedef struct IbgsSsimArgs {
  int32_t planes, height, width;
  const float* img1;      /* [planes,H,W] */
  const float* img2;      /* [planes,H,W] */
  float* ssim_map;        /* [planes,H,W] forward output */
  float* dm_dmu1;         /* [planes,H,W] forward outputs / backward inputs; all NULL = inference */
  float* dm_de11;
  float* dm_de12;
  float* dm_dmu2;         /* NULL unless img2 needs a gradient */
  const float* dL_dmap;   /* backward: [planes,H,W]; or ONE device float broadcast to every pixel when
                             dL_dmap_is_scalar (the `.mean()` case, no host read-back needed); or NULL (= 1) */
  int32_t dL_dmap_is_scalar;
  float dL_dmap_scale;    /* backward: multiplies dL_dmap */
  float* dL_dimg1;        /* backward output [planes,H,W] */
  float* dL_dimg2;        /* backward output or NULL */
} IbgsSsimArgs;
int ibgs_ssim_forward(const IbgsSsimArgs* args, void* stream);
int ibgs_ssim_backward(const IbgsSsimArgs* args, void* stream);

/* One-launch Adam step over flat arenas (optional fast path).  Replaces
 * gaussians.optimizer.step() + zero_grad (train.py:422-424) for torch.optim.Adam(l, lr=0.0, eps=1e-15) with the eight
 * per-Gaussian parameter groups of scene/gaussian_model.py:227-240: params / grads / exp_avg / exp_avg_sq are four flat
 * float arrays of one layout, every group a contiguous [offset, offset+count) range with its own learning rate.
 * Arithmetic: torch/optim/adam.py (amsgrad=False, weight_decay=0, maximize=False).  `step` is the 1-based count of this
 * update (torch increments state['step'] before using it).  grad_scale multiplies the gradients first (1/views for a
 * mean over a view batch); zero_grads != 0 clears the gradients in the same pass. */
#define IBGS_ADAM_MAX_GROUPS 16
typedef struct IbgsAdamGroup { int64_t offset, count; float lr; } IbgsAdamGroup;
typedef struct IbgsAdamArgs {
  float* params;
  float* grads;
  float* exp_avg;
  float* exp_avg_sq;
  int32_t num_groups;
  IbgsAdamGroup groups[IBGS_ADAM_MAX_GROUPS];
  float beta1, beta2, eps;
  int64_t step;
  float grad_scale;
  int32_t zero_grads;
} IbgsAdamArgs;
int ibgs_adam_step(const IbgsAdamArgs* args, void* stream);

/* Colour-aggregation front half (optional fast path).  Replaces, in ONE launch per direction,
 * the feature assembly of fuse_color (color_aggregation_network.py:196-206: valid = sum of a view's 4 camera features > 0,
 * residual = (warped - rendered) * valid, x = [residual, cam_feat]) and ColorFusionResidualNet.forward up to the conv
 * decoder's input (:121-131: per_view_mlp = Linear(7,32) ReLU Linear(32,32) ReLU per (pixel, view), mean / max over the
 * views, cat with ray_dir and c_3dgs).  The output is the decoder's input in NHWC with `channel_pitch` channels per pixel
 * (0..31 aggregated features, 32..34 ray, 35..37 rendered colour, the rest zero), float32 or bf16.
 * The backward recomputes the hidden layers; weight gradients are ACCUMULATED (atomicAdd) into d_w1 / d_b1 / d_w2 / d_b2
 * (the caller zeroes them), d_warped / d_rendered are written (either may be NULL: the detached case of :171-177). */
typedef struct IbgsColorFeatArgs {
  int32_t height, width;
  int32_t n_views;            /* nb_valid_warp_level, 1..5 */
  int32_t mode;               /* 0 = mean, 1 = max (feat_aggregate_mode) */
  int32_t channel_pitch;      /* multiple of 8, >= 40 */
  int32_t bf16;               /* cnn_input / g_cnn_input element type: 1 = bf16, 0 = float32 */
  const float* warped;        /* [n_views][3][H*W]  first n_views of out_warped_image */
  const float* cam_feat;      /* [n_views][4][H*W]  first n_views of out_cam_feat */
  const float* rendered;      /* [3][H*W] */
  const float* camera_ray;    /* [3][H*W] */
  const float* w1;            /* per_view_mlp[0].weight [32][7] */
  const float* b1;            /* [32] */
  const float* w2;            /* per_view_mlp[2].weight [32][32] */
  const float* b2;            /* [32] */
  void* cnn_input;            /* forward output [H*W][channel_pitch] */
  const void* g_cnn_input;    /* backward input  [H*W][channel_pitch] */
  float* d_warped;            /* backward output [n_views][3][H*W] or NULL */
  float* d_rendered;          /* backward output [3][H*W] or NULL */
  float* d_w1; float* d_b1; float* d_w2; float* d_b2;   /* backward, accumulated */
} IbgsColorFeatArgs;
int ibgs_color_features_forward(const IbgsColorFeatArgs* args, void* stream);
int ibgs_color_features_backward(const IbgsColorFeatArgs* args, void* stream);

/* NHWC glue of the colour network's conv decoder (ConvDecoderAE.forward, color_aggregation_network.py:51-68; optional fast
 * path): nn.MaxPool2d(2) and F.interpolate(mode="nearest") with their backward passes on [H][W][C] tensors, C a multiple
 * of 8, elements bf16 (bf16 != 0) or float32, 16-byte aligned.  Same results as torch (first maximum of a window wins,
 * NaN propagates; source index = min(int(floorf(dst * (float(in) / out))), in - 1)).
 *   maxpool2_forward:   y [H/2][W/2][C], idx [H/2][W/2][C] bytes (position 0..3 of the maximum in its window)
 *   maxpool2_backward:  gx [H][W][C] written completely (zeros for non-maxima and the odd last row / column)
 *   upsample_cat_forward: out [Ho][Wo][Ca+Cb] = cat(nearest-upsampled a [Hi][Wi][Ca], b [Ho][Wo][Cb]); b = NULL, Cb = 0: plain upsample
 *   upsample_backward:  ga [Hi][Wi][Ca] = sum over the output pixels that read it of g[..][0..Ca); g has `pitch` elements per pixel */
int ibgs_nhwc_maxpool2_forward(const void* x, void* y, uint8_t* idx, int32_t H, int32_t W, int32_t C, int32_t bf16, void* stream);
int ibgs_nhwc_maxpool2_backward(const void* gy, const uint8_t* idx, void* gx, int32_t H, int32_t W, int32_t C, int32_t bf16,
                                void* stream);
int ibgs_nhwc_upsample_cat_forward(const void* a, const void* b, void* out, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo,
                                   int32_t Ca, int32_t Cb, int32_t bf16, void* stream);
int ibgs_nhwc_upsample_backward(const void* g, void* ga, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, int32_t Ca,
                                int32_t pitch, int32_t bf16, void* stream);
/* ReLU backward of a conv + bias + ReLU layer fused with the bias gradient: gm[p][c] = y[p][c] > 0 ? g[p*pitch + c] : 0
 * (aten::threshold_backward; g may be a channel slice of a wider NHWC gradient: `pitch` elements per pixel) and
 * db[c] += sum_p gm[p][c] (float32, ACCUMULATED: the caller zeroes it). */
int ibgs_nhwc_relu_bias_backward(const void* g, int32_t pitch, const void* y, void* gm, float* db, int64_t npix, int32_t C,
                                 int32_t bf16, void* stream);

/* Unit normal map of a depth image (optional fast path): render_normal + the renormalisation of
 * gaussian_renderer/__init__.py:15-26,332-335 (utils/graphics_utils.py:38-75 with Camera.get_calib_matrix_nerf's pinhole
 * intrinsics): normal [3][H][W] from depth [H][W], zero on the 1-pixel border; the backward writes g_depth [H][W]. */
int ibgs_depth_normal_forward(const float* depth, float* normal, int32_t H, int32_t W, float fx, float fy, float cx, float cy,
                              void* stream);
int ibgs_depth_normal_backward(const float* depth, const float* g_normal, float* g_depth, int32_t H, int32_t W, float fx,
                               float fy, float cx, float cy, void* stream);

/* Per-view densification statistics (train.py:399-405 + GaussianModel.add_densification_stats,
 * scene/gaussian_model.py:600-604; optional fast path), for the Gaussians with radii > 0: max_radii2D = max(max_radii2D,
 * radii); xyz_gradient_accum += |viewspace_grad[:, :2]|; xyz_gradient_accum_abs += |viewspace_grad_abs[:, :2]|; denom += 1;
 * denom_abs += 1.  viewspace_grad / _abs are the [P,3] gradients of the two screen-space dummy tensors; the five
 * statistics arrays hold P floats each and are updated in place. */
int ibgs_densification_stats(int32_t P, const int32_t* radii, const float* viewspace_grad, const float* viewspace_grad_abs,
                             float* max_radii2D, float* xyz_gradient_accum, float* xyz_gradient_accum_abs, float* denom,
                             float* denom_abs, void* stream);

/* Host-buffer convenience entry points (what a non-torch caller binds -- cgo / JNI / ctypes on plain host arrays;
 * exercised by tests/test_gpu_host_api.py against the device entry points):
 * identical semantics, but every pointer in the structs is a HOST pointer; the library stages
 * through its own device arena (cudaMallocAsync) and copies results back before returning. */
int64_t ibgs_forward_h(IbgsForwardArgs* host_args);
/* One training view with host buffers: the forward above, then ibgs_backward with the host cotangents of `host_bw`
 * (dL_dout_*) while the state is still on the device; every non-NULL dL_d* of `host_bw` receives its gradient.  Only the
 * cotangent and gradient pointers of `host_bw` are read (its inputs are the forward's).  Returns num_rendered. */
int64_t ibgs_forward_backward_h(IbgsForwardArgs* host_args, IbgsBackwardArgs* host_bw);
int ibgs_dist2_h(int32_t P, const float* points_host, float* mean_dists_host);

/* The binning stage's own device primitives (csrc/sort.cu; they replace cub::DeviceRadixSort::SortPairs /
 * cub::DeviceScan::InclusiveSum of rasterizer_impl.cu:426,452-457 and simple_knn.cu:210-213), exported so that they can
 * be tested on their own.  ibgs_sort_pairs: stable ascending sort of n (key, value) pairs on the low `key_bits` bits of
 * uint16 (key_bytes 2) or uint32 (4) keys; values uint32, vals_in NULL = the item's index.  ibgs_scan_gather:
 * out[i] = sum_{j<=i} src[idx[j]].  All pointers are device pointers; `temp` must hold ibgs_sort_temp_bytes /
 * ibgs_scan_temp_bytes bytes. */
size_t ibgs_sort_temp_bytes(int64_t n, int key_bits, int key_bytes);
int ibgs_sort_pairs(const void* keys_in, const uint32_t* vals_in, void* keys_out, uint32_t* vals_out, int64_t n,
                    int key_bits, int key_bytes, void* temp, size_t temp_bytes, void* stream);
size_t ibgs_scan_temp_bytes(int64_t n);
int ibgs_scan_gather(int64_t n, const uint32_t* idx, const uint32_t* src, uint32_t* out, void* temp, size_t temp_bytes,
                     void* stream);

/* State-layout introspection (tests decode the buffers with this, mirroring how the reference's
 * GeometryState/ImageState/BinningState::fromChunk carve theirs, rasterizer_impl.cu:272-316).
 * Writes up to `max` byte offsets; returns the number of arrays in the buffer. `count` is P for GEOM,
 * R for BINNING/SCRATCH, W*H for IMAGE (aux = number of tiles for IMAGE, unused otherwise). */
int ibgs_state_layout(int which, size_t count, size_t aux, size_t* offsets, int max, size_t* total_bytes);

/* Tile-instance sort-key width: 32 + getHigherMsb(tiles) (rasterizer_impl.cu:152-167,449). */
int ibgs_sort_bits(int32_t num_tiles);

const char* ibgs_last_error(void);
int ibgs_abi_version(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches counter). */
int64_t ibgs_launch_count(void);
/* Built-in per-stage timer.  When enabled, every stage launched by ibgs_forward / ibgs_backward is
 * bracketed by CUDA events on the launching stream; ibgs_profile_read(stage, &ms, &n) returns the summed
 * device time and the number of timed launches (it waits for pending events).  Stage ids are
 * 0..ibgs_profile_stages()-1, named by ibgs_profile_name. */
void ibgs_profile_enable(int on);
void ibgs_profile_reset(void);
int ibgs_profile_read(int stage, double* ms_total, int64_t* count);
const char* ibgs_profile_name(int stage);
int ibgs_profile_stages(void);
/* Tuning knobs for tests and experiments: each tile renderer exists in two variants -- one pixel per lane
 * (8 warps per tile) and two pixels per lane (4 warps per tile, one shared-memory reduction per 64 pixels) -- and
 * ibgs_backward picks per view by the average tile-list length.  0 restores that choice, 1 / 2 force a variant.
 * Results agree within the gradient gate either way (only the summation order differs). */
int ibgs_set_backward_variant(int pixels_per_lane);
int ibgs_set_forward_variant(int pixels_per_lane);   /* same knob for the forward tile renderer */
/* Releases cached textures / arenas. */
void ibgs_release_cached(void);

#ifdef __cplusplus
}
#endif
#endif /* IBGS_B200_H_INCLUDED */
