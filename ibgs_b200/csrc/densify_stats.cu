// densify_stats.cu -- the per-view densification statistics, one launch.
//
// Reference behaviour (train.py:399-405 + GaussianModel.add_densification_stats, scene/gaussian_model.py:600-604), on the
// Gaussians the view saw (visibility_filter = radii > 0):
//     max_radii2D[mask]            = max(max_radii2D[mask], radii[mask])
//     xyz_gradient_accum[mask]     += |viewspace_points.grad[mask, :2]|          (2-norm of the screen-space gradient)
//     xyz_gradient_accum_abs[mask] += |viewspace_points_abs.grad[mask, :2]|
//     denom[mask] += 1;  denom_abs[mask] += 1
// Five boolean-mask indexing statements: each one compacts the mask first (nonzero + a host synchronisation to size the
// result), ~25 kernels and 5 syncs per view during the first 15 000 iterations.  Here: one pass over the Gaussians, no
// synchronisation, untouched rows for invisible Gaussians.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const int32_t* __restrict__ radii,
                                                            const float* __restrict__ g, const float* __restrict__ g_abs,
                                                            float* __restrict__ max_radii2D, float* __restrict__ accum,
                                                            float* __restrict__ accum_abs, float* __restrict__ denom,
                                                            float* __restrict__ denom_abs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int r = radii[i];
  if (r <= 0) return;
  max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
  const float gx = g[3 * (size_t)i], gy = g[3 * (size_t)i + 1];
  const float ax = g_abs[3 * (size_t)i], ay = g_abs[3 * (size_t)i + 1];
  accum[i] += sqrtf(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
  accum_abs[i] += sqrtf(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)));
  denom[i] += 1.0f;
  denom_abs[i] += 1.0f;
}

}  // namespace

extern "C" int ibgs_densification_stats(int32_t P, const int32_t* radii, const float* viewspace_grad,
                                        const float* viewspace_grad_abs, float* max_radii2D, float* xyz_gradient_accum,
                                        float* xyz_gradient_accum_abs, float* denom, float* denom_abs, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  if (P < 0) { ibgs_set_error("P must be >= 0"); return IBGS_EINVAL; }
  if (P == 0) return IBGS_OK;
  if (!radii || !viewspace_grad || !viewspace_grad_abs || !max_radii2D || !xyz_gradient_accum || !xyz_gradient_accum_abs ||
      !denom || !denom_abs) {
    ibgs_set_error("null pointer");
    return IBGS_EINVAL;
  }
  densify_stats_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, radii, viewspace_grad, viewspace_grad_abs, max_radii2D,
                                                       xyz_gradient_accum, xyz_gradient_accum_abs, denom, denom_abs);
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}
