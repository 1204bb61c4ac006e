// depth_normal.cu -- unit normal map of a rendered depth image, forward and backward (part of the per-view glue of
// gaussian_renderer.render()).
//
// Reference behaviour (torch ops): render_normal (gaussian_renderer/__init__.py:15-26) -> normal_from_depth_image ->
// depth2point_world + depth_pcd2normal (utils/graphics_utils.py:38-75): back-project every pixel with the pinhole
// intrinsics (Camera.get_calib_matrix_nerf, scene/cameras.py:115-118), n = cross(P(y,x+1) - P(y,x-1), P(y-1,x) - P(y+1,x)),
// F.normalize(n) (eps 1e-12), zero on the 1-pixel border; then n / (|n| + 1e-8) (gaussian_renderer/__init__.py:332-335).
// ~12 elementwise torch kernels forward and ~25 in autograd's backward over (3, H, W) tensors; here one kernel each.
// The backward recomputes the cross product per pixel and scatters the four point gradients into the depth gradient
// (4 float atomics per interior pixel).
#include "common.cuh"

namespace {

struct DnArgs {
  const float* depth;
  float* normal;          // forward out [3][H][W]
  const float* g_normal;  // backward in
  float* g_depth;         // backward out [H][W], zeroed before the launch
  int H, W;
  float fx, fy, cx, cy;
};

__device__ __forceinline__ void point(const DnArgs& a, int y, int x, float p[3]) {
  const float d = a.depth[(size_t)y * a.W + x];
  p[0] = d * (((float)x - a.cx) / a.fx);
  p[1] = d * (((float)y - a.cy) / a.fy);
  p[2] = d;
}

// n = cross(l2r, b2t); s1 = max(|n|, 1e-12); u = n / s1; s2 = |u| + 1e-8; out = u / s2
__device__ __forceinline__ void normal_at(const DnArgs& a, int y, int x, float l2r[3], float b2t[3], float n[3]) {
  float pr[3], pl[3], pt[3], pb[3];
  point(a, y, x + 1, pr);
  point(a, y, x - 1, pl);
  point(a, y - 1, x, pt);
  point(a, y + 1, x, pb);
#pragma unroll
  for (int k = 0; k < 3; k++) { l2r[k] = pr[k] - pl[k]; b2t[k] = pt[k] - pb[k]; }
  n[0] = l2r[1] * b2t[2] - l2r[2] * b2t[1];
  n[1] = l2r[2] * b2t[0] - l2r[0] * b2t[2];
  n[2] = l2r[0] * b2t[1] - l2r[1] * b2t[0];
}

__global__ void __launch_bounds__(256) depth_normal_forward_kernel(const DnArgs a) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= a.W) return;
  const size_t HW = (size_t)a.H * a.W, pix = (size_t)y * a.W + x;
  float o[3] = {0.f, 0.f, 0.f};
  if (x > 0 && x < a.W - 1 && y > 0 && y < a.H - 1) {
    float l2r[3], b2t[3], n[3];
    normal_at(a, y, x, l2r, b2t, n);
    const float s1 = fmaxf(sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1e-12f);
    const float u[3] = {n[0] / s1, n[1] / s1, n[2] / s1};
    const float s2 = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) + 1e-8f;
#pragma unroll
    for (int k = 0; k < 3; k++) o[k] = u[k] / s2;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) a.normal[k * HW + pix] = o[k];
}

__global__ void __launch_bounds__(256) depth_normal_backward_kernel(const DnArgs a) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x <= 0 || x >= a.W - 1 || y <= 0 || y >= a.H - 1) return;
  const size_t HW = (size_t)a.H * a.W, pix = (size_t)y * a.W + x;
  float l2r[3], b2t[3], n[3];
  normal_at(a, y, x, l2r, b2t, n);
  const float g[3] = {a.g_normal[pix], a.g_normal[HW + pix], a.g_normal[2 * HW + pix]};
  const float len = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  const float s1 = fmaxf(len, 1e-12f);
  const float u[3] = {n[0] / s1, n[1] / s1, n[2] / s1};
  const float lu = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  const float s2 = lu + 1e-8f;
  // out = u / s2, s2 = |u| + eps:  d u = g / s2 - u (g . u) / (s2^2 |u|)     (|u| = 0 only for a zero cross product)
  const float gu = g[0] * u[0] + g[1] * u[1] + g[2] * u[2];
  const float c2 = lu > 0.f ? gu / (s2 * s2 * lu) : 0.f;
  float du[3];
#pragma unroll
  for (int k = 0; k < 3; k++) du[k] = g[k] / s2 - u[k] * c2;
  // u = n / s1, s1 = max(|n|, eps):  d n = du / s1 - n (du . n) / (s1^2 |n|) when |n| > eps, du / eps otherwise
  float dn[3];
  if (len > 1e-12f) {
    const float dun = du[0] * n[0] + du[1] * n[1] + du[2] * n[2];
    const float c1 = dun / (s1 * s1 * len);
#pragma unroll
    for (int k = 0; k < 3; k++) dn[k] = du[k] / s1 - n[k] * c1;
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) dn[k] = du[k] / s1;
  }
  // n = l2r x b2t:  d l2r = b2t x dn,  d b2t = dn x l2r
  const float da[3] = {b2t[1] * dn[2] - b2t[2] * dn[1], b2t[2] * dn[0] - b2t[0] * dn[2], b2t[0] * dn[1] - b2t[1] * dn[0]};
  const float db[3] = {dn[1] * l2r[2] - dn[2] * l2r[1], dn[2] * l2r[0] - dn[0] * l2r[2], dn[0] * l2r[1] - dn[1] * l2r[0]};
  // P(y, x) = d * ((x - cx) / fx, (y - cy) / fy, 1):  d depth = dP . that direction
  const float rxr = ((float)(x + 1) - a.cx) / a.fx, rxl = ((float)(x - 1) - a.cx) / a.fx, rx = ((float)x - a.cx) / a.fx;
  const float ryt = ((float)(y - 1) - a.cy) / a.fy, ryb = ((float)(y + 1) - a.cy) / a.fy, ry = ((float)y - a.cy) / a.fy;
  atomicAdd(a.g_depth + pix + 1, da[0] * rxr + da[1] * ry + da[2]);
  atomicAdd(a.g_depth + pix - 1, -(da[0] * rxl + da[1] * ry + da[2]));
  atomicAdd(a.g_depth + pix - a.W, db[0] * rx + db[1] * ryt + db[2]);
  atomicAdd(a.g_depth + pix + a.W, -(db[0] * rx + db[1] * ryb + db[2]));
}

int fill(DnArgs& a, const float* depth, int H, int W, float fx, float fy, float cx, float cy) {
  if (H <= 0 || W <= 0 || !depth || fx == 0.f || fy == 0.f) { ibgs_set_error("bad depth_normal arguments"); return IBGS_EINVAL; }
  a.depth = depth; a.H = H; a.W = W; a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy;
  a.normal = nullptr; a.g_normal = nullptr; a.g_depth = nullptr;
  return IBGS_OK;
}

}  // namespace

extern "C" int ibgs_depth_normal_forward(const float* depth, float* normal, int32_t H, int32_t W, float fx, float fy, float cx,
                                         float cy, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  DnArgs a;
  int rc = fill(a, depth, H, W, fx, fy, cx, cy);
  if (rc != IBGS_OK) return rc;
  if (!normal) { ibgs_set_error("normal must not be NULL"); return IBGS_EINVAL; }
  a.normal = normal;
  depth_normal_forward_kernel<<<dim3((W + 255) / 256, H), 256, 0, s>>>(a);
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}

extern "C" int ibgs_depth_normal_backward(const float* depth, const float* g_normal, float* g_depth, int32_t H, int32_t W,
                                          float fx, float fy, float cx, float cy, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  DnArgs a;
  int rc = fill(a, depth, H, W, fx, fy, cx, cy);
  if (rc != IBGS_OK) return rc;
  if (!g_normal || !g_depth) { ibgs_set_error("g_normal / g_depth must not be NULL"); return IBGS_EINVAL; }
  a.g_normal = g_normal;
  a.g_depth = g_depth;
  CUDA_TRY(cudaMemsetAsync(g_depth, 0, (size_t)H * W * sizeof(float), s));
  COUNT_LAUNCH();
  depth_normal_backward_kernel<<<dim3((W + 255) / 256, H), 256, 0, s>>>(a);
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}
