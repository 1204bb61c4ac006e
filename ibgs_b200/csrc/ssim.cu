// ssim.cu -- fused SSIM map and its backward (the loss step next to the rasterizer).
//
// Reference behaviour: utils/loss_utils.py:24-65 (`gaussian`, `create_window`, `ssim`, `_ssim`), :67-90
// (`compute_photometric_ssim`) and :92-117 (`ssim2`): five depthwise 11x11 convolutions (zero padding 5, Gaussian
// window sigma 1.5, built in float32) of img1, img2, img1^2, img2^2, img1*img2, then
//   ssim = (2 mu1 mu2 + C1)(2 sigma12 + C2) / ((mu1^2 + mu2^2 + C1)(sigma1^2 + sigma2^2 + C2)),  C1 = 0.01^2, C2 = 0.03^2.
// train.py evaluates it six times per iteration (:302,330,355), each as 5 cuDNN depthwise convolutions + ~15
// elementwise kernels forward and about twice that backward, all round-tripping full-resolution planes through HBM.
//
// Here: ONE launch per direction.  A CTA owns a 32x32 tile of one plane: the (32+10)^2 halo tiles of both images are
// staged in shared memory once, the window is applied separably (11 horizontal taps into shared memory, 11 vertical
// taps into registers), and the SSIM value plus the partial derivatives of the map with respect to the convolution
// outputs are written in the same pass.  The backward convolves (cotangent x partial) planes with the same (symmetric,
// zero-padded) window: dL/dimg1 = W*(g dm/dmu1) + 2 img1 W*(g dm/de11) + img2 W*(g dm/de12), and symmetrically for img2.
// HBM traffic per pixel and plane: forward 8 B read + 4 B map + 12-16 B partials; backward 24-28 B read + 4-8 B written.
#include "common.cuh"

namespace {

constexpr int SS_BX = 32, SS_BY = 32, SS_R = 5;
constexpr int SS_SY = SS_BY + 2 * SS_R;        // 42 staged rows
constexpr int SS_SXP = 48;                     // staged row pitch (42 used; 16-byte aligned float4 reads)

struct SsimWindow { float w[11]; };

struct SsimArgs {
  int C, H, W;
  const float* img1;
  const float* img2;
  float* map;
  float* p_mu1; float* p_e11; float* p_e12; float* p_mu2;
  const float* g;      // cotangent of the map; one broadcast device scalar if g_scalar; NULL = 1
  int g_scalar;
  float g_scale;
  float* d_img1;
  float* d_img2;
  SsimWindow win;
};

// horizontal 11-tap pass over NA staged arrays: thread -> 4 adjacent output columns of one staged row
// (4 x LDS.128 per array), products formed on the fly by `Op`.
constexpr size_t ssim_smem_bytes(int na, int no) {
  return sizeof(float) * ((size_t)na * SS_SY * SS_SXP + (size_t)no * SS_SY * SS_BX);
}

template <int NA, int NO, typename Op>
__device__ __forceinline__ void horizontal_pass(const float (*raw)[SS_SY][SS_SXP], float (*hs)[SS_SY][SS_BX],
                                                const SsimWindow& win, Op op) {
  for (int item = threadIdx.x; item < SS_SY * (SS_BX / 4); item += blockDim.x) {
    const int row = item >> 3, c4 = (item & 7) * 4;
    float v[NA][16];
#pragma unroll
    for (int a = 0; a < NA; a++) {
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float4 t = *reinterpret_cast<const float4*>(&raw[a][row][c4 + 4 * q]);
        v[a][4 * q] = t.x; v[a][4 * q + 1] = t.y; v[a][4 * q + 2] = t.z; v[a][4 * q + 3] = t.w;
      }
    }
    float acc[NO][4];
#pragma unroll
    for (int o = 0; o < NO; o++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[o][j] = 0.f;
#pragma unroll
    for (int k = 0; k < 11; k++) {
      const float wk = win.w[k];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float in[NA], out[NO];
#pragma unroll
        for (int a = 0; a < NA; a++) in[a] = v[a][j + k];
        op(in, out);
#pragma unroll
        for (int o = 0; o < NO; o++) acc[o][j] = fmaf(wk, out[o], acc[o][j]);
      }
    }
#pragma unroll
    for (int o = 0; o < NO; o++)
      *reinterpret_cast<float4*>(&hs[o][row][c4]) = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
  }
}

// vertical 11-tap pass: thread (tx, ty) -> column tx, rows 4ty..4ty+3 of the tile
template <int NO>
__device__ __forceinline__ void vertical_pass(const float (*hs)[SS_SY][SS_BX], const SsimWindow& win, int tx, int ty,
                                              float (*res)[4]) {
#pragma unroll
  for (int o = 0; o < NO; o++) {
    float col[14];
#pragma unroll
    for (int j = 0; j < 14; j++) col[j] = hs[o][4 * ty + j][tx];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 11; k++) s = fmaf(win.w[k], col[r + k], s);
      res[o][r] = s;
    }
  }
}

// Staging: thread t takes elements t, t+256, ... of the [42][48] staged tile; (row, col) advance incrementally
// (256 = 5*48 + 16), so there is one division per thread, not one per element.  All global loads of the tile are
// issued before the first shared-memory store (two fully unrolled phases): with the loads inside one rolled loop the
// CTA paid a dependent global round trip per iteration.
template <int NA, typename Load>
__device__ __forceinline__ void stage_tiles(float (*raw)[SS_SY][SS_SXP], int x0, int y0, int W, int H, Load load) {
  constexpr int ITERS = (SS_SY * SS_SXP + 255) / 256;
  static_assert(256 % SS_SXP == 16 && 256 / SS_SXP == 5, "incremental (row, col) update below");
  float v[ITERS][NA];
  const int row0 = threadIdx.x / SS_SXP, col0 = threadIdx.x - row0 * SS_SXP;
  int row = row0, col = col0;
#pragma unroll
  for (int it = 0; it < ITERS; it++) {
    const int gx = x0 - SS_R + col, gy = y0 - SS_R + row;
    // zero padding (conv2d padding=5); columns 42..47 of the padded pitch are zero as well
    const bool in = row < SS_SY && col < SS_BX + 2 * SS_R && gx >= 0 && gx < W && gy >= 0 && gy < H;
#pragma unroll
    for (int a = 0; a < NA; a++) v[it][a] = 0.f;
    if (in) load((size_t)gy * W + gx, v[it]);
    col += 16;
    row += 5;
    if (col >= SS_SXP) { col -= SS_SXP; row += 1; }
  }
  row = row0;
  col = col0;
#pragma unroll
  for (int it = 0; it < ITERS; it++) {
    if (row < SS_SY) {
#pragma unroll
      for (int a = 0; a < NA; a++) raw[a][row][col] = v[it][a];
    }
    col += 16;
    row += 5;
    if (col >= SS_SXP) { col -= SS_SXP; row += 1; }
  }
}

// PART: 0 = map only (inference), 1 = + partials for d/dimg1, 2 = + partials for both images
template <int PART>
__global__ void __launch_bounds__(256) ssim_forward_kernel(const SsimArgs a) {
  extern __shared__ __align__(16) float ssim_smem[];
  auto raw = reinterpret_cast<float(*)[SS_SY][SS_SXP]>(ssim_smem);
  auto hs = reinterpret_cast<float(*)[SS_SY][SS_BX]>(ssim_smem + 2 * SS_SY * SS_SXP);
  const int x0 = blockIdx.x * SS_BX, y0 = blockIdx.y * SS_BY;
  const size_t plane = (size_t)blockIdx.z * a.H * a.W;
  const float* x = a.img1 + plane;
  const float* y = a.img2 + plane;
  stage_tiles<2>(raw, x0, y0, a.W, a.H, [&](size_t o, float* v) { v[0] = x[o]; v[1] = y[o]; });
  __syncthreads();
  horizontal_pass<2, 5>(raw, hs, a.win, [](const float* in, float* out) {
    out[0] = in[0]; out[1] = in[1]; out[2] = in[0] * in[0]; out[3] = in[1] * in[1]; out[4] = in[0] * in[1];
  });
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float res[5][4];
  vertical_pass<5>(hs, a.win, tx, ty, res);
  const int gx = x0 + tx;
  if (gx >= a.W) return;
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int gy = y0 + 4 * ty + r;
    if (gy >= a.H) break;
    const size_t o = plane + (size_t)gy * a.W + gx;
    const float mu1 = res[0][r], mu2 = res[1][r];
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float s1 = res[2][r] - mu1_sq, s2 = res[3][r] - mu2_sq, s12 = res[4][r] - mu12;
    const float A = 2.f * mu12 + C1, B = 2.f * s12 + C2, Cc = mu1_sq + mu2_sq + C1, D = s1 + s2 + C2;
    // loss_utils.py:60.  Two approximate reciprocals (MUFU.RCP, ~1 ulp) replace the four IEEE divisions of the
    // straightforward form: the map and its partials stay within a few ulp, far inside the 2e-5 / 1e-4 gates.
    const float inv_C = __fdividef(1.f, Cc), inv_D = __fdividef(1.f, D);
    const float inv_CD = inv_C * inv_D;
    const float m = (A * B) * inv_CD;
    a.map[o] = m;
    if (PART >= 1) {
      // partials of m with respect to the convolution outputs mu1, e11 = W*x^2, e12 = W*xy (and mu2, e22), with
      // sigma1^2 = e11 - mu1^2, sigma12 = e12 - mu1 mu2 substituted
      const float m_Cc = m * inv_C, m_D = m * inv_D;
      const float dB = 2.f * A * inv_CD;          // dm/de12 (= 2 * dm/dB)
      a.p_e11[o] = -m_D;
      a.p_e12[o] = dB;
      a.p_mu1[o] = 2.f * mu2 * B * inv_CD - 2.f * mu1 * m_Cc + 2.f * mu1 * m_D - mu2 * dB;
      if (PART >= 2) {  // dm/de22 = dm/de11 (the map depends on sigma1^2 + sigma2^2 only): no plane of its own
        a.p_mu2[o] = 2.f * mu1 * B * inv_CD - 2.f * mu2 * m_Cc + 2.f * mu2 * m_D - mu1 * dB;
      }
    }
  }
}

// BOTH: also the gradient with respect to img2
template <bool BOTH>
__global__ void __launch_bounds__(256, 4) ssim_backward_kernel(const SsimArgs a) {
  constexpr int NA = BOTH ? 4 : 3;
  extern __shared__ __align__(16) float ssim_smem[];
  auto raw = reinterpret_cast<float(*)[SS_SY][SS_SXP]>(ssim_smem);
  auto hs = reinterpret_cast<float(*)[SS_SY][SS_BX]>(ssim_smem + NA * SS_SY * SS_SXP);
  const int x0 = blockIdx.x * SS_BX, y0 = blockIdx.y * SS_BY;
  const size_t plane = (size_t)blockIdx.z * a.H * a.W;
  // this thread's own pixels of both images are needed only by the last statement: request them first
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int gx = x0 + tx;
  float xv[4], yv[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int gy = y0 + 4 * ty + r;
    const bool in = gx < a.W && gy < a.H;
    const size_t o = plane + (size_t)gy * a.W + gx;
    xv[r] = in ? a.img1[o] : 0.f;
    yv[r] = in ? a.img2[o] : 0.f;
  }
  const float gs = (a.g && a.g_scalar) ? a.g[0] * a.g_scale : a.g_scale;
  const float* gmap = (a.g && !a.g_scalar) ? a.g + plane : nullptr;
  stage_tiles<NA>(raw, x0, y0, a.W, a.H, [&](size_t o, float* v) {
    const float g = gmap ? gmap[o] * gs : gs;
    v[0] = g * a.p_mu1[plane + o];
    v[1] = g * a.p_e11[plane + o];
    v[2] = g * a.p_e12[plane + o];
    if (BOTH) v[3] = g * a.p_mu2[plane + o];
  });
  __syncthreads();
  horizontal_pass<NA, NA>(raw, hs, a.win, [](const float* in, float* out) {
#pragma unroll
    for (int i = 0; i < NA; i++) out[i] = in[i];
  });
  __syncthreads();
  float res[NA][4];
  vertical_pass<NA>(hs, a.win, tx, ty, res);
  if (gx >= a.W) return;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int gy = y0 + 4 * ty + r;
    if (gy >= a.H) break;
    const size_t o = plane + (size_t)gy * a.W + gx;
    const float x = xv[r], y = yv[r];
    a.d_img1[o] = res[0][r] + 2.f * x * res[1][r] + y * res[2][r];
    if (BOTH) a.d_img2[o] = res[3][r] + 2.f * y * res[1][r] + x * res[2][r];
  }
}

// utils/loss_utils.py:24-26: float32 exp values divided by their float32 sum
SsimWindow make_window() {
  SsimWindow w;
  float g[11], sum = 0.f;
  for (int i = 0; i < 11; i++) {
    g[i] = (float)exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5));
    sum += g[i];
  }
  for (int i = 0; i < 11; i++) w.w[i] = g[i] / sum;
  return w;
}

int check(const IbgsSsimArgs* a) {
  if (!a) { ibgs_set_error("args is NULL"); return IBGS_EINVAL; }
  if (a->planes < 0 || a->height <= 0 || a->width <= 0) {
    ibgs_set_error("bad SSIM shape: %d planes of %dx%d", a->planes, a->width, a->height);
    return IBGS_EINVAL;
  }
  if (a->planes > 65535) { ibgs_set_error("at most 65535 planes per call, got %d", a->planes); return IBGS_EINVAL; }
  return IBGS_OK;
}

SsimArgs convert(const IbgsSsimArgs* f) {
  SsimArgs a;
  a.C = f->planes; a.H = f->height; a.W = f->width;
  a.img1 = f->img1; a.img2 = f->img2; a.map = f->ssim_map;
  a.p_mu1 = f->dm_dmu1; a.p_e11 = f->dm_de11; a.p_e12 = f->dm_de12; a.p_mu2 = f->dm_dmu2;
  a.g = f->dL_dmap; a.g_scalar = f->dL_dmap_is_scalar; a.g_scale = f->dL_dmap_scale;
  a.d_img1 = f->dL_dimg1; a.d_img2 = f->dL_dimg2;
  a.win = make_window();
  return a;
}

}  // namespace

extern "C" int ibgs_ssim_forward(const IbgsSsimArgs* f, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  int rc = check(f);
  if (rc != IBGS_OK) return rc;
  if (f->planes == 0) return IBGS_OK;
  if (!f->img1 || !f->img2 || !f->ssim_map) { ibgs_set_error("img1 / img2 / ssim_map must not be NULL"); return IBGS_EINVAL; }
  const bool p1 = f->dm_dmu1 && f->dm_de11 && f->dm_de12;
  const bool p2 = p1 && f->dm_dmu2;
  if (!p1 && (f->dm_dmu1 || f->dm_de11 || f->dm_de12 || f->dm_dmu2)) {
    ibgs_set_error("partial-derivative planes: give dm_dmu1 + dm_de11 + dm_de12 (and optionally dm_dmu2) or none");
    return IBGS_EINVAL;
  }
  const SsimArgs a = convert(f);
  dim3 grid((a.W + SS_BX - 1) / SS_BX, (a.H + SS_BY - 1) / SS_BY, a.C);
  const size_t smem = ssim_smem_bytes(2, 5);
  ProfScope prof(PROF_SSIM_FWD, s);
  if (p2) ssim_forward_kernel<2><<<grid, 256, smem, s>>>(a);
  else if (p1) ssim_forward_kernel<1><<<grid, 256, smem, s>>>(a);
  else ssim_forward_kernel<0><<<grid, 256, smem, s>>>(a);
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}

extern "C" int ibgs_ssim_backward(const IbgsSsimArgs* f, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  int rc = check(f);
  if (rc != IBGS_OK) return rc;
  if (f->planes == 0) return IBGS_OK;
  if (!f->img1 || !f->img2 || !f->dm_dmu1 || !f->dm_de11 || !f->dm_de12 || !f->dL_dimg1) {
    ibgs_set_error("img1 / img2 / dm_dmu1 / dm_de11 / dm_de12 / dL_dimg1 must not be NULL");
    return IBGS_EINVAL;
  }
  const bool both = f->dL_dimg2 != nullptr;
  if (both && !f->dm_dmu2) {
    ibgs_set_error("dL_dimg2 needs dm_dmu2 from the forward");
    return IBGS_EINVAL;
  }
  const SsimArgs a = convert(f);
  dim3 grid((a.W + SS_BX - 1) / SS_BX, (a.H + SS_BY - 1) / SS_BY, a.C);
  ProfScope prof(PROF_SSIM_BWD, s);
  if (both) {
    // 4 staged + 4 filtered arrays = 54 KB of shared memory: above the 48 KB default, opt in (per device)
    const size_t smem = ssim_smem_bytes(4, 4);
    CUDA_TRY(cudaFuncSetAttribute(ssim_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ssim_backward_kernel<true><<<grid, 256, smem, s>>>(a);
  } else {
    ssim_backward_kernel<false><<<grid, 256, ssim_smem_bytes(3, 3), s>>>(a);
  }
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}
