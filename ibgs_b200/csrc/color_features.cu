// color_features.cu -- the per-pixel front half of the colour-aggregation step.
//
// Reference behaviour (PyTorch ops, no CUDA of its own):
//   color_aggregation_network.py:196-206   fuse_color: valid = sum(cam_feat, views' 4 features) > 0,
//                                          residual = (warped - rendered) * valid, x = cat(residual, cam_feat)  [7 per view]
//   :121-133                               ColorFusionResidualNet.forward: per_view_mlp = Linear(7,32) ReLU Linear(32,32) ReLU
//                                          on every (pixel, view), mean / max over the views, then
//                                          cnn_input = cat(aggregated[32], ray_dir[3], c_3dgs[3]) as a (1, 38, H, W) grid
// i.e. ~25 elementwise / cat / permute kernels, two SGEMMs over (pixels x views) rows and their backward passes, all of
// which stream (pixels x views x 32)-float intermediates through HBM.  Here: ONE kernel per direction.  Planar inputs as
// the rasterizer writes them are read once (coalesced), the two layers run in registers with the weights broadcast from
// shared memory (packed FP32: two views of a pixel per FFMA2), and the conv decoder's input leaves in its final layout --
// NHWC, channel count padded to a multiple of 8, optionally bf16 -- so the tensor-core convolutions that follow need
// no layout or dtype conversion pass.  The backward recomputes the hidden layers (nothing but the inputs is saved) and
// forms the weight gradients per warp in registers: the warp's 32 pixels are staged in a shared-memory tile and every lane
// walks them -- pass A: lane j accumulates column j of dW2 = g2 (x) h1; pass B (the g2 rows overwritten by g1): row j of
// dW1 = g1 (x) x -- with the matrix products in packed FP32 (two outputs per FFMA2 from the register pair an LDS.128
// delivers); the block's four warps are summed through shared memory and there is one atomicAdd per (block, weight).
// 128 registers, 10 KB of tile per warp -> 4 CTAs / SM (ncu: bound by shared-memory latency, not by issue slots).
//
// Algorithmic bytes per pixel (V views): forward 4*(7V + 6) in, 2*CP (bf16) out; backward the same in + 2*CP, 4*(3V + 3) out.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int F = 32;        // per_view_feat_dim
constexpr int XIN = 7;       // 3 residual + 4 camera features
constexpr int MAXV = 5;

struct CfArgs {
  int N;                     // pixels
  int V;                     // views
  int mode;                  // 0 mean, 1 max
  int CP;                    // channel pitch of cnn_input
  const float *warped, *cam_feat, *rendered, *ray;
  const float *w1, *b1, *w2, *b2;
  void* out;
  const void* g_out;
  float *d_warped, *d_rendered, *d_w1, *d_b1, *d_w2, *d_b2;
};

typedef float2 f2;
__device__ __forceinline__ f2 bc2(float s) { return make_float2(s, s); }

// shared-memory weight block.  Both directions walk the HIDDEN index j in a real loop and keep everything indexed by the
// output feature i in registers, so layer 2 is stored transposed: row j of w2t holds w2[0..31][j] (what h1[j] multiplies in
// the forward, and what dots with g2 in the backward).
struct Weights {
  float w1[F][8];            // [j][k], padded to 8 inputs
  float b1[F];
  float w2t[F][F];           // [j][i] = w2[i][j]
  float b2[F];
};

__device__ __forceinline__ void load_weights(Weights& s, const CfArgs& a, int tid, int nthreads) {
  for (int i = tid; i < F * 8; i += nthreads) s.w1[i >> 3][i & 7] = (i & 7) < XIN ? a.w1[(i >> 3) * XIN + (i & 7)] : 0.f;
  for (int i = tid; i < F * F; i += nthreads) s.w2t[i & 31][i >> 5] = a.w2[i];
  for (int i = tid; i < F; i += nthreads) { s.b1[i] = a.b1[i]; s.b2[i] = a.b2[i]; }
}

// x of one (pixel, view): color_aggregation_network.py:203-206
__device__ __forceinline__ float load_x(const CfArgs& a, int v, int p, const float rend[3], float x[8]) {
  const size_t N = (size_t)a.N;
  const float f0 = a.cam_feat[(4 * (size_t)v + 0) * N + p], f1 = a.cam_feat[(4 * (size_t)v + 1) * N + p];
  const float f2_ = a.cam_feat[(4 * (size_t)v + 2) * N + p], f3 = a.cam_feat[(4 * (size_t)v + 3) * N + p];
  const float valid = (((f0 + f1) + f2_) + f3) > 0.0f ? 1.0f : 0.0f;
#pragma unroll
  for (int c = 0; c < 3; c++) x[c] = (a.warped[(3 * (size_t)v + c) * N + p] - rend[c]) * valid;
  x[3] = f0; x[4] = f1; x[5] = f2_; x[6] = f3; x[7] = 0.f;
  return valid;
}

// hidden unit j of layer 1: relu(b1[j] + w1[j] . x), two views packed / one view
__device__ __forceinline__ f2 hidden1_2(const Weights& s, int j, const f2 x[8]) {
  const float4 wa = *reinterpret_cast<const float4*>(&s.w1[j][0]);
  const float4 wb = *reinterpret_cast<const float4*>(&s.w1[j][4]);
  f2 acc = bc2(s.b1[j]);
  acc = __ffma2_rn(bc2(wa.x), x[0], acc);
  acc = __ffma2_rn(bc2(wa.y), x[1], acc);
  acc = __ffma2_rn(bc2(wa.z), x[2], acc);
  acc = __ffma2_rn(bc2(wa.w), x[3], acc);
  acc = __ffma2_rn(bc2(wb.x), x[4], acc);
  acc = __ffma2_rn(bc2(wb.y), x[5], acc);
  acc = __ffma2_rn(bc2(wb.z), x[6], acc);
  return make_float2(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f));
}
__device__ __forceinline__ float hidden1(const Weights& s, int j, const float x[8]) {
  const float4 wa = *reinterpret_cast<const float4*>(&s.w1[j][0]);
  const float4 wb = *reinterpret_cast<const float4*>(&s.w1[j][4]);
  float acc = s.b1[j];
  acc = fmaf(wa.x, x[0], acc); acc = fmaf(wa.y, x[1], acc); acc = fmaf(wa.z, x[2], acc); acc = fmaf(wa.w, x[3], acc);
  acc = fmaf(wb.x, x[4], acc); acc = fmaf(wb.y, x[5], acc); acc = fmaf(wb.z, x[6], acc);
  return fmaxf(acc, 0.f);
}
// acc[i] += w2[i][j] * h for all 32 outputs (row j of w2t: 8 broadcast LDS.128)
__device__ __forceinline__ void layer2_add_2(const Weights& s, int j, f2 h, f2 acc[F]) {
#pragma unroll
  for (int i = 0; i < F; i += 4) {
    const float4 w = *reinterpret_cast<const float4*>(&s.w2t[j][i]);
    acc[i] = __ffma2_rn(bc2(w.x), h, acc[i]);
    acc[i + 1] = __ffma2_rn(bc2(w.y), h, acc[i + 1]);
    acc[i + 2] = __ffma2_rn(bc2(w.z), h, acc[i + 2]);
    acc[i + 3] = __ffma2_rn(bc2(w.w), h, acc[i + 3]);
  }
}
// one view, outputs packed in pairs (acc[i / 2] = features i, i + 1): two outputs per FFMA2, the weight pair is the
// register pair the LDS.128 delivered
__device__ __forceinline__ void layer2_add(const Weights& s, int j, float h, f2 acc[F / 2]) {
  const f2 hh = bc2(h);
#pragma unroll
  for (int i = 0; i < F; i += 4) {
    const float4 w = *reinterpret_cast<const float4*>(&s.w2t[j][i]);
    acc[i / 2] = __ffma2_rn(make_float2(w.x, w.y), hh, acc[i / 2]);
    acc[i / 2 + 1] = __ffma2_rn(make_float2(w.z, w.w), hh, acc[i / 2 + 1]);
  }
}
__device__ __forceinline__ float& el(f2* a, int i) { return (i & 1) ? a[i >> 1].y : a[i >> 1].x; }

template <bool BF16>
__device__ __forceinline__ void store_channels(void* out, size_t p, int CP, const float agg[F], const float ray[3],
                                               const float col[3]) {
  if (BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + p * CP;   // CP % 8 == 0: 16-byte aligned rows
    uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      __nv_bfloat162 h[4];
#pragma unroll
      for (int k = 0; k < 4; k++) h[k] = __floats2bfloat162_rn(agg[8 * q + 2 * k], agg[8 * q + 2 * k + 1]);
      o4[q] = *reinterpret_cast<uint4*>(h);
    }
    __nv_bfloat162 t[4];
    t[0] = __floats2bfloat162_rn(ray[0], ray[1]);
    t[1] = __floats2bfloat162_rn(ray[2], col[0]);
    t[2] = __floats2bfloat162_rn(col[1], col[2]);
    t[3] = __floats2bfloat162_rn(0.f, 0.f);
    o4[4] = *reinterpret_cast<uint4*>(t);
    for (int c = 40; c < CP; c += 8) o4[c >> 3] = make_uint4(0, 0, 0, 0);
  } else {
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + p * CP);
#pragma unroll
    for (int q = 0; q < 8; q++) o4[q] = make_float4(agg[4 * q], agg[4 * q + 1], agg[4 * q + 2], agg[4 * q + 3]);
    o4[8] = make_float4(ray[0], ray[1], ray[2], col[0]);
    o4[9] = make_float4(col[1], col[2], 0.f, 0.f);
    for (int c = 40; c < CP; c += 4) o4[c >> 2] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <bool BF16>
__global__ void __launch_bounds__(128) color_features_forward_kernel(const CfArgs a) {
  __shared__ __align__(16) Weights s;
  load_weights(s, a, threadIdx.x, blockDim.x);
  __syncthreads();
  const float inv_v = 1.0f / (float)a.V;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < a.N; p += gridDim.x * blockDim.x) {
    float rend[3], ray[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { rend[c] = a.rendered[(size_t)c * a.N + p]; ray[c] = a.ray[(size_t)c * a.N + p]; }
    float agg[F];
#pragma unroll
    for (int i = 0; i < F; i++) agg[i] = 0.f;   // hidden activations are >= 0: zero is also the identity of max
    for (int v = 0; v < a.V; v += 2) {
      const bool two = v + 1 < a.V;
      float xa[8], xb[8];
      load_x(a, v, p, rend, xa);
      load_x(a, two ? v + 1 : v, p, rend, xb);
      f2 x2[8];
#pragma unroll
      for (int k = 0; k < 8; k++) x2[k] = make_float2(xa[k], xb[k]);
      f2 acc[F];
#pragma unroll
      for (int i = 0; i < F; i++) acc[i] = bc2(s.b2[i]);
#pragma unroll 2
      for (int j = 0; j < F; j++) layer2_add_2(s, j, hidden1_2(s, j, x2), acc);
#pragma unroll
      for (int i = 0; i < F; i++) {
        const float first = fmaxf(acc[i].x, 0.f);
        const float second = two ? fmaxf(acc[i].y, 0.f) : 0.f;
        agg[i] = a.mode == 0 ? agg[i] + (first + second) : fmaxf(agg[i], fmaxf(first, second));
      }
    }
    if (a.mode == 0) {
#pragma unroll
      for (int i = 0; i < F; i++) agg[i] *= inv_v;
    }
    store_channels<BF16>(a.out, (size_t)p, a.CP, agg, ray, rend);
  }
}

// ---- backward -------------------------------------------------------------------------------------------------------
constexpr int BW_WARPS = 4;
constexpr int PITCH = 36;   // floats per staged row: 16-byte aligned, (4 p + c) mod 32 spreads a quarter-warp's STS.128
struct WarpTile {
  float h1[32][PITCH];      // relu(layer 1) of the warp's 32 pixels (one view)
  float g[32][PITCH];       // gradient at layer 2's pre-activation (pass A), then at layer 1's (pass B)
  float x[32][8];
};
struct BwdShared {
  Weights w;
  WarpTile tile[BW_WARPS];
};

template <bool BF16>
__device__ __forceinline__ void load_grad(const void* g, size_t p, int CP, float ga[F], float gcol[3]) {
  if (BF16) {
    const uint4* g4 = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(g) + p * CP);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint4 u = g4[q];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 t = __bfloat1622float2(h[k]);
        ga[8 * q + 2 * k] = t.x;
        ga[8 * q + 2 * k + 1] = t.y;
      }
    }
    const uint4 u = g4[4];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    gcol[0] = __bfloat1622float2(h[1]).y;
    const float2 t = __bfloat1622float2(h[2]);
    gcol[1] = t.x;
    gcol[2] = t.y;
  } else {
    const float4* g4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g) + p * CP);
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const float4 t = g4[q];
      ga[4 * q] = t.x; ga[4 * q + 1] = t.y; ga[4 * q + 2] = t.z; ga[4 * q + 3] = t.w;
    }
    gcol[0] = g4[8].w;
    const float4 t = g4[9];
    gcol[1] = t.x;
    gcol[2] = t.y;
  }
}

template <bool BF16>
__global__ void __launch_bounds__(32 * BW_WARPS, 4) color_features_backward_kernel(const CfArgs a) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  BwdShared& s = *reinterpret_cast<BwdShared*>(s_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  load_weights(s.w, a, tid, blockDim.x);
  __syncthreads();
  WarpTile& t = s.tile[warp];
  const float inv_v = 1.0f / (float)a.V;

  // this lane's slices of the weight gradients: column `lane` of dW2, row `lane` of dW1, element `lane` of the biases
  f2 dw2c[F / 2], dw1r[4];
  float db1 = 0.f, db2 = 0.f;
#pragma unroll
  for (int i = 0; i < F / 2; i++) dw2c[i] = bc2(0.f);
#pragma unroll
  for (int k = 0; k < 4; k++) dw1r[k] = bc2(0.f);

  const int nbatch = (a.N + 31) / 32;
  for (int batch = blockIdx.x * BW_WARPS + warp; batch < nbatch; batch += gridDim.x * BW_WARPS) {
    const int p = batch * 32 + lane;
    const bool live = p < a.N;
    const int pc = live ? p : a.N - 1;
    float rend[3], ga[F], gcol[3];
#pragma unroll
    for (int c = 0; c < 3; c++) rend[c] = a.rendered[(size_t)c * a.N + pc];
    load_grad<BF16>(a.g_out, (size_t)pc, a.CP, ga, gcol);
    if (!live) {
#pragma unroll
      for (int i = 0; i < F; i++) ga[i] = 0.f;
    }
    // max mode: the view every feature's maximum came from (first one on ties, like a running `>` comparison)
    uint32_t arg_b[4] = {0u, 0u, 0u, 0u};   // 8 features x 4 bits each
    if (a.mode == 1) {
      float best[F];
#pragma unroll
      for (int i = 0; i < F; i++) best[i] = 0.f;
      for (int v = 0; v < a.V; v++) {
        float xa[8];
        load_x(a, v, pc, rend, xa);
        f2 acc[F / 2];
#pragma unroll
        for (int i = 0; i < F; i++) el(acc, i) = s.w.b2[i];
#pragma unroll 2
        for (int j = 0; j < F; j++) layer2_add(s.w, j, hidden1(s.w, j, xa), acc);
#pragma unroll
        for (int i = 0; i < F; i++) {
          if (el(acc, i) > best[i]) {
            best[i] = el(acc, i);
            arg_b[i >> 3] = (arg_b[i >> 3] & ~(0xfu << (4 * (i & 7)))) | ((uint32_t)v << (4 * (i & 7)));
          }
        }
      }
    }
    float drend[3] = {gcol[0], gcol[1], gcol[2]};   // d cnn_input / d c_3dgs is the identity
    for (int v = 0; v < a.V; v++) {
      float x[8];
      const float valid = load_x(a, v, pc, rend, x);
      __syncwarp();   // the previous view's pass B is done with the tile
      // recompute: hidden layer 1 goes straight into this lane's row of the staging tile, layer 2 accumulates
      f2 g2[F / 2];
#pragma unroll
      for (int i = 0; i < F; i++) el(g2, i) = s.w.b2[i];
      for (int j0 = 0; j0 < F; j0 += 4) {
        float h[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          h[u] = hidden1(s.w, j0 + u, x);
          layer2_add(s.w, j0 + u, h[u], g2);
        }
        *reinterpret_cast<float4*>(&t.h1[lane][j0]) = make_float4(h[0], h[1], h[2], h[3]);
      }
      // g2 = d agg / d h2 * relu'
#pragma unroll
      for (int i = 0; i < F; i++) {
        const float g = a.mode == 0 ? ga[i] * inv_v : (((arg_b[i >> 3] >> (4 * (i & 7))) & 0xfu) == (uint32_t)v ? ga[i] : 0.f);
        el(g2, i) = el(g2, i) > 0.f ? g : 0.f;
      }
      // stage g2 and x next to h1, then pass A: dW2 += g2 (x) h1 over the warp's 32 pixels
#pragma unroll
      for (int q = 0; q < 8; q++)
        *reinterpret_cast<float4*>(&t.g[lane][4 * q]) = make_float4(g2[2 * q].x, g2[2 * q].y, g2[2 * q + 1].x, g2[2 * q + 1].y);
      *reinterpret_cast<float4*>(&t.x[lane][0]) = make_float4(x[0], x[1], x[2], x[3]);
      *reinterpret_cast<float4*>(&t.x[lane][4]) = make_float4(x[4], x[5], x[6], 0.f);
      __syncwarp();
      // (a dead lane's g2 / g1 rows are exact zeros: ga was zeroed)
#pragma unroll 4
      for (int q = 0; q < 32; q++) {
        const float hj = t.h1[q][lane];
        db2 += t.g[q][lane];
#pragma unroll
        for (int i = 0; i < F; i += 4) {
          const float4 g = *reinterpret_cast<const float4*>(&t.g[q][i]);
          dw2c[i / 2] = __ffma2_rn(make_float2(g.x, g.y), bc2(hj), dw2c[i / 2]);
          dw2c[i / 2 + 1] = __ffma2_rn(make_float2(g.z, g.w), bc2(hj), dw2c[i / 2 + 1]);
        }
      }
      __syncwarp();   // every lane is done with the g2 rows: the same tile now takes g1
      // g1[j] = relu'(h1[j]) * sum_i w2[i][j] g2[i];  d x[0..2] = sum_j w1[j][c] g1[j]  -> residual -> warped / rendered
      float dx[3] = {0.f, 0.f, 0.f};
      for (int j0 = 0; j0 < F; j0 += 4) {
        const float4 hq = *reinterpret_cast<const float4*>(&t.h1[lane][j0]);
        const float hh[4] = {hq.x, hq.y, hq.z, hq.w};
        float g1[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          f2 acc0 = bc2(0.f), acc1 = bc2(0.f);
#pragma unroll
          for (int i = 0; i < F; i += 4) {
            const float4 w = *reinterpret_cast<const float4*>(&s.w.w2t[j0 + u][i]);
            acc0 = __ffma2_rn(make_float2(w.x, w.y), g2[i / 2], acc0);
            acc1 = __ffma2_rn(make_float2(w.z, w.w), g2[i / 2 + 1], acc1);
          }
          const f2 sum2 = __fadd2_rn(acc0, acc1);
          g1[u] = hh[u] > 0.f ? sum2.x + sum2.y : 0.f;
          const float4 w = *reinterpret_cast<const float4*>(&s.w.w1[j0 + u][0]);
          dx[0] = fmaf(w.x, g1[u], dx[0]);
          dx[1] = fmaf(w.y, g1[u], dx[1]);
          dx[2] = fmaf(w.z, g1[u], dx[2]);
        }
        *reinterpret_cast<float4*>(&t.g[lane][j0]) = make_float4(g1[0], g1[1], g1[2], g1[3]);
      }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float d = dx[c] * valid;
        if (a.d_warped && live) a.d_warped[(3 * (size_t)v + c) * a.N + p] = d;
        drend[c] -= d;
      }
      __syncwarp();
      // pass B: dW1 += g1 (x) x
#pragma unroll 4
      for (int q = 0; q < 32; q++) {
        const float g1i = t.g[q][lane];
        db1 += g1i;
        const float4 xa = *reinterpret_cast<const float4*>(&t.x[q][0]);
        const float4 xb = *reinterpret_cast<const float4*>(&t.x[q][4]);
        dw1r[0] = __ffma2_rn(bc2(g1i), make_float2(xa.x, xa.y), dw1r[0]);
        dw1r[1] = __ffma2_rn(bc2(g1i), make_float2(xa.z, xa.w), dw1r[1]);
        dw1r[2] = __ffma2_rn(bc2(g1i), make_float2(xb.x, xb.y), dw1r[2]);
        dw1r[3] = __ffma2_rn(bc2(g1i), make_float2(xb.z, xb.w), dw1r[3]);
      }
    }
    if (a.d_rendered && live) {
#pragma unroll
      for (int c = 0; c < 3; c++) a.d_rendered[(size_t)c * a.N + p] = drend[c];
    }
  }
  // ---- flush: sum the block's four warps through shared memory (the staging tiles are free now), one atomic per
  // (block, weight) ----
  __syncthreads();
  float* red = reinterpret_cast<float*>(&s.tile[0]);   // [BW_WARPS][F + 8 + 2][32]
  constexpr int ROWS = F + 8 + 2;
#pragma unroll
  for (int i = 0; i < F; i++) red[(warp * ROWS + i) * 32 + lane] = el(dw2c, i);
#pragma unroll
  for (int k = 0; k < 8; k++) red[(warp * ROWS + F + k) * 32 + lane] = el(dw1r, k);
  red[(warp * ROWS + F + 8) * 32 + lane] = db1;
  red[(warp * ROWS + F + 9) * 32 + lane] = db2;
  __syncthreads();
  for (int e = tid; e < ROWS * 32; e += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < BW_WARPS; w++) v += red[w * ROWS * 32 + e];
    const int row = e >> 5, ln = e & 31;
    if (row < F) atomicAdd(a.d_w2 + row * F + ln, v);                    // dW2[i = row][j = ln]
    else if (row < F + XIN) atomicAdd(a.d_w1 + ln * XIN + (row - F), v); // dW1[i = ln][k]
    else if (row == F + 8) atomicAdd(a.d_b1 + ln, v);
    else if (row == F + 9) atomicAdd(a.d_b2 + ln, v);
  }
}

int fill(CfArgs& c, const IbgsColorFeatArgs* a) {
  if (!a) { ibgs_set_error("args is NULL"); return IBGS_EINVAL; }
  if (a->height <= 0 || a->width <= 0 || (long long)a->height * a->width > 0x7fffffffLL) {
    ibgs_set_error("bad image shape %dx%d", a->width, a->height);
    return IBGS_EINVAL;
  }
  if (a->n_views < 1 || a->n_views > MAXV) { ibgs_set_error("n_views must be 1..%d, got %d", MAXV, a->n_views); return IBGS_EINVAL; }
  if (a->mode != 0 && a->mode != 1) { ibgs_set_error("mode must be 0 (mean) or 1 (max), got %d", a->mode); return IBGS_EINVAL; }
  if (a->channel_pitch < 40 || a->channel_pitch % 8) {
    ibgs_set_error("channel_pitch must be a multiple of 8 and >= 40 (38 channels + padding), got %d", a->channel_pitch);
    return IBGS_EINVAL;
  }
  if (!a->warped || !a->cam_feat || !a->rendered || !a->camera_ray || !a->w1 || !a->b1 || !a->w2 || !a->b2) {
    ibgs_set_error("input pointers must not be NULL");
    return IBGS_EINVAL;
  }
  c.N = a->height * a->width;
  c.V = a->n_views;
  c.mode = a->mode;
  c.CP = a->channel_pitch;
  c.warped = a->warped; c.cam_feat = a->cam_feat; c.rendered = a->rendered; c.ray = a->camera_ray;
  c.w1 = a->w1; c.b1 = a->b1; c.w2 = a->w2; c.b2 = a->b2;
  c.out = a->cnn_input;
  c.g_out = a->g_cnn_input;
  c.d_warped = a->d_warped; c.d_rendered = a->d_rendered;
  c.d_w1 = a->d_w1; c.d_b1 = a->d_b1; c.d_w2 = a->d_w2; c.d_b2 = a->d_b2;
  return IBGS_OK;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

extern "C" int ibgs_color_features_forward(const IbgsColorFeatArgs* a, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  CfArgs c;
  int rc = fill(c, a);
  if (rc != IBGS_OK) return rc;
  if (!a->cnn_input) { ibgs_set_error("cnn_input must not be NULL"); return IBGS_EINVAL; }
  if ((uintptr_t)a->cnn_input % 16) { ibgs_set_error("cnn_input must be 16-byte aligned"); return IBGS_EINVAL; }
  const int blocks = min((c.N + 127) / 128, sm_count() * 8);
  ProfScope prof(PROF_COLORFEAT_FWD, s);
  if (a->bf16) color_features_forward_kernel<true><<<blocks, 128, 0, s>>>(c);
  else color_features_forward_kernel<false><<<blocks, 128, 0, s>>>(c);
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}

extern "C" int ibgs_color_features_backward(const IbgsColorFeatArgs* a, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  CfArgs c;
  int rc = fill(c, a);
  if (rc != IBGS_OK) return rc;
  if (!a->g_cnn_input || !a->d_w1 || !a->d_b1 || !a->d_w2 || !a->d_b2) {
    ibgs_set_error("g_cnn_input and the four weight-gradient pointers must not be NULL");
    return IBGS_EINVAL;
  }
  if ((uintptr_t)a->g_cnn_input % 16) { ibgs_set_error("g_cnn_input must be 16-byte aligned"); return IBGS_EINVAL; }
  const int smem = (int)sizeof(BwdShared);
  const int nbatch = (c.N + 31) / 32;
  const int blocks = min((nbatch + BW_WARPS - 1) / BW_WARPS, sm_count() * 4);
  ProfScope prof(PROF_COLORFEAT_BWD, s);
  if (a->bf16) {
    CUDA_TRY(cudaFuncSetAttribute(color_features_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    color_features_backward_kernel<true><<<blocks, 32 * BW_WARPS, smem, s>>>(c);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(color_features_backward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    color_features_backward_kernel<false><<<blocks, 32 * BW_WARPS, smem, s>>>(c);
  }
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}
