// nhwc_ops.cu -- the memory-bound glue between the conv decoder's convolutions, on the NHWC
// layout color_features.cu produces: 2x2 max pooling, nearest-neighbour upsampling, channel concatenation and the ReLU
// backward fused with the bias gradient, bf16 or float32.
//
// Reference behaviour (ConvDecoderAE.forward, color_aggregation_network.py:51-68, through torch):
//   p = nn.MaxPool2d(2)(e)                          kernel 2, stride 2, no padding, floor: out = in // 2; the FIRST maximum
//                                                   of a window in row-major order wins, NaN propagates
//   u = F.interpolate(b, size=e.shape[-2:], mode="nearest"); torch.cat([conv(u), e], 1)
//                                                   source index = min(int(floorf(dst * (float(in) / out))), in - 1)
// torch's generic NHWC kernels for these run far below the copy rate (max_pool_backward_nhwc 0.17 ms for 80 MB in + out,
// upsample 0.08 ms, a cat 0.12 ms, threshold_backward on the channel-sliced gradient of a cat 0.11 ms).  Here every
// thread moves 16-byte channel groups; one kernel serves the upsample (b = NULL), the concatenation (identity mapping) and
// both at once; sliced gradients are read in place through a pitch.
//
// All tensors are [H][W][C] with C a multiple of 8 elements; `pitch` arguments are in elements (a channel slice of a wider
// NHWC buffer is addressed by base pointer + pitch).
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

// 16-byte vector of T: 8 bf16 or 4 float
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __nv_bfloat16 v[8];
};
template <typename T>
__device__ __forceinline__ Vec<T> ld16(const T* p) {
  Vec<T> r;
  *reinterpret_cast<uint4*>(r.v) = *reinterpret_cast<const uint4*>(p);
  return r;
}
template <typename T>
__device__ __forceinline__ void st16(T* p, const Vec<T>& r) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(r.v);
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// ---- 2x2 max pool --------------------------------------------------------------------------------------------------
// thread = (output pixel, 16-byte channel group); idx: one byte per (output pixel, channel) = position 0..3 in the window
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_forward_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                               uint8_t* __restrict__ idx, int H, int W, int C) {
  constexpr int VN = Vec<T>::N;
  const int Ho = H >> 1, Wo = W >> 1, G = C / VN;
  const long long total = (long long)Ho * Wo * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    const long long op = t / G;
    const int ox = (int)(op % Wo), oy = (int)(op / Wo);
    const T* base = x + ((size_t)(2 * oy) * W + 2 * ox) * C + g * VN;
    Vec<T> best = ld16(base);
    uint8_t bi[VN];
#pragma unroll
    for (int k = 0; k < VN; k++) bi[k] = 0;
#pragma unroll
    for (int w = 1; w < 4; w++) {
      const Vec<T> c = ld16(base + ((size_t)(w >> 1) * W + (w & 1)) * C);
#pragma unroll
      for (int k = 0; k < VN; k++) {
        const float cv = to_f(c.v[k]), bv = to_f(best.v[k]);
        if (cv > bv || cv != cv) { best.v[k] = c.v[k]; bi[k] = (uint8_t)w; }   // torch: (val > max) || isnan(val)
      }
    }
    st16(y + (size_t)op * C + g * VN, best);
    uint8_t* ip = idx + (size_t)op * C + g * VN;
    if (VN == 8) {
      uint2 pk;
      pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      *reinterpret_cast<uint2*>(ip) = pk;
    } else {
      *reinterpret_cast<uint32_t*>(ip) = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    }
  }
}
// thread = (INPUT pixel, channel group): every byte of gx is written (zeros outside the pooled area and for non-maxima)
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_backward_kernel(const T* __restrict__ gy, const uint8_t* __restrict__ idx,
                                                                T* __restrict__ gx, int H, int W, int C) {
  constexpr int VN = Vec<T>::N;
  const int Ho = H >> 1, Wo = W >> 1, G = C / VN;
  const long long total = (long long)H * W * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    const long long ip = t / G;
    const int ix = (int)(ip % W), iy = (int)(ip / W);
    Vec<T> out;
#pragma unroll
    for (int k = 0; k < VN; k++) out.v[k] = from_f<T>(0.f);
    const int oy = iy >> 1, ox = ix >> 1;
    if (oy < Ho && ox < Wo) {
      const size_t op = (size_t)oy * Wo + ox;
      const Vec<T> gv = ld16(gy + op * C + g * VN);
      const uint8_t* ib = idx + op * C + g * VN;
      const int me = ((iy & 1) << 1) | (ix & 1);
      uint8_t bi[VN];
      if (VN == 8) {
        const uint2 pk = *reinterpret_cast<const uint2*>(ib);
#pragma unroll
        for (int k = 0; k < 4; k++) { bi[k] = (pk.x >> (8 * k)) & 0xff; bi[4 + k] = (pk.y >> (8 * k)) & 0xff; }
      } else {
        const uint32_t pk = *reinterpret_cast<const uint32_t*>(ib);
#pragma unroll
        for (int k = 0; k < 4; k++) bi[k] = (pk >> (8 * k)) & 0xff;
      }
#pragma unroll
      for (int k = 0; k < VN; k++)
        if (bi[k] == me) out.v[k] = gv.v[k];
    }
    st16(gx + (size_t)ip * C + g * VN, out);
  }
}

// ---- nearest upsample fused with the concatenation --------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  return min((int)floorf((float)dst * scale), in_size - 1);   // ATen nearest_neighbor_compute_source_index
}
// out[p][0..Ca) = a[src(p)][0..Ca) (a is the low-resolution tensor), out[p][Ca..Ca+Cb) = b[p] (b may be NULL: Cb = 0).
// thread = (output pixel, channel group of the concatenated row)
template <typename T>
__global__ void __launch_bounds__(256) upsample_cat_forward_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                                   T* __restrict__ out, int Hi, int Wi, int Ho, int Wo,
                                                                   int Ca, int Cb, float sh, float sw) {
  constexpr int VN = Vec<T>::N;
  const int Ga = Ca / VN, G = (Ca + Cb) / VN;
  const long long total = (long long)Ho * Wo * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    const long long op = t / G;
    const int ox = (int)(op % Wo), oy = (int)(op / Wo);
    Vec<T> v;
    if (g < Ga) {
      const int sy = nearest_src(oy, sh, Hi), sx = nearest_src(ox, sw, Wi);
      v = ld16(a + ((size_t)sy * Wi + sx) * Ca + g * VN);
    } else {
      v = ld16(b + (size_t)op * Cb + (g - Ga) * VN);
    }
    st16(out + (size_t)op * (Ca + Cb) + g * VN, v);
  }
}
// ga[s] = sum of g[p][0..Ca) over the output pixels p whose source is s; g has `pitch` elements per pixel.
// thread = (low-resolution pixel, channel group); sums in float, one rounding at the store (as ATen's accscalar_t)
template <typename T>
__global__ void __launch_bounds__(256) upsample_backward_kernel(const T* __restrict__ g, T* __restrict__ ga, int Hi, int Wi,
                                                                int Ho, int Wo, int Ca, int pitch, float sh, float sw) {
  constexpr int VN = Vec<T>::N;
  const int Ga = Ca / VN;
  const long long total = (long long)Hi * Wi * Ga;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int gq = (int)(t % Ga);
    const long long sp = t / Ga;
    const int sx = (int)(sp % Wi), sy = (int)(sp / Wi);
    // candidate destination rows / columns: around s / scale; the exact set is decided by the forward's own mapping
    const int y0 = max(0, (int)floorf((float)sy / sh) - 2), y1 = min(Ho - 1, (int)ceilf((float)(sy + 1) / sh) + 2);
    const int x0 = max(0, (int)floorf((float)sx / sw) - 2), x1 = min(Wo - 1, (int)ceilf((float)(sx + 1) / sw) + 2);
    float acc[VN];
#pragma unroll
    for (int k = 0; k < VN; k++) acc[k] = 0.f;
    for (int oy = y0; oy <= y1; oy++) {
      if (nearest_src(oy, sh, Hi) != sy) continue;
      for (int ox = x0; ox <= x1; ox++) {
        if (nearest_src(ox, sw, Wi) != sx) continue;
        const Vec<T> v = ld16(g + ((size_t)oy * Wo + ox) * pitch + gq * VN);
#pragma unroll
        for (int k = 0; k < VN; k++) acc[k] += to_f(v.v[k]);
      }
    }
    Vec<T> out;
#pragma unroll
    for (int k = 0; k < VN; k++) out.v[k] = from_f<T>(acc[k]);
    st16(ga + (size_t)sp * Ca + gq * VN, out);
  }
}

// ---- ReLU backward fused with the bias gradient ---------------------------------------------------------------------------
// gm[p][c] = y[p][c] > 0 ? g[p * pitch + c] : 0  (aten::threshold_backward on a possibly channel-sliced gradient) and
// db[c] += sum_p gm[p][c] (the bias gradient aten::convolution_backward would reduce in another pass).
// blockDim.x = G * k threads (G = channel groups per pixel): thread t always owns group t % G, so its channel sums stay in
// registers over the grid-stride loop; one shared-memory reduction and one global atomic per (block, channel) at the end.
#ifndef RELU_UNROLL
#define RELU_UNROLL 1
#endif
template <typename T>
__global__ void __launch_bounds__(256) relu_bias_backward_kernel(const T* __restrict__ g, int pitch, const T* __restrict__ y,
                                                                 T* __restrict__ gm, float* __restrict__ db, long long npix,
                                                                 int C) {
  constexpr int VN = Vec<T>::N;
  extern __shared__ float s_db[];
  const int G = C / VN;
  const int gq = threadIdx.x % G, lp = threadIdx.x / G, ppb = blockDim.x / G;   // pixels per block and step
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_db[c] = 0.f;
  __syncthreads();
  float acc[VN];
#pragma unroll
  for (int k = 0; k < VN; k++) acc[k] = 0.f;
  // RELU_UNROLL pixels per trip: that many pairs of independent 16-byte loads in flight per thread
  const long long step = (long long)gridDim.x * ppb;
  for (long long p0 = (long long)blockIdx.x * ppb + lp; p0 < npix; p0 += RELU_UNROLL * step) {
    Vec<T> gv[RELU_UNROLL], yv[RELU_UNROLL];
#pragma unroll
    for (int u = 0; u < RELU_UNROLL; u++) {
      const long long p = p0 + u * step;
      if (p < npix) {
        gv[u] = ld16(g + (size_t)p * pitch + gq * VN);
        yv[u] = ld16(y + (size_t)p * C + gq * VN);
      }
    }
#pragma unroll
    for (int u = 0; u < RELU_UNROLL; u++) {
      const long long p = p0 + u * step;
      if (p < npix) {
        Vec<T> o;
#pragma unroll
        for (int k = 0; k < VN; k++) {
          const bool on = to_f(yv[u].v[k]) > 0.f;
          o.v[k] = on ? gv[u].v[k] : from_f<T>(0.f);
          acc[k] += on ? to_f(gv[u].v[k]) : 0.f;
        }
        st16(gm + (size_t)p * C + gq * VN, o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VN; k++) atomicAdd(&s_db[gq * VN + k], acc[k]);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(db + c, s_db[c]);
}

int grid_for(long long total) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long want = (total + 255) / 256;
  return (int)std::max(1LL, std::min(want, (long long)sms * 16));
}

int check_common(int H, int W, int C, const void* p0, const void* p1) {
  if (H <= 0 || W <= 0 || C <= 0 || C % 8) { ibgs_set_error("bad NHWC shape %dx%dx%d (C must be a multiple of 8)", H, W, C); return IBGS_EINVAL; }
  if (!p0 || !p1) { ibgs_set_error("null pointer"); return IBGS_EINVAL; }
  if (((uintptr_t)p0 | (uintptr_t)p1) % 16) { ibgs_set_error("NHWC tensors must be 16-byte aligned"); return IBGS_EINVAL; }
  return IBGS_OK;
}

}  // namespace

extern "C" int ibgs_nhwc_maxpool2_forward(const void* x, void* y, uint8_t* idx, int32_t H, int32_t W, int32_t C,
                                          int32_t bf16, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  int rc = check_common(H, W, C, x, y);
  if (rc != IBGS_OK) return rc;
  if (!idx) { ibgs_set_error("idx must not be NULL"); return IBGS_EINVAL; }
  if (H < 2 || W < 2) return IBGS_OK;
  if (bf16) {
    const long long total = (long long)(H / 2) * (W / 2) * (C / 8);
    maxpool2_forward_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, idx, H, W, C);
  } else {
    const long long total = (long long)(H / 2) * (W / 2) * (C / 4);
    maxpool2_forward_kernel<float><<<grid_for(total), 256, 0, s>>>((const float*)x, (float*)y, idx, H, W, C);
  }
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}

extern "C" int ibgs_nhwc_maxpool2_backward(const void* gy, const uint8_t* idx, void* gx, int32_t H, int32_t W, int32_t C,
                                           int32_t bf16, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  int rc = check_common(H, W, C, gx, gx);
  if (rc != IBGS_OK) return rc;
  if ((H >= 2 && W >= 2) && (!gy || !idx)) { ibgs_set_error("null pointer"); return IBGS_EINVAL; }
  if (bf16) {
    const long long total = (long long)H * W * (C / 8);
    maxpool2_backward_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)gy, idx, (__nv_bfloat16*)gx, H, W, C);
  } else {
    const long long total = (long long)H * W * (C / 4);
    maxpool2_backward_kernel<float><<<grid_for(total), 256, 0, s>>>((const float*)gy, idx, (float*)gx, H, W, C);
  }
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}

extern "C" int ibgs_nhwc_upsample_cat_forward(const void* a, const void* b, void* out, int32_t Hi, int32_t Wi, int32_t Ho,
                                              int32_t Wo, int32_t Ca, int32_t Cb, int32_t bf16, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  int rc = check_common(Ho, Wo, Ca, a, out);
  if (rc != IBGS_OK) return rc;
  if (Hi <= 0 || Wi <= 0 || Cb < 0 || Cb % 8 || (Cb && !b) || (b && (uintptr_t)b % 16)) {
    ibgs_set_error("bad upsample_cat arguments");
    return IBGS_EINVAL;
  }
  const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  if (bf16) {
    const long long total = (long long)Ho * Wo * ((Ca + Cb) / 8);
    upsample_cat_forward_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                                              (__nv_bfloat16*)out, Hi, Wi, Ho, Wo, Ca, Cb, sh, sw);
  } else {
    const long long total = (long long)Ho * Wo * ((Ca + Cb) / 4);
    upsample_cat_forward_kernel<float><<<grid_for(total), 256, 0, s>>>((const float*)a, (const float*)b, (float*)out, Hi, Wi, Ho,
                                                                      Wo, Ca, Cb, sh, sw);
  }
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}

extern "C" int ibgs_nhwc_upsample_backward(const void* g, void* ga, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo,
                                           int32_t Ca, int32_t pitch, int32_t bf16, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  int rc = check_common(Ho, Wo, Ca, g, ga);
  if (rc != IBGS_OK) return rc;
  if (Hi <= 0 || Wi <= 0 || pitch < Ca || pitch % 8) { ibgs_set_error("bad upsample_backward arguments"); return IBGS_EINVAL; }
  const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  if (bf16) {
    const long long total = (long long)Hi * Wi * (Ca / 8);
    upsample_backward_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)g, (__nv_bfloat16*)ga, Hi, Wi, Ho,
                                                                           Wo, Ca, pitch, sh, sw);
  } else {
    const long long total = (long long)Hi * Wi * (Ca / 4);
    upsample_backward_kernel<float><<<grid_for(total), 256, 0, s>>>((const float*)g, (float*)ga, Hi, Wi, Ho, Wo, Ca, pitch, sh, sw);
  }
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}

extern "C" int ibgs_nhwc_relu_bias_backward(const void* g, int32_t pitch, const void* y, void* gm, float* db, int64_t npix,
                                            int32_t C, int32_t bf16, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  if (npix < 0 || C <= 0 || C % 8 || pitch < C || pitch % 8 || C > 2048) {
    ibgs_set_error("bad relu_bias_backward shape: npix %lld, C %d, pitch %d", (long long)npix, C, pitch);
    return IBGS_EINVAL;
  }
  if (npix == 0) return IBGS_OK;
  if (!g || !y || !gm || !db || (((uintptr_t)g | (uintptr_t)y | (uintptr_t)gm) % 16)) {
    ibgs_set_error("relu_bias_backward: null or misaligned pointer");
    return IBGS_EINVAL;
  }
  const int VN = bf16 ? 8 : 4, G = C / VN;
  const int threads = G * std::max(1, 256 / G);
  const int ppb = threads / G;
  const int blocks = grid_for(((npix + ppb - 1) / ppb) * 256) ;
  if (bf16)
    relu_bias_backward_kernel<__nv_bfloat16><<<blocks, threads, C * sizeof(float), s>>>((const __nv_bfloat16*)g, pitch, (const __nv_bfloat16*)y,
                                                                                     (__nv_bfloat16*)gm, db, npix, C);
  else
    relu_bias_backward_kernel<float><<<blocks, threads, C * sizeof(float), s>>>((const float*)g, pitch, (const float*)y, (float*)gm, db,
                                                                               npix, C);
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}
