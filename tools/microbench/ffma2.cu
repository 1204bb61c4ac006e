// Microbenchmark: FFMA vs FFMA2 (packed f32x2) issue throughput on sm_100a (dev tool).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
      else x[i] = __ffma2_rn(x[i], A, B);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; mode++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.0001f); else k<1><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.0001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 8 * 256 * (double)iters * 16;
      printf("mode %d (%s): %.3f ms  %.2f TFMA/s  = %.1f FMA/clk/SM @1.965GHz\n", mode, mode ? "FFMA2" : "FFMA", ms, fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
    }
  }
  return 0;
}
