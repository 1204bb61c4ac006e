// knn.cu -- distCUDA2: mean squared distance to the 3 nearest neighbours of every point.
//
// Reference behaviour: SimpleKNN::knn (submodules/simple-knn/simple_knn.cu:185-221): component-wise
// min/max with init {0,0,0} (:190-200, so min <= 0 <= max always), 30-bit Morton codes (:45-70),
// index sort by code (:210-213), AABB of every run of 1024 sorted points (:78-117), and a pruned brute
// force over the boxes (:147-183).  Output [P] float32, exact 3-NN (box pruning is conservative).
//
// Differences in structure, not in results: no host round-trips (the reference copies min and max back
// to the host, :197,200, and allocates five thrust vectors per call); bounds stay on the device as
// order-preserving integers; the Morton-sorted points are gathered once into a float4 array so the
// neighbour scan reads contiguous memory instead of points[indices[i]]; 256-thread CTAs.
#include "common.cuh"
#include <cfloat>

#define BOX_SIZE 1024

namespace {

struct MinMax {
  float3 minn;
  float3 maxx;
};

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void init_bounds_kernel(uint32_t* bounds) {
  if (threadIdx.x < 6) bounds[threadIdx.x] = f2ord(0.0f);  // reference init {0,0,0}, simple_knn.cu:191
}

__global__ void __launch_bounds__(256) bounds_kernel(int P, const float* __restrict__ pts, uint32_t* bounds) {
  float mn[3] = {0.f, 0.f, 0.f}, mx[3] = {0.f, 0.f, 0.f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const float v = pts[3 * (size_t)i + c];
      mn[c] = min(mn[c], v);
      mx[c] = max(mx[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      atomicMin(&bounds[c], f2ord(mn[c]));
      atomicMax(&bounds[3 + c], f2ord(mx[c]));
    }
  }
}

// reference simple_knn.cu:36-61
__device__ __forceinline__ uint32_t prepMorton(uint32_t x) {
  x = (x | (x << 16)) & 0x030000FF;
  x = (x | (x << 8)) & 0x0300F00F;
  x = (x | (x << 4)) & 0x030C30C3;
  x = (x | (x << 2)) & 0x09249249;
  return x;
}

__global__ void __launch_bounds__(256) morton_kernel(int P, const float* __restrict__ pts,
                                                     const uint32_t* __restrict__ bounds, uint32_t* codes,
                                                     uint32_t* indices) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float3 minn = {ord2f(bounds[0]), ord2f(bounds[1]), ord2f(bounds[2])};
  const float3 maxx = {ord2f(bounds[3]), ord2f(bounds[4]), ord2f(bounds[5])};
  const float3 coord = {pts[3 * (size_t)idx], pts[3 * (size_t)idx + 1], pts[3 * (size_t)idx + 2]};
  uint32_t x = prepMorton(((coord.x - minn.x) / (maxx.x - minn.x)) * ((1 << 10) - 1));
  uint32_t y = prepMorton(((coord.y - minn.y) / (maxx.y - minn.y)) * ((1 << 10) - 1));
  uint32_t z = prepMorton(((coord.z - minn.z) / (maxx.z - minn.z)) * ((1 << 10) - 1));
  codes[idx] = x | (y << 1) | (z << 2);
  indices[idx] = idx;
}

__global__ void __launch_bounds__(256) gather_kernel(int P, const float* __restrict__ pts,
                                                     const uint32_t* __restrict__ indices, float4* sorted) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const uint32_t s = indices[idx];
  sorted[idx] = make_float4(pts[3 * (size_t)s], pts[3 * (size_t)s + 1], pts[3 * (size_t)s + 2], 0.f);
}

// one CTA of 256 threads per box of 1024 sorted points
__global__ void __launch_bounds__(256) box_minmax_kernel(int P, const float4* __restrict__ sorted, MinMax* boxes) {
  __shared__ float s_red[8][6];
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  const int start = blockIdx.x * BOX_SIZE;
  for (int i = start + threadIdx.x; i < min(P, start + BOX_SIZE); i += 256) {
    const float4 p = sorted[i];
    mn[0] = min(mn[0], p.x); mn[1] = min(mn[1], p.y); mn[2] = min(mn[2], p.z);
    mx[0] = max(mx[0], p.x); mx[1] = max(mx[1], p.y); mx[2] = max(mx[2], p.z);
  }
#pragma unroll
  for (int c = 0; c < 3; c++)
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) { s_red[warp][c] = mn[c]; s_red[warp][3 + c] = mx[c]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        mn[c] = min(mn[c], s_red[w][c]);
        mx[c] = max(mx[c], s_red[w][3 + c]);
      }
    MinMax b;
    b.minn = {mn[0], mn[1], mn[2]};
    b.maxx = {mx[0], mx[1], mx[2]};
    boxes[blockIdx.x] = b;
  }
}

// reference simple_knn.cu:119-130
__device__ __forceinline__ float distBoxPoint(const MinMax& box, const float3& p) {
  float3 diff = {0, 0, 0};
  if (p.x < box.minn.x || p.x > box.maxx.x) diff.x = min(abs(p.x - box.minn.x), abs(p.x - box.maxx.x));
  if (p.y < box.minn.y || p.y > box.maxx.y) diff.y = min(abs(p.y - box.minn.y), abs(p.y - box.maxx.y));
  if (p.z < box.minn.z || p.z > box.maxx.z) diff.z = min(abs(p.z - box.minn.z), abs(p.z - box.maxx.z));
  return diff.x * diff.x + diff.y * diff.y + diff.z * diff.z;
}

// reference simple_knn.cu:132-145
__device__ __forceinline__ void updateKBest3(const float3& ref, const float4& point, float* knn) {
  float3 d = {point.x - ref.x, point.y - ref.y, point.z - ref.z};
  float dist = d.x * d.x + d.y * d.y + d.z * d.z;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (knn[j] > dist) {
      float t = knn[j];
      knn[j] = dist;
      dist = t;
    }
  }
}

// reference simple_knn.cu:147-183
__global__ void __launch_bounds__(256) box_mean_dist_kernel(int P, const float4* __restrict__ sorted,
                                                            const uint32_t* __restrict__ indices,
                                                            const MinMax* __restrict__ boxes, float* dists) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float4 p4 = sorted[idx];
  const float3 point = {p4.x, p4.y, p4.z};
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  for (int i = max(0, idx - 3); i <= min(P - 1, idx + 3); i++) {
    if (i == idx) continue;
    updateKBest3(point, sorted[i], best);
  }
  const float reject = best[2];
  best[0] = FLT_MAX;
  best[1] = FLT_MAX;
  best[2] = FLT_MAX;
  const int nb = (P + BOX_SIZE - 1) / BOX_SIZE;
  for (int b = 0; b < nb; b++) {
    const MinMax box = boxes[b];
    const float dist = distBoxPoint(box, point);
    if (dist > reject || dist > best[2]) continue;
    const int end = min(P, (b + 1) * BOX_SIZE);
    for (int i = b * BOX_SIZE; i < end; i++) {
      if (i == idx) continue;
      updateKBest3(point, sorted[i], best);
    }
  }
  dists[indices[idx]] = (best[0] + best[1] + best[2]) / 3.0f;
}

struct KnnScratch {
  uint32_t* bounds;
  uint32_t* codes;
  uint32_t* codes_sorted;
  uint32_t* idx;
  uint32_t* idx_sorted;
  float4* sorted;
  MinMax* boxes;
  char* sort_temp;
  SortPlan plan;
};

size_t carve_knn(KnnScratch& k, char* base, size_t P) {
  size_t off = 0;
  carve(off, k.bounds, base, 8);
  carve(off, k.codes, base, P);
  carve(off, k.codes_sorted, base, P);
  carve(off, k.idx, base, P);
  carve(off, k.idx_sorted, base, P);
  carve(off, k.sorted, base, P);
  carve(off, k.boxes, base, (P + BOX_SIZE - 1) / BOX_SIZE);
  k.plan = sort_plan(P, 30, 4);   // Morton codes occupy 30 bits: three 10-bit multisplit passes (sort.cu)
  off = align_up(off, 256);
  k.sort_temp = base + off;
  off += k.plan.bytes;
  return align_up(off, 256);
}

}  // namespace

extern "C" size_t ibgs_dist2_scratch_bytes(int32_t P) {
  KnnScratch k;
  return carve_knn(k, nullptr, (size_t)(P > 0 ? P : 1));
}

extern "C" int ibgs_dist2(int32_t P, const float* points, float* mean_dists, void* scratch,
                          size_t scratch_bytes, void* stream) {
  if (P < 0) { ibgs_set_error("P must be >= 0"); return IBGS_EINVAL; }
  if (P == 0) return IBGS_OK;
  if (!points || !mean_dists || !scratch) { ibgs_set_error("null pointer"); return IBGS_EINVAL; }
  cudaStream_t s = (cudaStream_t)stream;
  KnnScratch k;
  size_t need = carve_knn(k, (char*)scratch, (size_t)P);
  if (need > scratch_bytes) { ibgs_set_error("scratch too small: %zu < %zu", scratch_bytes, need); return IBGS_EINVAL; }
  const int blocks = (P + 255) / 256;
  init_bounds_kernel<<<1, 32, 0, s>>>(k.bounds);
  KERNEL_CHECK(0, s);
  bounds_kernel<<<min(blocks, 148 * 8), 256, 0, s>>>(P, points, k.bounds);
  KERNEL_CHECK(0, s);
  morton_kernel<<<blocks, 256, 0, s>>>(P, points, k.bounds, k.codes, k.idx);
  KERNEL_CHECK(0, s);
  // codes occupy 30 bits, so sorting bits [0,30) is identical to the reference's full 32-bit sort (:210-213)
  {
    int rc = sort_pairs(k.plan, k.codes, k.idx, k.codes_sorted, k.idx_sorted, nullptr, k.sort_temp, s, 0);
    if (rc != IBGS_OK) return rc;
  }
  gather_kernel<<<blocks, 256, 0, s>>>(P, points, k.idx_sorted, k.sorted);
  KERNEL_CHECK(0, s);
  const int nb = (P + BOX_SIZE - 1) / BOX_SIZE;
  box_minmax_kernel<<<nb, 256, 0, s>>>(P, k.sorted, k.boxes);
  KERNEL_CHECK(0, s);
  box_mean_dist_kernel<<<blocks, 256, 0, s>>>(P, k.sorted, k.idx_sorted, k.boxes, mean_dists);
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}
