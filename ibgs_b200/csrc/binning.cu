// binning.cu -- tile binning: offsets scan, key/value duplication, 64-bit tile|depth sort, tile ranges.
//
// Reference behaviour: cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:426), duplicateWithKeys
// (:187-228), getHigherMsb (:152-167), cub::DeviceRadixSort::SortPairs on bits [0, 32+msb) (:452-457),
// cudaMemset + identifyTileRanges (:459-467, :233-255).  Integer path: results are bit-exact by
// construction (same emission order, stable sort over the same bit range).
//
// The scan and the LSD radix sort are CCCL/CUB device primitives (library code, like the reference);
// duplication and range identification are this project's kernels: duplication is warp-cooperative
// so the 12 B/instance stores of one Gaussian's tile rectangle are issued by adjacent lanes instead of
// one thread walking the rectangle serially.
#include "common.cuh"
#include <cub/cub.cuh>

namespace {

// one warp per 32 Gaussians; rectangles with >= 8 tiles are written cooperatively by the whole warp
__global__ void __launch_bounds__(256) duplicate_with_keys_kernel(int P, const float4* __restrict__ rec,
                                                                  const float* __restrict__ depths,
                                                                  const uint32_t* __restrict__ offsets,
                                                                  uint64_t* __restrict__ keys,
                                                                  uint32_t* __restrict__ vals,
                                                                  const int* __restrict__ radii, dim3 grid) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  int radius = 0;
  uint32_t off = 0;
  uint2 rmin = {0, 0}, rmax = {0, 0};
  uint32_t dbits = 0;
  if (idx < P) radius = radii[idx];
  if (radius > 0) {
    off = (idx == 0) ? 0 : offsets[idx - 1];
    float4 q0 = rec[4 * (size_t)idx];
    getRect(make_float2(q0.x, q0.y), radius, rmin, rmax, grid);
    dbits = __float_as_uint(depths[idx]);
  }
  const uint32_t w = rmax.x - rmin.x;
  const uint32_t n = (radius > 0) ? w * (rmax.y - rmin.y) : 0;

  // small rectangles: the owning lane writes them itself (row-major y then x, rasterizer_impl.cu:215-226)
  const bool big = n >= 8;
  if (n > 0 && !big) {
    uint32_t o = off;
    for (uint32_t y = rmin.y; y < rmax.y; y++)
      for (uint32_t x = rmin.x; x < rmax.x; x++) {
        uint64_t key = y * grid.x + x;
        key <<= 32;
        key |= dbits;
        keys[o] = key;
        vals[o] = idx;
        o++;
      }
  }
  // large rectangles: all 32 lanes share the work, coalesced stores
  unsigned mask = __ballot_sync(0xffffffffu, big);
  while (mask) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    const uint32_t s_off = __shfl_sync(0xffffffffu, off, src);
    const uint32_t s_n = __shfl_sync(0xffffffffu, n, src);
    const uint32_t s_w = __shfl_sync(0xffffffffu, w, src);
    const uint32_t s_x0 = __shfl_sync(0xffffffffu, rmin.x, src);
    const uint32_t s_y0 = __shfl_sync(0xffffffffu, rmin.y, src);
    const uint32_t s_d = __shfl_sync(0xffffffffu, dbits, src);
    const uint32_t s_idx = __shfl_sync(0xffffffffu, (uint32_t)idx, src);
    for (uint32_t k = lane; k < s_n; k += 32) {
      const uint32_t y = s_y0 + k / s_w;
      const uint32_t x = s_x0 + k % s_w;
      uint64_t key = y * grid.x + x;
      key <<= 32;
      key |= s_d;
      keys[s_off + k] = key;
      vals[s_off + k] = s_idx;
    }
  }
}

// reference rasterizer_impl.cu:233-255
__global__ void identify_tile_ranges_kernel(int L, const uint64_t* __restrict__ keys, uint2* ranges) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L) return;
  uint32_t currtile = keys[idx] >> 32;
  if (idx == 0)
    ranges[currtile].x = 0;
  else {
    uint32_t prevtile = keys[idx - 1] >> 32;
    if (currtile != prevtile) {
      ranges[prevtile].y = idx;
      ranges[currtile].x = idx;
    }
  }
  if (idx == L - 1) ranges[currtile].y = L;
}

// reference rasterizer_impl.cu:152-167
uint32_t getHigherMsb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4;
  uint32_t step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb)
      msb += step;
    else
      msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

}  // namespace

extern "C" int ibgs_sort_bits(int32_t num_tiles) { return 32 + (int)getHigherMsb((uint32_t)num_tiles); }

size_t scan_temp_bytes(size_t P) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)P);
  return bytes;
}

int run_scan(const GeomState& g, size_t P, void* temp, size_t temp_bytes, cudaStream_t s) {
  ProfScope prof(PROF_SCAN, s);
  CUDA_TRY(cub::DeviceScan::InclusiveSum(temp, temp_bytes, g.tiles_touched, g.point_offsets, (int)P, s));
  g_launch_count += 2;
  return IBGS_OK;
}

size_t carve_scratch(ScratchState& sc, char* base, size_t R, int end_bit) {
  size_t off = 0;
  carve(off, sc.keys_unsorted, base, R);
  carve(off, sc.keys_sorted, base, R);
  carve(off, sc.vals_unsorted, base, R);
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)R, 0, end_bit);
  sc.sort_temp_bytes = bytes;
  off = align_up(off, 256);
  sc.sort_temp = base + off;
  off += bytes;
  return align_up(off, 256);
}

int run_binning(const IbgsForwardArgs& a, const GeomState& g, const ImageState& im, char* scratch_base,
                size_t scratch_bytes, BinningState& b, int64_t R, dim3 grid, cudaStream_t s) {
  const int debug = a.view.debug;
  const int end_bit = ibgs_sort_bits((int32_t)(grid.x * grid.y));
  ScratchState sc;
  size_t need = carve_scratch(sc, scratch_base, (size_t)R, end_bit);
  if (need > scratch_bytes) {
    ibgs_set_error("scratch too small: %zu < %zu", scratch_bytes, need);
    return IBGS_EINVAL;
  }
  {
    ProfScope prof(PROF_DUPLICATE, s);
    duplicate_with_keys_kernel<<<(a.P + 255) / 256, 256, 0, s>>>(a.P, g.rec, g.depths, g.point_offsets,
                                                                 sc.keys_unsorted, sc.vals_unsorted, a.radii,
                                                                 grid);
    KERNEL_CHECK(debug, s);
  }
  if (R > 0) {
    ProfScope prof(PROF_SORT, s);
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(sc.sort_temp, sc.sort_temp_bytes, sc.keys_unsorted,
                                             sc.keys_sorted, sc.vals_unsorted, b.point_list, (int)R, 0,
                                             end_bit, s));
    g_launch_count += (end_bit + 7) / 8 + 1;
  }
  {
    ProfScope prof(PROF_RANGES, s);
    CUDA_TRY(cudaMemsetAsync(im.ranges, 0, (size_t)grid.x * grid.y * sizeof(uint2), s));
    if (R > 0) {
      identify_tile_ranges_kernel<<<(int)((R + 255) / 256), 256, 0, s>>>((int)R, sc.keys_sorted, im.ranges);
      KERNEL_CHECK(debug, s);
    }
  }
  return IBGS_OK;
}
