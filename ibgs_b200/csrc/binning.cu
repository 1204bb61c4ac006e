// binning.cu -- tile binning: depth order of the Gaussians, offsets scan, tile-instance emission, tile sort, ranges.
//
// Reference behaviour: cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:426), duplicateWithKeys
// (:187-228), getHigherMsb (:152-167), cub::DeviceRadixSort::SortPairs of (tile<<32 | depth bits, Gaussian id)
// on bits [0, 32+msb) (:452-457), cudaMemset + identifyTileRanges (:459-467, :233-255).  Its result -- the list
// `point_list` ordered by tile, then by depth bits, ties by Gaussian id (stable sort, Gaussian-major emission)
// -- and the tile ranges are what every later stage consumes; they are reproduced here bit for bit.
//
// How (this project's own decomposition; the radix sorts and the scan are CCCL/CUB device primitives, library
// code like in the reference).  A 64-bit LSD sort of all R instances is 6 passes over 24-byte pairs at R = 15 M.
// The depth half of the key is a property of the GAUSSIAN, not of the instance, so it is sorted once per
// Gaussian instead of once per instance:
//   1. stable sort of the P Gaussians by depth bits (32-bit keys, P << R items); culled Gaussians carry the key
//      0xFFFFFFFF and land behind every visible one;
//   2. inclusive scan of tiles_touched in that order -> write offsets;
//   3. emission in depth order: instance = (tile id u32, Gaussian id u32), warp-cooperative for big rectangles;
//   4. stable sort of the R instances by tile id only: ceil(msb(T)/8) = 2 passes over 8-byte pairs;
//   5. tile ranges from the sorted tile ids.
// Stability of (4) keeps each tile's instances in emission order = ascending depth bits, ties in ascending
// Gaussian id (stability of (1)) -- exactly the reference's order.  The reference's sorted 64-bit keys are
// (tile id << 32 | depth bits of point_list[i]); tests rebuild them from this state and compare bit-exactly.
#include "common.cuh"
#include <cub/cub.cuh>

namespace {

struct GatherTiles {
  const uint32_t* tiles;
  __host__ __device__ __forceinline__ uint32_t operator()(const uint32_t& g) const { return tiles[g]; }
};

// One thread per position in depth order; one warp per 32 positions; rectangles with >= 8 tiles are written
// cooperatively by the whole warp.  Tile order inside a rectangle is row-major like rasterizer_impl.cu:215-226
// (irrelevant for the result -- a Gaussian appears at most once per tile -- but kept).
// BATCH (ibgs_forward_depth_batch): the P items are (view, Gaussian) pairs, item id = view * items_per_view + Gaussian;
// the tile id carries the view: view * tiles_per_view + tile.
template <typename TileT, bool BATCH>
__global__ void __launch_bounds__(256) emit_instances_kernel(int P, const uint32_t* __restrict__ order,
                                                             const uint32_t* __restrict__ tiles_touched,
                                                             const uint32_t* __restrict__ offsets,
                                                             const float4* __restrict__ rec,
                                                             const int* __restrict__ radii,
                                                             TileT* __restrict__ tile_ids,
                                                             uint32_t* __restrict__ vals, dim3 grid,
                                                             uint32_t items_per_view, uint32_t tiles_per_view) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  uint32_t gid = 0, n = 0, off = 0, tbase = 0;
  uint2 rmin = {0, 0}, rmax = {0, 0};
  if (pos < P) {
    gid = order[pos];
    n = tiles_touched[gid];
  }
  if (n > 0) {
    off = (pos == 0) ? 0 : offsets[pos - 1];
    const float4 q0 = rec[4 * (size_t)gid];
    getRect(make_float2(q0.x, q0.y), radii[gid], rmin, rmax, grid);
    if (BATCH) tbase = (gid / items_per_view) * tiles_per_view;
  }
  const uint32_t w = rmax.x - rmin.x;
  const bool big = n >= 8;
  if (n > 0 && !big) {
    uint32_t o = off;
    for (uint32_t y = rmin.y; y < rmax.y; y++)
      for (uint32_t x = rmin.x; x < rmax.x; x++) {
        tile_ids[o] = (TileT)(tbase + y * grid.x + x);
        vals[o] = gid;
        o++;
      }
  }
  unsigned mask = __ballot_sync(0xffffffffu, big);
  while (mask) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    const uint32_t s_off = __shfl_sync(0xffffffffu, off, src);
    const uint32_t s_n = __shfl_sync(0xffffffffu, n, src);
    const uint32_t s_w = __shfl_sync(0xffffffffu, w, src);
    const uint32_t s_x0 = __shfl_sync(0xffffffffu, rmin.x, src);
    const uint32_t s_y0 = __shfl_sync(0xffffffffu, rmin.y, src);
    const uint32_t s_gid = __shfl_sync(0xffffffffu, gid, src);
    const uint32_t s_tb = BATCH ? __shfl_sync(0xffffffffu, tbase, src) : 0u;
    for (uint32_t k = lane; k < s_n; k += 32) {
      const uint32_t y = s_y0 + k / s_w;
      const uint32_t x = s_x0 + k % s_w;
      tile_ids[s_off + k] = (TileT)(s_tb + y * grid.x + x);
      vals[s_off + k] = s_gid;
    }
  }
}

// reference rasterizer_impl.cu:233-255 (on 16- / 32-bit tile ids instead of the high half of 64-bit keys).  One thread
// takes the 16 bytes of ids starting at a 16-byte boundary (8 uint16 or 4 uint32: one vector load) plus the id before
// them; the reference's one-thread-per-instance form is launch- and tail-bound at 15 M instances.
template <typename TileT>
__global__ void __launch_bounds__(256) identify_tile_ranges_kernel(int L, const TileT* __restrict__ tile_ids, uint2* ranges) {
  constexpr int PER = 16 / sizeof(TileT);
  const int base = (blockIdx.x * blockDim.x + threadIdx.x) * PER;
  if (base >= L) return;
  TileT ids[PER];
  if (base + PER <= L) {
    *reinterpret_cast<uint4*>(ids) = *reinterpret_cast<const uint4*>(tile_ids + base);
  } else {
#pragma unroll
    for (int k = 0; k < PER; k++) ids[k] = (base + k < L) ? tile_ids[base + k] : (TileT)0;
  }
  uint32_t prevtile = (base == 0) ? 0u : (uint32_t)tile_ids[base - 1];
#pragma unroll
  for (int k = 0; k < PER; k++) {
    const int idx = base + k;
    if (idx < L) {
      const uint32_t currtile = ids[k];
      if (idx == 0) {
        ranges[currtile].x = 0;
      } else if (currtile != prevtile) {
        ranges[prevtile].y = idx;
        ranges[currtile].x = idx;
      }
      if (idx == L - 1) ranges[currtile].y = L;
      prevtile = currtile;
    }
  }
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

size_t carve_order(OrderState& o, char* base, size_t P) {
  size_t off = 0;
  carve(off, o.iota, base, P);
  carve(off, o.keys_sorted, base, P);
  carve(off, o.order, base, P);
  carve(off, o.offsets, base, P);
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)P, 0, 32);
  cub::TransformInputIterator<uint32_t, GatherTiles, const uint32_t*> it((const uint32_t*)nullptr, GatherTiles{nullptr});
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, it, (uint32_t*)nullptr, (int)P);
  o.temp_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  off = align_up(off, 256);
  o.temp = base + off;
  off += o.temp_bytes;
  return align_up(off, 256);
}

// steps 1-2: depth order of the Gaussians + write offsets in that order
int run_depth_order(const GeomState& g, const OrderState& o, size_t P, cudaStream_t s) {
  {
    ProfScope prof(PROF_GSORT, s);
    size_t tb = o.temp_bytes;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(o.temp, tb, reinterpret_cast<const uint32_t*>(g.depths), o.keys_sorted,
                                             o.iota, o.order, (int)P, 0, 32, s));
    g_launch_count += 5;
  }
  {
    ProfScope prof(PROF_SCAN, s);
    size_t tb = o.temp_bytes;
    cub::TransformInputIterator<uint32_t, GatherTiles, const uint32_t*> it(o.order, GatherTiles{g.tiles_touched});
    CUDA_TRY(cub::DeviceScan::InclusiveSum(o.temp, tb, it, o.offsets, (int)P, s));
    g_launch_count += 2;
  }
  return IBGS_OK;
}

// Tile ids are stored as uint16 whenever the image has at most 65536 tiles (every resolution up to 4096x4096):
// the tile sort then moves 6 instead of 8 bytes per instance and pass.  The arrays are carved for the wider type.
size_t carve_scratch(ScratchState& sc, char* base, size_t R, int tile_bits) {
  size_t off = 0;
  carve(off, sc.tiles_unsorted, base, R);
  carve(off, sc.tiles_sorted, base, R);
  carve(off, sc.vals_unsorted, base, R);
  size_t bytes = 0, bytes16 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)R, 0, tile_bits);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes16, (uint16_t*)nullptr, (uint16_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)R, 0, tile_bits < 16 ? tile_bits : 16);
  if (bytes16 > bytes) bytes = bytes16;
  sc.sort_temp_bytes = bytes;
  off = align_up(off, 256);
  sc.sort_temp = base + off;
  off += bytes;
  return align_up(off, 256);
}

// steps 3-5
int run_binning(const IbgsForwardArgs& a, const GeomState& g, const OrderState& o, const ImageState& im,
                char* scratch_base, size_t scratch_bytes, BinningState& b, int64_t R, dim3 grid, cudaStream_t s) {
  return run_binning_items(a.P, a.radii, a.view.debug, 1, g, o, im.ranges, scratch_base, scratch_bytes, b, R, grid, s);
}

// P items; with views > 1 they are (view, Gaussian) pairs (item id = view * P/views + Gaussian) and ranges has
// views * tiles entries
int run_binning_items(int P, const int* radii, int debug, int views, const GeomState& g, const OrderState& o,
                      uint2* ranges, char* scratch_base, size_t scratch_bytes, BinningState& b, int64_t R, dim3 grid,
                      cudaStream_t s) {
  const uint32_t tiles_per_view = grid.x * grid.y;
  const uint32_t items_per_view = (uint32_t)(P / views);
  const int tile_bits = ibgs_sort_bits((int32_t)(tiles_per_view * views)) - 32;
  ScratchState sc;
  size_t need = carve_scratch(sc, scratch_base, (size_t)R, tile_bits);
  if (need > scratch_bytes) {
    ibgs_set_error("scratch too small: %zu < %zu", scratch_bytes, need);
    return IBGS_EINVAL;
  }
  const bool narrow = tile_bits <= 16;
  uint16_t* t16_unsorted = reinterpret_cast<uint16_t*>(sc.tiles_unsorted);
  uint16_t* t16_sorted = reinterpret_cast<uint16_t*>(sc.tiles_sorted);
  {
    ProfScope prof(PROF_DUPLICATE, s);
    const int nb = (P + 255) / 256;
#define EMIT(T, BATCH, dst)                                                                                          \
  emit_instances_kernel<T, BATCH><<<nb, 256, 0, s>>>(P, o.order, g.tiles_touched, o.offsets, g.rec, radii, dst,      \
                                                     sc.vals_unsorted, grid, items_per_view, tiles_per_view)
    if (views > 1) {
      if (narrow) EMIT(uint16_t, true, t16_unsorted); else EMIT(uint32_t, true, sc.tiles_unsorted);
    } else {
      if (narrow) EMIT(uint16_t, false, t16_unsorted); else EMIT(uint32_t, false, sc.tiles_unsorted);
    }
#undef EMIT
    KERNEL_CHECK(debug, s);
  }
  if (R > 0) {
    ProfScope prof(PROF_SORT, s);
    if (narrow)
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(sc.sort_temp, sc.sort_temp_bytes, t16_unsorted, t16_sorted,
                                               sc.vals_unsorted, b.point_list, (int)R, 0, tile_bits, s));
    else
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(sc.sort_temp, sc.sort_temp_bytes, sc.tiles_unsorted, sc.tiles_sorted,
                                               sc.vals_unsorted, b.point_list, (int)R, 0, tile_bits, s));
    g_launch_count += (tile_bits + 7) / 8 + 1;
  }
  {
    ProfScope prof(PROF_RANGES, s);
    CUDA_TRY(cudaMemsetAsync(ranges, 0, (size_t)tiles_per_view * views * sizeof(uint2), s));
    if (R > 0) {
      if (narrow)
        identify_tile_ranges_kernel<uint16_t><<<(int)((R + 256 * 8 - 1) / (256 * 8)), 256, 0, s>>>((int)R, t16_sorted, ranges);
      else
        identify_tile_ranges_kernel<uint32_t><<<(int)((R + 256 * 4 - 1) / (256 * 4)), 256, 0, s>>>((int)R, sc.tiles_sorted, ranges);
      KERNEL_CHECK(debug, s);
    }
  }
  return IBGS_OK;
}
