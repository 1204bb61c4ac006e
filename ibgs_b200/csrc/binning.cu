// binning.cu -- tile binning: depth order of the Gaussians, offsets scan, tile-instance emission, tile sort, ranges.
//
// Reference behaviour: cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:426), duplicateWithKeys
// (:187-228), getHigherMsb (:152-167), cub::DeviceRadixSort::SortPairs of (tile<<32 | depth bits, Gaussian id)
// on bits [0, 32+msb) (:452-457), cudaMemset + identifyTileRanges (:459-467, :233-255).  Its result -- the list
// `point_list` ordered by tile, then by depth bits, ties by Gaussian id (stable sort, Gaussian-major emission)
// -- and the tile ranges are what every later stage consumes; they are reproduced here bit for bit.
//
// How (this project's own decomposition, on this project's own sort / scan primitives, sort.cu -- no library code).
// A 64-bit LSD sort of all R instances is 6 passes over 24-byte pairs at R = 15 M.  The depth half of the key is a
// property of the GAUSSIAN, not of the instance, so it is sorted once per Gaussian instead of once per instance:
//   1. stable sort of the P Gaussians by depth bits (32-bit keys, P << R items, four 8-bit passes of sort.cu);
//      culled Gaussians carry the key 0xFFFFFFFF and land behind every visible one;
//   2. inclusive scan of tiles_touched in that order -> write offsets; offsets[P-1] = num_rendered;
//   3. emission in depth order: instance = (tile id u16/u32, Gaussian id u32), warp-cooperative for big rectangles;
//   4. stable sort of the R instances by tile id only: ceil(msb(T) / 8) passes (two at 1080p: 13 bits -> 7 + 6) over
//      (u16 tile, u32 id) pairs -- (u32, u32) above 65536 tiles or in batched depth renders;
//   5. tile ranges: identify_tile_ranges_kernel on the sorted tile ids (a single-pass sort emits them itself).
// Stability of (4) keeps each tile's instances in emission order = ascending depth bits, ties in ascending
// Gaussian id (stability of (1)) -- exactly the reference's order.  The reference's sorted 64-bit keys are
// (tile id << 32 | depth bits of point_list[i]); tests rebuild them from this state and compare bit-exactly.
// Steps 3-4 are queued BEFORE the host knows num_rendered (run_binning_begin: the kernels read it on the device and are
// bounded by the scratch capacity), so the GPU keeps working while the 4-byte read-back is in flight (api.cu).
#include "common.cuh"

namespace {

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
                                                             uint32_t items_per_view, uint32_t tiles_per_view,
                                                             uint32_t cap) {
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
    // speculative launch: the scratch holds `cap` instances.  If num_rendered turns out larger the host re-runs the
    // emission into a bigger scratch; until then nothing may be written past the end
    if ((uint64_t)off + n > cap) n = 0;
  }
  if (n > 0) {
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
__global__ void __launch_bounds__(256) identify_tile_ranges_kernel(const uint32_t* __restrict__ n_dev, uint32_t n_cap,
                                                                   const TileT* __restrict__ tile_ids, uint2* ranges) {
  constexpr int PER = 16 / sizeof(TileT);
  const int L = (int)min(*n_dev, n_cap);
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
  carve(off, o.keys_sorted, base, P);
  carve(off, o.order, base, P);
  carve(off, o.offsets, base, P);
  o.plan = sort_plan(P, 32, 4);
  off = align_up(off, 256);
  o.temp = base + off;
  o.scan_off = o.plan.bytes;
  off += o.plan.bytes + scan_temp_bytes(P);
  return align_up(off, 256);
}

// steps 1-2: depth order of the Gaussians + write offsets in that order
int run_depth_order(const GeomState& g, const OrderState& o, size_t P, cudaStream_t s) {
  {
    ProfScope prof(PROF_GSORT, s);
    // values = Gaussian ids = positions in the input: no identity-permutation array
    int rc = sort_pairs(o.plan, g.depths, nullptr, o.keys_sorted, o.order, nullptr, o.temp, s, 0);
    if (rc != IBGS_OK) return rc;
  }
  {
    ProfScope prof(PROF_SCAN, s);
    int rc = scan_gather_inclusive(P, o.order, g.tiles_touched, o.offsets, o.temp + o.scan_off, s, 0);
    if (rc != IBGS_OK) return rc;
  }
  return IBGS_OK;
}

// Tile ids are stored as uint16 whenever the image has at most 65536 tiles (every resolution up to 4096x4096):
// the tile sort then moves 6 instead of 8 bytes per instance.  The arrays are carved for the wider type.
size_t carve_scratch(ScratchState& sc, char* base, size_t cap, int tile_bits) {
  size_t off = 0;
  carve(off, sc.tiles_unsorted, base, cap);
  carve(off, sc.tiles_sorted, base, cap);
  carve(off, sc.vals_unsorted, base, cap);
  sc.plan = sort_plan(cap, tile_bits, tile_bits <= 16 ? 2 : 4);
  off = align_up(off, 256);
  sc.sort_temp = base + off;
  off += sc.plan.bytes;
  return align_up(off, 256);
}

// P items; with views > 1 they are (view, Gaussian) pairs (item id = view * P/views + Gaussian) and ranges has
// views * tiles entries
int run_binning_begin(int P, const int* radii, int debug, int views, const GeomState& g, const OrderState& o,
                      ScratchState& sc, size_t cap, dim3 grid, cudaStream_t s) {
  const uint32_t tiles_per_view = grid.x * grid.y;
  const uint32_t items_per_view = (uint32_t)(P / views);
  const bool narrow = sc.plan.key_bytes == 2;
  uint16_t* t16_unsorted = reinterpret_cast<uint16_t*>(sc.tiles_unsorted);
  {
    ProfScope prof(PROF_DUPLICATE, s);
    const int nb = (P + 255) / 256;
#define EMIT(T, BATCH, dst)                                                                                          \
  emit_instances_kernel<T, BATCH><<<nb, 256, 0, s>>>(P, o.order, g.tiles_touched, o.offsets, g.rec, radii, dst,      \
                                                     sc.vals_unsorted, grid, items_per_view, tiles_per_view,         \
                                                     (uint32_t)cap)
    if (views > 1) {
      if (narrow) EMIT(uint16_t, true, t16_unsorted); else EMIT(uint32_t, true, sc.tiles_unsorted);
    } else {
      if (narrow) EMIT(uint16_t, false, t16_unsorted); else EMIT(uint32_t, false, sc.tiles_unsorted);
    }
#undef EMIT
    KERNEL_CHECK(debug, s);
  }
  if (cap > 0) {
    ProfScope prof(PROF_SORT_FRONT, s);
    // the output pointers of the last scatter are not needed yet (single pass) / are the scratch's own (two passes)
    int rc = sort_pairs_begin(sc.plan, sc.tiles_unsorted, sc.vals_unsorted, sc.tiles_sorted, nullptr, o.offsets + (P - 1),
                              sc.sort_temp, s, debug);
    if (rc != IBGS_OK) return rc;
  }
  return IBGS_OK;
}

int run_binning_finish(int P, int debug, int views, const OrderState& o, ScratchState& sc, uint2* ranges, BinningState& b,
                       dim3 grid, cudaStream_t s) {
  const uint32_t tiles = grid.x * grid.y * (uint32_t)views;
  const uint32_t* n_dev = o.offsets + (P - 1);
  const bool one_pass = sc.plan.npass == 1;
  if (sc.plan.n_cap > 0) {
    ProfScope prof(PROF_SORT, s);
    // one pass: the sorted tile ids are not written at all (every store of the scatter is a separate 32-byte sector
    // request -- the pass is bound by their rate -- and the ranges already say which tile a list position belongs to)
    int rc = sort_pairs_finish(sc.plan, sc.tiles_unsorted, sc.vals_unsorted, one_pass ? nullptr : (void*)sc.tiles_sorted,
                               b.point_list, n_dev, sc.sort_temp, one_pass ? ranges : nullptr, tiles, s, debug);
    if (rc != IBGS_OK) return rc;
  }
  if (!one_pass || sc.plan.n_cap == 0) {
    ProfScope prof(PROF_RANGES, s);
    CUDA_TRY(cudaMemsetAsync(ranges, 0, (size_t)tiles * sizeof(uint2), s));
    COUNT_LAUNCH();
    if (sc.plan.n_cap > 0) {
      const size_t R = sc.plan.n_cap;   // grid sized for the capacity, the kernel stops at *n_dev
      if (sc.plan.key_bytes == 2)
        identify_tile_ranges_kernel<uint16_t><<<(int)((R + 256 * 8 - 1) / (256 * 8)), 256, 0, s>>>(
            n_dev, (uint32_t)R, reinterpret_cast<const uint16_t*>(sc.tiles_sorted), ranges);
      else
        identify_tile_ranges_kernel<uint32_t><<<(int)((R + 256 * 4 - 1) / (256 * 4)), 256, 0, s>>>(
            n_dev, (uint32_t)R, sc.tiles_sorted, ranges);
      KERNEL_CHECK(debug, s);
    }
  }
  return IBGS_OK;
}
