// sort.cu -- this project's own device-wide primitives for the binning stage: a stable LSD radix sort of (key, value)
// pairs (digits of up to 8 bits), and the gathered inclusive scan of tiles_touched.
//
// Reference behaviour they replace: cub::DeviceRadixSort::SortPairs (rasterizer_impl.cu:452-457; simple_knn.cu:210-213)
// and cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:426).  Both are fully specified (stable ascending sort on a bit
// range; prefix sum), so any correct implementation is bit-identical; round 1 called the same CUB primitives.
//
// Design (deterministic, no atomics on global memory, no decoupled look-back): one pass over a digit of <= 8 bits is
//   hist     one CTA per 4096-item tile: digit histogram in shared memory -> row `tile` of a [tiles][bins] matrix;
//   colsum   column sums per segment of 32 tiles; thread = (bin, segment);
//   colscan  the matrix is scanned down its columns in place (exclusive): cell (tile, bin) = number of items with that
//            digit in earlier tiles (earlier segments' sums + a 32-long serial scan per thread); column totals;
//   reorder  one CTA per tile ranks its items stably (per-warp ballot matching, row by row), sorts them by digit inside
//            shared memory and copies every run of equal digits to its final place with coalesced stores.
// A tile id has 13 bits at 1080p -> two passes (7 + 6 bits) over the 15 M instances; the depth key is sorted per
// Gaussian (3 M items, four 8-bit passes), not per instance.  Round 2 first tried ONE pass with 2^13 bins (one warp per
// 16 K-item chunk, running offsets in 32 KB of shared memory, __match_any_sync groups): correct, but every store is its
// own 32-byte sector request and every step pays a 32-address shared-memory atomic -- 0.45 ms for the tile sort against
// 0.26 ms for the library it replaced (profiles/NOTES.md).  Sorting inside shared memory first is what makes the
// global stores coalesce.
// With `ranges_out` a single-pass sort also emits the tile ranges (first / one-past-last list position of every tile;
// empty tiles keep (0,0) like rasterizer_impl.cu:459-467) straight from the bin totals.
#include "common.cuh"

namespace {

constexpr int SEG = 32;  // chunks per column-scan segment

__device__ __forceinline__ uint32_t load_key(const void* keys, int key_bytes, size_t i) {
  return key_bytes == 2 ? (uint32_t)reinterpret_cast<const uint16_t*>(keys)[i] : reinterpret_cast<const uint32_t*>(keys)[i];
}

template <int KEY_BYTES>
__global__ void __launch_bounds__(256) sort_hist_kernel(const void* __restrict__ keys, const uint32_t* __restrict__ n_dev,
                                                        uint32_t n_cap, int shift, uint32_t mask, int chunk,
                                                        uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t s_bins[];
  const uint32_t NB = mask + 1;
  const uint32_t n = n_dev ? min(*n_dev, n_cap) : n_cap;
  const size_t begin = (size_t)blockIdx.x * chunk;
  if (begin >= n) return;
  const size_t end = min((size_t)n, begin + chunk);
  for (uint32_t b = threadIdx.x; b < NB; b += blockDim.x) s_bins[b] = 0;
  __syncthreads();
  if (KEY_BYTES == 2 && ((begin & 7) == 0)) {
    // 8 uint16 keys per 16-byte load
    const uint4* k4 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(keys) + begin);
    const size_t nvec = (end - begin) / 8;
    for (size_t v = threadIdx.x; v < nvec; v += blockDim.x) {
      const uint4 q = k4[v];
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        atomicAdd(&s_bins[((w[j] & 0xffffu) >> shift) & mask], 1u);
        atomicAdd(&s_bins[((w[j] >> 16) >> shift) & mask], 1u);
      }
    }
    for (size_t i = begin + nvec * 8 + threadIdx.x; i < end; i += blockDim.x)
      atomicAdd(&s_bins[(load_key(keys, 2, i) >> shift) & mask], 1u);
  } else {
    for (size_t i = begin + threadIdx.x; i < end; i += blockDim.x)
      atomicAdd(&s_bins[(load_key(keys, KEY_BYTES, i) >> shift) & mask], 1u);
  }
  __syncthreads();
  uint32_t* row = hist + (size_t)blockIdx.x * NB;
  for (uint32_t b = threadIdx.x; b < NB; b += blockDim.x) row[b] = s_bins[b];
}

// column sums per segment of SEG chunks: seg_total[seg][bin]; thread = (bin, segment)
__global__ void __launch_bounds__(256) sort_colsum_kernel(const uint32_t* __restrict__ n_dev, uint32_t n_cap, int chunk,
                                                          uint32_t NB, const uint32_t* __restrict__ hist,
                                                          uint32_t* __restrict__ seg_total) {
  const uint32_t n = n_dev ? min(*n_dev, n_cap) : n_cap;
  const uint32_t nchunks = (uint32_t)(((size_t)n + chunk - 1) / chunk);
  const uint32_t bin = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t seg = blockIdx.y;
  const uint32_t c0 = seg * SEG;
  if (bin >= NB) return;
  uint32_t sum = 0;
  if (c0 < nchunks) {
    const uint32_t c1 = min(nchunks, c0 + SEG);
#pragma unroll 8
    for (uint32_t c = c0; c < c1; c++) sum += hist[(size_t)c * NB + bin];
  }
  seg_total[(size_t)seg * NB + bin] = sum;
}

// exclusive scan down the columns of hist[chunks][NB], in place; thread = (bin, segment)
__global__ void __launch_bounds__(256) sort_colscan_kernel(const uint32_t* __restrict__ n_dev, uint32_t n_cap, int chunk,
                                                           uint32_t NB, uint32_t* __restrict__ hist,
                                                           const uint32_t* __restrict__ seg_total,
                                                           uint32_t* __restrict__ bin_total) {
  const uint32_t n = n_dev ? min(*n_dev, n_cap) : n_cap;
  const uint32_t nchunks = (uint32_t)(((size_t)n + chunk - 1) / chunk);
  const uint32_t bin = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t seg = blockIdx.y;
  const uint32_t c0 = seg * SEG;
  if (bin >= NB) return;
  if (c0 >= nchunks) {
    if (seg == 0) bin_total[bin] = 0;   // n == 0
    return;
  }
  uint32_t running = 0;
  for (uint32_t sg = 0; sg < seg; sg++) running += seg_total[(size_t)sg * NB + bin];
  const uint32_t c1 = min(nchunks, c0 + SEG);
  uint32_t v[SEG];
#pragma unroll
  for (int k = 0; k < SEG; k++) v[k] = (c0 + k < c1) ? hist[(size_t)(c0 + k) * NB + bin] : 0u;
#pragma unroll
  for (int k = 0; k < SEG; k++) {
    if (c0 + k < c1) hist[(size_t)(c0 + k) * NB + bin] = running;
    running += v[k];
  }
  if (c1 == nchunks) bin_total[bin] = running;   // the last live segment holds the column total
}

// Reorder + scatter of one 4096-item tile by a digit of <= 8 bits.  VAL_IS_INDEX: the value of item i is i itself (first
// pass of an index sort: no identity-permutation array).
//   1. warp w owns a contiguous 1/8 of the tile as RT_ITEMS rows of 32; row by row (= input order) the lanes with
//      equal digits find each other with one ballot per digit bit, the group's first lane bumps the warp's private
//      counter of that digit, every lane learns its rank among the warp's items with the same digit;
//   2. a block scan over (digit, warp) turns the counters into local positions: the tile's items, sorted stably by
//      digit, are written into shared memory;
//   3. thread t copies items t, t + 256, ... to  bin_start[digit] + (items of that digit in earlier tiles: the
//      column-scanned matrix) + (position inside the tile's run of that digit): runs of equal digits are contiguous
//      in shared memory AND in the output, so the global stores are coalesced.
#ifndef IBGS_SORT_ITEMS
#define IBGS_SORT_ITEMS 8
#endif
#ifndef IBGS_SORT_CTAS
#define IBGS_SORT_CTAS 4
#endif
constexpr int RT_ITEMS = IBGS_SORT_ITEMS;    // rows per warp
constexpr int RT_TILE = 256 * RT_ITEMS;      // items per CTA
template <int KEY_BYTES, bool VAL_IS_INDEX>
__global__ void __launch_bounds__(256, IBGS_SORT_CTAS) sort_reorder_kernel(const void* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                          void* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                          const uint32_t* __restrict__ n_dev, uint32_t n_cap, int shift,
                                                          int nbits, const uint32_t* __restrict__ colbase,
                                                          const uint32_t* __restrict__ bin_total, uint2* ranges_out,
                                                          uint32_t num_ranges) {
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ uint32_t s_wh[8][256];        // per-warp digit counters, then: items of the digit in earlier warps
  __shared__ uint32_t s_dstart[256];       // first local position of each digit
  __shared__ uint32_t s_gbase[256];        // global position of the tile's first item of each digit
  __shared__ uint32_t s_scan[2][8];
  __shared__ uint32_t s_key[RT_TILE];
  __shared__ uint32_t s_val[RT_TILE];
  const uint32_t NB = 1u << nbits, mask = NB - 1;
  const uint32_t n = n_dev ? min(*n_dev, n_cap) : n_cap;
  const size_t tile_begin = (size_t)blockIdx.x * RT_TILE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool live = tile_begin < n;
  const uint32_t cnt = live ? (uint32_t)min((size_t)RT_TILE, (size_t)n - tile_begin) : 0u;
#pragma unroll
  for (int w = 0; w < 8; w++) s_wh[w][tid] = 0;
  __syncthreads();

  uint32_t key[RT_ITEMS], val[RT_ITEMS], rank[RT_ITEMS];
#pragma unroll
  for (int r = 0; r < RT_ITEMS; r++) {
    const uint32_t idx = warp * (32 * RT_ITEMS) + r * 32 + lane;
    const size_t i = tile_begin + idx;
    key[r] = (idx < cnt) ? load_key(keys_in, KEY_BYTES, i) : 0u;
    val[r] = VAL_IS_INDEX ? (uint32_t)i : ((idx < cnt) ? vals_in[i] : 0u);
  }
  // (1) ranks inside the warp, row by row
#pragma unroll
  for (int r = 0; r < RT_ITEMS; r++) {
    const uint32_t idx = warp * (32 * RT_ITEMS) + r * 32 + lane;
    const bool valid = idx < cnt;
    const uint32_t d = (key[r] >> shift) & mask;
    unsigned peers = __ballot_sync(FULL, valid);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k < nbits) {
        const bool bit = (d >> k) & 1u;
        const unsigned b = __ballot_sync(FULL, bit);
        peers &= bit ? b : ~b;
      }
    }
    const int leader = __ffs(peers) - 1;   // (peers contains this lane whenever valid)
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = s_wh[warp][d];
      s_wh[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(FULL, old, valid ? leader : 0);
    rank[r] = old + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  // (2) thread t = digit t: prefix over the warps, then block scans over the digits (tile-local counts and global totals)
  uint32_t tot = 0, gtot = 0;
  if ((uint32_t)tid < NB) {
#pragma unroll
    for (int w = 0; w < 8; w++) {
      const uint32_t c = s_wh[w][tid];
      s_wh[w][tid] = tot;
      tot += c;
    }
    gtot = bin_total[tid];
  }
  uint32_t inc0 = tot, inc1 = gtot;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    const uint32_t u0 = __shfl_up_sync(FULL, inc0, dd), u1 = __shfl_up_sync(FULL, inc1, dd);
    if (lane >= dd) { inc0 += u0; inc1 += u1; }
  }
  if (lane == 31) { s_scan[0][warp] = inc0; s_scan[1][warp] = inc1; }
  __syncthreads();
  uint32_t off0 = 0, off1 = 0;
  for (int w = 0; w < warp; w++) { off0 += s_scan[0][w]; off1 += s_scan[1][w]; }
  const uint32_t dstart = off0 + inc0 - tot;     // exclusive
  const uint32_t gstart = off1 + inc1 - gtot;    // bin_start of the digit
  if ((uint32_t)tid < NB) {
    s_dstart[tid] = dstart;
    s_gbase[tid] = gstart + (live ? colbase[(size_t)blockIdx.x * NB + tid] : 0u);
    if (blockIdx.x == 0 && ranges_out && (uint32_t)tid < num_ranges)
      ranges_out[tid] = gtot ? make_uint2(gstart, gstart + gtot) : make_uint2(0u, 0u);
  }
  __syncthreads();
  if (!live) return;
  // sorted order inside the tile -> shared memory
#pragma unroll
  for (int r = 0; r < RT_ITEMS; r++) {
    const uint32_t idx = warp * (32 * RT_ITEMS) + r * 32 + lane;
    if (idx < cnt) {
      const uint32_t d = (key[r] >> shift) & mask;
      const uint32_t pos = s_dstart[d] + s_wh[warp][d] + rank[r];
      s_key[pos] = key[r];
      s_val[pos] = val[r];
    }
  }
  __syncthreads();
  // (3) coalesced copy-out
#pragma unroll
  for (int k = 0; k < RT_ITEMS; k++) {
    const uint32_t idx = k * 256 + tid;
    if (idx < cnt) {
      const uint32_t kk = s_key[idx];
      const uint32_t d = (kk >> shift) & mask;
      const size_t pos = (size_t)s_gbase[d] + (idx - s_dstart[d]);
      if (keys_out) {
        if (KEY_BYTES == 2) reinterpret_cast<uint16_t*>(keys_out)[pos] = (uint16_t)kk;
        else reinterpret_cast<uint32_t*>(keys_out)[pos] = kk;
      }
      vals_out[pos] = s_val[idx];
    }
  }
}

// ---- gathered inclusive scan: out[i] = sum_{j<=i} src[idx[j]] ------------------------------------------------------
constexpr int SCAN_ITEMS = 16;                       // per thread
constexpr int SCAN_BLOCK = 256 * SCAN_ITEMS;         // per CTA
__global__ void __launch_bounds__(256) scan_blocksum_kernel(uint32_t n, const uint32_t* __restrict__ idx,
                                                            const uint32_t* __restrict__ src, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t s_w[8];
  const size_t base = (size_t)blockIdx.x * SCAN_BLOCK;
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const size_t i = base + (size_t)k * 256 + threadIdx.x;
    if (i < n) sum += src[idx[i]];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += s_w[w];
    block_sums[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) scan_apply_kernel(uint32_t n, const uint32_t* __restrict__ idx,
                                                         const uint32_t* __restrict__ src,
                                                         const uint32_t* __restrict__ block_sums, uint32_t* __restrict__ out) {
  __shared__ uint32_t s_w[8];
  __shared__ uint32_t s_prefix;
  // sum of the earlier CTAs' totals (<= a few thousand values)
  uint32_t p = 0;
  for (uint32_t b = threadIdx.x; b < blockIdx.x; b += 256) p += block_sums[b];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) p += __shfl_xor_sync(0xffffffffu, p, d);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += s_w[w];
    s_prefix = t;
  }
  __syncthreads();
  // thread t owns SCAN_ITEMS consecutive items
  const size_t base = (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const size_t i = base + k;
    v[k] = (i < n) ? src[idx[i]] : 0u;
    sum += v[k];
  }
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if ((int)(threadIdx.x & 31) >= d) incl += t;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t wp = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); w++) wp += s_w[w];
  uint32_t run = s_prefix + wp + incl - sum;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const size_t i = base + k;
    run += v[k];
    if (i < n) out[i] = run;
  }
}

}  // namespace

// ---- host side ----------------------------------------------------------------------------------------------------
SortPlan sort_plan(size_t n_cap, int key_bits, int key_bytes) {
  SortPlan p;
  p.key_bytes = key_bytes;
  p.n_cap = n_cap;
  if (key_bits < 1) key_bits = 1;
  const int maxd = 8;                                      // digit width: 256 counters per warp in shared memory
  p.npass = (key_bits + maxd - 1) / maxd;
  int left = key_bits, shift = 0;
  for (int i = 0; i < p.npass; i++) {
    const int b = (left + (p.npass - i) - 1) / (p.npass - i);   // spread the bits evenly over the passes
    p.shift[i] = shift;
    p.bits[i] = b;
    shift += b;
    left -= b;
  }
  int maxbits = 0;
  for (int i = 0; i < p.npass; i++) maxbits = p.bits[i] > maxbits ? p.bits[i] : maxbits;
  p.chunk = RT_TILE;                                       // one CTA reorders 4096 items
  p.nchunks = (n_cap + p.chunk - 1) / p.chunk;
  if (p.nchunks == 0) p.nchunks = 1;
  const size_t NBmax = (size_t)1 << maxbits;
  const size_t nseg = (p.nchunks + SEG - 1) / SEG;
  size_t off = 0;
  auto take = [&](size_t bytes) { off = align_up(off, 256); size_t o = off; off += bytes; return o; };
  p.hist_off = take(p.nchunks * NBmax * 4);
  p.seg_off = take((nseg + 1) * NBmax * 4);                // seg_total[nseg][NB] | bin_total[NB]
  const int ntmp = p.npass > 2 ? 2 : (p.npass > 1 ? 1 : 0);
  p.tmp_stride_k = align_up(n_cap * key_bytes, 256);
  p.tmp_stride_v = align_up(n_cap * 4, 256);
  p.keys_tmp_off = ntmp ? take(ntmp * p.tmp_stride_k) : 0;
  p.vals_tmp_off = ntmp ? take(ntmp * p.tmp_stride_v) : 0;
  p.bytes = align_up(off, 256);
  return p;
}

namespace {
struct PassIO { const void* kin; const uint32_t* vin; void* kout; uint32_t* vout; bool index; };

// input / output of pass i: passes ping-pong between the temp buffers so that the LAST pass writes (keys_out, vals_out)
PassIO pass_io(const SortPlan& p, int i, const void* keys_in, const uint32_t* vals_in, void* keys_out, uint32_t* vals_out,
               char* temp) {
  // intermediate passes alternate between two temp buffers; ONLY the last pass writes (keys_out, vals_out), so
  // sort_pairs_begin never needs the outputs
  void* kt[2] = {temp + p.keys_tmp_off, temp + p.keys_tmp_off + p.tmp_stride_k};
  uint32_t* vt[2] = {reinterpret_cast<uint32_t*>(temp + p.vals_tmp_off),
                     reinterpret_cast<uint32_t*>(temp + p.vals_tmp_off + p.tmp_stride_v)};
  PassIO io;
  const bool last = i == p.npass - 1;
  io.kout = last ? keys_out : kt[i & 1];
  io.vout = last ? vals_out : vt[i & 1];
  if (i == 0) {
    io.kin = keys_in;
    io.vin = vals_in;
    io.index = vals_in == nullptr;
  } else {
    io.kin = kt[(i - 1) & 1];
    io.vin = vt[(i - 1) & 1];
    io.index = false;
  }
  return io;
}

int run_pass_front(const SortPlan& p, int i, const void* kin, const uint32_t* n_dev, char* temp, cudaStream_t s, int debug) {
  const uint32_t NB = 1u << p.bits[i], mask = NB - 1;
  uint32_t* hist = reinterpret_cast<uint32_t*>(temp + p.hist_off);
  uint32_t* seg_total = reinterpret_cast<uint32_t*>(temp + p.seg_off);
  const size_t nseg = (p.nchunks + SEG - 1) / SEG;
  uint32_t* bin_total = seg_total + nseg * NB;
  if (p.key_bytes == 2)
    sort_hist_kernel<2><<<(unsigned)p.nchunks, 256, NB * 4, s>>>(kin, n_dev, (uint32_t)p.n_cap, p.shift[i], mask, p.chunk, hist);
  else
    sort_hist_kernel<4><<<(unsigned)p.nchunks, 256, NB * 4, s>>>(kin, n_dev, (uint32_t)p.n_cap, p.shift[i], mask, p.chunk, hist);
  KERNEL_CHECK(debug, s);
  dim3 g((NB + 255) / 256, (unsigned)nseg, 1);
  sort_colsum_kernel<<<g, 256, 0, s>>>(n_dev, (uint32_t)p.n_cap, p.chunk, NB, hist, seg_total);
  KERNEL_CHECK(debug, s);
  sort_colscan_kernel<<<g, 256, 0, s>>>(n_dev, (uint32_t)p.n_cap, p.chunk, NB, hist, seg_total, bin_total);
  KERNEL_CHECK(debug, s);
  return IBGS_OK;
}

int run_pass_scatter(const SortPlan& p, int i, const PassIO& io, const uint32_t* n_dev, char* temp, uint2* ranges_out,
                     uint32_t num_ranges, cudaStream_t s, int debug) {
  const uint32_t NB = 1u << p.bits[i];
  const uint32_t* hist = reinterpret_cast<const uint32_t*>(temp + p.hist_off);
  const size_t nseg = (p.nchunks + SEG - 1) / SEG;
  const uint32_t* bin_total = reinterpret_cast<const uint32_t*>(temp + p.seg_off) + nseg * NB;
#define SCATTER(KB, IDX)                                                                                              \
  sort_reorder_kernel<KB, IDX><<<(unsigned)p.nchunks, 256, 0, s>>>(io.kin, io.vin, io.kout, io.vout, n_dev,            \
                                                                  (uint32_t)p.n_cap, p.shift[i], p.bits[i], hist,      \
                                                                  bin_total, ranges_out, num_ranges)
  if (p.key_bytes == 2) { if (io.index) SCATTER(2, true); else SCATTER(2, false); }
  else { if (io.index) SCATTER(4, true); else SCATTER(4, false); }
#undef SCATTER
  KERNEL_CHECK(debug, s);
  return IBGS_OK;
}
}  // namespace

int sort_pairs_begin(const SortPlan& p, const void* keys_in, const uint32_t* vals_in, void* keys_out, uint32_t* vals_out,
                     const uint32_t* n_dev, char* temp, cudaStream_t s, int debug) {
  if (p.n_cap == 0) return IBGS_OK;
  for (int i = 0; i < p.npass; i++) {
    const PassIO io = pass_io(p, i, keys_in, vals_in, keys_out, vals_out, temp);
    int rc = run_pass_front(p, i, io.kin, n_dev, temp, s, debug);
    if (rc != IBGS_OK) return rc;
    if (i == p.npass - 1) break;   // the last scatter is sort_pairs_finish's
    rc = run_pass_scatter(p, i, io, n_dev, temp, nullptr, 0, s, debug);
    if (rc != IBGS_OK) return rc;
  }
  return IBGS_OK;
}

int sort_pairs_finish(const SortPlan& p, const void* keys_in, const uint32_t* vals_in, void* keys_out, uint32_t* vals_out,
                      const uint32_t* n_dev, char* temp, uint2* ranges_out, uint32_t num_ranges, cudaStream_t s, int debug) {
  if (p.n_cap == 0) return IBGS_OK;
  const PassIO io = pass_io(p, p.npass - 1, keys_in, vals_in, keys_out, vals_out, temp);
  return run_pass_scatter(p, p.npass - 1, io, n_dev, temp, p.npass == 1 ? ranges_out : nullptr, num_ranges, s, debug);
}

int sort_pairs(const SortPlan& p, const void* keys_in, const uint32_t* vals_in, void* keys_out, uint32_t* vals_out,
               const uint32_t* n_dev, char* temp, cudaStream_t s, int debug) {
  int rc = sort_pairs_begin(p, keys_in, vals_in, keys_out, vals_out, n_dev, temp, s, debug);
  if (rc != IBGS_OK) return rc;
  return sort_pairs_finish(p, keys_in, vals_in, keys_out, vals_out, n_dev, temp, nullptr, 0, s, debug);
}

size_t scan_temp_bytes(size_t n) { return align_up(((n + SCAN_BLOCK - 1) / SCAN_BLOCK + 1) * 4, 256); }

int scan_gather_inclusive(size_t n, const uint32_t* idx, const uint32_t* src, uint32_t* out, void* temp, cudaStream_t s,
                          int debug) {
  if (n == 0) return IBGS_OK;
  const unsigned nb = (unsigned)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
  uint32_t* block_sums = reinterpret_cast<uint32_t*>(temp);
  scan_blocksum_kernel<<<nb, 256, 0, s>>>((uint32_t)n, idx, src, block_sums);
  KERNEL_CHECK(debug, s);
  scan_apply_kernel<<<nb, 256, 0, s>>>((uint32_t)n, idx, src, block_sums, out);
  KERNEL_CHECK(debug, s);
  return IBGS_OK;
}

// ---- C ABI (include/ibgs_b200.h) ---------------------------------------------------------------------------------------
extern "C" size_t ibgs_sort_temp_bytes(int64_t n, int key_bits, int key_bytes) {
  if (n < 0 || (key_bytes != 2 && key_bytes != 4)) return 0;
  return sort_plan((size_t)n, key_bits, key_bytes).bytes;
}
extern "C" int ibgs_sort_pairs(const void* keys_in, const uint32_t* vals_in, void* keys_out, uint32_t* vals_out, int64_t n,
                               int key_bits, int key_bytes, void* temp, size_t temp_bytes, void* stream) {
  if (n < 0 || n > 0x7fffffffLL) { ibgs_set_error("n must be in [0, 2^31), got %lld", (long long)n); return IBGS_EINVAL; }
  if (key_bytes != 2 && key_bytes != 4) { ibgs_set_error("key_bytes must be 2 or 4, got %d", key_bytes); return IBGS_EINVAL; }
  if (key_bits < 1 || key_bits > 8 * key_bytes) { ibgs_set_error("key_bits must be in [1,%d], got %d", 8 * key_bytes, key_bits); return IBGS_EINVAL; }
  if (n == 0) return IBGS_OK;
  if (!keys_in || !keys_out || !vals_out || !temp) { ibgs_set_error("null pointer"); return IBGS_EINVAL; }
  const SortPlan p = sort_plan((size_t)n, key_bits, key_bytes);
  if (temp_bytes < p.bytes) { ibgs_set_error("temp too small: %zu < %zu", temp_bytes, p.bytes); return IBGS_EINVAL; }
  return sort_pairs(p, keys_in, vals_in, keys_out, vals_out, nullptr, (char*)temp, (cudaStream_t)stream, 0);
}
extern "C" size_t ibgs_scan_temp_bytes(int64_t n) { return n < 0 ? 0 : scan_temp_bytes((size_t)n); }
extern "C" int ibgs_scan_gather(int64_t n, const uint32_t* idx, const uint32_t* src, uint32_t* out, void* temp,
                                size_t temp_bytes, void* stream) {
  if (n < 0 || n > 0x7fffffffLL) { ibgs_set_error("n must be in [0, 2^31), got %lld", (long long)n); return IBGS_EINVAL; }
  if (n == 0) return IBGS_OK;
  if (!idx || !src || !out || !temp) { ibgs_set_error("null pointer"); return IBGS_EINVAL; }
  if (temp_bytes < scan_temp_bytes((size_t)n)) { ibgs_set_error("temp too small"); return IBGS_EINVAL; }
  return scan_gather_inclusive((size_t)n, idx, src, out, temp, (cudaStream_t)stream, 0);
}
