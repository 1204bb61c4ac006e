// api.cu -- extern "C" entry points declared in include/ibgs_b200.h.
//
// Orchestration counterpart of CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (cuda_rasterizer/rasterizer_impl.cu:258-270,320-515,519-666).  Differences from the reference
// orchestration: every launch goes to the caller's stream; no cudaMalloc / cudaFree /
// cudaDeviceSynchronize on the path (the reference does all three per call for its textures,
// :80-99,135-148); the single unavoidable host read -- num_rendered, which sizes the binning buffers
// (:429-434) -- is an async copy into pinned memory followed by a stream (not device) synchronise.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <atomic>
#include <mutex>

std::atomic<long long> g_launch_count{0};

namespace {
thread_local char g_err[1024] = "";
// Pinned host words for the num_rendered read-back (256 bytes), ONE SLOT PER HOST THREAD: two threads driving two
// streams / GPUs concurrently must not see each other's count between the async copy and the host read.
thread_local int32_t* g_pinned_count = nullptr;
std::mutex g_prof_mutex;   // the per-stage event timer below is process-global

// num_rendered of the last forward per (P, W, H, views) shape: sizes the speculative binning scratch of the next one
struct RHint { int P, W, H, V; int64_t R; };
std::mutex g_hint_mutex;
RHint g_hints[8] = {};
int g_hint_next = 0;
thread_local cudaEvent_t g_count_event = nullptr;   // marks the end of the num_rendered read-back (one per host thread)
}  // namespace

int64_t r_hint_get(int P, int W, int H, int V) {
  std::lock_guard<std::mutex> lock(g_hint_mutex);
  for (const RHint& h : g_hints)
    if (h.P == P && h.W == W && h.H == H && h.V == V) return h.R;
  return 0;
}
void r_hint_set(int P, int W, int H, int V, int64_t R) {
  std::lock_guard<std::mutex> lock(g_hint_mutex);
  for (RHint& h : g_hints)
    if (h.P == P && h.W == W && h.H == H && h.V == V) { h.R = R; return; }
  g_hints[g_hint_next] = RHint{P, W, H, V, R};
  g_hint_next = (g_hint_next + 1) % 8;
}
static int count_event(cudaEvent_t* ev) {
  if (!g_count_event) CUDA_TRY(cudaEventCreateWithFlags(&g_count_event, cudaEventDisableTiming));
  *ev = g_count_event;
  return IBGS_OK;
}

// ---- per-stage event timer -------------------------------------------------------------------------
namespace {
struct ProfPair { cudaEvent_t a, b; int id; };
bool g_prof_on = false;
std::vector<ProfPair> g_prof_pending;
std::vector<ProfPair> g_prof_free;
ProfPair g_prof_open[PROF_COUNT];
bool g_prof_is_open[PROF_COUNT] = {false};
double g_prof_ms[PROF_COUNT] = {0};
long long g_prof_n[PROF_COUNT] = {0};
const char* g_prof_names[PROF_COUNT] = {"preprocess", "depth_order_sort", "scan", "duplicate_with_keys", "radix_sort", "identify_tile_ranges",
                                        "texture_fill", "render_forward", "render_backward", "preprocess_backward",
                                        "ssim_forward", "ssim_backward", "tile_sort_histograms",
                                        "color_features_forward", "color_features_backward"};
void prof_drain() {
  for (auto& p : g_prof_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      g_prof_ms[p.id] += ms;
      g_prof_n[p.id] += 1;
    }
    g_prof_free.push_back(p);
  }
  g_prof_pending.clear();
}
}  // namespace

void prof_begin(int id, cudaStream_t s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  ProfPair p;
  if (!g_prof_free.empty()) {
    p = g_prof_free.back();
    g_prof_free.pop_back();
  } else {
    cudaEventCreate(&p.a);
    cudaEventCreate(&p.b);
  }
  p.id = id;
  cudaEventRecord(p.a, s);
  g_prof_open[id] = p;
  g_prof_is_open[id] = true;
}
void prof_end(int id, cudaStream_t s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (!g_prof_is_open[id]) return;
  cudaEventRecord(g_prof_open[id].b, s);
  g_prof_pending.push_back(g_prof_open[id]);
  g_prof_is_open[id] = false;
  if (g_prof_pending.size() > 4096) prof_drain();
}
extern "C" void ibgs_profile_enable(int on) { g_prof_on = on != 0; }
extern "C" void ibgs_profile_reset(void) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  prof_drain();
  for (int i = 0; i < PROF_COUNT; i++) { g_prof_ms[i] = 0; g_prof_n[i] = 0; }
}
extern "C" int ibgs_profile_read(int id, double* ms_total, int64_t* count) {
  if (id < 0 || id >= PROF_COUNT) return IBGS_EINVAL;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  prof_drain();
  if (ms_total) *ms_total = g_prof_ms[id];
  if (count) *count = g_prof_n[id];
  return IBGS_OK;
}
extern "C" const char* ibgs_profile_name(int id) { return (id >= 0 && id < PROF_COUNT) ? g_prof_names[id] : ""; }
extern "C" int ibgs_profile_stages(void) { return PROF_COUNT; }

void ibgs_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* ibgs_last_error(void) { return g_err; }
extern "C" int ibgs_abi_version(void) { return IBGS_ABI_VERSION; }
extern "C" int64_t ibgs_launch_count(void) { return g_launch_count.load(); }
extern "C" void ibgs_release_cached(void) {
  textures_release_all();
  if (g_pinned_count) {
    cudaFreeHost(g_pinned_count);
    g_pinned_count = nullptr;
  }
  if (g_count_event) {
    cudaEventDestroy(g_count_event);
    g_count_event = nullptr;
  }
  {
    std::lock_guard<std::mutex> lock(g_hint_mutex);
    for (RHint& h : g_hints) h = RHint{};
  }
}

static int check_view(const IbgsView& v, bool need_src) {
  if (v.image_width <= 0 || v.image_height <= 0) {
    ibgs_set_error("image size must be positive, got %dx%d", v.image_width, v.image_height);
    return IBGS_EINVAL;
  }
  if (!v.bg || !v.viewmatrix || !v.projmatrix || !v.campos) {
    ibgs_set_error("bg / viewmatrix / projmatrix / campos must not be NULL");
    return IBGS_EINVAL;
  }
  if (v.buffer_length < 1 || v.buffer_length > MAX_BL) {
    ibgs_set_error("buffer_length must be in [1,%d], got %d", MAX_BL, v.buffer_length);
    return IBGS_EINVAL;
  }
  if (v.nb_src_images < 0 || v.nb_src_images > MAX_SRC) {
    ibgs_set_error("nb_src_images must be in [0,%d], got %d", MAX_SRC, v.nb_src_images);
    return IBGS_EINVAL;
  }
  if (need_src && v.nb_src_images > 0 &&
      (!v.ref_to_src_list || !v.src_cam_pos || !v.src_images || !v.src_rendered_depths)) {
    ibgs_set_error("render_geo needs ref_to_src_list / src_cam_pos / src_images / src_rendered_depths");
    return IBGS_EINVAL;
  }
  return IBGS_OK;
}

extern "C" int ibgs_state_layout(int which, size_t count, size_t aux, size_t* offsets, int max,
                                 size_t* total_bytes) {
  std::vector<size_t> offs;
  size_t total = 0;
  char* base = nullptr;
  if (which == IBGS_BUF_GEOM) {
    GeomState g;
    total = carve_geom(g, base, count);
    offs = {(size_t)g.rec, (size_t)g.depths, (size_t)g.tiles_touched, (size_t)g.clamped};
  } else if (which == IBGS_BUF_IMAGE) {
    ImageState s;
    total = carve_image(s, base, count, aux);
    offs = {(size_t)s.final_T, (size_t)s.n_contrib, (size_t)s.sum_w, (size_t)s.low, (size_t)s.high,
            (size_t)s.valid_idx, (size_t)s.valid_w, (size_t)s.ranges};
  } else if (which == IBGS_BUF_BINNING) {
    BinningState b;
    total = carve_binning(b, base, count);
    offs = {(size_t)b.point_list};
  } else if (which == IBGS_BUF_SCRATCH) {
    // the (second) forward scratch request: binning temporaries for R=count instances, aux = number of tiles
    ScratchState sc;
    total = carve_scratch(sc, base, count, ibgs_sort_bits((int32_t)(aux ? aux : 1)) - 32);
    offs = {(size_t)sc.tiles_unsorted, (size_t)sc.tiles_sorted, (size_t)sc.vals_unsorted, (size_t)sc.sort_temp};
    // (count = the CAPACITY the scratch was carved for: IbgsForwardArgs.scratch_capacity_out)
  } else {
    ibgs_set_error("unknown buffer id %d", which);
    return IBGS_EINVAL;
  }
  for (int i = 0; i < (int)offs.size() && i < max; i++) offsets[i] = offs[i];
  if (total_bytes) *total_bytes = total;
  return (int)offs.size();
}

extern "C" int64_t ibgs_forward(IbgsForwardArgs* a, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  if (!a) { ibgs_set_error("args is NULL"); return IBGS_EINVAL; }
  a->tex_generation_out = 0;
  a->scratch_capacity_out = 0;
  const int P = a->P;
  if (P < 0) { ibgs_set_error("P must be >= 0"); return IBGS_EINVAL; }
  if (P == 0) return 0;  // rasterize_points.cu:101-102: outputs stay zero, rendered = 0
  const IbgsView& v = a->view;
  int rc = check_view(v, v.render_geo != 0);
  if (rc != IBGS_OK) return rc;
  if (!a->means3D || !a->opacities || !a->radii || !a->alloc) {
    ibgs_set_error("means3D / opacities / radii / alloc must not be NULL");
    return IBGS_EINVAL;
  }
  if (!a->cov3D_precomp && (!a->scales || !a->rotations)) {
    ibgs_set_error("provide scales+rotations or cov3D_precomp");
    return IBGS_EINVAL;
  }
  if (!a->colors_precomp && !a->shs && !v.render_depth_only) {
    ibgs_set_error("provide shs or colors_precomp");
    return IBGS_EINVAL;
  }
  if (a->shs_rest && (!a->shs || v.sh_coeffs < 2)) {
    ibgs_set_error("shs_rest needs shs (the DC part) and sh_coeffs >= 2");
    return IBGS_EINVAL;
  }
  if ((v.render_geo || v.render_depth_only) && !a->all_map) {
    ibgs_set_error("render_geo / render_depth_only need all_map");
    return IBGS_EINVAL;
  }
  if (!ibgs_aligned16(a->rotations)) {
    ibgs_set_error("rotations must be 16-byte aligned (read as float4); pass an aligned copy");
    return IBGS_EINVAL;
  }
  const int W = v.image_width, H = v.image_height;
  const float focal_y = H / (2.0f * v.tanfovy);  // rasterizer_impl.cu:362-363
  const float focal_x = W / (2.0f * v.tanfovx);
  dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, 1);
  const size_t N = (size_t)W * H, T = (size_t)grid.x * grid.y;

  GeomState g;
  ImageState im;
  size_t geom_bytes = carve_geom(g, nullptr, (size_t)P);
  size_t image_bytes = carve_image(im, nullptr, N, T);
  char* geom_base = (char*)a->alloc(a->alloc_user, IBGS_BUF_GEOM, geom_bytes);
  char* image_base = (char*)a->alloc(a->alloc_user, IBGS_BUF_IMAGE, image_bytes);
  if (!geom_base || !image_base) { ibgs_set_error("allocator returned NULL"); return IBGS_EALLOC; }
  carve_geom(g, geom_base, (size_t)P);
  carve_image(im, image_base, N, T);

  // textures first: their fill overlaps nothing it depends on and keeps the stream busy
  TexPair tex = {0, 0};
  if (v.render_geo) {
    rc = textures_acquire(W, H, v.nb_src_images, v.src_images, v.src_rendered_depths, s, &tex,
                          &a->tex_generation_out, 0);
    if (rc != IBGS_OK) return rc;
  }

  if (!g_pinned_count) CUDA_TRY(cudaHostAlloc((void**)&g_pinned_count, 256, cudaHostAllocDefault));
  cudaEvent_t ev = nullptr;
  rc = count_event(&ev);
  if (rc != IBGS_OK) return rc;

  // We do not know R yet: the first scratch request holds the P-sized depth-order state, the second one the R-sized
  // binning temporaries.  Both stay alive until the forward's last launch is enqueued.
  OrderState ord;
  const size_t order_bytes = carve_order(ord, nullptr, (size_t)P);
  char* order_base = (char*)a->alloc(a->alloc_user, IBGS_BUF_SCRATCH, order_bytes);
  if (!order_base) { ibgs_set_error("allocator returned NULL"); return IBGS_EALLOC; }
  carve_order(ord, order_base, (size_t)P);

  rc = launch_preprocess(*a, g, focal_x, focal_y, grid, s);
  if (rc != IBGS_OK) return rc;
  rc = run_depth_order(g, ord, (size_t)P, s);
  if (rc != IBGS_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(g_pinned_count, ord.offsets + (P - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaEventRecord(ev, s));

  // Speculative binning: emission and the front half of the tile sort are queued NOW, into a scratch sized from the
  // instance count this (P, W, H) shape produced last time, so the GPU is busy while the host waits for the 4 bytes.
  // The kernels read R on the device and never touch memory beyond the capacity; if R turns out larger (first call,
  // scene change) they are simply queued again with the exact size -- the reference's order of events.
  const int tile_bits = ibgs_sort_bits((int32_t)T) - 32;
  ScratchState sc;
  size_t cap = 0;
  static const bool no_spec = getenv("IBGS_NO_SPECULATION") != nullptr;   // (A/B switch for profiles/NOTES.md)
  const int64_t hint = no_spec ? 0 : r_hint_get(P, W, H, 1);
  if (hint > 0) {
    cap = (size_t)(hint + hint / 8 + 4096);
    if (cap > 0x7fffffffull) cap = 0x7fffffffull;
    const size_t bytes = carve_scratch(sc, nullptr, cap, tile_bits);
    char* base = (char*)a->alloc(a->alloc_user, IBGS_BUF_SCRATCH, bytes);
    if (!base) { ibgs_set_error("allocator returned NULL"); return IBGS_EALLOC; }
    carve_scratch(sc, base, cap, tile_bits);
    rc = run_binning_begin(P, a->radii, v.debug, 1, g, ord, sc, cap, grid, s);
    if (rc != IBGS_OK) return rc;
  }
  CUDA_TRY(cudaEventSynchronize(ev));   // the copy only -- NOT the kernels queued behind it
  const uint32_t R_u = *(volatile uint32_t*)g_pinned_count;
  if (R_u > 0x7fffffffu) {
    // the reference stores this in an int (rasterizer_impl.cu:429) and would overflow
    ibgs_set_error("num_rendered %u exceeds the int32 limit of the reference layout", R_u);
    return IBGS_ELIMIT;
  }
  const int64_t R = (int64_t)R_u;
  r_hint_set(P, W, H, 1, R);
  if ((size_t)R > cap || hint <= 0) {
    cap = (size_t)R;
    const size_t bytes = carve_scratch(sc, nullptr, cap, tile_bits);
    char* base = (char*)a->alloc(a->alloc_user, IBGS_BUF_SCRATCH, bytes);
    if (!base) { ibgs_set_error("allocator returned NULL"); return IBGS_EALLOC; }
    carve_scratch(sc, base, cap, tile_bits);
    rc = run_binning_begin(P, a->radii, v.debug, 1, g, ord, sc, cap, grid, s);
    if (rc != IBGS_OK) return rc;
  }
  a->scratch_capacity_out = (int64_t)cap;

  BinningState b;
  size_t bin_bytes = carve_binning(b, nullptr, (size_t)R);
  char* bin_base = (char*)a->alloc(a->alloc_user, IBGS_BUF_BINNING, bin_bytes);
  if (!bin_base) { ibgs_set_error("allocator returned NULL"); return IBGS_EALLOC; }
  carve_binning(b, bin_base, (size_t)R);
  rc = run_binning_finish(P, v.debug, 1, ord, sc, im.ranges, b, grid, s);
  if (rc != IBGS_OK) return rc;

  rc = launch_render_forward(*a, g, im, b, tex, focal_x, focal_y, grid, R, s);
  if (rc != IBGS_OK) return rc;
  return R;
}

// Batched depth-only forward: the V*P (view, Gaussian) items are binned as ONE list -- a single stable depth sort over
// all items interleaves the views, the tile id of an instance carries its view (view * T + tile), and the stable tile
// sort brings every (view, tile) list back into ascending depth order with ties in ascending Gaussian id, i.e. exactly
// the list a separate ibgs_forward call of that view builds.
extern "C" int64_t ibgs_forward_depth_batch(IbgsDepthBatchArgs* a, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  if (!a) { ibgs_set_error("args is NULL"); return IBGS_EINVAL; }
  const int P = a->P, V = a->V;
  if (P < 0) { ibgs_set_error("P must be >= 0"); return IBGS_EINVAL; }
  if (V < 1 || V > IBGS_MAX_DEPTH_BATCH) {
    ibgs_set_error("V must be in [1,%d], got %d", IBGS_MAX_DEPTH_BATCH, V);
    return IBGS_EINVAL;
  }
  if (a->num_rendered) for (int v = 0; v < V; v++) a->num_rendered[v] = 0;
  if (P == 0) return 0;  // like rasterize_points.cu:101-102: the caller's zero-filled outputs stay
  if (a->image_width <= 0 || a->image_height <= 0) {
    ibgs_set_error("image size must be positive, got %dx%d", a->image_width, a->image_height);
    return IBGS_EINVAL;
  }
  if (a->buffer_length < 1 || a->buffer_length > MAX_BL) {
    ibgs_set_error("buffer_length must be in [1,%d], got %d", MAX_BL, a->buffer_length);
    return IBGS_EINVAL;
  }
  if (!a->viewmatrices || !a->projmatrices || !a->means3D || !a->opacities || !a->out_depths || !a->alloc) {
    ibgs_set_error("viewmatrices / projmatrices / means3D / opacities / out_depths / alloc must not be NULL");
    return IBGS_EINVAL;
  }
  if (!a->cov3D_precomp && (!a->scales || !a->rotations)) {
    ibgs_set_error("provide scales+rotations or cov3D_precomp");
    return IBGS_EINVAL;
  }
  if (!a->all_maps && (!a->normals || !a->camera_centers)) {
    ibgs_set_error("provide all_maps, or normals + camera_centers");
    return IBGS_EINVAL;
  }
  if (!ibgs_aligned16(a->rotations)) {
    ibgs_set_error("rotations must be 16-byte aligned (read as float4); pass an aligned copy");
    return IBGS_EINVAL;
  }
  if ((int64_t)P * V > 0x7fffffffLL) {
    ibgs_set_error("P*V = %lld exceeds the int32 item limit", (long long)P * V);
    return IBGS_ELIMIT;
  }
  const int W = a->image_width, H = a->image_height;
  const float focal_y = H / (2.0f * a->tanfovy);
  const float focal_x = W / (2.0f * a->tanfovx);
  dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, 1);
  const size_t T = (size_t)grid.x * grid.y;
  const size_t items = (size_t)P * V;

  // one scratch request for the item-sized state: records / depths / tiles_touched, radii (when the caller does not
  // want them), tile ranges, per-view counters, depth-order state
  GeomState g;
  OrderState ord;
  const size_t geom_bytes = carve_geom(g, nullptr, items);
  const size_t order_bytes = carve_order(ord, nullptr, items);
  const size_t radii_bytes = a->radii ? 0 : align_up(items * sizeof(int32_t), 256);
  const size_t ranges_bytes = align_up(T * V * sizeof(uint2), 256);
  const size_t counts_bytes = 256;
  char* base = (char*)a->alloc(a->alloc_user, IBGS_BUF_SCRATCH,
                               geom_bytes + order_bytes + radii_bytes + ranges_bytes + counts_bytes);
  if (!base) { ibgs_set_error("allocator returned NULL"); return IBGS_EALLOC; }
  carve_geom(g, base, items);
  carve_order(ord, base + geom_bytes, items);
  int32_t* radii = a->radii ? a->radii : (int32_t*)(base + geom_bytes + order_bytes);
  uint2* ranges = (uint2*)(base + geom_bytes + order_bytes + radii_bytes);
  unsigned long long* counts = (unsigned long long*)(base + geom_bytes + order_bytes + radii_bytes + ranges_bytes);
  static_assert(IBGS_MAX_DEPTH_BATCH * sizeof(unsigned long long) <= 256, "counter block");

  if (!g_pinned_count) CUDA_TRY(cudaHostAlloc((void**)&g_pinned_count, 256, cudaHostAllocDefault));
  CUDA_TRY(cudaMemsetAsync(counts, 0, counts_bytes, s));
  COUNT_LAUNCH();
  int rc = launch_preprocess_depth_batch(*a, g, radii, counts, focal_x, focal_y, grid, s);
  if (rc != IBGS_OK) return rc;
  rc = run_depth_order(g, ord, items, s);
  if (rc != IBGS_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(g_pinned_count, counts, V * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  unsigned long long R_u = 0;
  for (int v = 0; v < V; v++) {
    const unsigned long long rv = ((volatile unsigned long long*)g_pinned_count)[v];
    if (a->num_rendered) a->num_rendered[v] = (int64_t)rv;
    R_u += rv;
  }
  if (R_u > 0x7fffffffull) {
    ibgs_set_error("num_rendered %llu over the %d views exceeds the int32 limit; render fewer views per call", R_u, V);
    return IBGS_ELIMIT;
  }
  const int64_t R = (int64_t)R_u;

  BinningState b;
  ScratchState sc;
  const size_t bin_bytes = carve_binning(b, nullptr, (size_t)R);
  const size_t scratch_bytes = carve_scratch(sc, nullptr, (size_t)R, ibgs_sort_bits((int32_t)(T * V)) - 32);
  char* base2 = (char*)a->alloc(a->alloc_user, IBGS_BUF_SCRATCH, bin_bytes + scratch_bytes);
  if (!base2) { ibgs_set_error("allocator returned NULL"); return IBGS_EALLOC; }
  carve_binning(b, base2, (size_t)R);
  carve_scratch(sc, base2 + bin_bytes, (size_t)R, ibgs_sort_bits((int32_t)(T * V)) - 32);
  rc = run_binning_begin((int)items, radii, a->debug, V, g, ord, sc, (size_t)R, grid, s);
  if (rc != IBGS_OK) return rc;
  rc = run_binning_finish((int)items, a->debug, V, ord, sc, ranges, b, grid, s);
  if (rc != IBGS_OK) return rc;
  rc = launch_render_depth_batch(*a, g, ranges, b, focal_x, focal_y, grid, R, s);
  if (rc != IBGS_OK) return rc;
  return R;
}

extern "C" int ibgs_backward(IbgsBackwardArgs* a, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  if (!a) { ibgs_set_error("args is NULL"); return IBGS_EINVAL; }
  const int P = a->P;
  if (P < 0) { ibgs_set_error("P must be >= 0"); return IBGS_EINVAL; }
  if (P == 0) return IBGS_OK;
  const IbgsView& v = a->view;
  int rc = check_view(v, v.render_geo != 0);
  if (rc != IBGS_OK) return rc;
  if (!a->geom_buffer || !a->binning_buffer || !a->image_buffer || !a->alloc || !a->means3D || !a->radii ||
      !a->dL_dout_color) {
    ibgs_set_error("state buffers / alloc / means3D / radii / dL_dout_color must not be NULL");
    return IBGS_EINVAL;
  }
  if (v.render_geo && (!a->dL_dout_normal_map || !a->dL_dout_median_intersected_depth ||
                       !a->dL_dout_warped_image || !a->out_median_intersected_depth || !a->out_warped_image)) {
    ibgs_set_error("render_geo backward needs the normal/depth/warped cotangents and saved outputs");
    return IBGS_EINVAL;
  }
  if (!a->dL_dmeans3D || !a->dL_dmeans2D || !a->dL_dmeans2D_abs || !a->dL_dcolors || !a->dL_dopacity ||
      !a->dL_dscales || !a->dL_drotations || !a->dL_dall_map || (a->shs && !a->dL_dsh) ||
      (a->shs && a->shs_rest && !a->dL_dsh_rest)) {
    ibgs_set_error("gradient output pointers must not be NULL");
    return IBGS_EINVAL;
  }
  if (!ibgs_aligned16(a->rotations) || !ibgs_aligned16(a->dL_drotations)) {
    ibgs_set_error("rotations / dL_drotations must be 16-byte aligned (accessed as float4); pass aligned tensors");
    return IBGS_EINVAL;
  }
  const int W = v.image_width, H = v.image_height;
  const float focal_y = H / (2.0f * v.tanfovy);  // rasterizer_impl.cu:578-579
  const float focal_x = W / (2.0f * v.tanfovx);
  dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, 1);
  const size_t N = (size_t)W * H, T = (size_t)grid.x * grid.y;

  GeomState g;
  ImageState im;
  BinningState b;
  carve_geom(g, (char*)a->geom_buffer, (size_t)P);
  carve_image(im, (char*)a->image_buffer, N, T);
  carve_binning(b, (char*)a->binning_buffer, (size_t)a->R);

  // one scratch request: [P][16] accumulation arena (zeroed) | per-pixel median-pair lists (render_geo only)
  const size_t arena_bytes = align_up((size_t)P * 64, 256);
  const size_t ent_bytes = render_backward_scratch_bytes(N, v.buffer_length, v.render_geo);
  float4* arena = (float4*)a->alloc(a->alloc_user, IBGS_BUF_SCRATCH, arena_bytes + ent_bytes);
  if (!arena) { ibgs_set_error("allocator returned NULL"); return IBGS_EALLOC; }
  void* ent_scratch = ent_bytes ? (void*)((char*)arena + arena_bytes) : nullptr;
  CUDA_TRY(cudaMemsetAsync(arena, 0, arena_bytes, s));
  COUNT_LAUNCH();

  TexPair tex = {0, 0};
  if (v.render_geo) {
    int64_t gen = 0;
    rc = textures_acquire(W, H, v.nb_src_images, v.src_images, v.src_rendered_depths, s, &tex, &gen,
                          a->tex_generation);
    if (rc != IBGS_OK) return rc;
  }
  rc = launch_render_backward(*a, g, im, b, tex, focal_x, focal_y, grid, arena, ent_scratch, s);
  if (rc != IBGS_OK) return rc;
  rc = launch_preprocess_backward(*a, g, arena, focal_x, focal_y, s);
  return rc;
}

extern "C" int ibgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                                 const float* projmatrix, uint8_t* present, void* stream) {
  if (P < 0) { ibgs_set_error("P must be >= 0"); return IBGS_EINVAL; }
  if (P == 0) return IBGS_OK;
  if (!means3D || !viewmatrix || !present) { ibgs_set_error("null pointer"); return IBGS_EINVAL; }
  return launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, (cudaStream_t)stream);
}
