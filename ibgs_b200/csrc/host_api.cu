// host_api.cu -- host-buffer convenience entry points (ibgs_forward_h, ibgs_forward_backward_h, ibgs_dist2_h).
//
// What a caller without a device allocator binds (cgo / JNI / ctypes on plain host arrays): every
// pointer in the argument struct is a HOST pointer; inputs are staged to the device, the device entry
// point runs on a private stream, outputs are copied back before returning.  State buffers are freed
// on return: ibgs_forward_h is the inference call (reference analogue: render.py's no_grad loop,
// render.py:297); ibgs_forward_backward_h runs one training view -- forward, then backward with the
// caller's cotangents while the state is still on the device -- and returns outputs and gradients.
#include "common.cuh"
#include <cstring>
#include <vector>

namespace {

struct DevPool {
  cudaStream_t s = nullptr;
  std::vector<void*> ptrs;
  void* get(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    if (cudaMallocAsync(&p, bytes, s) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    return p;
  }
  template <typename T>
  const T* up(const T* host, size_t count) {
    if (!host) return nullptr;
    void* p = get(count * sizeof(T));
    if (!p) return nullptr;
    if (cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s) != cudaSuccess) return nullptr;
    return (const T*)p;
  }
  void release() {
    for (void* p : ptrs) cudaFreeAsync(p, s);
    ptrs.clear();
  }
};

struct HostCall {
  DevPool pool;
  void* state[4] = {nullptr, nullptr, nullptr, nullptr};   // last GEOM / BINNING / IMAGE buffer handed to the library
};

void* pool_alloc(void* user, int which, size_t bytes) {
  HostCall* c = (HostCall*)user;
  void* p = c->pool.get(bytes);
  if (which == IBGS_BUF_GEOM || which == IBGS_BUF_BINNING || which == IBGS_BUF_IMAGE) c->state[which & 3] = p;
  return p;
}

template <typename T>
struct Out {
  T* host;
  T* dev;
  size_t count;
};


// stages the forward's host inputs, allocates device outputs, runs ibgs_forward on call.pool.s and queues the copies of the
// outputs back to the host; `d` keeps the device-side argument struct for a backward that may follow
int64_t forward_staged(HostCall& call, IbgsForwardArgs* h, IbgsForwardArgs& d) {
  DevPool& pool = call.pool;
  const size_t P = (size_t)h->P;
  const IbgsView& hv = h->view;
  const size_t N = (size_t)hv.image_width * hv.image_height;
  const size_t nb = (size_t)hv.nb_src_images;
  const size_t M = (size_t)hv.sh_coeffs;

  d = *h;
  d.view.bg = pool.up(hv.bg, 3);
  d.view.viewmatrix = pool.up(hv.viewmatrix, 16);
  d.view.projmatrix = pool.up(hv.projmatrix, 16);
  d.view.campos = pool.up(hv.campos, 3);
  d.view.ref_to_src_list = pool.up(hv.ref_to_src_list, nb * 16);
  d.view.src_cam_pos = pool.up(hv.src_cam_pos, nb * 3);
  d.view.src_images = pool.up(hv.src_images, nb * 3 * N);
  d.view.src_rendered_depths = pool.up(hv.src_rendered_depths, nb * N);
  d.means3D = pool.up(h->means3D, P * 3);
  d.shs = pool.up(h->shs, h->shs_rest ? P * 3 : P * M * 3);
  d.shs_rest = pool.up(h->shs_rest, M > 1 ? P * (M - 1) * 3 : 0);
  d.colors_precomp = pool.up(h->colors_precomp, P * 3);
  d.opacities = pool.up(h->opacities, P);
  d.scales = pool.up(h->scales, P * 3);
  d.rotations = pool.up(h->rotations, P * 4);
  d.cov3D_precomp = pool.up(h->cov3D_precomp, P * 6);
  d.all_map = pool.up(h->all_map, P * 5);

  std::vector<Out<float>> fouts = {{h->out_color, nullptr, 3 * N},
                                   {h->out_normal_map, nullptr, 3 * N},
                                   {h->out_median_intersected_depth, nullptr, N},
                                   {h->out_cam_feat, nullptr, 4 * MAX_SRC * N},
                                   {h->out_warped_image, nullptr, 3 * MAX_SRC * N},
                                   {h->out_min_depth_diff, nullptr, N},
                                   {h->out_camera_ray, nullptr, 3 * N}};
  for (auto& o : fouts) {
    o.dev = (float*)pool.get(o.count * sizeof(float));
    if (!o.dev) { ibgs_set_error("device allocation failed"); return IBGS_EALLOC; }
    cudaMemsetAsync(o.dev, 0, o.count * sizeof(float), pool.s);
  }
  int32_t* radii_d = (int32_t*)pool.get((P ? P : 1) * sizeof(int32_t));
  int32_t* mask_d = (int32_t*)pool.get(N * sizeof(int32_t));
  if (!radii_d || !mask_d) { ibgs_set_error("device allocation failed"); return IBGS_EALLOC; }
  cudaMemsetAsync(radii_d, 0, (P ? P : 1) * sizeof(int32_t), pool.s);
  cudaMemsetAsync(mask_d, 0, N * sizeof(int32_t), pool.s);
  d.out_color = fouts[0].dev;
  d.out_normal_map = fouts[1].dev;
  d.out_median_intersected_depth = fouts[2].dev;
  d.out_cam_feat = fouts[3].dev;
  d.out_warped_image = fouts[4].dev;
  d.out_min_depth_diff = fouts[5].dev;
  d.out_camera_ray = fouts[6].dev;
  d.radii = radii_d;
  d.out_use_first_src_frame = mask_d;
  d.alloc = pool_alloc;
  d.alloc_user = &call;

  int64_t R = ibgs_forward(&d, pool.s);
  if (R >= 0) {
    for (auto& o : fouts)
      if (o.host) cudaMemcpyAsync(o.host, o.dev, o.count * sizeof(float), cudaMemcpyDeviceToHost, pool.s);
    if (h->radii) cudaMemcpyAsync(h->radii, radii_d, P * sizeof(int32_t), cudaMemcpyDeviceToHost, pool.s);
    if (h->out_use_first_src_frame)
      cudaMemcpyAsync(h->out_use_first_src_frame, mask_d, N * sizeof(int32_t), cudaMemcpyDeviceToHost, pool.s);
    h->tex_generation_out = d.tex_generation_out;
    h->scratch_capacity_out = d.scratch_capacity_out;
  }
  return R;
}

int64_t finish(HostCall& call, int64_t rc, const char* who) {
  call.pool.release();
  cudaError_t e = cudaStreamSynchronize(call.pool.s);
  cudaStreamDestroy(call.pool.s);
  if (rc >= 0 && e != cudaSuccess) {
    ibgs_set_error("%s: %s", who, cudaGetErrorString(e));
    return IBGS_ECUDA;
  }
  return rc;
}

}  // namespace

extern "C" int64_t ibgs_forward_h(IbgsForwardArgs* h) {
  if (!h) { ibgs_set_error("args is NULL"); return IBGS_EINVAL; }
  HostCall call;
  CUDA_TRY(cudaStreamCreateWithFlags(&call.pool.s, cudaStreamNonBlocking));
  IbgsForwardArgs d;
  const int64_t R = forward_staged(call, h, d);
  return finish(call, R, "ibgs_forward_h");
}

extern "C" int64_t ibgs_forward_backward_h(IbgsForwardArgs* h, IbgsBackwardArgs* hb) {
  if (!h || !hb) { ibgs_set_error("args is NULL"); return IBGS_EINVAL; }
  if (!hb->dL_dout_color) { ibgs_set_error("dL_dout_color must be given"); return IBGS_EINVAL; }
  HostCall call;
  CUDA_TRY(cudaStreamCreateWithFlags(&call.pool.s, cudaStreamNonBlocking));
  DevPool& pool = call.pool;
  IbgsForwardArgs d;
  const int64_t R = forward_staged(call, h, d);
  if (R < 0 || h->P == 0) return finish(call, R, "ibgs_forward_backward_h");

  const size_t P = (size_t)h->P;
  const size_t N = (size_t)h->view.image_width * h->view.image_height;
  const size_t M = (size_t)h->view.sh_coeffs;
  const bool split = h->shs_rest != nullptr;
  IbgsBackwardArgs b;
  memset(&b, 0, sizeof(b));
  b.P = h->P;
  b.R = R;
  b.view = d.view;
  b.means3D = d.means3D; b.shs = d.shs; b.shs_rest = d.shs_rest; b.colors_precomp = d.colors_precomp;
  b.scales = d.scales; b.rotations = d.rotations; b.cov3D_precomp = d.cov3D_precomp; b.all_map = d.all_map;
  b.radii = d.radii;
  b.out_median_intersected_depth = d.out_median_intersected_depth;
  b.out_warped_image = d.out_warped_image;
  b.geom_buffer = call.state[IBGS_BUF_GEOM & 3];
  b.binning_buffer = call.state[IBGS_BUF_BINNING & 3];
  b.image_buffer = call.state[IBGS_BUF_IMAGE & 3];
  b.tex_generation = d.tex_generation_out;
  b.dL_dout_color = pool.up(hb->dL_dout_color, 3 * N);
  b.dL_dout_normal_map = pool.up(hb->dL_dout_normal_map, 3 * N);
  b.dL_dout_median_intersected_depth = pool.up(hb->dL_dout_median_intersected_depth, N);
  b.dL_dout_warped_image = pool.up(hb->dL_dout_warped_image, 3 * MAX_SRC * N);
  if (h->view.render_geo && (!b.dL_dout_normal_map || !b.dL_dout_median_intersected_depth || !b.dL_dout_warped_image)) {
    ibgs_set_error("render_geo needs the normal / depth / warped cotangents");
    return finish(call, IBGS_EINVAL, "ibgs_forward_backward_h");
  }
  const size_t sh_rows = split ? 1 : M;
  std::vector<Out<float>> gouts = {{hb->dL_dmeans3D, nullptr, P * 3},        {hb->dL_dmeans2D, nullptr, P * 3},
                                   {hb->dL_dmeans2D_abs, nullptr, P * 3},    {hb->dL_dcolors, nullptr, P * 3},
                                   {hb->dL_dopacity, nullptr, P},            {hb->dL_dcov3D, nullptr, h->cov3D_precomp ? P * 6 : 0},
                                   {hb->dL_dsh, nullptr, M ? P * sh_rows * 3 : 0},
                                   {hb->dL_dsh_rest, nullptr, split ? P * (M - 1) * 3 : 0},
                                   {hb->dL_dscales, nullptr, P * 3},         {hb->dL_drotations, nullptr, P * 4},
                                   {hb->dL_dall_map, nullptr, P * 5}};
  for (auto& o : gouts) {
    if (!o.count) continue;
    o.dev = (float*)pool.get(o.count * sizeof(float));
    if (!o.dev) { ibgs_set_error("device allocation failed"); return finish(call, IBGS_EALLOC, "ibgs_forward_backward_h"); }
  }
  b.dL_dmeans3D = gouts[0].dev; b.dL_dmeans2D = gouts[1].dev; b.dL_dmeans2D_abs = gouts[2].dev; b.dL_dcolors = gouts[3].dev;
  b.dL_dopacity = gouts[4].dev; b.dL_dcov3D = gouts[5].dev; b.dL_dsh = gouts[6].dev; b.dL_dsh_rest = gouts[7].dev;
  b.dL_dscales = gouts[8].dev; b.dL_drotations = gouts[9].dev; b.dL_dall_map = gouts[10].dev;
  b.alloc = pool_alloc;
  b.alloc_user = &call;
  b.accumulate_mask = 0;
  int rc = ibgs_backward(&b, pool.s);
  if (rc == IBGS_OK) {
    for (auto& o : gouts)
      if (o.host && o.dev) cudaMemcpyAsync(o.host, o.dev, o.count * sizeof(float), cudaMemcpyDeviceToHost, pool.s);
  }
  return finish(call, rc == IBGS_OK ? R : rc, "ibgs_forward_backward_h");
}

extern "C" int ibgs_dist2_h(int32_t P, const float* points_host, float* mean_dists_host) {
  if (P < 0) { ibgs_set_error("P must be >= 0"); return IBGS_EINVAL; }
  if (P == 0) return IBGS_OK;
  if (!points_host || !mean_dists_host) { ibgs_set_error("null pointer"); return IBGS_EINVAL; }
  DevPool pool;
  CUDA_TRY(cudaStreamCreateWithFlags(&pool.s, cudaStreamNonBlocking));
  const float* pts = pool.up(points_host, (size_t)P * 3);
  float* out = (float*)pool.get((size_t)P * sizeof(float));
  size_t sb = ibgs_dist2_scratch_bytes(P);
  void* scratch = pool.get(sb);
  int rc = IBGS_EALLOC;
  if (pts && out && scratch) {
    rc = ibgs_dist2(P, pts, out, scratch, sb, pool.s);
    if (rc == IBGS_OK) cudaMemcpyAsync(mean_dists_host, out, (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, pool.s);
  } else {
    ibgs_set_error("device allocation failed");
  }
  pool.release();
  cudaError_t e = cudaStreamSynchronize(pool.s);
  cudaStreamDestroy(pool.s);
  if (rc == IBGS_OK && e != cudaSuccess) {
    ibgs_set_error("ibgs_dist2_h: %s", cudaGetErrorString(e));
    return IBGS_ECUDA;
  }
  return rc;
}
