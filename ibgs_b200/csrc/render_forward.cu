// render_forward.cu -- forward tile renderer.
//
// Reference behaviour: FORWARD::renderCUDA<3,5> (cuda_rasterizer/forward.cu:303-665): per 16x16 tile,
// front-to-back alpha blend of colour and plane normal, the median-buffer plane-intersection depth
// (ring of ceil(BL/2) entries with T>0.5 + first floor(BL/2) entries with T<=0.5, :450-463), the
// depth-only variant (:466-489) and the multi-view warp epilogue (:512-663).
//
// Kernel structure (this project's own):
//   * one CTA (256 threads) per tile, but the eight warps are AUTONOMOUS: warp w owns an 8x4 pixel
//     sub-tile and walks the tile's sorted list on its own, 32 instances per step, with no CTA barrier
//     in the loop (the reference synchronises the whole CTA twice per 256 instances, forward.cu:405,414,
//     so every warp waits for the slowest one);
//   * per step each lane fetches ONE 64-byte per-Gaussian record with cp.async (LDGSTS, L1-allocating so
//     the other seven warps of the tile hit L1) into a per-warp double buffer in shared memory -- the
//     copy of step c+1 is in flight while step c is blended; colour and plane parameters ride in the
//     record, so the pair loop never touches global memory (the reference re-reads features / all_map
//     from global for every blended pair, forward.cu:433-448);
//   * each lane tests its own Gaussian against the warp's sub-tile with the conservative alpha>=1/255
//     extent stored in the record, the warp ballots, and only the survivors are walked; rejected
//     Gaussians are exactly ones the reference would skip at forward.cu:425 for all 32 pixels, so
//     per-pixel results are unchanged;
//   * warp-level early termination: a warp leaves as soon as its 32 pixels are done;
//   * the median ring lives in registers (compile-time BL, updates written as selects; the depth-only ring is a
//     shift register), not in a dynamically indexed local array (the reference kernel carries a 192-byte local
//     stack for it);
//   * template PPL: one pixel per lane (8 warps per tile) or two pixels per lane (4 warps per tile, 8x8 pixels per
//     warp); the second variant is taken for depth-only launches with long tile lists, where there is no epilogue
//     to pay for twice per lane.
#include "common.cuh"

namespace {

struct FwdArgs {
  const uint2* ranges;
  const uint32_t* point_list;
  const float4* rec;
  int W, H;
  float focal_x, focal_y, cx, cy;
  const float* viewmatrix;
  const float* ref_to_src_list;
  const float* src_cam_pos;
  cudaTextureObject_t texColor, texDepth;
  int nb_src;
  const float* cam_pos;
  const float* bg;
  float depth_error_threshold;
  float* final_T;
  uint32_t* n_contrib;
  float* sum_w;
  uint32_t* low;
  uint32_t* high;
  int32_t* valid_idx;
  float* valid_w;
  float* out_color;
  float* out_normal;
  float* out_depth;
  float* out_cam_feat;
  float* out_warped;
  float* out_min_depth_diff;
  float* out_camera_ray;
  int32_t* out_mask;
};

enum { MODE_COLOR = 0, MODE_GEO = 1, MODE_DEPTH = 2 };

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Per-pixel epilogue (forward.cu:496-664): final transmittance / contributor count, colour + background, depth-only
// quotient, and in render_geo mode the multi-view warp block.  One call per pixel of the lane.
template <int MODE, int BL>
__device__ __forceinline__ void render_forward_epilogue(const FwdArgs& a, const float* s_ref_to_src,
                                                        const float* s_src_cam_pos, uint32_t pix_id, float pixfx,
                                                        float pixfy, int HW, float T, uint32_t last_contributor,
                                                        const float (&C)[3], const float (&normal_accum)[3],
                                                        const float (&zb)[BL], const float (&wb)[BL],
                                                        const float (&db)[BL], const uint32_t (&cb)[BL],
                                                        float total_buffer_weight, float weighted_depth_sum) {
  const int W = a.W, H = a.H;
  const float epsilon = 1.0e-8f;
  if (MODE != MODE_DEPTH || a.final_T != nullptr) {  // the batched depth launch keeps no image state
    a.final_T[pix_id] = T;
    a.n_contrib[pix_id] = last_contributor;
  }

  if (MODE != MODE_DEPTH) {
#pragma unroll
    for (int ch = 0; ch < 3; ch++) a.out_color[ch * HW + pix_id] = C[ch] + T * a.bg[ch];
  }
  if (MODE == MODE_DEPTH) {
    a.out_depth[pix_id] = weighted_depth_sum / (total_buffer_weight + epsilon);  // forward.cu:508
  }
  if (MODE == MODE_GEO) {
    // forward.cu:512-663
    const float inv_focal_x = 1.0f / a.focal_x;
    const float inv_focal_y = 1.0f / a.focal_y;
    const float pix_diff_x = pixfx - a.cx;
    const float pix_diff_y = pixfy - a.cy;
    const float focal_x = a.focal_x, focal_y = a.focal_y, cx = a.cx, cy = a.cy;
    const int nb_src = a.nb_src;
    float median_intersected_depth = 0.0f;
    float total_buffer_weight_local = 0.0f;
    float total_w_src[MAX_SRC];
    float warped_color_all[MAX_SRC * 3];
#pragma unroll
    for (int s = 0; s < MAX_SRC; s++) {
      total_w_src[s] = 0.f;
      warped_color_all[3 * s] = 0.f; warped_color_all[3 * s + 1] = 0.f; warped_color_all[3 * s + 2] = 0.f;
    }
    uint32_t low_c = cb[0], high_c = cb[0];
#pragma unroll
    for (int i = 0; i < BL; i++) {
      const float weight = wb[i];
      if (weight != 0.0f) {
        const float idepth = -db[i] / zb[i];  // the postponed forward.cu:439-442 division, same operands
        const float3 ipt = {pix_diff_x * idepth * inv_focal_x, pix_diff_y * idepth * inv_focal_y, idepth};
#pragma unroll
        for (int s = 0; s < MAX_SRC; s++) {
          if (s < nb_src) {
            const float* m = &s_ref_to_src[s * 16];
            const float tx = m[0] * ipt.x + m[1] * ipt.y + m[2] * ipt.z + m[3] * 1.0f;
            const float ty = m[4] * ipt.x + m[5] * ipt.y + m[6] * ipt.z + m[7] * 1.0f;
            const float tz = m[8] * ipt.x + m[9] * ipt.y + m[10] * ipt.z + m[11] * 1.0f;
            const float inv_z = 1.0f / (tz + epsilon);
            const float2 pp = {tx * focal_x * inv_z + cx, ty * focal_y * inv_z + cy};
            const bool in_bounds = (pp.x >= 0.0f && pp.x <= (float)(W - 1) && pp.y >= 0.0f &&
                                    pp.y <= (float)(H - 1));
            if (in_bounds) {
              const float4 texC = tex2DLayered<float4>(a.texColor, pp.x + 0.5f, pp.y + 0.5f, s);
              warped_color_all[s * 3] += weight * texC.x;
              warped_color_all[s * 3 + 1] += weight * texC.y;
              warped_color_all[s * 3 + 2] += weight * texC.z;
              total_w_src[s] += weight;
            }
          }
        }
        total_buffer_weight_local += weight;
        median_intersected_depth += weight * idepth;
        low_c = min(low_c, cb[i]);
        high_c = max(high_c, cb[i]);
      }
    }
    a.low[pix_id] = low_c;
    a.high[pix_id] = high_c;
    a.sum_w[pix_id] = total_buffer_weight_local;
    median_intersected_depth /= (total_buffer_weight_local + epsilon);
    const float3 mpt = {pix_diff_x * median_intersected_depth * inv_focal_x,
                        pix_diff_y * median_intersected_depth * inv_focal_y, median_intersected_depth};

    const float* vm = a.viewmatrix;
    const float3 translation = {vm[12], vm[13], vm[14]};
    const float3 pcw = {mpt.x - translation.x, mpt.y - translation.y, mpt.z - translation.z};
    const float3 mpw = {vm[0] * pcw.x + vm[1] * pcw.y + vm[2] * pcw.z,
                        vm[4] * pcw.x + vm[5] * pcw.y + vm[6] * pcw.z,
                        vm[8] * pcw.x + vm[9] * pcw.y + vm[10] * pcw.z};
    const float cam0 = a.cam_pos[0], cam1 = a.cam_pos[1], cam2 = a.cam_pos[2];
    float3 ray_dir = {mpw.x - cam0, mpw.y - cam1, mpw.z - cam2};
    const float ray_len = sqrtf(ray_dir.x * ray_dir.x + ray_dir.y * ray_dir.y + ray_dir.z * ray_dir.z) + epsilon;
    ray_dir.x /= ray_len;
    ray_dir.y /= ray_len;
    ray_dir.z /= ray_len;
    a.out_camera_ray[0 * HW + pix_id] = ray_dir.x;
    a.out_camera_ray[1 * HW + pix_id] = ray_dir.y;
    a.out_camera_ray[2 * HW + pix_id] = ray_dir.z;

    int valid_src_count = 0;
    bool first_valid = false;
    float min_depth_error = 1.0f;
#pragma unroll
    for (int s = 0; s < MAX_SRC; s++) {
      if (s < nb_src) {
        const float* m = &s_ref_to_src[s * 16];
        const float tx = m[0] * mpt.x + m[1] * mpt.y + m[2] * mpt.z + m[3] * 1.0f;
        const float ty = m[4] * mpt.x + m[5] * mpt.y + m[6] * mpt.z + m[7] * 1.0f;
        const float tz = m[8] * mpt.x + m[9] * mpt.y + m[10] * mpt.z + m[11] * 1.0f;
        const float inv_z = 1.0f / (tz + epsilon);
        const float2 pp = {tx * focal_x * inv_z + cx, ty * focal_y * inv_z + cy};
        const bool in_bounds = (pp.x >= 0.0f && pp.x <= (float)(W - 1) && pp.y >= 0.0f &&
                                pp.y <= (float)(H - 1));
        float warped_depth = 0.0f;
        if (in_bounds) warped_depth = tex2DLayered<float>(a.texDepth, pp.x + 0.5f, pp.y + 0.5f, s);
        const float depth_error = fabsf(warped_depth - tz) * inv_z;
        if (warped_depth > 0.0f && depth_error < a.depth_error_threshold) {
          const float inv_weight = 1.0f / (total_w_src[s] + epsilon);
          const float cam[3] = {cam0, cam1, cam2};
#pragma unroll
          for (int c = 0; c < 3; c++) {
            warped_color_all[s * 3 + c] *= inv_weight;
            a.out_cam_feat[valid_src_count * 4 * HW + c * HW + pix_id] = cam[c] - s_src_cam_pos[s * 3 + c];
            a.out_warped[valid_src_count * 3 * HW + c * HW + pix_id] = warped_color_all[s * 3 + c];
          }
          float3 sd = {mpw.x - s_src_cam_pos[s * 3], mpw.y - s_src_cam_pos[s * 3 + 1],
                       mpw.z - s_src_cam_pos[s * 3 + 2]};
          const float sl = sqrtf(sd.x * sd.x + sd.y * sd.y + sd.z * sd.z) + epsilon;
          sd.x /= sl;
          sd.y /= sl;
          sd.z /= sl;
          const float ray_dir_diff = sd.x * ray_dir.x + sd.y * ray_dir.y + sd.z * ray_dir.z;
          a.out_cam_feat[valid_src_count * 4 * HW + 3 * HW + pix_id] = ray_dir_diff;
          if (s == 0) first_valid = true;
          a.valid_idx[valid_src_count * HW + pix_id] = s;
          a.valid_w[valid_src_count * HW + pix_id] = total_w_src[s];
          valid_src_count++;
          min_depth_error = min(min_depth_error, depth_error);
        }
      }
    }
    if (valid_src_count <= MAX_SRC - 1) a.valid_idx[valid_src_count * HW + pix_id] = -1;
    // Unused slots and the first-source mask are written as zeros HERE: in render_geo mode this kernel writes every
    // word of every output, so the caller needs no zero-filled tensors (the reference binding memsets all nine
    // outputs with torch::full, rasterize_points.cu:80-90, and relies on that for these slots).
    a.out_mask[pix_id] = first_valid ? 1 : 0;
    for (int k = valid_src_count; k < MAX_SRC; k++) {
#pragma unroll
      for (int c = 0; c < 4; c++) a.out_cam_feat[(k * 4 + c) * HW + pix_id] = 0.0f;
#pragma unroll
      for (int c = 0; c < 3; c++) a.out_warped[(k * 3 + c) * HW + pix_id] = 0.0f;
    }
    a.out_min_depth_diff[pix_id] = min_depth_error;
    a.out_depth[pix_id] = median_intersected_depth;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) a.out_normal[ch * HW + pix_id] = normal_accum[ch];
  }
}

// PPL = pixels per lane.  PPL == 1: eight warps per tile, warp = 8x4 pixels.  PPL == 2: four warps per tile, warp =
// 8x8 pixels, lane = the two pixels (x, y) and (x, y + 4): half as many warps pay the per-step overhead (record fetch,
// cull, vote) of a tile, and the two pixels' blend chains interleave.  Per-pixel arithmetic and order are unchanged.
template <int MODE, int BL, int PPL>
__global__ void __launch_bounds__(256 / PPL, PPL == 1 ? 4 : 4) render_forward_kernel(const FwdArgs a) {
  constexpr int BEFORE = (BL + 1) / 2;  // forward.cu:384
  constexpr int BELOW = BL - BEFORE;    // forward.cu:385
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NW = 8 / PPL;

  // per-warp double buffer: [warp][buf][quad][lane]
  __shared__ float4 s_rec[NW][2][4][32];
  __shared__ float s_ref_to_src[MAX_SRC * 16];
  __shared__ float s_src_cam_pos[MAX_SRC * 3];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int W = a.W, H = a.H;
  // warp w -> 8 x (4 PPL) sub-tile (2 across); lane -> pixel(s) inside it
  const int sub_x0 = blockIdx.x * TILE + (warp & 1) * 8;
  const int sub_y0 = blockIdx.y * TILE + (warp >> 1) * 4 * PPL;
  const unsigned px = (unsigned)(sub_x0 + (lane & 7));
  const float pixfx = (float)px;
  const float rayx = (pixfx - a.cx) / a.focal_x;  // forward.cu:352
  unsigned py[PPL];
  uint32_t pix_id[PPL];
  float pixfy[PPL], rayy[PPL];
  bool inside[PPL], done[PPL];
#pragma unroll
  for (int q = 0; q < PPL; q++) {
    py[q] = (unsigned)(sub_y0 + (lane >> 3) + 4 * q);
    // blockIdx.z = view of a batched depth-only launch (0 otherwise): its tile ranges and output plane follow view z-1's
    pix_id[q] = (MODE == MODE_DEPTH ? blockIdx.z * (uint32_t)(W * H) : 0u) + W * py[q] + px;
    pixfy[q] = (float)py[q];
    rayy[q] = (pixfy[q] - a.cy) / a.focal_y;
    inside[q] = px < (unsigned)W && py[q] < (unsigned)H;
    done[q] = !inside[q];
  }

  // sub-tile bounds for the cull test
  const float wx0 = (float)sub_x0, wx1 = (float)(sub_x0 + 7);
  const float wy0 = (float)sub_y0, wy1 = (float)(sub_y0 + 4 * PPL - 1);

  const uint2 range = a.ranges[((MODE == MODE_DEPTH ? blockIdx.z * gridDim.y : 0u) + blockIdx.y) * gridDim.x + blockIdx.x];
  const int total = (int)(range.y - range.x);

  if (MODE == MODE_GEO) {
    if (tid < a.nb_src * 16) s_ref_to_src[tid] = a.ref_to_src_list[tid];
    if (tid < a.nb_src * 3) s_src_cam_pos[tid] = a.src_cam_pos[tid];
    __syncthreads();  // the only CTA barrier: epilogue constants
  }

  const float epsilon = 1.0e-8f;
  float T[PPL];
  uint32_t last_contributor[PPL];
  float C[PPL][3], normal_accum[PPL][3];
  float zb[PPL][BL], wb[PPL][BL], db[PPL][BL];   // GEO: zb = depth denominator, db = plane distance; DEPTH: zb = depth
  uint32_t cb[PPL][BL];
  int before_ptr[PPL], below_count[PPL];
  float total_buffer_weight[PPL], weighted_depth_sum[PPL];
  // depth-only with BELOW==0 (BL==1): the reference `break`s out of the current 256-instance batch only
  // (forward.cu:484-488) and resumes with the next one
  bool brk[PPL];
#pragma unroll
  for (int q = 0; q < PPL; q++) {
    T[q] = 1.0f;
    last_contributor[q] = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) { C[q][i] = 0.f; normal_accum[q][i] = 0.f; }
#pragma unroll
    for (int k = 0; k < BL; k++) { zb[q][k] = 0.f; wb[q][k] = 0.f; db[q][k] = 0.f; cb[q][k] = 0u; }
    before_ptr[q] = 0;
    below_count[q] = 0;
    total_buffer_weight[q] = 0.0f;
    weighted_depth_sum[q] = 0.0f;
    brk[q] = false;
  }

  const uint32_t* plist = a.point_list + range.x;
  const int nchunks = (total + 31) >> 5;
  float4(*wrec)[4][32] = s_rec[warp];

  bool all_done = true;
#pragma unroll
  for (int q = 0; q < PPL; q++) all_done = all_done && done[q];

  // software pipeline: ids two steps ahead (register), records one step ahead (cp.async)
  uint32_t id_issue = (lane < total) ? plist[lane] : 0u;
  if (nchunks > 0 && !__all_sync(FULL, all_done)) {
    if (lane < total) {
      const float4* r = a.rec + 4 * (size_t)id_issue;
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (k < 3 || MODE != MODE_COLOR) cp_async16(&wrec[0][k][lane], r + k);
    }
    cp_async_commit();
    id_issue = (32 + lane < total) ? plist[32 + lane] : 0u;

    for (int c = 0; c < nchunks; c++) {
      const int buf = c & 1;
      const int c0 = c << 5;
      if (c + 1 < nchunks) {
        if (c0 + 32 + lane < total) {
          const float4* r = a.rec + 4 * (size_t)id_issue;
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (k < 3 || MODE != MODE_COLOR) cp_async16(&wrec[buf ^ 1][k][lane], r + k);
        }
        id_issue = (c0 + 64 + lane < total) ? plist[c0 + 64 + lane] : 0u;
      }
      cp_async_commit();
      cp_async_wait<1>();   // everything but the newest group has landed -> step c is in shared memory
      __syncwarp();
      if ((c0 & (TILE_PIX - 1)) == 0) {  // new 256-instance batch of the reference
#pragma unroll
        for (int q = 0; q < PPL; q++) brk[q] = false;
      }

      const int j = c0 + lane;
      bool keep = false;
      if (j < total) {
        const float4 q0 = wrec[buf][0][lane];
        const float4 q1 = wrec[buf][1][lane];
        keep = subtile_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, wx0, wx1, wy0, wy1);
      }
      unsigned m = __ballot_sync(FULL, keep);
      // render_geo: once EVERY pixel of the warp has T <= 0.5 and its "below" half of the median buffer full (or is
      // finished), no later pair can enter the buffer (forward.cu:450-463: T only falls, the below part only fills), so
      // the rest of the list is walked by a copy of the loop without the plane-depth sign test and the ring selects
      // (the select-based ring + its tests were most of the ALU-pipe load: ALU 52 % vs FMA 25 % in round 1's ncu).
      bool frozen = false;
      if (MODE == MODE_GEO) {
        bool lane_frozen = true;
#pragma unroll
        for (int q = 0; q < PPL; q++)
          lane_frozen = lane_frozen && (done[q] || (!(T[q] > 0.5f) && below_count[q] >= BELOW));
        frozen = __all_sync(FULL, lane_frozen);
      }
      if (frozen) {
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t contributor = (uint32_t)(c0 + b + 1);  // forward.cu:417
        const float4 g0 = wrec[buf][0][b];
        const float4 g1 = wrec[buf][1][b];
        const float dx = g0.x - pixfx;
#pragma unroll
        for (int q = 0; q < PPL; q++) {
          // (nested ifs instead of `continue`: the q loop must unroll completely so that the per-pixel arrays stay
          // in registers)
          const float dy = g0.y - pixfy[q];
          // con_o = (g0.z, g0.w, g1.x, g1.y); forward.cu:421-427
          const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
          const float alpha = min(0.99f, g1.y * exp_power(power));
          const float test_T = T[q] * (1.0f - alpha);
          const bool blend = !(done[q] || brk[q]) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
          if (blend && test_T < 0.0001f) done[q] = true;
          if (blend && !(test_T < 0.0001f)) {
          const float aT = alpha * T[q];

          const float4 g2 = wrec[buf][2][b];
          if (MODE != MODE_DEPTH) {
            C[q][0] += g2.x * aT;
            C[q][1] += g2.y * aT;
            C[q][2] += g2.z * aT;
          }
          if (MODE != MODE_COLOR) {
            const float4 g3 = wrec[buf][3][b];
            // forward.cu:439-442: intersected_depth = -d / (n.ray + eps)
            if (MODE == MODE_GEO) {
              normal_accum[q][0] += g3.x * aT;
              normal_accum[q][1] += g3.y * aT;
              normal_accum[q][2] += g3.z * aT;
            } else {  // MODE_DEPTH, forward.cu:466-489 (never reached: the lean walk is render_geo only)
              const float z_den = g3.x * rayx + g3.y * rayy[q] + g3.z + epsilon;
              const float intersected_depth = -g2.w / z_den;
              if (intersected_depth > 0.0f) {
                if (T[q] > 0.5f) {
                  // The depth-only ring is only ever consulted for the entry it evicts, so it is kept as a shift
                  // register (newest entry in slot 0, the evicted one falls out of slot BEFORE-1; empty slots hold
                  // weight 0) instead of the reference's cyclic pointer: same evicted values in the same order, and
                  // no dynamically indexed array (which the compiler would place in local memory).
                  const float old_w = wb[q][BEFORE - 1], old_z = zb[q][BEFORE - 1];
#pragma unroll
                  for (int k = BEFORE - 1; k > 0; k--) { zb[q][k] = zb[q][k - 1]; wb[q][k] = wb[q][k - 1]; }
                  zb[q][0] = intersected_depth;
                  wb[q][0] = aT;
                  total_buffer_weight[q] -= old_w;
                  weighted_depth_sum[q] -= old_w * old_z;
                  total_buffer_weight[q] += aT;
                  weighted_depth_sum[q] += aT * intersected_depth;
                } else if (below_count[q] < BELOW) {
                  below_count[q]++;
                  total_buffer_weight[q] += aT;
                  weighted_depth_sum[q] += aT * intersected_depth;
                }
                if (below_count[q] == BELOW) {
                  // BELOW>0: T<=0.5 from here on, the sums are final -> the pixel is finished.
                  // BELOW==0: reference semantics = leave this batch, resume at the next.
                  if (BELOW > 0) done[q] = true; else brk[q] = true;
                }
              }
            }
          }
          T[q] = test_T;
          last_contributor[q] = contributor;
          }
        }
      }
      } else {
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t contributor = (uint32_t)(c0 + b + 1);  // forward.cu:417
        const float4 g0 = wrec[buf][0][b];
        const float4 g1 = wrec[buf][1][b];
        const float dx = g0.x - pixfx;
#pragma unroll
        for (int q = 0; q < PPL; q++) {
          // (nested ifs instead of `continue`: the q loop must unroll completely so that the per-pixel arrays stay
          // in registers)
          const float dy = g0.y - pixfy[q];
          // con_o = (g0.z, g0.w, g1.x, g1.y); forward.cu:421-427
          const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
          const float alpha = min(0.99f, g1.y * exp_power(power));
          const float test_T = T[q] * (1.0f - alpha);
          const bool blend = !(done[q] || brk[q]) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
          if (blend && test_T < 0.0001f) done[q] = true;
          if (blend && !(test_T < 0.0001f)) {
          const float aT = alpha * T[q];

          const float4 g2 = wrec[buf][2][b];
          if (MODE != MODE_DEPTH) {
            C[q][0] += g2.x * aT;
            C[q][1] += g2.y * aT;
            C[q][2] += g2.z * aT;
          }
          if (MODE != MODE_COLOR) {
            const float4 g3 = wrec[buf][3][b];
            // forward.cu:439-442: intersected_depth = -d / (n.ray + eps)
            const float z_den = g3.x * rayx + g3.y * rayy[q] + g3.z + epsilon;
            if (MODE == MODE_GEO) {
              normal_accum[q][0] += g3.x * aT;
              normal_accum[q][1] += g3.y * aT;
              normal_accum[q][2] += g3.z * aT;
              // The ring only ever needs the depths of the entries it still holds at the end, so the IEEE
              // division is postponed to the epilogue: the ring stores the denominator (the numerator -d is
              // re-read there).  "depth > 0" is decided from the operand signs: -d/den > 0 <=> d and den are
              // non-zero with opposite signs (den = 0 gives +-inf of the wrong sign or NaN, NaN compares false;
              // the quotient cannot underflow to 0 for finite plane parameters and a ray inside the frustum).
              const bool need = (T[q] > 0.5f) || (below_count[q] < BELOW);
              const bool z_pos = (g2.w > 0.0f && z_den < 0.0f) || (g2.w < 0.0f && z_den > 0.0f);
              if (need && z_pos) {
                if (T[q] > 0.5f) {
#pragma unroll
                  for (int k = 0; k < BEFORE; k++) {   // selects, not indexed stores: the ring must stay in registers
                    const bool hit = before_ptr[q] == k;
                    zb[q][k] = hit ? z_den : zb[q][k];
                    db[q][k] = hit ? g2.w : db[q][k];
                    wb[q][k] = hit ? aT : wb[q][k];
                    cb[q][k] = hit ? contributor : cb[q][k];
                  }
                  before_ptr[q] = (before_ptr[q] + 1) % BEFORE;
                } else {
#pragma unroll
                  for (int k = 0; k < BELOW; k++) {
                    const bool hit = below_count[q] == k;
                    zb[q][BEFORE + k] = hit ? z_den : zb[q][BEFORE + k];
                    db[q][BEFORE + k] = hit ? g2.w : db[q][BEFORE + k];
                    wb[q][BEFORE + k] = hit ? aT : wb[q][BEFORE + k];
                    cb[q][BEFORE + k] = hit ? contributor : cb[q][BEFORE + k];
                  }
                  below_count[q]++;
                }
              }
            } else {  // MODE_DEPTH, forward.cu:466-489
              const float intersected_depth = -g2.w / z_den;
              if (intersected_depth > 0.0f) {
                if (T[q] > 0.5f) {
                  // The depth-only ring is only ever consulted for the entry it evicts, so it is kept as a shift
                  // register (newest entry in slot 0, the evicted one falls out of slot BEFORE-1; empty slots hold
                  // weight 0) instead of the reference's cyclic pointer: same evicted values in the same order, and
                  // no dynamically indexed array (which the compiler would place in local memory).
                  const float old_w = wb[q][BEFORE - 1], old_z = zb[q][BEFORE - 1];
#pragma unroll
                  for (int k = BEFORE - 1; k > 0; k--) { zb[q][k] = zb[q][k - 1]; wb[q][k] = wb[q][k - 1]; }
                  zb[q][0] = intersected_depth;
                  wb[q][0] = aT;
                  total_buffer_weight[q] -= old_w;
                  weighted_depth_sum[q] -= old_w * old_z;
                  total_buffer_weight[q] += aT;
                  weighted_depth_sum[q] += aT * intersected_depth;
                } else if (below_count[q] < BELOW) {
                  below_count[q]++;
                  total_buffer_weight[q] += aT;
                  weighted_depth_sum[q] += aT * intersected_depth;
                }
                if (below_count[q] == BELOW) {
                  // BELOW>0: T<=0.5 from here on, the sums are final -> the pixel is finished.
                  // BELOW==0: reference semantics = leave this batch, resume at the next.
                  if (BELOW > 0) done[q] = true; else brk[q] = true;
                }
              }
            }
          }
          T[q] = test_T;
          last_contributor[q] = contributor;
          }
        }
      }
      }
      all_done = true;
#pragma unroll
      for (int q = 0; q < PPL; q++) all_done = all_done && done[q];
      if (__all_sync(FULL, all_done)) break;   // warp-level early termination
      __syncwarp();  // every lane is done reading buf before the step after next overwrites it (a vote is not a
                     // memory-ordering barrier; compute-sanitizer racecheck flags the re-use without this)
    }
    cp_async_wait<0>();
  }

  const int HW = H * W;
#pragma unroll
  for (int q = 0; q < PPL; q++) {
    if (inside[q])
      render_forward_epilogue<MODE, BL>(a, s_ref_to_src, s_src_cam_pos, pix_id[q], pixfx, pixfy[q], HW, T[q],
                                        last_contributor[q], C[q], normal_accum[q], zb[q], wb[q], db[q], cb[q],
                                        total_buffer_weight[q], weighted_depth_sum[q]);
  }
}

// average tile-list length from which the two-pixels-per-lane variant is used (see render_backward.cu)
constexpr long long FWD_DENSE_LIST = 512;
int g_fwd_variant = 0;   // 0 = choose per view, 1 / 2 = forced (ibgs_set_forward_variant)

template <int MODE>
int dispatch_bl(int BL, dim3 grid, cudaStream_t s, const FwdArgs& fa, bool two_per_lane) {
  switch (BL) {
#define CASE_BL(n)                                                                              \
  case n:                                                                                       \
    if (two_per_lane) render_forward_kernel<MODE, n, 2><<<grid, 128, 0, s>>>(fa);               \
    else render_forward_kernel<MODE, n, 1><<<grid, 256, 0, s>>>(fa);                            \
    break;
    CASE_BL(1) CASE_BL(2) CASE_BL(3) CASE_BL(4) CASE_BL(5) CASE_BL(6) CASE_BL(7) CASE_BL(8)
#undef CASE_BL
    default:
      ibgs_set_error("buffer_length must be in [1,%d], got %d", MAX_BL, BL);
      return IBGS_EINVAL;
  }
  return IBGS_OK;
}

// Measured (3M Gaussians @1080p): the two-pixel variant wins in depth-only mode on long tile lists (render stage of
// 4 batched views 0.98 -> 0.84 ms) and LOSES in render_geo / colour mode (1.97 -> 2.31 ms: 106 registers, and the
// texture-bound epilogue runs twice per lane on half as many warps), so only depth-only launches take it by default.
bool forward_two_per_lane(int mode, long long R, dim3 grid, int views) {
  if (g_fwd_variant) return g_fwd_variant == 2;
  return mode == MODE_DEPTH && R > FWD_DENSE_LIST * (long long)grid.x * grid.y * views;
}

}  // namespace

extern "C" int ibgs_set_forward_variant(int pixels_per_lane) {
  if (pixels_per_lane < 0 || pixels_per_lane > 2) {
    ibgs_set_error("pixels_per_lane must be 0 (choose per view), 1 or 2, got %d", pixels_per_lane);
    return IBGS_EINVAL;
  }
  g_fwd_variant = pixels_per_lane;
  return IBGS_OK;
}

int launch_render_forward(const IbgsForwardArgs& f, const GeomState& g, const ImageState& im,
                          const BinningState& b, TexPair tex, float focal_x, float focal_y, dim3 grid, int64_t R,
                          cudaStream_t s) {
  FwdArgs fa;
  fa.ranges = im.ranges;
  fa.point_list = b.point_list;
  fa.rec = g.rec;
  fa.W = f.view.image_width;
  fa.H = f.view.image_height;
  fa.focal_x = focal_x;
  fa.focal_y = focal_y;
  fa.cx = float(fa.W * 0.5f);  // rasterizer_impl.cu:477
  fa.cy = float(fa.H * 0.5f);
  fa.viewmatrix = f.view.viewmatrix;
  fa.ref_to_src_list = f.view.ref_to_src_list;
  fa.src_cam_pos = f.view.src_cam_pos;
  fa.texColor = tex.color;
  fa.texDepth = tex.depth;
  fa.nb_src = f.view.nb_src_images;
  fa.cam_pos = f.view.campos;
  fa.bg = f.view.bg;
  fa.depth_error_threshold = f.view.depth_error_threshold;
  fa.final_T = im.final_T;
  fa.n_contrib = im.n_contrib;
  fa.sum_w = im.sum_w;
  fa.low = im.low;
  fa.high = im.high;
  fa.valid_idx = im.valid_idx;
  fa.valid_w = im.valid_w;
  fa.out_color = f.out_color;
  fa.out_normal = f.out_normal_map;
  fa.out_depth = f.out_median_intersected_depth;
  fa.out_cam_feat = f.out_cam_feat;
  fa.out_warped = f.out_warped_image;
  fa.out_min_depth_diff = f.out_min_depth_diff;
  fa.out_camera_ray = f.out_camera_ray;
  fa.out_mask = f.out_use_first_src_frame;

  int rc;
  const bool two = forward_two_per_lane(f.view.render_geo ? MODE_GEO : (f.view.render_depth_only ? MODE_DEPTH : MODE_COLOR), R, grid, 1);
  ProfScope prof(PROF_RENDER_FWD, s);
  // the reference evaluates render_geo before render_depth_only inside one kernel; with both set it
  // does both (forward.cu:445,466).  That combination is never produced by the callers
  // (gaussian_renderer/__init__.py:94-116,277-299); render_geo wins here.
  if (f.view.render_geo)
    rc = dispatch_bl<MODE_GEO>(f.view.buffer_length, grid, s, fa, two);
  else if (f.view.render_depth_only)
    rc = dispatch_bl<MODE_DEPTH>(f.view.buffer_length, grid, s, fa, two);
  else {
    if (two) render_forward_kernel<MODE_COLOR, 1, 2><<<grid, 128, 0, s>>>(fa);
    else render_forward_kernel<MODE_COLOR, 1, 1><<<grid, 256, 0, s>>>(fa);
    rc = IBGS_OK;
  }
  if (rc != IBGS_OK) return rc;
  KERNEL_CHECK(f.view.debug, s);
  return IBGS_OK;
}

// ibgs_forward_depth_batch: V depth-only views in one launch (gridDim.z = V); ranges has V * tiles entries, rec holds
// the V * P (view, Gaussian) records, out_depths is [V,1,H,W].
int launch_render_depth_batch(const IbgsDepthBatchArgs& f, const GeomState& g, const uint2* ranges,
                              const BinningState& b, float focal_x, float focal_y, dim3 grid, int64_t R,
                              cudaStream_t s) {
  FwdArgs fa = {};
  fa.ranges = ranges;
  fa.point_list = b.point_list;
  fa.rec = g.rec;
  fa.W = f.image_width;
  fa.H = f.image_height;
  fa.focal_x = focal_x;
  fa.focal_y = focal_y;
  fa.cx = float(fa.W * 0.5f);
  fa.cy = float(fa.H * 0.5f);
  fa.out_depth = f.out_depths;
  ProfScope prof(PROF_RENDER_FWD, s);
  int rc = dispatch_bl<MODE_DEPTH>(f.buffer_length, dim3(grid.x, grid.y, (unsigned)f.V), s, fa,
                                   forward_two_per_lane(MODE_DEPTH, R, grid, f.V));
  if (rc != IBGS_OK) return rc;
  KERNEL_CHECK(f.debug, s);
  return IBGS_OK;
}
