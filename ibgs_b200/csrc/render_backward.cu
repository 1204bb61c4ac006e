// render_backward.cu -- backward tile renderer.
//
// Reference behaviour: BACKWARD::renderCUDA<3,5> (cuda_rasterizer/backward.cu:496-807) with
// bilinearInterpolateBackward (:55-109): back-to-front re-walk of each tile list, running
// "accum_rec" recurrences for colour and normal, the median-buffer depth / warped-colour terms for
// pairs whose contributor index lies in [low-1, high-1] (:693-767, including the accumulate-inside-
// the-view-loop quirk at :757-763 and the integer-coordinate linear-filter taps at :62-79), and 16
// float atomicAdd per blended pair into per-Gaussian gradient arrays (:673,770,793-804).
//
// Kernel structure (this project's own):
//   * one CTA per tile, eight AUTONOMOUS warps (8x4 pixel sub-tile each) walking the tile list back to
//     front, 32 instances per step, records fetched with cp.async into a per-warp double buffer -- no CTA
//     barrier in the loop (see render_forward.cu); steps behind the warp's last contributor are skipped
//     without being loaded;
//   * each lane culls its own Gaussian (alpha extent vs sub-tile, contributor >= max n_contrib of the
//     warp); only survivors are evaluated;
//   * the 15 per-pair gradient terms are reduced across the 32 pixels of the warp by transposing them through a
//     conflict-free shared-memory tile (warp_reduce16_smem: 4 STS.128 + 16 LDS + 1 SHFL per Gaussian) and 16
//     lanes then issue ONE red.global.add.f32 each into the Gaussian's 64-byte row of a [P][16] arena: one
//     64-byte reduction per (warp, surviving Gaussian) instead of 16 same-address atomics per (pixel, Gaussian);
//   * the median-buffer terms (backward.cu:693-767: texture taps, per-view loops, per-pixel loads) touch
//     at most buffer_length pairs per pixel but sit in the middle of the reference's pair loop, where they
//     diverge the warp.  Here the pair loop only RECORDS those pairs (Gaussian id, T, colour/normal part of
//     dL/dalpha) into a per-pixel list in global scratch; after the loop every lane walks its own pixels' lists
//     (phase B, median_pair_backward) and adds their whole gradient with vector reductions.  Phase B runs as the
//     tail of the same kernel: it is texture / dependent-load latency bound, and as a tail its waits overlap the
//     pair loops of the other CTAs on the SM (a separate kernel for it was measured and is slower);
//   * two variants (template PPL): one pixel per lane (8 warps per tile) or two pixels per lane (4 warps per tile,
//     the two pixels' terms add in registers, one reduction per 64 pixels); launch_render_backward picks per view.
// Summation order differs from the reference (which is itself run-to-run nondeterministic); the
// parity gate for gradients is relative L2 <= 1e-3.
#include "common.cuh"

// ---- build-time knobs of the packed two-pixel kernel (tools/build_variant.py -D...) ----
#ifndef IBGS_BWD_PACKED_CTAS
#define IBGS_BWD_PACKED_CTAS 5
#endif
#ifndef IBGS_ABL
#define IBGS_ABL 0        // kernel-cost ablations for profiles/NOTES.md (1: no phase B, 2: no reduction / RED, 3: no stage B)
#endif

namespace {

struct BwdArgs {
  const uint2* ranges;
  const uint32_t* point_list;
  const float4* rec;
  int W, H;
  float fx, fy;
  const float* bg;
  const float* ref_to_src_list;
  cudaTextureObject_t texColor;
  int nb_src;
  const float* depth_pixels;    // out_median_intersected_depth
  const float* warped_pixels;   // out_warped_image
  const float* final_T;
  const uint32_t* n_contrib;
  const float* sum_w;
  const uint32_t* low;
  const uint32_t* high;
  const int32_t* valid_idx;
  const float* valid_w;
  const float* dL_dpixels;
  const float* dL_dnormals;
  const float* dL_ddepths;
  const float* dL_dwarped;
  float4* arena;  // [P][4]
  // recorded median-buffer pairs: three slot-major [MAXE][N] planes (Gaussian id bits, T, colour+normal part of
  // dL/dalpha) `ent_stride` floats apart; written by the pair loop, read back by the same thread in phase B
  float* ent;
  size_t ent_stride;
};

// reference backward.cu:55-109
__forceinline__ __device__ float2 bilinearInterpolateBackward(int src_idx, cudaTextureObject_t texColor,
                                                              float2 uv, float3 dL_dwarped_color) {
  float u = uv.x + 0.5f;
  float v = uv.y + 0.5f;
  int u0 = (int)floorf(u);
  int v0 = (int)floorf(v);
  int u1 = u0 + 1;
  int v1 = v0 + 1;
  float fu = u - (float)u0;
  float fv = v - (float)v0;
  float fu1 = 1.0f - fu;
  float fv1 = 1.0f - fv;
  float4 C00 = tex2DLayered<float4>(texColor, (float)u0, (float)v0, src_idx);
  float4 C01 = tex2DLayered<float4>(texColor, (float)u1, (float)v0, src_idx);
  float4 C10 = tex2DLayered<float4>(texColor, (float)u0, (float)v1, src_idx);
  float4 C11 = tex2DLayered<float4>(texColor, (float)u1, (float)v1, src_idx);
  float3 dI_du, dI_dv;
  dI_du.x = -fv1 * C00.x + fv1 * C01.x - fv * C10.x + fv * C11.x;
  dI_du.y = -fv1 * C00.y + fv1 * C01.y - fv * C10.y + fv * C11.y;
  dI_du.z = -fv1 * C00.z + fv1 * C01.z - fv * C10.z + fv * C11.z;
  dI_dv.x = -fu1 * C00.x - fu * C01.x + fu1 * C10.x + fu * C11.x;
  dI_dv.y = -fu1 * C00.y - fu * C01.y + fu1 * C10.y + fu * C11.y;
  dI_dv.z = -fu1 * C00.z - fu * C01.z + fu1 * C10.z + fu * C11.z;
  float du = dL_dwarped_color.x * dI_du.x + dL_dwarped_color.y * dI_du.y + dL_dwarped_color.z * dI_du.z;
  float dv = dL_dwarped_color.x * dI_dv.x + dL_dwarped_color.y * dI_dv.y + dL_dwarped_color.z * dI_dv.z;
  return make_float2(du, dv);
}

// Warp reduction of 16 per-lane values through shared memory.
// Every lane holds 16 partial gradient terms of ONE Gaussian (its own pixel's share); the warp needs the 16
// column sums.  A shuffle butterfly costs 31 SHFL + 62 SEL + 31 FADD per Gaussian; transposing through shared
// memory costs 4 STS.128 + 16 LDS + 16 FADD + 1 SHFL:
//   * lane l stores its 16 values as row l of a [32][16] float tile (4 x STS.128).  Rows are 20 words apart
//     (l*20 mod 32 walks all eight 4-bank groups over a quarter-warp, so the 128-bit stores are conflict free) and
//     rows 16..31 sit 16 words further (the two half-warps then read disjoint bank halves);
//   * lane (s = l&15, h = l>>4) sums column s over rows 16h..16h+15 with four independent partial sums (16 scalar
//     LDS at immediate offsets from one base register);
//   * one xor-16 shuffle adds the two halves.  On return lanes 0..15 hold the total of slot `lane`.
// Both addresses are 32-bit shared-window offsets computed once per thread and used through inline PTX, so the
// per-Gaussian cost carries no address arithmetic.
#define RED_ROW_WORDS 20
#define RED_WARP_FLOATS (32 * RED_ROW_WORDS + 16)
__device__ __forceinline__ uint32_t red_store_addr(const float* red, int lane) {
  return (uint32_t)__cvta_generic_to_shared(red + lane * RED_ROW_WORDS + (lane >> 4) * 16);
}
__device__ __forceinline__ uint32_t red_load_addr(const float* red, int lane) {
  return (uint32_t)__cvta_generic_to_shared(red + (lane >> 4) * (16 * RED_ROW_WORDS + 16) + (lane & 15));
}
template <int OFF>
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(addr), "n"(OFF), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
template <int OFF>
__device__ __forceinline__ float lds32(uint32_t addr) {
  float x;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(x) : "r"(addr), "n"(OFF) : "memory");
  return x;
}
__forceinline__ __device__ float warp_reduce16_smem(const float (&v)[16], uint32_t st_addr, uint32_t ld_addr) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int RB = RED_ROW_WORDS * 4;  // row pitch in bytes
  sts128<0>(st_addr, v[0], v[1], v[2], v[3]);
  sts128<16>(st_addr, v[4], v[5], v[6], v[7]);
  sts128<32>(st_addr, v[8], v[9], v[10], v[11]);
  sts128<48>(st_addr, v[12], v[13], v[14], v[15]);
  __syncwarp();
  const float s0 = (lds32<0 * RB>(ld_addr) + lds32<1 * RB>(ld_addr)) + (lds32<2 * RB>(ld_addr) + lds32<3 * RB>(ld_addr));
  const float s1 = (lds32<4 * RB>(ld_addr) + lds32<5 * RB>(ld_addr)) + (lds32<6 * RB>(ld_addr) + lds32<7 * RB>(ld_addr));
  const float s2 = (lds32<8 * RB>(ld_addr) + lds32<9 * RB>(ld_addr)) + (lds32<10 * RB>(ld_addr) + lds32<11 * RB>(ld_addr));
  const float s3 = (lds32<12 * RB>(ld_addr) + lds32<13 * RB>(ld_addr)) + (lds32<14 * RB>(ld_addr) + lds32<15 * RB>(ld_addr));
  float sum = (s0 + s1) + (s2 + s3);
  sum += __shfl_xor_sync(FULL, sum, 16);
  __syncwarp();  // all reads done before the next Gaussian's rows are stored
  return sum;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}


// ---- packed FP32 (sm_100 FFMA2 / FMUL2 / FADD2): one instruction works on the two pixels of a lane ----
// The SASS forms take a 64-bit register pair OR a scalar register broadcast to both halves (Rn.F32) OR an immediate
// per operand, so "scalar x pair" products need no packing moves.  Same FMA-pipe throughput as two scalar FFMAs, HALF
// the issue slots -- and this loop is bound by instruction issue / dependent latency, not by the FMA pipe.
typedef float2 f2;
__device__ __forceinline__ f2 bc(float s) { return make_float2(s, s); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float hsum(f2 a) { return a.x + a.y; }
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// arena slots: 0,1 dmean2D.xy | 2,3 |dmean2D|.xy | 4,5,6 dconic x,y,w | 7 dopacity |
//              8,9,10 dcolor | 11 dall_map[4] | 12,13,14 dall_map[0..2] | 15 unused
// Slots 0-6 are accumulated UNSCALED: the per-view constants 0.5*W, 0.5*H (backward.cu:606-607,793-797) and the
// -0.5 of the conic terms (:799-801) are applied once per Gaussian by preprocess_backward_kernel instead of once
// per pixel-Gaussian pair here.
// MAXE = number of median-buffer pairs a pixel can record (buffer_length + 1 spare for a pair that sits
// exactly on the alpha threshold and is decided differently than in the forward)

// ---------------------------------------------------------------------------------------------------------
// phase B: the recorded median-buffer pairs of ONE pixel (backward.cu:693-767, 773-804).  Runs as the tail of the pair
// kernels: every lane walks its own pixels' lists after the loop.
// Round 2 measurements on the headline scene (profiles/NOTES.md): compiled out, the round-1 tail cost 0.61 ms of the
// 1.98 ms kernel for ~4 % of the pairs -- 1230 warp instructions per recorded pair, because every per-pixel value was
// re-loaded and the double-precision ray re-derived for every entry and all NSRC view bodies ran whatever the number of
// valid views.  Here the per-pixel values are loaded / derived ONCE per pixel and only valid, in-bounds views are walked:
// 0.48 ms.  The same body as a separate launch (one thread per pixel, or one per pixel-entry) was measured at
// 0.69-0.83 ms: it is bound by texture / dependent-load latency, and as a tail its waits overlap the other CTAs' loops.
// ---------------------------------------------------------------------------------------------------------
template <int NSRC>
__device__ __forceinline__ void median_pixel_backward(const BwdArgs& a, unsigned px, unsigned py, uint32_t pix_id,
                                                      int n_ent) {
  const float* s_ref_to_src = a.ref_to_src_list;   // <= 80 floats every thread of the grid reads: L1 / constant-like
  const int W = a.W, H = a.H;
  const int HW = H * W;
  // ---- per-pixel values, loaded / derived once ----
  int sidx[NSRC];
#pragma unroll
  for (int mm = 0; mm < NSRC; mm++) sidx[mm] = a.valid_idx[mm * HW + pix_id];
  // valid views are compacted into slots 0..nvalid-1, terminated by -1 when fewer than MAX_SRC (forward.cu:655);
  // slots past the terminator are uninitialised and never used
  int nvalid = NSRC;
#pragma unroll
  for (int mm = NSRC - 1; mm >= 0; mm--)
    if (sidx[mm] == -1) nvalid = mm;
  const float sum_w = a.sum_w[pix_id];
  const float depth_pix = a.depth_pixels[pix_id];
  const float dL_ddepth = a.dL_ddepths[pix_id];
  const float T_final = a.final_T[pix_id];
  float bg_dot_dpixel = 0;
#pragma unroll
  for (int i = 0; i < 3; i++) bg_dot_dpixel += a.bg[i] * a.dL_dpixels[i * HW + pix_id];
  const float2 pixf = {(float)px, (float)py};
  const float fx = a.fx, fy = a.fy;
  // backward.cu:545-547 (double on purpose: W*0.5 is a double expression there)
  const float2 ray = {(float)((pixf.x - W * 0.5) / fx), (float)((pixf.y - H * 0.5) / fy)};
  const float cx = float(W * 0.5f);
  const float cy = float(H * 0.5f);
  const float inv_sumw = __fdividef(1.f, sum_w);
  const float A_val = (pixf.x - cx) / fx;
  const float B_val = (pixf.y - cy) / fy;
  // dL/dwarped scaled by 1 / (sum of the view's median weights), (warped pixel) likewise: what the entries need
  float dLw_s[NSRC][3], wpix[NSRC][3];
#pragma unroll
  for (int mm = 0; mm < NSRC; mm++) {
    const bool live = mm < nvalid;
    const float inv_vw = live ? __fdividef(1.f, a.valid_w[mm * HW + pix_id]) : 0.f;
#pragma unroll
    for (int n_i = 0; n_i < 3; n_i++) {
      dLw_s[mm][n_i] = live ? a.dL_dwarped[mm * 3 * HW + n_i * HW + pix_id] * inv_vw : 0.f;
      wpix[mm][n_i] = live ? a.warped_pixels[mm * 3 * HW + n_i * HW + pix_id] : 0.f;
    }
    if (!live) sidx[mm] = 0;
  }

#pragma unroll 1
  for (int e = 0; e < n_ent; e++) {
    const float* ent = a.ent + ((size_t)e * HW + pix_id);
    const uint32_t gid = __float_as_uint(ent[0]);
    const float Te = ent[a.ent_stride];
    float dL_dalpha = ent[2 * a.ent_stride];
    const float4* r = a.rec + 4 * (size_t)gid;
    const float4 g0 = __ldg(r + 0), g1 = __ldg(r + 1), g2 = __ldg(r + 2), g3 = __ldg(r + 3);

    const float2 d = {g0.x - pixf.x, g0.y - pixf.y};
    const float power = -0.5f * (g0.z * d.x * d.x + g1.x * d.y * d.y) - g0.w * d.x * d.y;
    const float G = __expf(power);
    const float alpha = min(0.99f, g1.y * G);
    const float dchannel_dcolor = alpha * Te;
    float dL_dall_map_temp[5] = {0.f, 0.f, 0.f, 0.f, 0.f};

    const float3 normal_gauss = {g3.x, g3.y, g3.z};
    const float distance_gauss = g2.w;
    const float tmp_gauss = (normal_gauss.x * ray.x + normal_gauss.y * ray.y + normal_gauss.z + 1.0e-8);
    const float tmp_gauss2 = distance_gauss / (tmp_gauss * tmp_gauss);
    const float intersected_depth =
        -distance_gauss / (normal_gauss.x * ray.x + normal_gauss.y * ray.y + normal_gauss.z + 1.0e-8);
    const float3 ip = {(pixf.x - cx) * intersected_depth / fx, (pixf.y - cy) * intersected_depth / fy,
                       intersected_depth};
    float dL_dz = dL_ddepth * dchannel_dcolor * inv_sumw;
    dL_dalpha += dL_ddepth * (intersected_depth - depth_pix) * inv_sumw;
#pragma unroll
    for (int mm = 0; mm < NSRC; mm++) {
      const int src_idx = sidx[mm];
      const float* r2s = &s_ref_to_src[src_idx * 16];
      const float3 tp = {r2s[0] * ip.x + r2s[1] * ip.y + r2s[2] * ip.z + r2s[3],
                         r2s[4] * ip.x + r2s[5] * ip.y + r2s[6] * ip.z + r2s[7],
                         r2s[8] * ip.x + r2s[9] * ip.y + r2s[10] * ip.z + r2s[11]};
      float2 pp = {(tp.x * fx / tp.z) + cx, (tp.y * fy / tp.z) + cy};
      const bool ok = (mm < nvalid) && (pp.x >= 0 && pp.x <= W - 1 && pp.y >= 0 && pp.y <= H - 1);  // NaN -> false
      if (!ok) continue;
      const float4 texC = tex2DLayered<float4>(a.texColor, pp.x + 0.5f, pp.y + 0.5f, src_idx);
      const float wc[3] = {texC.x, texC.y, texC.z};
      float dLc[3];
      float alpha_term = 0.f;
#pragma unroll
      for (int n_i = 0; n_i < 3; n_i++) {
        dLc[n_i] = dLw_s[mm][n_i] * dchannel_dcolor;
        alpha_term += dLw_s[mm][n_i] * (wc[n_i] - wpix[mm][n_i]);
      }
      const float U = r2s[0] * A_val + r2s[1] * B_val + r2s[2];
      const float V = r2s[4] * A_val + r2s[5] * B_val + r2s[6];
      const float W_coeff = r2s[8] * A_val + r2s[9] * B_val + r2s[10];
      const float r0 = r2s[3], r1 = r2s[7], r2 = r2s[11];
      const float denom = (W_coeff * intersected_depth + r2);
      const float inv_d2 = __fdividef(1.f, denom * denom);
      const float dp_x_dd = fx * (U * r2 - W_coeff * r0) * inv_d2;
      const float dp_y_dd = fy * (V * r2 - W_coeff * r1) * inv_d2;
      const float2 dpp = bilinearInterpolateBackward(src_idx, a.texColor, pp, make_float3(dLc[0], dLc[1], dLc[2]));
      const float from_color = dpp.x * dp_x_dd + dpp.y * dp_y_dd;
      if (ok) {  // pure register arithmetic: compiles to selects
        dL_dalpha += alpha_term;
        dL_dz += from_color;
        // accumulated inside the view loop, exactly like backward.cu:757-763
        dL_dall_map_temp[4] += (-dL_dz / tmp_gauss);
        dL_dall_map_temp[0] += dL_dz * tmp_gauss2 * ray.x;
        dL_dall_map_temp[1] += dL_dz * tmp_gauss2 * ray.y;
        dL_dall_map_temp[2] += dL_dz * tmp_gauss2;
      }
    }
    dL_dalpha *= Te;
    dL_dalpha += (-T_final * rcp_approx(1.f - alpha)) * bg_dot_dpixel;
    // same UNSCALED slot convention as the pair loop (see there)
    const float dL_dG = g1.y * dL_dalpha;
    const float gdx = G * d.x;
    const float gdy = G * d.y;
    const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
    const float dG_ddely = -gdy * g1.x - gdx * g0.w;
    float4 f0, f1;
    f0.x = dL_dG * dG_ddelx;
    f0.y = dL_dG * dG_ddely;
    f0.z = fabs(f0.x);
    f0.w = fabs(f0.y);
    const float hx = dL_dG * gdx, hy = dL_dG * gdy;
    f1.x = hx * d.x;
    f1.y = hx * d.y;
    f1.z = hy * d.y;
    f1.w = G * dL_dalpha;
    float4* dst = a.arena + 4 * (size_t)gid;
    atomicAdd(dst + 0, f0);
    atomicAdd(dst + 1, f1);
    if (dL_dall_map_temp[4] != 0.f || dL_dall_map_temp[0] != 0.f || dL_dall_map_temp[1] != 0.f ||
        dL_dall_map_temp[2] != 0.f) {
      atomicAdd(reinterpret_cast<float*>(a.arena) + 16 * (size_t)gid + 11, dL_dall_map_temp[4]);
      atomicAdd(dst + 3, make_float4(dL_dall_map_temp[0], dL_dall_map_temp[1], dL_dall_map_temp[2], 0.f));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// kernel A: the pair loop
// ---------------------------------------------------------------------------------------------------------
// PPL = pixels per lane.  PPL == 1: eight warps per tile, warp = 8x4 pixels.  PPL == 2: four warps per tile, warp =
// 8x8 pixels, lane = the two pixels (x, y) and (x, y + 4): the two pixels of a lane contribute to the SAME Gaussian, so
// their 15 terms are added in registers and ONE shared-memory reduction serves 64 pixels instead of 32; the two
// per-pixel dependency chains also interleave in the instruction stream.
#ifndef IBGS_BWD_PPL2_CTAS
#define IBGS_BWD_PPL2_CTAS 5   // 96 registers; measured 4 (120 regs): +1.4 %, 6 (80 regs, spills): +6 %
#endif
template <bool GEO, int MAXE, int NSRC, int PPL>
__global__ void __launch_bounds__(256 / PPL, PPL == 1 ? 3 : IBGS_BWD_PPL2_CTAS) render_backward_pairs_kernel(const BwdArgs a) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NW = 8 / PPL;  // warps per CTA (= per tile)
  // dynamic shared memory: record double buffers | reduction tiles   (no CTA barrier anywhere in this kernel)
  extern __shared__ float4 s_dyn[];
  float4(*s_rec)[2][4][32] = reinterpret_cast<float4(*)[2][4][32]>(s_dyn);
  float* s_red_all = reinterpret_cast<float*>(s_dyn + NW * 2 * 4 * 32);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int W = a.W, H = a.H;
  const int HW = H * W;
  const int sub_x0 = blockIdx.x * TILE + (warp & 1) * 8;
  const int sub_y0 = blockIdx.y * TILE + (warp >> 1) * 4 * PPL;
  const unsigned px = (unsigned)(sub_x0 + (lane & 7));
  const float pixfx = (float)px;
  unsigned py[PPL];
  uint32_t pix_id[PPL];
  float pixfy[PPL];
  bool inside[PPL];
#pragma unroll
  for (int q = 0; q < PPL; q++) {
    py[q] = (unsigned)(sub_y0 + (lane >> 3) + 4 * q);
    pix_id[q] = W * py[q] + px;
    pixfy[q] = (float)py[q];
    inside[q] = px < (unsigned)W && py[q] < (unsigned)H;
  }

  const uint2 range = a.ranges[blockIdx.y * gridDim.x + blockIdx.x];
  const int total = (int)(range.y - range.x);

  float T[PPL], bgT[PPL];
  uint32_t last_contributor[PPL], median_lo[PPL], median_hi[PPL];
  float accum_rec[PPL][3], accum_nrm[PPL][3], dL_dpixel[PPL][3], dL_dnormal[PPL][3];
  float rayy[PPL];
  int ent_n[PPL];
  uint32_t lane_max_contrib = 0;
#pragma unroll
  for (int q = 0; q < PPL; q++) {
    const float T_final = inside[q] ? a.final_T[pix_id[q]] : 0;
    T[q] = T_final;
    last_contributor[q] = inside[q] ? a.n_contrib[pix_id[q]] : 0;
    lane_max_contrib = max(lane_max_contrib, last_contributor[q]);
    // backward.cu:693 compares the unsigned contributor with (int) low-1 / high-1: same unsigned wrap here
    median_lo[q] = (GEO && inside[q]) ? (uint32_t)((int)a.low[pix_id[q]] - 1) : 1u;
    median_hi[q] = (GEO && inside[q]) ? (uint32_t)((int)a.high[pix_id[q]] - 1) : 0u;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      accum_rec[q][i] = 0.f;
      accum_nrm[q][i] = 0.f;
      dL_dpixel[q][i] = inside[q] ? a.dL_dpixels[i * HW + pix_id[q]] : 0.f;
      dL_dnormal[q][i] = (GEO && inside[q]) ? a.dL_dnormals[i * HW + pix_id[q]] : 0.f;
    }
    float bg_dot_dpixel = 0;  // backward.cu:779-781
#pragma unroll
    for (int i = 0; i < 3; i++) bg_dot_dpixel += a.bg[i] * dL_dpixel[q][i];
    bgT[q] = -T_final * bg_dot_dpixel;
    // the pair loop only needs the SIGN of the plane depth (recorded pairs are re-evaluated exactly, in the
    // reference's double/float mix, by phase B)
    rayy[q] = (pixfy[q] - 0.5f * (float)H) / a.fy;
    ent_n[q] = 0;
  }
  const float rayx = (pixfx - 0.5f * (float)W) / a.fx;
  const uint32_t warp_max_contrib = __reduce_max_sync(FULL, lane_max_contrib);

  float* arena_f = reinterpret_cast<float*>(a.arena);
  float4(*wrec)[4][32] = s_rec[warp];
  const uint32_t red_st = red_store_addr(s_red_all + warp * RED_WARP_FLOATS, lane);
  const uint32_t red_ld = red_load_addr(s_red_all + warp * RED_WARP_FLOATS, lane);

  // list position p (0 = back of the list) holds contributor index total-1-p; positions with
  // contributor >= warp_max_contrib cannot contribute to any pixel of this warp: start after them
  const int first_pos = total - (int)min((uint32_t)total, warp_max_contrib);
  const int c_begin = first_pos >> 5;
  const int nchunks = (total + 31) >> 5;
  const uint32_t* plist_back = a.point_list + range.y - 1;  // plist_back[-p]

  if (c_begin < nchunks) {
    const float wx0 = (float)sub_x0, wx1 = (float)(sub_x0 + 7);
    const float wy0 = (float)sub_y0, wy1 = (float)(sub_y0 + 4 * PPL - 1);
    uint32_t id_next = 0u;  // id of this lane's Gaussian in the step being fetched
    {
      const int p = (c_begin << 5) + lane;
      if (p < total) {
        id_next = plist_back[-p];
        const float4* r = a.rec + 4 * (size_t)id_next;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (k < 3 || GEO) cp_async16(&wrec[c_begin & 1][k][lane], r + k);
      }
      cp_async_commit();
    }
    uint32_t id_issue = 0u;
    {
      const int p = ((c_begin + 1) << 5) + lane;
      if (p < total) id_issue = plist_back[-p];
    }

    for (int c = c_begin; c < nchunks; c++) {
      const int buf = c & 1;
      const int c0 = c << 5;
      const uint32_t id_cur = id_next;
      if (c + 1 < nchunks) {
        id_next = id_issue;
        if (c0 + 32 + lane < total) {
          const float4* r = a.rec + 4 * (size_t)id_issue;
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (k < 3 || GEO) cp_async16(&wrec[buf ^ 1][k][lane], r + k);
        }
        const int p2 = c0 + 64 + lane;
        id_issue = (p2 < total) ? plist_back[-p2] : 0u;
      }
      cp_async_commit();
      cp_async_wait<1>();
      __syncwarp();

      const int j = c0 + lane;
      bool keep = false;
      if (j < total) {
        const uint32_t contributor_j = (uint32_t)(total - 1 - j);
        const float4 q0 = wrec[buf][0][lane];
        const float4 q1 = wrec[buf][1][lane];
        keep = subtile_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, wx0, wx1, wy0, wy1) && (contributor_j < warp_max_contrib);
      }
      unsigned m = __ballot_sync(FULL, keep);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t contributor = (uint32_t)(total - 1 - c0 - b);  // backward.cu:636
        const float4 g0 = wrec[buf][0][b];
        const float4 g1 = wrec[buf][1][b];
        const float dx = g0.x - pixfx;
        float dy[PPL], G[PPL], alpha[PPL];
        bool active[PPL];
        bool any_active = false;
#pragma unroll
        for (int q = 0; q < PPL; q++) {
          dy[q] = g0.y - pixfy[q];
          const float power = -0.5f * (g0.z * dx * dx + g1.x * dy[q] * dy[q]) - g0.w * dx * dy[q];
          // The reference evaluates exp() here and __expf in the forward (backward.cu:648, forward.cu:424).  The fast
          // one is used in both passes: same accept/reject decisions as our (and the reference's) forward, 2 ulp
          // from exp() -- five orders of magnitude below the 1e-3 gradient gate.
          G[q] = exp_power(power);
          alpha[q] = min(0.99f, g1.y * G[q]);
          active[q] = inside[q] && !(contributor >= last_contributor[q]) && !(power > 0.0f) &&
                      !(alpha[q] < 1.0f / 255.0f);
          any_active = any_active || active[q];
        }
        if (!__any_sync(FULL, any_active)) continue;
        const uint32_t gid = __shfl_sync(FULL, id_cur, b);

        float v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = 0.0f;

#pragma unroll
        for (int q = 0; q < PPL; q++) {
          if (active[q]) {
            // 1/(1-alpha) once, approximate reciprocal (2 ulp): gradients are compared at 1e-3 relative L2 and
            // summed in a different order than the reference anyway
            const float one_m_alpha = 1.f - alpha[q];
            const float rinv = rcp_approx(one_m_alpha);  // one_m_alpha in [0.01, 1]: no denormal rescue path needed
            T[q] = T[q] * rinv;
            const float dchannel_dcolor = alpha[q] * T[q];
            float dL_dalpha = 0.0f;
            const float4 g2 = wrec[buf][2][b];
            const float col[3] = {g2.x, g2.y, g2.z};
            // accum_rec holds the blend of everything BEHIND this pair; the reference folds the previous pair in
            // at the top of the next iteration (last_alpha / last_color, backward.cu:661-668) -- same values,
            // folded here at the bottom of the current one instead, which frees seven registers
            // (alpha*c + (1-alpha)*accum is evaluated as accum + alpha*(c - accum): the difference is needed anyway)
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
              const float diff = col[ch] - accum_rec[q][ch];
              const float dL_dchannel = dL_dpixel[q][ch];
              dL_dalpha += diff * dL_dchannel;
              v[8 + ch] += dchannel_dcolor * dL_dchannel;
              accum_rec[q][ch] += alpha[q] * diff;
            }
            bool deferred = false;
            if (GEO) {
              const float4 g3 = wrec[buf][3][b];
              const float nrm[3] = {g3.x, g3.y, g3.z};
#pragma unroll
              for (int ch = 0; ch < 3; ch++) {
                const float diff = nrm[ch] - accum_nrm[q][ch];
                const float dL_dchannel = dL_dnormal[q][ch];
                dL_dalpha += diff * dL_dchannel;
                v[12 + ch] += dchannel_dcolor * dL_dchannel;
                accum_nrm[q][ch] += alpha[q] * diff;
              }
              // median buffer, backward.cu:693-701: a pair in range with a valid plane depth is only RECORDED
              // here (Gaussian id, T, colour+normal part of dL/dalpha); phase B does the rest
              if (contributor >= median_lo[q] && contributor <= median_hi[q]) {
                // intersected_depth = -d / (n.ray + 1e-8) > 0  <=>  d and the denominator have opposite signs
                const float z_den = g3.x * rayx + g3.y * rayy[q] + g3.z + 1.0e-8f;
                const bool z_pos = (g2.w > 0.0f && z_den < 0.0f) || (g2.w < 0.0f && z_den > 0.0f);
                if (z_pos && ent_n[q] < MAXE) {
                  float* e = a.ent + ((size_t)ent_n[q] * HW + pix_id[q]);
                  e[0] = __uint_as_float(gid);
                  e[a.ent_stride] = T[q];
                  e[2 * a.ent_stride] = dL_dalpha;
                  ent_n[q]++;
                  deferred = true;
                }
              }
            }
            if (!deferred) {
              dL_dalpha *= T[q];
              dL_dalpha += bgT[q] * rinv;
              const float dL_dG = g1.y * dL_dalpha;
              const float gdx = G[q] * dx;
              const float gdy = G[q] * dy[q];
              const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
              const float dG_ddely = -gdy * g1.x - gdx * g0.w;
              const float t0 = dL_dG * dG_ddelx, t1 = dL_dG * dG_ddely;
              v[0] += t0;
              v[1] += t1;
              v[2] += fabs(t0);
              v[3] += fabs(t1);
              const float hx = dL_dG * gdx, hy = dL_dG * gdy;
              v[4] += hx * dx;
              v[5] += hx * dy[q];
              v[6] += hy * dy[q];
              v[7] += G[q] * dL_dalpha;
            }
          }
        }

        const float total_v = warp_reduce16_smem(v, red_st, red_ld);
        if (lane < 16) {
          if (GEO ? (lane != 11 && lane != 15) : (lane < 11)) atomicAdd(arena_f + 16 * (size_t)gid + lane, total_v);
        }
      }
      __syncwarp();  // all lanes are done with buf before the step after next overwrites it
    }
    cp_async_wait<0>();
  }
  // ---- phase B: this lane's recorded median-buffer pairs (its own writes above: no fence needed) ----
  if (GEO) {
#pragma unroll 1
    for (int q = 0; q < PPL; q++)
      if (ent_n[q]) median_pixel_backward<NSRC>(a, px, py[q], pix_id[q], ent_n[q]);
  }
}



// stage A of one survivor: everything that depends only on the Gaussian's record and the lane's two pixels
struct SurvA {
  int b;           // lane that holds the record in the step's buffer
  uint32_t gid;    // Gaussian id
  float dx;
  f2 dy, alpha, G; // alpha and G are zero for a pixel that does not accept the pair
  bool any;        // some pixel of the warp accepts the pair
  bool rec0, rec1; // accepted AND inside the pixel's median-buffer contributor range (backward.cu:693)
};
__device__ __forceinline__ void stage_a(SurvA& s, unsigned& m, const float4 (*rec)[32], uint32_t id_cur, int contrib_base,
                                        float pixfx, f2 npy, uint32_t last0, uint32_t last1, uint32_t mlo0,
                                        uint32_t mhi0, uint32_t mlo1, uint32_t mhi1) {
  const int b = m ? (__ffs(m) - 1) : 0;
  m &= m - 1;
  const uint32_t contributor = (uint32_t)(contrib_base - b);  // backward.cu:636
  const float4 g0 = rec[0][b];
  const float4 g1 = rec[1][b];
  const float dx = g0.x - pixfx;
  const f2 dy = add2(bc(g0.y), npy);
  // power = -0.5 (A dx^2 + C dy^2) - B dx dy   (forward.cu:421)
  const float adx2 = g0.z * dx * dx;
  const float nbdx = -(g0.w * dx);
  const f2 power = fma2(fma2(mul2(bc(g1.x), dy), dy, bc(adx2)), bc(-0.5f), mul2(bc(nbdx), dy));
  // same fast exponential as the forward (exp_power, common.cuh); the reference uses exp() here and __expf there
  // (backward.cu:648, forward.cu:424): 2 ulp apart, five orders of magnitude below the 1e-3 gradient gate
  const f2 pl = mul2(power, bc(1.4426950408889634f));
  f2 G = make_float2(ex2_approx(pl.x), ex2_approx(pl.y));
  f2 alpha = mul2(bc(g1.y), G);
  alpha.x = fminf(0.99f, alpha.x);
  alpha.y = fminf(0.99f, alpha.y);
  // (a pixel outside the image has last contributor 0 and never accepts)
  const bool act0 = (contributor < last0) && !(power.x > 0.0f) && !(alpha.x < 1.0f / 255.0f);
  const bool act1 = (contributor < last1) && !(power.y > 0.0f) && !(alpha.y < 1.0f / 255.0f);
  s.any = __any_sync(0xffffffffu, act0 || act1);
  s.gid = __shfl_sync(0xffffffffu, id_cur, b);
  s.b = b;
  s.dx = dx;
  s.dy = dy;
  s.alpha = make_float2(act0 ? alpha.x : 0.f, act1 ? alpha.y : 0.f);
  s.G = make_float2(act0 ? G.x : 0.f, act1 ? G.y : 0.f);
  s.rec0 = act0 && contributor >= mlo0 && contributor <= mhi0;
  s.rec1 = act1 && contributor >= mlo1 && contributor <= mhi1;
}

// ---------------------------------------------------------------------------------------------------------
// kernel A2: the pair loop with TWO pixels per lane in PACKED FP32
// ---------------------------------------------------------------------------------------------------------
// Four warps per tile, warp = 8x8 pixels, lane = the pixels (x, y) and (x, y + 4).  Everything per-pixel lives in
// float2 registers (.x = upper pixel, .y = lower pixel) and the blend / gradient arithmetic of the two pixels is ONE
// stream of FFMA2 / FMUL2 / FADD2 instructions.  There is no per-pixel branch: a pixel that does not take part in a
// pair (outside the image, behind its last contributor, power > 0, alpha < 1/255) runs the same instructions with
// alpha = G = 0, which leaves its state (T, accum) unchanged and adds exact zeros to every sum.  The two pixels'
// terms are added in registers, so one shared-memory transpose reduction serves 64 pixels.
template <bool GEO, int MAXE, int NSRC>
__global__ void __launch_bounds__(128, IBGS_BWD_PACKED_CTAS) render_backward_pairs2_kernel(const BwdArgs a) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NW = 4;
  extern __shared__ float4 s_dyn[];
  float4(*s_rec)[2][4][32] = reinterpret_cast<float4(*)[2][4][32]>(s_dyn);
  float* s_red_all = reinterpret_cast<float*>(s_dyn + NW * 2 * 4 * 32);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int W = a.W, H = a.H;
  const int HW = H * W;
  const int sub_x0 = blockIdx.x * TILE + (warp & 1) * 8;
  const int sub_y0 = blockIdx.y * TILE + (warp >> 1) * 8;
  const unsigned px = (unsigned)(sub_x0 + (lane & 7));
  const float pixfx = (float)px;
  const unsigned py0 = (unsigned)(sub_y0 + (lane >> 3)), py1 = py0 + 4;
  const uint32_t pix0 = W * py0 + px, pix1 = W * py1 + px;
  const bool in0 = px < (unsigned)W && py0 < (unsigned)H;
  const bool in1 = px < (unsigned)W && py1 < (unsigned)H;

  const uint2 range = a.ranges[blockIdx.y * gridDim.x + blockIdx.x];
  const int total = (int)(range.y - range.x);

  f2 T, bgT, arec[3], anrm[3], dLp[3], dLn[3];
  T.x = in0 ? a.final_T[pix0] : 0.f;
  T.y = in1 ? a.final_T[pix1] : 0.f;
  // a pixel outside the image has last contributor 0: no pair ever passes `contributor < last`
  const uint32_t last0 = in0 ? a.n_contrib[pix0] : 0u, last1 = in1 ? a.n_contrib[pix1] : 0u;
  // backward.cu:693 compares the unsigned contributor with (int) low-1 / high-1: same unsigned wrap here
  const uint32_t mlo0 = (GEO && in0) ? (uint32_t)((int)a.low[pix0] - 1) : 1u;
  const uint32_t mhi0 = (GEO && in0) ? (uint32_t)((int)a.high[pix0] - 1) : 0u;
  const uint32_t mlo1 = (GEO && in1) ? (uint32_t)((int)a.low[pix1] - 1) : 1u;
  const uint32_t mhi1 = (GEO && in1) ? (uint32_t)((int)a.high[pix1] - 1) : 0u;
  float bgd0 = 0.f, bgd1 = 0.f;  // backward.cu:779-781
#pragma unroll
  for (int i = 0; i < 3; i++) {
    arec[i] = bc(0.f);
    anrm[i] = bc(0.f);
    dLp[i].x = in0 ? a.dL_dpixels[i * HW + pix0] : 0.f;
    dLp[i].y = in1 ? a.dL_dpixels[i * HW + pix1] : 0.f;
    dLn[i].x = (GEO && in0) ? a.dL_dnormals[i * HW + pix0] : 0.f;
    dLn[i].y = (GEO && in1) ? a.dL_dnormals[i * HW + pix1] : 0.f;
    bgd0 += a.bg[i] * dLp[i].x;
    bgd1 += a.bg[i] * dLp[i].y;
  }
  bgT.x = -T.x * bgd0;
  bgT.y = -T.y * bgd1;
  // the pair loop only needs the SIGN of the plane depth (recorded pairs are re-evaluated exactly by phase B)
  const float rayx = (pixfx - 0.5f * (float)W) / a.fx;
  const f2 rayy = make_float2(((float)py0 - 0.5f * (float)H) / a.fy, ((float)py1 - 0.5f * (float)H) / a.fy);
  const f2 npy = make_float2(-(float)py0, -(float)py1);
  int ent_n0 = 0, ent_n1 = 0;
  const uint32_t warp_max_contrib = __reduce_max_sync(FULL, max(last0, last1));

  float* arena_f = reinterpret_cast<float*>(a.arena);
  float4(*wrec)[4][32] = s_rec[warp];
  const uint32_t red_st = red_store_addr(s_red_all + warp * RED_WARP_FLOATS, lane);
  const uint32_t red_ld = red_load_addr(s_red_all + warp * RED_WARP_FLOATS, lane);

  const int first_pos = total - (int)min((uint32_t)total, warp_max_contrib);
  const int c_begin = first_pos >> 5;
  const int nchunks = (total + 31) >> 5;
  const uint32_t* plist_back = a.point_list + range.y - 1;  // plist_back[-p]

  if (c_begin < nchunks) {
    const float wx0 = (float)sub_x0, wx1 = (float)(sub_x0 + 7);
    const float wy0 = (float)sub_y0, wy1 = (float)(sub_y0 + 7);
    uint32_t id_next = 0u;
    {
      const int p = (c_begin << 5) + lane;
      if (p < total) {
        id_next = plist_back[-p];
        const float4* r = a.rec + 4 * (size_t)id_next;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (k < 3 || GEO) cp_async16(&wrec[c_begin & 1][k][lane], r + k);
      }
      cp_async_commit();
    }
    uint32_t id_issue = 0u;
    {
      const int p = ((c_begin + 1) << 5) + lane;
      if (p < total) id_issue = plist_back[-p];
    }

    for (int c = c_begin; c < nchunks; c++) {
      const int buf = c & 1;
      const int c0 = c << 5;
      const uint32_t id_cur = id_next;
      if (c + 1 < nchunks) {
        id_next = id_issue;
        if (c0 + 32 + lane < total) {
          const float4* r = a.rec + 4 * (size_t)id_issue;
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (k < 3 || GEO) cp_async16(&wrec[buf ^ 1][k][lane], r + k);
        }
        const int p2 = c0 + 64 + lane;
        id_issue = (p2 < total) ? plist_back[-p2] : 0u;
      }
      cp_async_commit();
      cp_async_wait<1>();
      __syncwarp();

      const int j = c0 + lane;
      bool keep = false;
      if (j < total) {
        const uint32_t contributor_j = (uint32_t)(total - 1 - j);
        const float4 q0 = wrec[buf][0][lane];
        const float4 q1 = wrec[buf][1][lane];
        keep = subtile_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q1.z, wx0, wx1, wy0, wy1) && (contributor_j < warp_max_contrib);
      }
      unsigned m = __ballot_sync(FULL, keep);
      // ---- plain walk: stage A, stage B and the reduction of one survivor after the other ----------------------------
      while (m) {
        SurvA cur;
        stage_a(cur, m, wrec[buf], id_cur, total - 1 - c0, pixfx, npy, last0, last1, mlo0, mhi0, mlo1, mhi1);
        if (!cur.any) continue;
#if IBGS_ABL == 3   // ablation: no stage B, no reduction (results wrong on purpose)
        T = fma2(T, cur.alpha, cur.G);
        continue;
#endif
        // ---- stage B --------------------------------------------------------------------------------------------
        const int b = cur.b;
        const float4 g0 = wrec[buf][0][b];
        const float4 g1 = wrec[buf][1][b];
        const float dx = cur.dx;
        const f2 dy = cur.dy;
        f2 alpha = cur.alpha, G = cur.G;
        float v[16];
        v[11] = 0.f;
        v[15] = 0.f;
        // 1/(1-alpha), approximate reciprocal (2 ulp): gradients are compared at 1e-3 relative L2 and summed in a
        // different order than the reference anyway; 1-alpha is in [0.01, 1] (exactly 1 for a masked pixel)
        const f2 oma = fma2(alpha, bc(-1.f), bc(1.f));
        const f2 rinv = make_float2(rcp_approx(oma.x), rcp_approx(oma.y));
        T = mul2(T, rinv);
        const f2 dcd = mul2(alpha, T);   // dchannel_dcolor
        f2 dLda;
        {
          // accum holds the blend of everything BEHIND this pair (folded at the bottom of the iteration instead of the
          // reference's last_alpha / last_color at the top of the next one, backward.cu:661-668): same values
          const float4 g2 = wrec[buf][2][b];
          const float col[3] = {g2.x, g2.y, g2.z};
#pragma unroll
          for (int ch = 0; ch < 3; ch++) {
            const f2 diff = fma2(arec[ch], bc(-1.f), bc(col[ch]));
            dLda = ch == 0 ? mul2(diff, dLp[ch]) : fma2(diff, dLp[ch], dLda);
            v[8 + ch] = hsum(mul2(dcd, dLp[ch]));
            arec[ch] = fma2(alpha, diff, arec[ch]);
          }
          if (GEO) {
            const float4 g3 = wrec[buf][3][b];
            const float nrm[3] = {g3.x, g3.y, g3.z};
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
              const f2 diff = fma2(anrm[ch], bc(-1.f), bc(nrm[ch]));
              dLda = fma2(diff, dLn[ch], dLda);
              v[12 + ch] = hsum(mul2(dcd, dLn[ch]));
              anrm[ch] = fma2(alpha, diff, anrm[ch]);
            }
            // median buffer, backward.cu:693-701: a pair in range with a valid plane depth is only RECORDED here
            // (Gaussian id, T, colour+normal part of dL/dalpha); phase B does the rest
            if (cur.rec0 || cur.rec1) {
              // intersected_depth = -d / (n.ray + 1e-8) > 0  <=>  d and the denominator have opposite signs
              const f2 zden = add2(fma2(bc(g3.y), rayy, bc(g3.x * rayx)), bc(g3.z + 1.0e-8f));
              const bool zp0 = (g2.w > 0.0f && zden.x < 0.0f) || (g2.w < 0.0f && zden.x > 0.0f);
              const bool zp1 = (g2.w > 0.0f && zden.y < 0.0f) || (g2.w < 0.0f && zden.y > 0.0f);
              if (cur.rec0 && zp0 && ent_n0 < MAXE) {
                float* e = a.ent + ((size_t)ent_n0 * HW + pix0);
                e[0] = __uint_as_float(cur.gid);
                e[a.ent_stride] = T.x;
                e[2 * a.ent_stride] = dLda.x;
                ent_n0++;
                G.x = 0.f;   // a recorded pixel contributes its mean / conic / opacity terms in phase B, not here
              }
              if (cur.rec1 && zp1 && ent_n1 < MAXE) {
                float* e = a.ent + ((size_t)ent_n1 * HW + pix1);
                e[0] = __uint_as_float(cur.gid);
                e[a.ent_stride] = T.y;
                e[2 * a.ent_stride] = dLda.y;
                ent_n1++;
                G.y = 0.f;
              }
            }
          }
        }
        dLda = fma2(dLda, T, mul2(bgT, rinv));
        const f2 dLdG = mul2(bc(g1.y), dLda);
        const f2 gdx = mul2(G, bc(dx));
        const f2 gdy = mul2(G, dy);
        const f2 ddx = fma2(gdx, bc(-g0.z), mul2(gdy, bc(-g0.w)));   // dG/ddelx = -gdx A - gdy B
        const f2 ddy = fma2(gdy, bc(-g1.x), mul2(gdx, bc(-g0.w)));   // dG/ddely = -gdy C - gdx B
        const f2 t0 = mul2(dLdG, ddx), t1 = mul2(dLdG, ddy);
        v[0] = t0.x + t0.y;
        v[1] = t1.x + t1.y;
        v[2] = fabsf(t0.x) + fabsf(t0.y);
        v[3] = fabsf(t1.x) + fabsf(t1.y);
        const f2 hx = mul2(dLdG, gdx), hy = mul2(dLdG, gdy);
        v[4] = hsum(mul2(hx, bc(dx)));
        v[5] = hsum(mul2(hx, dy));
        v[6] = hsum(mul2(hy, dy));
        v[7] = hsum(mul2(G, dLda));
        if (!GEO) { v[12] = 0.f; v[13] = 0.f; v[14] = 0.f; }
#if IBGS_ABL == 2   // ablation: no reduction, no RED (results wrong on purpose)
        float acc_abl = 0.f;
#pragma unroll
        for (int k = 0; k < 16; k++) acc_abl += v[k];
        if (acc_abl == 123.456f) arena_f[lane] = acc_abl;
#else
        const float total_v = warp_reduce16_smem(v, red_st, red_ld);
        if (lane < 16 && (GEO ? (lane != 11 && lane != 15) : (lane < 11)))
          atomicAdd(arena_f + 16 * (size_t)cur.gid + lane, total_v);
#endif
      }
      __syncwarp();  // all lanes are done with buf before the step after next overwrites it
    }
    cp_async_wait<0>();
  }
  // ---- phase B: this lane's recorded median-buffer pairs (its own writes above: no fence needed) ----
  if (GEO && IBGS_ABL != 1) {
    if (ent_n0) median_pixel_backward<NSRC>(a, px, py0, pix0, ent_n0);
    if (ent_n1) median_pixel_backward<NSRC>(a, px, py1, pix1, ent_n1);
  }
}

// IBGS_BWD_PPL forces one variant (kernel experiments); by default launch_render_backward picks per view
#ifndef IBGS_BWD_PPL
#define IBGS_BWD_PPL 0
#endif
#ifndef IBGS_BWD_PACKED
#define IBGS_BWD_PACKED 1   // two-pixels-per-lane launches take the packed-FP32 kernel (0: the scalar PPL = 2 template)
#endif
constexpr int bwd_smem_bytes(int ppl) {
  return (8 / ppl) * (2 * 4 * 32 * 16 + RED_WARP_FLOATS * 4);
}
constexpr int bwd2_smem_bytes() { return bwd_smem_bytes(2); }
// average tile-list length from which the 2-pixels-per-lane variant is taken.  Measured at 1080p (forward+backward per
// view, one / two pixels per lane): uniform scene 25 instances per tile 1.35 / 1.45 ms, 126: 2.07 / 2.14, 504: 4.28 /
// 4.25; scene with a dense centre 158: 2.37 / 2.32, 315: 3.25 / 3.04, 630: 4.21 / 3.90, 1894: 5.14 / 4.73.  Short
// lists are dominated by the per-pixel set-up and phase B, which the 4-warp CTAs serialise.
constexpr long long BWD_DENSE_LIST = 256;
int g_bwd_variant = IBGS_BWD_PPL;   // 0 = choose per view, 1 / 2 = forced (ibgs_set_backward_variant)

inline int max_entries(int buffer_length) { return buffer_length <= 4 ? 5 : 9; }

}  // namespace

extern "C" int ibgs_set_backward_variant(int pixels_per_lane) {
  if (pixels_per_lane < 0 || pixels_per_lane > 2) {
    ibgs_set_error("pixels_per_lane must be 0 (choose per view), 1 or 2, got %d", pixels_per_lane);
    return IBGS_EINVAL;
  }
  g_bwd_variant = pixels_per_lane;
  return IBGS_OK;
}

size_t render_backward_scratch_bytes(size_t N, int buffer_length, int render_geo) {
  if (!render_geo) return 0;
  const size_t E = (size_t)max_entries(buffer_length);
  return 3 * align_up(E * N * 4, 256);
}

int launch_render_backward(const IbgsBackwardArgs& f, const GeomState& g, const ImageState& im,
                           const BinningState& b, TexPair tex, float focal_x, float focal_y, dim3 grid,
                           float4* arena, void* ent_scratch, cudaStream_t s) {
  BwdArgs a;
  a.ranges = im.ranges;
  a.point_list = b.point_list;
  a.rec = g.rec;
  a.W = f.view.image_width;
  a.H = f.view.image_height;
  a.fx = focal_x;
  a.fy = focal_y;
  a.bg = f.view.bg;
  a.ref_to_src_list = f.view.ref_to_src_list;
  a.texColor = tex.color;
  a.nb_src = f.view.nb_src_images;
  a.depth_pixels = f.out_median_intersected_depth;
  a.warped_pixels = f.out_warped_image;
  a.final_T = im.final_T;
  a.n_contrib = im.n_contrib;
  a.sum_w = im.sum_w;
  a.low = im.low;
  a.high = im.high;
  a.valid_idx = im.valid_idx;
  a.valid_w = im.valid_w;
  a.dL_dpixels = f.dL_dout_color;
  a.dL_dnormals = f.dL_dout_normal_map;
  a.dL_ddepths = f.dL_dout_median_intersected_depth;
  a.dL_dwarped = f.dL_dout_warped_image;
  a.arena = arena;
  a.ent = nullptr;
  a.ent_stride = 0;
  const size_t N = (size_t)a.W * a.H;
  const int E = max_entries(f.view.buffer_length);
  if (f.view.render_geo) {
    if (!ent_scratch) { ibgs_set_error("render_geo backward needs its entry scratch"); return IBGS_EINVAL; }
    char* p = (char*)ent_scratch;
    a.ent = (float*)p;
    a.ent_stride = align_up((size_t)E * N * 4, 256) / 4;
  }
  ProfScope prof(PROF_RENDER_BWD, s);
  // > 48 KB of shared memory needs the opt-in attribute (per device; cheap enough to set per launch)
#define LAUNCH_PAIRS_PPL(PPL, ...)                                                                             \
  do {                                                                                                         \
    CUDA_TRY(cudaFuncSetAttribute(render_backward_pairs_kernel<__VA_ARGS__, PPL>,                              \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem_bytes(PPL)));          \
    render_backward_pairs_kernel<__VA_ARGS__, PPL><<<grid, 256 / PPL, bwd_smem_bytes(PPL), s>>>(a);            \
  } while (0)
#define LAUNCH_PAIRS2(...)                                                                                     \
  do {                                                                                                         \
    CUDA_TRY(cudaFuncSetAttribute(render_backward_pairs2_kernel<__VA_ARGS__>,                                  \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, bwd2_smem_bytes()));            \
    render_backward_pairs2_kernel<__VA_ARGS__><<<grid, 128, bwd2_smem_bytes(), s>>>(a);                        \
  } while (0)
#if IBGS_BWD_PACKED
#define LAUNCH_PAIRS(...)                                                                                      \
  do {                                                                                                         \
    if (two_per_lane) LAUNCH_PAIRS2(__VA_ARGS__); else LAUNCH_PAIRS_PPL(1, __VA_ARGS__);                       \
  } while (0)
#else
#define LAUNCH_PAIRS(...)                                                                                      \
  do {                                                                                                         \
    if (two_per_lane) LAUNCH_PAIRS_PPL(2, __VA_ARGS__); else LAUNCH_PAIRS_PPL(1, __VA_ARGS__);                 \
  } while (0)
#endif
  const long long tiles = (long long)grid.x * grid.y;
  const bool two_per_lane = g_bwd_variant ? (g_bwd_variant == 2) : ((long long)f.R > BWD_DENSE_LIST * tiles);
  if (f.view.render_geo) {
    const bool few = f.view.nb_src_images <= 4;  // the callers' default is 4 source views (arguments/__init__.py:127)
    if (f.view.buffer_length <= 4) {
      if (few) LAUNCH_PAIRS(true, 5, 4); else LAUNCH_PAIRS(true, 5, MAX_SRC);
    } else {
      if (few) LAUNCH_PAIRS(true, 9, 4); else LAUNCH_PAIRS(true, 9, MAX_SRC);
    }
  } else {
    LAUNCH_PAIRS(false, 1, 1);
  }
#undef LAUNCH_PAIRS_PPL
#undef LAUNCH_PAIRS2
#undef LAUNCH_PAIRS
  KERNEL_CHECK(f.view.debug, s);
  return IBGS_OK;
}
