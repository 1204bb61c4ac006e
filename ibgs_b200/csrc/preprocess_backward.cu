// preprocess_backward.cu -- per-Gaussian backward: conic -> cov2D -> cov3D -> (scale, rotation), 2D mean ->
// 3D mean, colour -> SH (+ view-direction term into the mean).
//
// Reference behaviour: computeCov2DCUDA (cuda_rasterizer/backward.cu:241-371) followed by
// BACKWARD::preprocessCUDA<3> (:443-493) with computeColorFromSH backward (:116-235), computeCov3D
// backward (:375-438) and dnormvdv (auxiliary.h:111-121).
//
// Here both reference kernels are one launch that reads the [P][16] accumulation arena written by the
// backward renderer and writes EVERY output row (zeros for culled Gaussians), so the caller needs no
// zero-filled gradient tensors (the reference binding memsets eleven of them, rasterize_points.cu:209-219)
// and dL_dcov3D never round-trips through memory unless cov3D was precomputed.
//
// Memory access: one thread per Gaussian, but the two wide per-Gaussian rows -- the SH coefficients (read) and
// their gradient (written), 12*M bytes each, i.e. 2 x 108 B of the ~480 B a Gaussian moves at M = 9 -- go
// through a per-warp shared-memory tile: the warp copies its 32 rows with fully coalesced 128-byte accesses and
// every lane then works on its own row at an odd word stride (bank-conflict free).  Culled Gaussians' SH rows
// are not read at all; their gradient rows are written as zeros by the same coalesced copy.
#include "common.cuh"

namespace {

#define PB_THREADS 128
#define PB_WARPS (PB_THREADS / 32)
// 8 CTAs of 128 threads per SM = 64 registers (a few spilled words): the kernel is HBM-latency bound and wants the
// warps -- measured in accumulate mode 0.54 ms at 96 registers, 0.41 at 80, 0.37 at 64, 0.45 at 48
#ifndef PB_MIN_CTAS
#define PB_MIN_CTAS 8
#endif

struct PBArgs {
  int P, D, M;
  const float* means3D;
  const int* radii;
  const float* shs;
  const float* shs_rest;
  const uint8_t* clamped;
  const float* scales;
  const float* rotations;
  float scale_modifier;
  const float* cov3D_precomp;
  const float* view;
  const float* proj;
  float h_x, h_y, tan_fovx, tan_fovy;
  float half_w, half_h;
  uint32_t magic;  // floor(2^32 / (3M)) + 1: word index -> row by multiply-high
  uint32_t magic_dc, magic_rest;  // the same for rows of 3 and 3(M-1) words (split SH tensors)
  int vec_ok;      // the SH tensors and their gradients are 16-byte aligned
  const float* campos;
  const float4* arena;
  float* dL_dmeans3D;
  float* dL_dmeans2D;
  float* dL_dmeans2D_abs;
  float* dL_dcolors;
  float* dL_dopacity;
  float* dL_dcov3D;
  float* dL_dsh;
  float* dL_dsh_rest;
  float* dL_dscales;
  float* dL_drotations;
  float* dL_dall_map;
  unsigned acc;   // IBGS_ACC_* bits: outputs that are accumulated into (+=) instead of written
};

// gradient output word: plain store, or += when the caller asked for accumulation into this output
// (IbgsBackwardArgs.accumulate_mask: gradient accumulation over a view batch without a separate add pass)
template <bool ACC>
__device__ __forceinline__ void put(float* p, float v, unsigned acc, unsigned bit) {
  if (ACC && (acc & bit)) *p += v; else *p = v;
}

// reference auxiliary.h:111-121
__forceinline__ __device__ float3 dnormvdv(float3 v, float3 dv) {
  float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
  float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
  float3 r;
  r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
  r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
  r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
  return r;
}

template <bool ACC>
__forceinline__ __device__ void write_zero_row(const PBArgs& a, int idx) {
  const unsigned acc = ACC ? a.acc : 0u;   // accumulated outputs keep their value for a culled Gaussian
  float* p;
  if (!(acc & IBGS_ACC_MEANS3D)) { p = a.dL_dmeans3D + 3 * (size_t)idx; p[0] = p[1] = p[2] = 0.f; }
  if (!(acc & IBGS_ACC_MEANS2D)) { p = a.dL_dmeans2D + 3 * (size_t)idx; p[0] = p[1] = p[2] = 0.f; }
  if (!(acc & IBGS_ACC_MEANS2D_ABS)) { p = a.dL_dmeans2D_abs + 3 * (size_t)idx; p[0] = p[1] = p[2] = 0.f; }
  p = a.dL_dcolors + 3 * (size_t)idx; p[0] = p[1] = p[2] = 0.f;
  if (!(acc & IBGS_ACC_OPACITY)) a.dL_dopacity[idx] = 0.f;
  if (a.dL_dcov3D) {
    p = a.dL_dcov3D + 6 * (size_t)idx;
    for (int i = 0; i < 6; i++) p[i] = 0.f;
  }
  if (!(acc & IBGS_ACC_SCALES)) { p = a.dL_dscales + 3 * (size_t)idx; p[0] = p[1] = p[2] = 0.f; }
  if (!(acc & IBGS_ACC_ROTATIONS)) reinterpret_cast<float4*>(a.dL_drotations)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!(acc & IBGS_ACC_ALL_MAP)) {
    p = a.dL_dall_map + 5 * (size_t)idx;
    for (int i = 0; i < 5; i++) p[i] = 0.f;
  }
}

// warp-cooperative copy between a contiguous block of 32 global rows of `Lsrc` words (16-byte aligned because it
// starts at a multiple of 32 rows) and columns [col0, col0 + Lsrc) of the warp's shared-memory tile (row stride Ls
// words).  The block is moved as float4 (fully coalesced 512-byte requests); word i of the block belongs to row
// i / Lsrc, computed with a multiply-high (`magic` = floor(2^32 / Lsrc) + 1, exact for i < 2^16).
template <int STORE>   // 0: global -> tile, 1: tile -> global, 2: global += tile
__device__ __forceinline__ void tile_copy(float* tile, float* gblock, int nwords, int Lsrc, int Ls, int col0,
                                          uint32_t magic, bool vec_ok, int lane) {
  const int pad = Ls - Lsrc;
  tile += col0;
  // nwords is a multiple of 4 for a full block of 32 rows; the tail (and everything, if the caller's tensors are
  // not 16-byte aligned) goes through the scalar loop below
  const int nvec = vec_ok ? (nwords >> 2) : 0;
  float4* g4 = reinterpret_cast<float4*>(gblock);
  for (int i4 = lane; i4 < nvec; i4 += 32) {
    const int i = i4 << 2;
    int t[4];
#pragma unroll
    for (int j = 0; j < 4; j++) t[j] = i + j + pad * (int)__umulhi((uint32_t)(i + j), magic);
    if (STORE == 2) {
      const float4 o = g4[i4];
      g4[i4] = make_float4(o.x + tile[t[0]], o.y + tile[t[1]], o.z + tile[t[2]], o.w + tile[t[3]]);
    } else if (STORE == 1) {
      g4[i4] = make_float4(tile[t[0]], tile[t[1]], tile[t[2]], tile[t[3]]);
    } else {
      const float4 v = g4[i4];
      tile[t[0]] = v.x; tile[t[1]] = v.y; tile[t[2]] = v.z; tile[t[3]] = v.w;
    }
  }
  for (int i = (nvec << 2) + lane; i < nwords; i += 32) {  // partial last block (P not a multiple of 32)
    const int t = i + pad * (int)__umulhi((uint32_t)i, magic);
    if (STORE == 2) gblock[i] += tile[t]; else if (STORE == 1) gblock[i] = tile[t]; else tile[t] = gblock[i];
  }
}

// everything for one visible Gaussian; `shrow` is its SH row in the warp's shared-memory tile: coefficients on
// entry, their gradient on exit (every word of the row is overwritten)
template <bool ACC>
__device__ __forceinline__ void preprocess_backward_one(const PBArgs& a, int idx, float* shrow) {
  const unsigned acc = ACC ? a.acc : 0u;
  // slots 0-6 arrive unscaled from the tile renderer (render_backward.cu): apply 0.5*W, 0.5*H
  // (backward.cu:606-607) and the -0.5 of the conic terms (:799-801) once per Gaussian
  float4 a0 = a.arena[4 * (size_t)idx + 0];
  float4 a1 = a.arena[4 * (size_t)idx + 1];
  a0.x *= a.half_w; a0.y *= a.half_h; a0.z *= a.half_w; a0.w *= a.half_h;
  a1.x *= -0.5f; a1.y *= -0.5f; a1.z *= -0.5f;
  const float4 a2 = a.arena[4 * (size_t)idx + 2];
  const float4 a3 = a.arena[4 * (size_t)idx + 3];

  // straight copies of what the renderer accumulated
  {
    float* p = a.dL_dmeans2D + 3 * (size_t)idx;
    put<ACC>(p, a0.x, acc, IBGS_ACC_MEANS2D); put<ACC>(p + 1, a0.y, acc, IBGS_ACC_MEANS2D); put<ACC>(p + 2, 0.f, acc, IBGS_ACC_MEANS2D);
    p = a.dL_dmeans2D_abs + 3 * (size_t)idx;
    put<ACC>(p, a0.z, acc, IBGS_ACC_MEANS2D_ABS); put<ACC>(p + 1, a0.w, acc, IBGS_ACC_MEANS2D_ABS); put<ACC>(p + 2, 0.f, acc, IBGS_ACC_MEANS2D_ABS);
    p = a.dL_dcolors + 3 * (size_t)idx; p[0] = a2.x; p[1] = a2.y; p[2] = a2.z;
    put<ACC>(a.dL_dopacity + idx, a1.w, acc, IBGS_ACC_OPACITY);
    p = a.dL_dall_map + 5 * (size_t)idx;
    put<ACC>(p, a3.x, acc, IBGS_ACC_ALL_MAP); put<ACC>(p + 1, a3.y, acc, IBGS_ACC_ALL_MAP); put<ACC>(p + 2, a3.z, acc, IBGS_ACC_ALL_MAP);
    put<ACC>(p + 3, 0.f, acc, IBGS_ACC_ALL_MAP); put<ACC>(p + 4, a2.w, acc, IBGS_ACC_ALL_MAP);
  }

  const float3 mean = {a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2]};
  const float mod = a.scale_modifier;

  // forward recomputation of cov3D (forward.cu:156-190) unless it was given
  float cov3D[6];
  float4 rot = make_float4(0.f, 0.f, 0.f, 0.f);
  float3 scl = {0.f, 0.f, 0.f};
  M3 R, Mm;
  if (a.cov3D_precomp != nullptr) {
    for (int i = 0; i < 6; i++) cov3D[i] = a.cov3D_precomp[6 * (size_t)idx + i];
  } else {
    rot = reinterpret_cast<const float4*>(a.rotations)[idx];
    scl = {a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2]};
    const float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
    R = m3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
           2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
           2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    M3 S = m3(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
    S.m[0][0] = mod * scl.x;
    S.m[1][1] = mod * scl.y;
    S.m[2][2] = mod * scl.z;
    Mm = m3_mul(S, R);
    M3 Sigma = m3_mul(m3_t(Mm), Mm);
    cov3D[0] = Sigma.m[0][0]; cov3D[1] = Sigma.m[0][1]; cov3D[2] = Sigma.m[0][2];
    cov3D[3] = Sigma.m[1][1]; cov3D[4] = Sigma.m[1][2]; cov3D[5] = Sigma.m[2][2];
  }

  // ---- computeCov2DCUDA, backward.cu:241-371 ----
  const float3 dL_dconic = {a1.x, a1.y, a1.z};  // (x, y, w) of the 2x2, backward.cu:262
  float3 t = transformPoint4x3(mean, a.view);
  const float limx = 1.3f * a.tan_fovx;
  const float limy = 1.3f * a.tan_fovy;
  const float txtz = t.x / t.z;
  const float tytz = t.y / t.z;
  t.x = min(limx, max(-limx, txtz)) * t.z;
  t.y = min(limy, max(-limy, tytz)) * t.z;
  const float x_grad_mul = txtz < -limx || txtz > limx ? 0 : 1;
  const float y_grad_mul = tytz < -limy || tytz > limy ? 0 : 1;
  const float h_x = a.h_x, h_y = a.h_y;

  M3 J = m3(h_x / t.z, 0.0f, -(h_x * t.x) / (t.z * t.z), 0.0f, h_y / t.z, -(h_y * t.y) / (t.z * t.z), 0, 0, 0);
  M3 Wm = m3(a.view[0], a.view[4], a.view[8], a.view[1], a.view[5], a.view[9], a.view[2], a.view[6], a.view[10]);
  M3 Vrk = m3(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
  M3 T = m3_mul(Wm, J);
  M3 cov2D = m3_mul(m3_mul(m3_t(T), m3_t(Vrk)), T);

  const float ca = cov2D.m[0][0] += 0.3f;
  const float cb = cov2D.m[0][1];
  const float cc = cov2D.m[1][1] += 0.3f;
  const float denom = ca * cc - cb * cb;
  float dL_da = 0, dL_db = 0, dL_dc = 0;
  const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
  float dL_dcov[6];
  if (denom2inv != 0) {
    dL_da = denom2inv * (-cc * cc * dL_dconic.x + 2 * cb * cc * dL_dconic.y + (denom - ca * cc) * dL_dconic.z);
    dL_dc = denom2inv * (-ca * ca * dL_dconic.z + 2 * ca * cb * dL_dconic.y + (denom - ca * cc) * dL_dconic.x);
    dL_db = denom2inv * 2 * (cb * cc * dL_dconic.x - (denom + 2 * cb * cb) * dL_dconic.y + ca * cb * dL_dconic.z);
    dL_dcov[0] = (T.m[0][0] * T.m[0][0] * dL_da + T.m[0][0] * T.m[1][0] * dL_db + T.m[1][0] * T.m[1][0] * dL_dc);
    dL_dcov[3] = (T.m[0][1] * T.m[0][1] * dL_da + T.m[0][1] * T.m[1][1] * dL_db + T.m[1][1] * T.m[1][1] * dL_dc);
    dL_dcov[5] = (T.m[0][2] * T.m[0][2] * dL_da + T.m[0][2] * T.m[1][2] * dL_db + T.m[1][2] * T.m[1][2] * dL_dc);
    dL_dcov[1] = 2 * T.m[0][0] * T.m[0][1] * dL_da + (T.m[0][0] * T.m[1][1] + T.m[0][1] * T.m[1][0]) * dL_db +
                 2 * T.m[1][0] * T.m[1][1] * dL_dc;
    dL_dcov[2] = 2 * T.m[0][0] * T.m[0][2] * dL_da + (T.m[0][0] * T.m[1][2] + T.m[0][2] * T.m[1][0]) * dL_db +
                 2 * T.m[1][0] * T.m[1][2] * dL_dc;
    dL_dcov[4] = 2 * T.m[0][2] * T.m[0][1] * dL_da + (T.m[0][1] * T.m[1][2] + T.m[0][2] * T.m[1][1]) * dL_db +
                 2 * T.m[1][1] * T.m[1][2] * dL_dc;
  } else {
    for (int i = 0; i < 6; i++) dL_dcov[i] = 0;
  }
  if (a.dL_dcov3D) {
    float* p = a.dL_dcov3D + 6 * (size_t)idx;
    for (int i = 0; i < 6; i++) p[i] = dL_dcov[i];
  }

  const float dL_dT00 = 2 * (T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_da +
                        (T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_db;
  const float dL_dT01 = 2 * (T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_da +
                        (T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_db;
  const float dL_dT02 = 2 * (T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_da +
                        (T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_db;
  const float dL_dT10 = 2 * (T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_dc +
                        (T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_db;
  const float dL_dT11 = 2 * (T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_dc +
                        (T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_db;
  const float dL_dT12 = 2 * (T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_dc +
                        (T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_db;

  const float dL_dJ00 = Wm.m[0][0] * dL_dT00 + Wm.m[0][1] * dL_dT01 + Wm.m[0][2] * dL_dT02;
  const float dL_dJ02 = Wm.m[2][0] * dL_dT00 + Wm.m[2][1] * dL_dT01 + Wm.m[2][2] * dL_dT02;
  const float dL_dJ11 = Wm.m[1][0] * dL_dT10 + Wm.m[1][1] * dL_dT11 + Wm.m[1][2] * dL_dT12;
  const float dL_dJ12 = Wm.m[2][0] * dL_dT10 + Wm.m[2][1] * dL_dT11 + Wm.m[2][2] * dL_dT12;

  const float tz = 1.f / t.z;
  const float tz2 = tz * tz;
  const float tz3 = tz2 * tz;
  const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
  const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
  const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
                       (2 * h_y * t.y) * tz3 * dL_dJ12;
  // transformVec4x3Transpose, auxiliary.h:93-101
  float3 dL_dmean_acc = {a.view[0] * dL_dtx + a.view[1] * dL_dty + a.view[2] * dL_dtz,
                         a.view[4] * dL_dtx + a.view[5] * dL_dty + a.view[6] * dL_dtz,
                         a.view[8] * dL_dtx + a.view[9] * dL_dty + a.view[10] * dL_dtz};

  // ---- BACKWARD::preprocessCUDA, backward.cu:467-484 ----
  const float* proj = a.proj;
  const float3 m = mean;
  float4 m_hom = transformPoint4x4(m, proj);
  float m_w = 1.0f / (m_hom.w + 0.0000001f);
  const float dmx = a0.x, dmy = a0.y;
  float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
  float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
  float3 dL_dmean;
  dL_dmean.x = (proj[0] * m_w - proj[3] * mul1) * dmx + (proj[1] * m_w - proj[3] * mul2) * dmy;
  dL_dmean.y = (proj[4] * m_w - proj[7] * mul1) * dmx + (proj[5] * m_w - proj[7] * mul2) * dmy;
  dL_dmean.z = (proj[8] * m_w - proj[11] * mul1) * dmx + (proj[9] * m_w - proj[11] * mul2) * dmy;
  dL_dmean_acc.x += dL_dmean.x;
  dL_dmean_acc.y += dL_dmean.y;
  dL_dmean_acc.z += dL_dmean.z;

  // ---- SH backward, backward.cu:116-235 ----
  if (a.shs) {
    const float3 campos = {a.campos[0], a.campos[1], a.campos[2]};
    const float3 dir_orig = {m.x - campos.x, m.y - campos.y, m.z - campos.z};
    const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
    // The gradient row overwrites the coefficient row in place (same shared-memory words), so each degree band
    // first pulls its coefficients into registers, then writes its gradients.
    float* dL_dsh = shrow;
    const uint8_t cl = a.clamped[idx];
    float dRGB[3] = {a2.x, a2.y, a2.z};
#pragma unroll
    for (int c = 0; c < 3; c++) dRGB[c] *= ((cl >> c) & 1) ? 0 : 1;
    float dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
    const int deg = a.D;
    int written = 1;
#pragma unroll
    for (int c = 0; c < 3; c++) dL_dsh[c] = SH_C0 * dRGB[c];
    if (deg > 0) {
      float sh[12];
#pragma unroll
      for (int k = 3; k < 12; k++) sh[k] = shrow[k];
      const float d1 = -SH_C1 * y, d2 = SH_C1 * z, d3 = -SH_C1 * x;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        dL_dsh[3 + c] = d1 * dRGB[c];
        dL_dsh[6 + c] = d2 * dRGB[c];
        dL_dsh[9 + c] = d3 * dRGB[c];
        dRGBdx[c] = -SH_C1 * sh[9 + c];
        dRGBdy[c] = -SH_C1 * sh[3 + c];
        dRGBdz[c] = SH_C1 * sh[6 + c];
      }
      written = 4;
      if (deg > 1) {
        float sh2[27];
#pragma unroll
        for (int k = 12; k < 27; k++) sh2[k] = shrow[k];
        const float xx = x * x, yy = y * y, zz = z * z;
        const float xy = x * y, yz = y * z, xz = x * z;
        const float d4 = SH_C2[0] * xy, d5 = SH_C2[1] * yz, d6 = SH_C2[2] * (2.f * zz - xx - yy),
                    d7 = SH_C2[3] * xz, d8 = SH_C2[4] * (xx - yy);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          dL_dsh[12 + c] = d4 * dRGB[c];
          dL_dsh[15 + c] = d5 * dRGB[c];
          dL_dsh[18 + c] = d6 * dRGB[c];
          dL_dsh[21 + c] = d7 * dRGB[c];
          dL_dsh[24 + c] = d8 * dRGB[c];
          dRGBdx[c] += SH_C2[0] * y * sh2[12 + c] + SH_C2[2] * 2.f * -x * sh2[18 + c] + SH_C2[3] * z * sh2[21 + c] +
                       SH_C2[4] * 2.f * x * sh2[24 + c];
          dRGBdy[c] += SH_C2[0] * x * sh2[12 + c] + SH_C2[1] * z * sh2[15 + c] + SH_C2[2] * 2.f * -y * sh2[18 + c] +
                       SH_C2[4] * 2.f * -y * sh2[24 + c];
          dRGBdz[c] += SH_C2[1] * y * sh2[15 + c] + SH_C2[2] * 2.f * 2.f * z * sh2[18 + c] + SH_C2[3] * x * sh2[21 + c];
        }
        written = 9;
        if (deg > 2) {
          float sh3[48];
#pragma unroll
          for (int k = 27; k < 48; k++) sh3[k] = shrow[k];
          const float d9 = SH_C3[0] * y * (3.f * xx - yy), d10 = SH_C3[1] * xy * z,
                      d11 = SH_C3[2] * y * (4.f * zz - xx - yy),
                      d12 = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy),
                      d13 = SH_C3[4] * x * (4.f * zz - xx - yy), d14 = SH_C3[5] * z * (xx - yy),
                      d15 = SH_C3[6] * x * (xx - 3.f * yy);
#pragma unroll
          for (int c = 0; c < 3; c++) {
            dL_dsh[27 + c] = d9 * dRGB[c];
            dL_dsh[30 + c] = d10 * dRGB[c];
            dL_dsh[33 + c] = d11 * dRGB[c];
            dL_dsh[36 + c] = d12 * dRGB[c];
            dL_dsh[39 + c] = d13 * dRGB[c];
            dL_dsh[42 + c] = d14 * dRGB[c];
            dL_dsh[45 + c] = d15 * dRGB[c];
            dRGBdx[c] += (SH_C3[0] * sh3[27 + c] * 3.f * 2.f * xy + SH_C3[1] * sh3[30 + c] * yz +
                          SH_C3[2] * sh3[33 + c] * -2.f * xy + SH_C3[3] * sh3[36 + c] * -3.f * 2.f * xz +
                          SH_C3[4] * sh3[39 + c] * (-3.f * xx + 4.f * zz - yy) + SH_C3[5] * sh3[42 + c] * 2.f * xz +
                          SH_C3[6] * sh3[45 + c] * 3.f * (xx - yy));
            dRGBdy[c] += (SH_C3[0] * sh3[27 + c] * 3.f * (xx - yy) + SH_C3[1] * sh3[30 + c] * xz +
                          SH_C3[2] * sh3[33 + c] * (-3.f * yy + 4.f * zz - xx) +
                          SH_C3[3] * sh3[36 + c] * -3.f * 2.f * yz + SH_C3[4] * sh3[39 + c] * -2.f * xy +
                          SH_C3[5] * sh3[42 + c] * -2.f * yz + SH_C3[6] * sh3[45 + c] * -3.f * 2.f * xy);
            dRGBdz[c] += (SH_C3[1] * sh3[30 + c] * xy + SH_C3[2] * sh3[33 + c] * 4.f * 2.f * yz +
                          SH_C3[3] * sh3[36 + c] * 3.f * (2.f * zz - xx - yy) +
                          SH_C3[4] * sh3[39 + c] * 4.f * 2.f * xz + SH_C3[5] * sh3[42 + c] * (xx - yy));
          }
          written = 16;
        }
      }
    }
    // coefficients above the active degree get no gradient (reference leaves its zero fill)
    for (int k = written * 3; k < a.M * 3; k++) dL_dsh[k] = 0.f;

    const float3 dL_ddir = {dRGBdx[0] * dRGB[0] + dRGBdx[1] * dRGB[1] + dRGBdx[2] * dRGB[2],
                            dRGBdy[0] * dRGB[0] + dRGBdy[1] * dRGB[1] + dRGBdy[2] * dRGB[2],
                            dRGBdz[0] * dRGB[0] + dRGBdz[1] * dRGB[1] + dRGBdz[2] * dRGB[2]};
    const float3 dm = dnormvdv(dir_orig, dL_ddir);
    dL_dmean_acc.x += dm.x;
    dL_dmean_acc.y += dm.y;
    dL_dmean_acc.z += dm.z;
  }
  {
    float* p = a.dL_dmeans3D + 3 * (size_t)idx;
    put<ACC>(p, dL_dmean_acc.x, acc, IBGS_ACC_MEANS3D); put<ACC>(p + 1, dL_dmean_acc.y, acc, IBGS_ACC_MEANS3D);
    put<ACC>(p + 2, dL_dmean_acc.z, acc, IBGS_ACC_MEANS3D);
  }

  // ---- cov3D -> scale / rotation, backward.cu:375-438 ----
  if (a.scales) {
    const float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
    const float3 s = {mod * scl.x, mod * scl.y, mod * scl.z};
    M3 dL_dSigma = m3(dL_dcov[0], 0.5f * dL_dcov[1], 0.5f * dL_dcov[2],
                      0.5f * dL_dcov[1], dL_dcov[3], 0.5f * dL_dcov[4],
                      0.5f * dL_dcov[2], 0.5f * dL_dcov[4], dL_dcov[5]);
    // dL_dM = 2.0f * M * dL_dSigma (scalar*matrix first, glm operator order)
    M3 M2;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int w = 0; w < 3; w++) M2.m[c][w] = 2.0f * Mm.m[c][w];
    M3 dL_dM = m3_mul(M2, dL_dSigma);
    M3 Rt = m3_t(R);
    M3 dL_dMt = m3_t(dL_dM);
    float3 dscale;
    dscale.x = Rt.m[0][0] * dL_dMt.m[0][0] + Rt.m[0][1] * dL_dMt.m[0][1] + Rt.m[0][2] * dL_dMt.m[0][2];
    dscale.y = Rt.m[1][0] * dL_dMt.m[1][0] + Rt.m[1][1] * dL_dMt.m[1][1] + Rt.m[1][2] * dL_dMt.m[1][2];
    dscale.z = Rt.m[2][0] * dL_dMt.m[2][0] + Rt.m[2][1] * dL_dMt.m[2][1] + Rt.m[2][2] * dL_dMt.m[2][2];
    float* ps = a.dL_dscales + 3 * (size_t)idx;
    put<ACC>(ps, dscale.x, acc, IBGS_ACC_SCALES); put<ACC>(ps + 1, dscale.y, acc, IBGS_ACC_SCALES);
    put<ACC>(ps + 2, dscale.z, acc, IBGS_ACC_SCALES);
#pragma unroll
    for (int w = 0; w < 3; w++) {
      dL_dMt.m[0][w] *= s.x;
      dL_dMt.m[1][w] *= s.y;
      dL_dMt.m[2][w] *= s.z;
    }
    float4 dq;
    dq.x = 2 * z * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * y * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) +
           2 * x * (dL_dMt.m[1][2] - dL_dMt.m[2][1]);
    dq.y = 2 * y * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * z * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) +
           2 * r * (dL_dMt.m[1][2] - dL_dMt.m[2][1]) - 4 * x * (dL_dMt.m[2][2] + dL_dMt.m[1][1]);
    dq.z = 2 * x * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * r * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) +
           2 * z * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * y * (dL_dMt.m[2][2] + dL_dMt.m[0][0]);
    dq.w = 2 * r * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * x * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) +
           2 * y * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * z * (dL_dMt.m[1][1] + dL_dMt.m[0][0]);
    float4* pq = reinterpret_cast<float4*>(a.dL_drotations) + idx;  // gradient w.r.t. the un-normalised quaternion (:437)
    if (ACC && (acc & IBGS_ACC_ROTATIONS)) {
      const float4 o = *pq;
      dq = make_float4(o.x + dq.x, o.y + dq.y, o.z + dq.z, o.w + dq.w);
    }
    *pq = dq;
  } else {
    if (!(acc & IBGS_ACC_SCALES)) { float* ps = a.dL_dscales + 3 * (size_t)idx; ps[0] = ps[1] = ps[2] = 0.f; }
    if (!(acc & IBGS_ACC_ROTATIONS)) reinterpret_cast<float4*>(a.dL_drotations)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <bool ACC>
__global__ void __launch_bounds__(PB_THREADS, PB_MIN_CTAS) preprocess_backward_kernel(const PBArgs a) {
  extern __shared__ float s_tiles[];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int L = a.M * 3;        // words per SH row
  const int Ls = L | 1;         // odd stride: conflict-free row access
  float* tile = s_tiles + (threadIdx.x >> 5) * (32 * Ls);
  const size_t row0 = (size_t)idx - lane;
  if (row0 >= (size_t)a.P) return;  // whole warp out of range
  const int rows = min(32, a.P - (int)row0);
  const bool in_range = idx < a.P;
  const bool visible = in_range && (a.radii[idx] > 0);
  const bool any_visible = __any_sync(0xffffffffu, visible);
  const bool has_sh = a.shs != nullptr;
  // SH rows: one [P,M,3] tensor, or DC + coefficients 1..M-1 as two tensors (columns 0..2 and 3.. of the tile)
  const bool split = a.shs_rest != nullptr;
  const bool vec = a.vec_ok != 0;
  if (has_sh && any_visible) {
    if (split) {
      tile_copy<0>(tile, const_cast<float*>(a.shs) + row0 * 3, rows * 3, 3, Ls, 0, a.magic_dc, vec, lane);
      if (L > 3)
        tile_copy<0>(tile, const_cast<float*>(a.shs_rest) + row0 * (L - 3), rows * (L - 3), L - 3, Ls, 3,
                         a.magic_rest, vec, lane);
    } else {
      tile_copy<0>(tile, const_cast<float*>(a.shs) + row0 * L, rows * L, L, Ls, 0, a.magic, vec, lane);
    }
    __syncwarp();
  }
  if (visible) {
    preprocess_backward_one<ACC>(a, idx, tile + lane * Ls);
  } else if (in_range) {
    write_zero_row<ACC>(a, idx);
    if (has_sh) {
      float* row = tile + lane * Ls;
      for (int k = 0; k < L; k++) row[k] = 0.f;
    }
  }
  if (has_sh) {
    __syncwarp();
    // accumulated SH gradients: a block of 32 culled Gaussians has nothing to add
    const bool acc_dc = ACC && (a.acc & IBGS_ACC_SH), acc_rest = ACC && (a.acc & IBGS_ACC_SH_REST);
    if (split) {
      if (acc_dc) { if (any_visible) tile_copy<2>(tile, a.dL_dsh + row0 * 3, rows * 3, 3, Ls, 0, a.magic_dc, vec, lane); }
      else tile_copy<1>(tile, a.dL_dsh + row0 * 3, rows * 3, 3, Ls, 0, a.magic_dc, vec, lane);
      if (L > 3) {
        if (acc_rest) { if (any_visible) tile_copy<2>(tile, a.dL_dsh_rest + row0 * (L - 3), rows * (L - 3), L - 3, Ls, 3, a.magic_rest, vec, lane); }
        else tile_copy<1>(tile, a.dL_dsh_rest + row0 * (L - 3), rows * (L - 3), L - 3, Ls, 3, a.magic_rest, vec, lane);
      }
    } else {
      if (acc_dc) { if (any_visible) tile_copy<2>(tile, a.dL_dsh + row0 * L, rows * L, L, Ls, 0, a.magic, vec, lane); }
      else tile_copy<1>(tile, a.dL_dsh + row0 * L, rows * L, L, Ls, 0, a.magic, vec, lane);
    }
  }
}

}  // namespace

int launch_preprocess_backward(const IbgsBackwardArgs& f, const GeomState& g, const float4* arena,
                               float focal_x, float focal_y, cudaStream_t s) {
  PBArgs a;
  a.P = f.P;
  a.D = f.view.sh_degree;
  a.M = f.view.sh_coeffs;
  a.means3D = f.means3D;
  a.radii = f.radii;
  a.shs = f.shs;
  a.shs_rest = f.shs_rest;
  a.clamped = g.clamped;
  a.scales = f.scales;
  a.rotations = f.rotations;
  a.scale_modifier = f.view.scale_modifier;
  a.cov3D_precomp = f.cov3D_precomp;
  a.view = f.view.viewmatrix;
  a.proj = f.view.projmatrix;
  a.magic = (a.M > 0) ? (uint32_t)(0x100000000ull / (uint64_t)(a.M * 3)) + 1u : 0u;
  a.magic_dc = (uint32_t)(0x100000000ull / 3ull) + 1u;
  a.magic_rest = (a.M > 1) ? (uint32_t)(0x100000000ull / (uint64_t)((a.M - 1) * 3)) + 1u : 0u;
  a.vec_ok = ((((uintptr_t)f.shs) | ((uintptr_t)f.dL_dsh) | ((uintptr_t)f.shs_rest) | ((uintptr_t)f.dL_dsh_rest)) & 15u) == 0;
  a.half_w = (float)(0.5 * f.view.image_width);
  a.half_h = (float)(0.5 * f.view.image_height);
  a.h_x = focal_x;
  a.h_y = focal_y;
  a.tan_fovx = f.view.tanfovx;
  a.tan_fovy = f.view.tanfovy;
  a.campos = f.view.campos;
  a.arena = arena;
  a.dL_dmeans3D = f.dL_dmeans3D;
  a.dL_dmeans2D = f.dL_dmeans2D;
  a.dL_dmeans2D_abs = f.dL_dmeans2D_abs;
  a.dL_dcolors = f.dL_dcolors;
  a.dL_dopacity = f.dL_dopacity;
  a.dL_dcov3D = f.dL_dcov3D;
  a.dL_dsh = (f.shs != nullptr) ? f.dL_dsh : nullptr;
  a.dL_dsh_rest = (f.shs != nullptr && f.shs_rest != nullptr) ? f.dL_dsh_rest : nullptr;
  a.dL_dscales = f.dL_dscales;
  a.dL_drotations = f.dL_drotations;
  a.dL_dall_map = f.dL_dall_map;
  a.acc = f.accumulate_mask;
  ProfScope prof(PROF_PREPROCESS_BWD, s);
  const size_t smem = (size_t)PB_WARPS * 32 * ((a.M * 3) | 1) * sizeof(float);
  if (a.acc) preprocess_backward_kernel<true><<<(f.P + PB_THREADS - 1) / PB_THREADS, PB_THREADS, smem, s>>>(a);
  else preprocess_backward_kernel<false><<<(f.P + PB_THREADS - 1) / PB_THREADS, PB_THREADS, smem, s>>>(a);
  KERNEL_CHECK(f.view.debug, s);
  return IBGS_OK;
}
