// preprocess.cu -- per-Gaussian forward preprocess and markVisible.
//
// Reference behaviour: FORWARD::preprocessCUDA<3> (cuda_rasterizer/forward.cu:194-295) with
// in_frustum (auxiliary.h:143-168), computeCov3D (forward.cu:156-190), computeCov2D (:112-151),
// computeColorFromSH (:58-109), ndc2Pix / getRect (auxiliary.h:45-60); checkFrustum
// (rasterizer_impl.cu:171-183).
//
// Design: one thread per Gaussian, one 64-byte render record written per visible Gaussian
// (4 x st.global.v4) so that the tile renderers gather ONE aligned record per tile instance
// instead of the reference's four separate arrays + per-pair global re-reads.  The geometric
// expression trees (projection, cov3D, cov2D, conic, radius, ndc2Pix) are kept operation for
// operation because radii / tile keys / blend decisions must match the reference bit for bit.
#include "common.cuh"
#include <cstdio>

namespace {

struct PreArgs {
  int P, D, M;
  const float* means3D;
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* opacities;
  const float* shs;
  const float* shs_rest;
  const float* cov3D_precomp;
  const float* colors_precomp;
  const float* all_map;
  const float* viewmatrix;
  const float* projmatrix;
  const float* campos;
  int W, H;
  float tan_fovx, tan_fovy, focal_x, focal_y;
  int* radii;
  float4* rec;
  float* depths;
  uint32_t* tiles_touched;
  uint8_t* clamped;
  dim3 grid;
  int prefiltered;
  int render_depth_only;
};

// reference auxiliary.h:45-48 (double arithmetic on purpose: float v and int S promote)
__forceinline__ __device__ float ndc2Pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// reference forward.cu:156-190.  The quaternion is used as given (normalisation is commented out
// in the reference, :165).
__forceinline__ __device__ void computeCov3D(float sx, float sy, float sz, float mod, float4 rot,
                                             float* cov3D) {
  M3 S = m3(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
  S.m[0][0] = mod * sx;
  S.m[1][1] = mod * sy;
  S.m[2][2] = mod * sz;
  float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
  M3 R = m3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
            2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
            2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
  M3 Mm = m3_mul(S, R);
  M3 Sigma = m3_mul(m3_t(Mm), Mm);
  cov3D[0] = Sigma.m[0][0];
  cov3D[1] = Sigma.m[0][1];
  cov3D[2] = Sigma.m[0][2];
  cov3D[3] = Sigma.m[1][1];
  cov3D[4] = Sigma.m[1][2];
  cov3D[5] = Sigma.m[2][2];
}

// reference forward.cu:112-151
__forceinline__ __device__ float3 computeCov2D(const float3& mean, float focal_x, float focal_y,
                                               float tan_fovx, float tan_fovy, const float* cov3D,
                                               const float* viewmatrix) {
  float3 t = transformPoint4x3(mean, viewmatrix);
  const float limx = 1.3f * tan_fovx;
  const float limy = 1.3f * tan_fovy;
  const float txtz = t.x / t.z;
  const float tytz = t.y / t.z;
  t.x = min(limx, max(-limx, txtz)) * t.z;
  t.y = min(limy, max(-limy, tytz)) * t.z;

  M3 J = m3(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z),
            0.0f, focal_y / t.z, -(focal_y * t.y) / (t.z * t.z),
            0, 0, 0);
  M3 W = m3(viewmatrix[0], viewmatrix[4], viewmatrix[8],
            viewmatrix[1], viewmatrix[5], viewmatrix[9],
            viewmatrix[2], viewmatrix[6], viewmatrix[10]);
  M3 T = m3_mul(W, J);
  M3 Vrk = m3(cov3D[0], cov3D[1], cov3D[2],
              cov3D[1], cov3D[3], cov3D[4],
              cov3D[2], cov3D[4], cov3D[5]);
  M3 cov = m3_mul(m3_mul(m3_t(T), m3_t(Vrk)), T);
  cov.m[0][0] += 0.3f;
  cov.m[1][1] += 0.3f;
  return {float(cov.m[0][0]), float(cov.m[0][1]), float(cov.m[1][1])};
}

// reference forward.cu:58-109 (vec3 arithmetic written out per channel in glm's evaluation order)
__forceinline__ __device__ float3 computeColorFromSH(int idx, int deg, int max_coeffs, float3 pos,
                                                     float3 campos, const float* shs, const float* shs_rest,
                                                     uint8_t* clamped_bits) {
  float3 dir = {pos.x - campos.x, pos.y - campos.y, pos.z - campos.z};
  float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
  dir.x = dir.x / len;
  dir.y = dir.y / len;
  dir.z = dir.z / len;
  // coefficient k of channel c is sh0[c] for k = 0 and sh[3k + c] for k >= 1: one [P,M,3] tensor, or the DC part and
  // coefficients 1..M-1 as two tensors (the -3 makes the same 3k + c index valid for the second tensor)
  const float* sh0 = shs_rest ? shs + (size_t)idx * 3 : shs + (size_t)idx * max_coeffs * 3;
  const float* sh = shs_rest ? shs_rest + (size_t)idx * (max_coeffs - 1) * 3 - 3 : sh0;
  float res[3];
#pragma unroll
  for (int c = 0; c < 3; c++) res[c] = SH_C0 * sh0[c];
  if (deg > 0) {
    float x = dir.x, y = dir.y, z = dir.z;
#pragma unroll
    for (int c = 0; c < 3; c++)
      res[c] = res[c] - SH_C1 * y * sh[3 + c] + SH_C1 * z * sh[6 + c] - SH_C1 * x * sh[9 + c];
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z;
      float xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
      for (int c = 0; c < 3; c++)
        res[c] = res[c] + SH_C2[0] * xy * sh[12 + c] + SH_C2[1] * yz * sh[15 + c] +
                 SH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + c] + SH_C2[3] * xz * sh[21 + c] +
                 SH_C2[4] * (xx - yy) * sh[24 + c];
      if (deg > 2) {
#pragma unroll
        for (int c = 0; c < 3; c++)
          res[c] = res[c] + SH_C3[0] * y * (3.0f * xx - yy) * sh[27 + c] +
                   SH_C3[1] * xy * z * sh[30 + c] + SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + c] +
                   SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
                   SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + c] +
                   SH_C3[5] * z * (xx - yy) * sh[42 + c] + SH_C3[6] * x * (xx - 3.0f * yy) * sh[45 + c];
      }
    }
  }
  uint8_t bits = 0;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    res[c] += 0.5f;
    if (res[c] < 0) bits |= (1u << c);
    res[c] = fmaxf(res[c], 0.0f);
  }
  *clamped_bits = bits;
  return {res[0], res[1], res[2]};
}

__global__ void __launch_bounds__(256) preprocess_kernel(const PreArgs a) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.P) return;

  // forward.cu:227-228
  a.radii[idx] = 0;
  a.tiles_touched[idx] = 0;
  // binning.cu sorts the Gaussians by the bit pattern of depths[]: culled ones get the largest key
  reinterpret_cast<uint32_t*>(a.depths)[idx] = 0xFFFFFFFFu;

  // in_frustum, auxiliary.h:143-168
  float3 p_orig = {a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2]};
  float3 p_view = transformPoint4x3(p_orig, a.viewmatrix);
  if (p_view.z <= 0.2f) {
    if (a.prefiltered) {
      printf("Point is filtered although prefiltered is set. This shouldn't happen!");
      __trap();
    }
    return;
  }
  float4 p_hom = transformPoint4x4(p_orig, a.projmatrix);
  float p_w = 1.0f / (p_hom.w + 0.0000001f);
  float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};

  float cov3D_local[6];
  const float* cov3D;
  if (a.cov3D_precomp != nullptr) {
    cov3D = a.cov3D_precomp + (size_t)idx * 6;
  } else {
    float4 rot = reinterpret_cast<const float4*>(a.rotations)[idx];
    computeCov3D(a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2], a.scale_modifier, rot,
                 cov3D_local);
    cov3D = cov3D_local;
  }

  float3 cov = computeCov2D(p_orig, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy, cov3D, a.viewmatrix);

  // forward.cu:258-276
  float det = (cov.x * cov.z - cov.y * cov.y);
  if (det == 0.0f) return;
  float det_inv = 1.f / det;
  float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};

  float mid = 0.5f * (cov.x + cov.z);
  float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
  float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
  float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
  float2 point_image = {ndc2Pix(p_proj.x, a.W), ndc2Pix(p_proj.y, a.H)};
  uint2 rect_min, rect_max;
  getRect(point_image, my_radius, rect_min, rect_max, a.grid);
  if ((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y) == 0) return;

  // features: SH colour, precomputed colour, or none (depth-only), forward.cu:280-286
  float3 feat = {0.f, 0.f, 0.f};
  if (a.colors_precomp != nullptr) {
    feat = {a.colors_precomp[3 * idx], a.colors_precomp[3 * idx + 1], a.colors_precomp[3 * idx + 2]};
  } else if (!a.render_depth_only) {
    float3 campos = {a.campos[0], a.campos[1], a.campos[2]};
    uint8_t bits;
    feat = computeColorFromSH(idx, a.D, a.M, p_orig, campos, a.shs, a.shs_rest, &bits);
    a.clamped[idx] = bits;
  }

  const float opacity = a.opacities[idx];

  // Cull threshold for the tile renderers: o*exp(power) >= 1/255  <=>  -power <= ln(255 o) =: tau.  A sub-tile
  // whose minimum of -power exceeds tau cannot blend this Gaussian (the reference would reject every pair at
  // forward.cu:425); the 0.02 margin covers __expf / rounding error.  NaN never culls (comparisons false).
  const float tau = __logf(255.0f * opacity) + 0.02f;

  float4 plane_n = {0.f, 0.f, 0.f, 0.f};
  float plane_d = 0.f;
  if (a.all_map != nullptr) {
    const float* am = a.all_map + (size_t)idx * 5;
    plane_n = {am[0], am[1], am[2], 0.f};
    plane_d = am[4];
  }

  // forward.cu:289-294
  a.depths[idx] = p_view.z;
  a.radii[idx] = my_radius;
  float4* rec = a.rec + 4 * (size_t)idx;
  rec[0] = {point_image.x, point_image.y, conic.x, conic.y};
  rec[1] = {conic.z, opacity, tau, 0.f};
  rec[2] = {feat.x, feat.y, feat.z, plane_d};
  rec[3] = plane_n;
  a.tiles_touched[idx] = (rect_max.y - rect_min.y) * (rect_max.x - rect_min.x);
}

// ---- batched depth-only preprocess (ibgs_forward_depth_batch) -------------------------------
// One thread per Gaussian reads its parameters ONCE (the 3D covariance does not depend on the view) and then walks
// the V views, writing the (view, Gaussian) item v*P + idx of every per-item array.  Per view the arithmetic is
// preprocess_kernel's, operation for operation, so radii / depths / tiles_touched / records are bit-identical to V
// separate depth-only calls given the same plane parameters.
struct PreBatchArgs {
  int P, V;
  const float* means3D;
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* opacities;
  const float* cov3D_precomp;
  const float* all_maps;
  const float* normals;
  const float* offsets;
  const float* camera_centers;
  const float* viewmatrices;
  const float* projmatrices;
  int W, H;
  float tan_fovx, tan_fovy, focal_x, focal_y;
  int* radii;
  float4* rec;
  float* depths;
  uint32_t* tiles_touched;
  unsigned long long* counts;
  dim3 grid;
  int prefiltered;
};

__global__ void __launch_bounds__(256) preprocess_depth_batch_kernel(const PreBatchArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = idx < a.P;
  const int i = valid ? idx : 0;

  const float3 p_orig = {a.means3D[3 * (size_t)i], a.means3D[3 * (size_t)i + 1], a.means3D[3 * (size_t)i + 2]};
  float cov3D[6];
  if (a.cov3D_precomp != nullptr) {
#pragma unroll
    for (int k = 0; k < 6; k++) cov3D[k] = a.cov3D_precomp[(size_t)i * 6 + k];
  } else {
    const float4 rot = reinterpret_cast<const float4*>(a.rotations)[i];
    computeCov3D(a.scales[3 * (size_t)i], a.scales[3 * (size_t)i + 1], a.scales[3 * (size_t)i + 2], a.scale_modifier, rot,
                 cov3D);
  }
  const float opacity = a.opacities[i];
  const float tau = __logf(255.0f * opacity) + 0.02f;
  float nrm[3] = {0.f, 0.f, 0.f};
  float off = 0.f;
  if (a.all_maps == nullptr) {
    nrm[0] = a.normals[3 * (size_t)i]; nrm[1] = a.normals[3 * (size_t)i + 1]; nrm[2] = a.normals[3 * (size_t)i + 2];
    if (a.offsets != nullptr) off = a.offsets[i];
  }
  const float pw[3] = {p_orig.x, p_orig.y, p_orig.z};

  for (int v = 0; v < a.V; v++) {
    const float* viewmatrix = a.viewmatrices + 16 * v;
    const float* projmatrix = a.projmatrices + 16 * v;
    const size_t o = (size_t)v * a.P + i;
    uint32_t tiles = 0;
    int radius_out = 0;
    uint32_t depth_bits = 0xFFFFFFFFu;
    do {
      if (!valid) break;
      const float3 p_view = transformPoint4x3(p_orig, viewmatrix);
      if (p_view.z <= 0.2f) {
        if (a.prefiltered) {
          printf("Point is filtered although prefiltered is set. This shouldn't happen!");
          __trap();
        }
        break;
      }
      const float4 p_hom = transformPoint4x4(p_orig, projmatrix);
      const float p_w = 1.0f / (p_hom.w + 0.0000001f);
      const float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};
      const float3 cov = computeCov2D(p_orig, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy, cov3D, viewmatrix);
      const float det = (cov.x * cov.z - cov.y * cov.y);
      if (det == 0.0f) break;
      const float det_inv = 1.f / det;
      const float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};
      const float mid = 0.5f * (cov.x + cov.z);
      const float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
      const float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
      const float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
      const float2 point_image = {ndc2Pix(p_proj.x, a.W), ndc2Pix(p_proj.y, a.H)};
      uint2 rect_min, rect_max;
      getRect(point_image, my_radius, rect_min, rect_max, a.grid);
      const uint32_t area = (rect_max.x - rect_min.x) * (rect_max.y - rect_min.y);
      if (area == 0) break;

      float4 plane_n;
      float plane_d;
      if (a.all_maps != nullptr) {
        const float* am = a.all_maps + o * 5;
        plane_n = {am[0], am[1], am[2], 0.f};
        plane_d = am[4];
      } else {
        const PlaneTerms t = plane_terms(nrm, off, pw, viewmatrix, a.camera_centers + 3 * v);
        plane_n = {t.ln[0], t.ln[1], t.ln[2], 0.f};
        plane_d = fabsf(t.u);
      }
      depth_bits = __float_as_uint(p_view.z);
      radius_out = my_radius;
      float4* rec = a.rec + 4 * o;
      rec[0] = {point_image.x, point_image.y, conic.x, conic.y};
      rec[1] = {conic.z, opacity, tau, 0.f};
      rec[2] = {0.f, 0.f, 0.f, plane_d};
      rec[3] = plane_n;
      tiles = area;
    } while (false);
    if (valid) {
      a.radii[o] = radius_out;
      a.tiles_touched[o] = tiles;
      reinterpret_cast<uint32_t*>(a.depths)[o] = depth_bits;
    }
    if (a.counts != nullptr) {
      const uint32_t wsum = __reduce_add_sync(0xffffffffu, tiles);
      if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(a.counts + v, (unsigned long long)wsum);
    }
  }
}

// reference rasterizer_impl.cu:171-183
__global__ void mark_visible_kernel(int P, const float* means3D, const float* view, uint8_t* present) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  float3 p = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
  float3 p_view = transformPoint4x3(p, view);
  present[idx] = (p_view.z <= 0.2f) ? 0 : 1;
}

}  // namespace

int launch_preprocess(const IbgsForwardArgs& f, const GeomState& g, float focal_x, float focal_y,
                      dim3 grid, cudaStream_t s) {
  PreArgs a;
  a.P = f.P;
  a.D = f.view.sh_degree;
  a.M = f.view.sh_coeffs;
  a.means3D = f.means3D;
  a.scales = f.scales;
  a.scale_modifier = f.view.scale_modifier;
  a.rotations = f.rotations;
  a.opacities = f.opacities;
  a.shs = f.shs;
  a.shs_rest = f.shs_rest;
  a.cov3D_precomp = f.cov3D_precomp;
  a.colors_precomp = f.colors_precomp;
  a.all_map = f.all_map;
  a.viewmatrix = f.view.viewmatrix;
  a.projmatrix = f.view.projmatrix;
  a.campos = f.view.campos;
  a.W = f.view.image_width;
  a.H = f.view.image_height;
  a.tan_fovx = f.view.tanfovx;
  a.tan_fovy = f.view.tanfovy;
  a.focal_x = focal_x;
  a.focal_y = focal_y;
  a.radii = f.radii;
  a.rec = g.rec;
  a.depths = g.depths;
  a.tiles_touched = g.tiles_touched;
  a.clamped = g.clamped;
  a.grid = grid;
  a.prefiltered = f.view.prefiltered;
  a.render_depth_only = f.view.render_depth_only;
  ProfScope prof(PROF_PREPROCESS, s);
  preprocess_kernel<<<(f.P + 255) / 256, 256, 0, s>>>(a);
  KERNEL_CHECK(f.view.debug, s);
  return IBGS_OK;
}

int launch_preprocess_depth_batch(const IbgsDepthBatchArgs& f, const GeomState& g, int* radii,
                                  unsigned long long* counts, float focal_x, float focal_y, dim3 grid,
                                  cudaStream_t s) {
  PreBatchArgs a;
  a.P = f.P;
  a.V = f.V;
  a.means3D = f.means3D;
  a.scales = f.scales;
  a.scale_modifier = f.scale_modifier;
  a.rotations = f.rotations;
  a.opacities = f.opacities;
  a.cov3D_precomp = f.cov3D_precomp;
  a.all_maps = f.all_maps;
  a.normals = f.normals;
  a.offsets = f.offsets;
  a.camera_centers = f.camera_centers;
  a.viewmatrices = f.viewmatrices;
  a.projmatrices = f.projmatrices;
  a.W = f.image_width;
  a.H = f.image_height;
  a.tan_fovx = f.tanfovx;
  a.tan_fovy = f.tanfovy;
  a.focal_x = focal_x;
  a.focal_y = focal_y;
  a.radii = radii;
  a.rec = g.rec;
  a.depths = g.depths;
  a.tiles_touched = g.tiles_touched;
  a.counts = counts;
  a.grid = grid;
  a.prefiltered = f.prefiltered;
  ProfScope prof(PROF_PREPROCESS, s);
  preprocess_depth_batch_kernel<<<(f.P + 255) / 256, 256, 0, s>>>(a);
  KERNEL_CHECK(f.debug, s);
  return IBGS_OK;
}

int launch_mark_visible(int P, const float* means3D, const float* view, const float* proj,
                        uint8_t* present, cudaStream_t s) {
  (void)proj;  // the reference computes p_hom but only tests view-space z (auxiliary.h:156-158)
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, present);
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}
