// prologue.cu -- fused per-view parameter prologue of gaussian_renderer.render() and its backward.
//
// Reference behaviour: before every rasterizer call the reference runs ~30 small
// PyTorch kernels over all P Gaussians -- the activations of GaussianModel's getters
// (scene/gaussian_model.py:127-147: exp on _scaling, F.normalize on _rotation, sigmoid on _opacity, torch.cat of
// _features_dc / _features_rest), the learnt plane normal (get_normal, scene/gaussian_model.py:166-173: normalise,
// flip towards the camera, flip the offset with it) and the all_map construction
// (gaussian_renderer/__init__.py:304-315: rotate the normal into the camera frame, plane distance in the camera
// frame, abs) -- and autograd replays twice as many in the backward.  Here each direction is two launches:
// one thread per Gaussian for everything but the SH rows, and a flat coalesced copy that concatenates / splits them.
// Arithmetic is float32 in the same operation order as the torch expressions it replaces.
#include "common.cuh"

namespace {

struct ProArgs {
  int P, L;  // L = 3*K words per concatenated SH row
  const float *xyz, *opacity_raw, *scaling_raw, *rotation_raw, *fdc, *frest, *normal_raw, *offset, *view, *campos;
  float *opacity, *scales, *rotations, *shs, *all_map;
  // backward
  const float *g_opacity, *g_scales, *g_rotations, *g_shs, *g_all_map;
  float *d_xyz, *d_opacity_raw, *d_scaling_raw, *d_rotation_raw, *d_fdc, *d_frest, *d_normal_raw, *d_offset;
  int smallest_axis;     // plane normal = shortest axis of the Gaussian instead of the learnt normal
  uint32_t magic;        // floor(2^32 / L) + 1
  uint32_t magic_limit;  // word indices below 2^32 / L divide exactly by multiply-high with `magic`
};


// ---- learnt_normal = False: the plane normal is the Gaussian's shortest axis (scene/gaussian_model.py:149-161) --------
// get_smallest_axis: column argmin(get_scaling) of quaternion_to_matrix(get_rotation) (pytorch3d convention: real part
// first, R = I + (2 / |q|^2) A(q) -- the quaternion it receives is already normalised, the 2 / |q|^2 is kept as written).
// A(q) column c, rows 0..2, for q = (r, i, j, k)
__device__ __forceinline__ void quat_column(const float q[4], int c, float a[3]) {
  const float r = q[0], i = q[1], j = q[2], k = q[3];
  if (c == 0) { a[0] = -(j * j + k * k); a[1] = i * j + k * r; a[2] = i * k - j * r; }
  else if (c == 1) { a[0] = i * j - k * r; a[1] = -(i * i + k * k); a[2] = j * k + i * r; }
  else { a[0] = i * k + j * r; a[1] = j * k - i * r; a[2] = -(i * i + j * j); }
}
// d a[row] / d q[m] of the column above
__device__ __forceinline__ void quat_column_grad(const float q[4], int c, float da[3][4]) {
  const float r = q[0], i = q[1], j = q[2], k = q[3];
  if (c == 0) {
    da[0][0] = 0.f; da[0][1] = 0.f; da[0][2] = -2.f * j; da[0][3] = -2.f * k;
    da[1][0] = k;   da[1][1] = j;   da[1][2] = i;        da[1][3] = r;
    da[2][0] = -j;  da[2][1] = k;   da[2][2] = -r;       da[2][3] = i;
  } else if (c == 1) {
    da[0][0] = -k;  da[0][1] = j;        da[0][2] = i;   da[0][3] = -r;
    da[1][0] = 0.f; da[1][1] = -2.f * i; da[1][2] = 0.f; da[1][3] = -2.f * k;
    da[2][0] = i;   da[2][1] = r;        da[2][2] = k;   da[2][3] = j;
  } else {
    da[0][0] = j;   da[0][1] = k;        da[0][2] = r;        da[0][3] = i;
    da[1][0] = -i;  da[1][1] = -r;       da[1][2] = k;        da[1][3] = j;
    da[2][0] = 0.f; da[2][1] = -2.f * i; da[2][2] = -2.f * j; da[2][3] = 0.f;
  }
}
__device__ __forceinline__ int argmin3(float a, float b, float c) {   // first minimum, like torch.min(dim)
  int m = 0;
  float v = a;
  if (b < v) { v = b; m = 1; }
  if (c < v) { m = 2; }
  return m;
}
// unit world normal (no normalisation step in this mode) -> flip towards the camera, camera-frame normal, plane distance
// (get_normal_w_smallest_axis + gaussian_renderer/__init__.py:307-311 without the offset)
__device__ __forceinline__ PlaneTerms plane_terms_unit(const float* nh, const float* p, const float* V, const float* cam) {
  PlaneTerms t;
  t.inv_len = 1.0f;
#pragma unroll
  for (int i = 0; i < 3; i++) t.nh[i] = nh[i];
  const float d = __fadd_rn(__fadd_rn(__fmul_rn(nh[0], cam[0] - p[0]), __fmul_rn(nh[1], cam[1] - p[1])),
                            __fmul_rn(nh[2], cam[2] - p[2]));
  t.sgn = (d < 0.0f) ? -1.0f : 1.0f;
#pragma unroll
  for (int i = 0; i < 3; i++) t.ng[i] = (d < 0.0f) ? -nh[i] : nh[i];
#pragma unroll
  for (int j = 0; j < 3; j++) t.ln[j] = t.ng[0] * V[0 * 4 + j] + t.ng[1] * V[1 * 4 + j] + t.ng[2] * V[2 * 4 + j];
  const float gd = -(t.ng[0] * p[0] + t.ng[1] * p[1] + t.ng[2] * p[2]);
  t.u = gd - (t.ln[0] * V[12] + t.ln[1] * V[13] + t.ln[2] * V[14]);
  return t;
}

__global__ void __launch_bounds__(256) prologue_forward_kernel(const ProArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.P) return;
  a.opacity[i] = 1.0f / (1.0f + expf(-a.opacity_raw[i]));  // torch.sigmoid
  float r[4];
  float sc[3];
#pragma unroll
  for (int k = 0; k < 3; k++) { sc[k] = expf(a.scaling_raw[3 * (size_t)i + k]); a.scales[3 * (size_t)i + k] = sc[k]; }
  const float4 q = reinterpret_cast<const float4*>(a.rotation_raw)[i];
  r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w;
  const float qn = fmaxf(sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]), 1e-12f);  // F.normalize eps
  reinterpret_cast<float4*>(a.rotations)[i] = make_float4(r[0] / qn, r[1] / qn, r[2] / qn, r[3] / qn);
  if (a.all_map) {
    const float p[3] = {a.xyz[3 * (size_t)i], a.xyz[3 * (size_t)i + 1], a.xyz[3 * (size_t)i + 2]};
    PlaneTerms t;
    if (a.smallest_axis) {
      const float y[4] = {r[0] / qn, r[1] / qn, r[2] / qn, r[3] / qn};   // get_rotation
      const float two_s = 2.0f / (y[0] * y[0] + y[1] * y[1] + y[2] * y[2] + y[3] * y[3]);
      const int c = argmin3(sc[0], sc[1], sc[2]);
      float col[3];
      quat_column(y, c, col);
      float nh[3];
#pragma unroll
      for (int k = 0; k < 3; k++) nh[k] = two_s * col[k];
      nh[c] = 1.0f + nh[c];                                               // diagonal entry: 1 - two_s (..)
      t = plane_terms_unit(nh, p, a.view, a.campos);
    } else {
      const float n[3] = {a.normal_raw[3 * (size_t)i], a.normal_raw[3 * (size_t)i + 1], a.normal_raw[3 * (size_t)i + 2]};
      t = plane_terms(n, a.offset[i], p, a.view, a.campos);
    }
    float* o = a.all_map + 5 * (size_t)i;
    o[0] = t.ln[0]; o[1] = t.ln[1]; o[2] = t.ln[2]; o[3] = 1.0f; o[4] = fabsf(t.u);
  }
}

// shs[p][0] = features_dc[p][0], shs[p][1..K-1] = features_rest[p]  (torch.cat, scene/gaussian_model.py:139-143);
// one thread per output word, consecutive threads -> consecutive words on both sides
template <bool SPLIT>
__global__ void __launch_bounds__(256) sh_rows_kernel(const ProArgs a) {
  const size_t total = (size_t)a.P * a.L;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    // row = i / L: multiply-high where that is exact (error term i / 2^32 < 1 / L), else a real division
    const size_t p = (i < a.magic_limit) ? (size_t)__umulhi((uint32_t)i, a.magic) : i / (size_t)a.L;
    const int k = (int)(i - p * a.L);
    if (SPLIT) {
      const float g = a.g_shs ? a.g_shs[i] : 0.0f;
      if (k < 3) a.d_fdc[3 * p + k] = g; else a.d_frest[p * (a.L - 3) + (k - 3)] = g;
    } else {
      a.shs[i] = (k < 3) ? a.fdc[3 * p + k] : a.frest[p * (a.L - 3) + (k - 3)];
    }
  }
}

__global__ void __launch_bounds__(256) prologue_backward_kernel(const ProArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.P) return;
  {
    const float o = 1.0f / (1.0f + expf(-a.opacity_raw[i]));
    a.d_opacity_raw[i] = (a.g_opacity ? a.g_opacity[i] : 0.0f) * (1.0f - o) * o;  // sigmoid_backward: g*(1-y)*y
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const size_t e = 3 * (size_t)i + k;
    a.d_scaling_raw[e] = (a.g_scales ? a.g_scales[e] : 0.0f) * expf(a.scaling_raw[e]);
  }
  float dx[3] = {0.f, 0.f, 0.f};
  float g_y_extra[4] = {0.f, 0.f, 0.f, 0.f};   // gradient reaching the NORMALISED quaternion through the plane normal
  const float4 q = reinterpret_cast<const float4*>(a.rotation_raw)[i];
  const float len = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  if (a.normal_raw || a.smallest_axis) {
    const float p[3] = {a.xyz[3 * (size_t)i], a.xyz[3 * (size_t)i + 1], a.xyz[3 * (size_t)i + 2]};
    const float* V = a.view;
    PlaneTerms t;
    float y[4] = {0.f, 0.f, 0.f, 0.f}, two_s = 0.f;
    int c = 0;
    if (a.smallest_axis) {
      const float qn = fmaxf(len, 1e-12f);
      y[0] = q.x / qn; y[1] = q.y / qn; y[2] = q.z / qn; y[3] = q.w / qn;
      two_s = 2.0f / (y[0] * y[0] + y[1] * y[1] + y[2] * y[2] + y[3] * y[3]);
      c = argmin3(expf(a.scaling_raw[3 * (size_t)i]), expf(a.scaling_raw[3 * (size_t)i + 1]),
                  expf(a.scaling_raw[3 * (size_t)i + 2]));
      float col[3], nh[3];
      quat_column(y, c, col);
#pragma unroll
      for (int k = 0; k < 3; k++) nh[k] = two_s * col[k];
      nh[c] = 1.0f + nh[c];
      t = plane_terms_unit(nh, p, V, a.campos);
    } else {
      const float n[3] = {a.normal_raw[3 * (size_t)i], a.normal_raw[3 * (size_t)i + 1], a.normal_raw[3 * (size_t)i + 2]};
      t = plane_terms(n, a.offset[i], p, V, a.campos);
    }
    float g_ln[3] = {0.f, 0.f, 0.f};
    float g_ld = 0.f;
    if (a.g_all_map) {
      const float* g = a.g_all_map + 5 * (size_t)i;
      g_ln[0] = g[0]; g_ln[1] = g[1]; g_ln[2] = g[2]; g_ld = g[4];
    }
    // all_map[4] = |u|, u = gd - ln.t  (torch.abs backward: g * sign(u))
    const float g_u = g_ld * ((t.u > 0.0f) ? 1.0f : ((t.u < 0.0f) ? -1.0f : 0.0f));
#pragma unroll
    for (int j = 0; j < 3; j++) g_ln[j] -= g_u * V[12 + j];
    float g_ng[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      g_ng[k] = V[k * 4 + 0] * g_ln[0] + V[k * 4 + 1] * g_ln[1] + V[k * 4 + 2] * g_ln[2];  // ln = ng @ V3
      g_ng[k] -= g_u * p[k];                                                              // gd = -ng.p + og
      dx[k] = -g_u * t.ng[k];
    }
    float g_nh[3];
#pragma unroll
    for (int k = 0; k < 3; k++) g_nh[k] = t.sgn * g_ng[k];
    if (a.smallest_axis) {
      // nh[row] = [row == c] + two_s * A[row][c](y), two_s = 2 / (y.y):  d two_s / d y[m] = -two_s^2 y[m]
      float col[3], dcol[3][4];
      quat_column(y, c, col);
      quat_column_grad(y, c, dcol);
      const float ga = g_nh[0] * col[0] + g_nh[1] * col[1] + g_nh[2] * col[2];
#pragma unroll
      for (int m = 0; m < 4; m++)
        g_y_extra[m] = two_s * (g_nh[0] * dcol[0][m] + g_nh[1] * dcol[1][m] + g_nh[2] * dcol[2][m]) - two_s * two_s * y[m] * ga;
    } else {
      a.d_offset[i] = g_u * t.sgn;
      const float nhg = t.nh[0] * g_nh[0] + t.nh[1] * g_nh[1] + t.nh[2] * g_nh[2];
#pragma unroll
      for (int k = 0; k < 3; k++) a.d_normal_raw[3 * (size_t)i + k] = (g_nh[k] - t.nh[k] * nhg) * t.inv_len;
    }
  }
  {
    float4 g = a.g_rotations ? reinterpret_cast<const float4*>(a.g_rotations)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    g.x += g_y_extra[0]; g.y += g_y_extra[1]; g.z += g_y_extra[2]; g.w += g_y_extra[3];
    float4 d;
    if (len > 1e-12f) {  // y = r/len: d = (g - y (y.g)) / len
      const float inv = 1.0f / len;
      const float4 y = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
      const float yg = y.x * g.x + y.y * g.y + y.z * g.z + y.w * g.w;
      d = make_float4((g.x - y.x * yg) * inv, (g.y - y.y * yg) * inv, (g.z - y.z * yg) * inv, (g.w - y.w * yg) * inv);
    } else {             // clamped branch of F.normalize: y = r / eps
      d = make_float4(g.x * 1e12f, g.y * 1e12f, g.z * 1e12f, g.w * 1e12f);
    }
    reinterpret_cast<float4*>(a.d_rotation_raw)[i] = d;
  }
  if (a.d_xyz) {
#pragma unroll
    for (int k = 0; k < 3; k++) a.d_xyz[3 * (size_t)i + k] = dx[k];
  }
}

int fill(ProArgs& p, const IbgsPrologueArgs& a) {
  if (a.P < 0 || a.sh_rest < 0) { ibgs_set_error("P and sh_rest must be >= 0"); return IBGS_EINVAL; }
  if (!a.xyz || !a.opacity_raw || !a.scaling_raw || !a.rotation_raw) {
    ibgs_set_error("xyz / opacity_raw / scaling_raw / rotation_raw must not be NULL");
    return IBGS_EINVAL;
  }
  if ((a.normal_raw != nullptr) != (a.offset != nullptr) ||
      ((a.normal_raw || a.smallest_axis_normal) && (!a.world_view_transform || !a.camera_center))) {
    ibgs_set_error("normal_raw and offset come together and need world_view_transform + camera_center");
    return IBGS_EINVAL;
  }
  if (a.smallest_axis_normal && a.normal_raw) {
    ibgs_set_error("smallest_axis_normal excludes normal_raw / offset");
    return IBGS_EINVAL;
  }
  p.smallest_axis = a.smallest_axis_normal != 0;
  p.P = a.P;
  p.L = 3 * (a.sh_rest + 1);
  p.magic = (uint32_t)(0x100000000ull / (uint64_t)p.L) + 1u;
  p.magic_limit = (uint32_t)(0x100000000ull / (uint64_t)p.L);
  p.xyz = a.xyz; p.opacity_raw = a.opacity_raw; p.scaling_raw = a.scaling_raw; p.rotation_raw = a.rotation_raw;
  p.fdc = a.features_dc; p.frest = a.features_rest; p.normal_raw = a.normal_raw; p.offset = a.offset;
  p.view = a.world_view_transform; p.campos = a.camera_center;
  p.opacity = a.opacity; p.scales = a.scales; p.rotations = a.rotations; p.shs = a.shs; p.all_map = a.all_map;
  p.g_opacity = a.g_opacity; p.g_scales = a.g_scales; p.g_rotations = a.g_rotations; p.g_shs = a.g_shs;
  p.g_all_map = a.g_all_map;
  p.d_xyz = a.d_xyz; p.d_opacity_raw = a.d_opacity_raw; p.d_scaling_raw = a.d_scaling_raw;
  p.d_rotation_raw = a.d_rotation_raw; p.d_fdc = a.d_features_dc; p.d_frest = a.d_features_rest;
  p.d_normal_raw = a.d_normal_raw; p.d_offset = a.d_offset;
  return IBGS_OK;
}

int copy_grid(const ProArgs& p) {
  const size_t total = (size_t)p.P * p.L;
  size_t blocks = (total + 255) / 256;
  const size_t cap = 148 * 32;  // grid-stride beyond that: 32 resident CTAs' worth per SM is plenty for a copy
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

extern "C" int ibgs_prologue_forward(const IbgsPrologueArgs* a, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  if (!a) { ibgs_set_error("args is NULL"); return IBGS_EINVAL; }
  ProArgs p;
  int rc = fill(p, *a);
  if (rc != IBGS_OK) return rc;
  if (a->P == 0) return IBGS_OK;
  if (!a->opacity || !a->scales || !a->rotations || ((a->normal_raw || a->smallest_axis_normal) && !a->all_map)) {
    ibgs_set_error("output pointers must not be NULL");
    return IBGS_EINVAL;
  }
  if (!ibgs_aligned16(a->rotation_raw) || !ibgs_aligned16(a->rotations)) {
    ibgs_set_error("rotation_raw / rotations must be 16-byte aligned (they are accessed as float4)");
    return IBGS_EINVAL;
  }
  if (!a->normal_raw && !a->smallest_axis_normal) p.all_map = nullptr;
  prologue_forward_kernel<<<(a->P + 255) / 256, 256, 0, s>>>(p);
  KERNEL_CHECK(0, s);
  if (a->shs) {  // NULL: the caller hands features_dc / features_rest to the rasterizer in place (shs_rest)
    sh_rows_kernel<false><<<copy_grid(p), 256, 0, s>>>(p);
    KERNEL_CHECK(0, s);
  }
  return IBGS_OK;
}

extern "C" int ibgs_prologue_backward(const IbgsPrologueArgs* a, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  if (!a) { ibgs_set_error("args is NULL"); return IBGS_EINVAL; }
  ProArgs p;
  int rc = fill(p, *a);
  if (rc != IBGS_OK) return rc;
  if (a->P == 0) return IBGS_OK;
  const bool with_sh = a->d_features_dc != nullptr;  // NULL: SH gradients do not pass through the prologue
  if (!a->d_opacity_raw || !a->d_scaling_raw || !a->d_rotation_raw ||
      (with_sh && a->sh_rest > 0 && !a->d_features_rest) ||
      (a->normal_raw && (!a->d_normal_raw || !a->d_offset || !a->d_xyz)) || (a->smallest_axis_normal && !a->d_xyz)) {
    ibgs_set_error("gradient output pointers must not be NULL");
    return IBGS_EINVAL;
  }
  if (!ibgs_aligned16(a->rotation_raw) || !ibgs_aligned16(a->g_rotations) || !ibgs_aligned16(a->d_rotation_raw)) {
    ibgs_set_error("rotation_raw / g_rotations / d_rotation_raw must be 16-byte aligned (they are accessed as float4)");
    return IBGS_EINVAL;
  }
  prologue_backward_kernel<<<(a->P + 255) / 256, 256, 0, s>>>(p);
  KERNEL_CHECK(0, s);
  if (with_sh) {
    sh_rows_kernel<true><<<copy_grid(p), 256, 0, s>>>(p);
    KERNEL_CHECK(0, s);
  }
  return IBGS_OK;
}
