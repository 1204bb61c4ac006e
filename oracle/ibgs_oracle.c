/*
 * ibgs_oracle.c -- CPU restatement (float64 arithmetic, float32 I/O) of the IBGS planar Gaussian
 * rasterizer hot path.  TEST INFRASTRUCTURE ONLY: it is the checker for the CUDA path in tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg; nothing in ibgs_b200/ links or calls it.
 *
 * Pinning: the reference ships no tests or golden vectors for this path, so the
 * oracle is pinned against outputs of the UNMODIFIED reference CUDA extension (oracle/_ref, built from
 * /root/reference by oracle/build_ref.py) captured on a B200 into tests/golden/ by
 * tests/golden/make_golden.py; tests/test_oracle_golden.py checks it against them on the CPU.
 *
 * What each function follows in the reference (submodules/diff-plane-rasterization/cuda_rasterizer):
 *   o_preprocess      forward.cu:194-295 (+ :58-109 SH, :112-151 cov2D, :156-190 cov3D), auxiliary.h:45-60,143-168
 *   o_bin             rasterizer_impl.cu:187-255 (+ sort :452-457, getHigherMsb :152-167)
 *   o_render          forward.cu:303-665
 *   o_render_backward backward.cu:496-807 (+ :55-109 bilinear taps)
 *   o_preprocess_backward backward.cu:241-371, :375-438, :443-493, :116-235, auxiliary.h:111-121
 *   o_dist2           submodules/simple-knn/simple_knn.cu:147-183 (as exact brute force)
 * Texture taps emulate the CUDA linear filter: unnormalised coordinates, clamp addressing, interpolation
 * weights held in 1.8 fixed point (CUDA C Programming Guide, "Linear Filtering").
 *
 * Float64 makes different last-bit rounding decisions than the float32 GPU kernels, so comparisons
 * against CUDA are statistical (fraction of matching pixels / relative L2), never bit-exact; bit-exact
 * parity is asserted against oracle/_ref on the GPU box instead.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16
#define MAXS 5
#define MAXBL 8

typedef struct {
  int32_t P, D, M, W, H, nb_src, BL, render_geo, depth_only;
  double tanfovx, tanfovy, scale_modifier, thr;
  const float *bg, *means3D, *shs, *colors_precomp, *opacities, *scales, *rotations, *cov3D_precomp,
      *all_map, *view, *proj, *campos, *ref_to_src, *src_cam_pos, *src_images, *src_depths;
} OIn;

typedef struct { /* per-Gaussian state, float64 */
  int32_t* radii;
  float* depths;       /* float32: its bit pattern is the low half of the sort key */
  double* means2D;     /* [P,2] */
  double* conic_o;     /* [P,4] */
  double* rgb;         /* [P,3] */
  uint8_t* clamped;    /* [P,3] */
  double* cov3D;       /* [P,6] */
  uint32_t* tiles;     /* [P] */
  uint32_t* offsets;   /* [P] inclusive scan */
} OGeom;

typedef struct { /* per-pixel state */
  double* final_T;
  uint32_t* n_contrib;
  double* sum_w;
  uint32_t *low, *high;
  int32_t* valid_idx;  /* [5,N] */
  double* valid_w;     /* [5,N] */
} OImg;

typedef struct { /* outputs, float32 CHW, pre-zeroed by the caller */
  float *color, *normal, *depth, *cam_feat, *warped, *min_depth_diff, *camera_ray;
  int32_t* mask;
} OOut;

typedef struct { /* gradients, float64, pre-zeroed */
  double *means3D, *means2D, *means2D_abs, *colors, *opacity, *cov3D, *sh, *scales, *rots, *all_map, *conic;
} OGrad;

static const double F02 = (double)0.2f, F03 = (double)0.3f, F13 = (double)1.3f, F099 = (double)0.99f,
                    F255 = (double)(1.0f / 255.0f), F1E4 = (double)0.0001f, F1E7 = (double)0.0000001f,
                    F01 = (double)0.1f, EPS = (double)1.0e-8f;

/* ---- small helpers ---------------------------------------------------------------------------- */
static void xf43(const double* p, const float* m, double* o) {
  for (int r = 0; r < 3; r++) o[r] = m[r] * p[0] + m[4 + r] * p[1] + m[8 + r] * p[2] + m[12 + r];
}
static void xf44(const double* p, const float* m, double* o) {
  for (int r = 0; r < 4; r++) o[r] = m[r] * p[0] + m[4 + r] * p[1] + m[8 + r] * p[2] + m[12 + r];
}
/* 3x3 in glm storage m[col][row] */
typedef struct { double m[3][3]; } M3;
static M3 mm(const M3* A, const M3* B) {
  M3 r;
  for (int c = 0; c < 3; c++)
    for (int w = 0; w < 3; w++) r.m[c][w] = A->m[0][w] * B->m[c][0] + A->m[1][w] * B->m[c][1] + A->m[2][w] * B->m[c][2];
  return r;
}
static M3 mt(const M3* A) {
  M3 r;
  for (int c = 0; c < 3; c++)
    for (int w = 0; w < 3; w++) r.m[c][w] = A->m[w][c];
  return r;
}
static void rot_from_quat(const double* q, M3* R) {
  double r = q[0], x = q[1], y = q[2], z = q[3];
  double v[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                 2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                 2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)};
  for (int c = 0; c < 3; c++)
    for (int w = 0; w < 3; w++) R->m[c][w] = v[3 * c + w];
}
static void rect_of(double px, double py, int radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
  /* auxiliary.h:50-60: (int) truncates toward zero */
  int a = (int)((px - radius) / TILE), b = (int)((py - radius) / TILE);
  int c = (int)((px + radius + TILE - 1) / TILE), d = (int)((py + radius + TILE - 1) / TILE);
  *x0 = a < 0 ? 0 : (a > gx ? gx : a);
  *y0 = b < 0 ? 0 : (b > gy ? gy : b);
  *x1 = c < 0 ? 0 : (c > gx ? gx : c);
  *y1 = d < 0 ? 0 : (d > gy ? gy : d);
}

static const double C0 = 0.28209479177387814, C1 = 0.4886025119029199;
static const double C2[5] = {1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792,
                             0.5462742152960396};
static const double C3[7] = {-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                             -0.4570457994644658, 1.445305721320277, -0.5900435899266435};

/* SH basis values b[k] for direction (x,y,z); returns number of coefficients for degree deg */
static int sh_basis(int deg, double x, double y, double z, double* b) {
  b[0] = C0;
  if (deg < 1) return 1;
  b[1] = -C1 * y; b[2] = C1 * z; b[3] = -C1 * x;
  if (deg < 2) return 4;
  double xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  b[4] = C2[0] * xy; b[5] = C2[1] * yz; b[6] = C2[2] * (2 * zz - xx - yy); b[7] = C2[3] * xz; b[8] = C2[4] * (xx - yy);
  if (deg < 3) return 9;
  b[9] = C3[0] * y * (3 * xx - yy); b[10] = C3[1] * xy * z; b[11] = C3[2] * y * (4 * zz - xx - yy);
  b[12] = C3[3] * z * (2 * zz - 3 * xx - 3 * yy); b[13] = C3[4] * x * (4 * zz - xx - yy);
  b[14] = C3[5] * z * (xx - yy); b[15] = C3[6] * x * (xx - 3 * yy);
  return 16;
}

static void cov3d_of(const OIn* in, int i, double* cov, M3* Rout, M3* Mout) {
  if (in->cov3D_precomp) {
    for (int k = 0; k < 6; k++) cov[k] = in->cov3D_precomp[6 * (size_t)i + k];
    return;
  }
  double q[4] = {in->rotations[4 * (size_t)i], in->rotations[4 * (size_t)i + 1], in->rotations[4 * (size_t)i + 2],
                 in->rotations[4 * (size_t)i + 3]};
  M3 R, S, Mx, Mt, Sg;
  rot_from_quat(q, &R);
  memset(&S, 0, sizeof(S));
  for (int k = 0; k < 3; k++) S.m[k][k] = in->scale_modifier * in->scales[3 * (size_t)i + k];
  Mx = mm(&S, &R);
  Mt = mt(&Mx);
  Sg = mm(&Mt, &Mx);
  cov[0] = Sg.m[0][0]; cov[1] = Sg.m[0][1]; cov[2] = Sg.m[0][2]; cov[3] = Sg.m[1][1]; cov[4] = Sg.m[1][2]; cov[5] = Sg.m[2][2];
  if (Rout) *Rout = R;
  if (Mout) *Mout = Mx;
}

/* cov2D forward pieces shared with the backward: t (clamped), T, Vrk, cov (a,b,c after +0.3) */
static void cov2d_of(const OIn* in, const double* mean, const double* cov3D, double fx, double fy, double* t,
                     double* txtz, double* tytz, M3* T, M3* Vrk, M3* Wm, double* abc) {
  xf43(mean, in->view, t);
  double limx = F13 * in->tanfovx, limy = F13 * in->tanfovy;
  *txtz = t[0] / t[2];
  *tytz = t[1] / t[2];
  t[0] = fmin(limx, fmax(-limx, *txtz)) * t[2];
  t[1] = fmin(limy, fmax(-limy, *tytz)) * t[2];
  M3 J;
  memset(&J, 0, sizeof(J));
  J.m[0][0] = fx / t[2]; J.m[0][2] = -(fx * t[0]) / (t[2] * t[2]);
  J.m[1][1] = fy / t[2]; J.m[1][2] = -(fy * t[1]) / (t[2] * t[2]);
  const float* v = in->view;
  double wv[9] = {v[0], v[4], v[8], v[1], v[5], v[9], v[2], v[6], v[10]};
  for (int c = 0; c < 3; c++)
    for (int w = 0; w < 3; w++) Wm->m[c][w] = wv[3 * c + w];
  *T = mm(Wm, &J);
  double vv[9] = {cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]};
  for (int c = 0; c < 3; c++)
    for (int w = 0; w < 3; w++) Vrk->m[c][w] = vv[3 * c + w];
  M3 Tt = mt(T), Vt = mt(Vrk), A = mm(&Tt, &Vt), cov = mm(&A, T);
  abc[0] = cov.m[0][0] + F03;
  abc[1] = cov.m[0][1];
  abc[2] = cov.m[1][1] + F03;
}

/* ---- forward: per-Gaussian ---------------------------------------------------------------------- */
int64_t o_preprocess(const OIn* in, OGeom* g) {
  const int P = in->P, W = in->W, H = in->H;
  const double fy = H / (2.0 * in->tanfovy), fx = W / (2.0 * in->tanfovx);
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    g->radii[i] = 0;
    g->tiles[i] = 0;
    double p[3] = {in->means3D[3 * (size_t)i], in->means3D[3 * (size_t)i + 1], in->means3D[3 * (size_t)i + 2]};
    double pv[3], ph[4];
    xf43(p, in->view, pv);
    if ((float)pv[2] <= 0.2f) continue; /* the GPU decides on the float32 depth */
    xf44(p, in->proj, ph);
    double pw = 1.0 / (ph[3] + F1E7);
    double proj[3] = {ph[0] * pw, ph[1] * pw, ph[2] * pw};
    double* cov3 = g->cov3D + 6 * (size_t)i;
    cov3d_of(in, i, cov3, NULL, NULL);
    double t[3], a, b, abc[3];
    M3 T, Vrk, Wm;
    cov2d_of(in, p, cov3, fx, fy, t, &a, &b, &T, &Vrk, &Wm, abc);
    double det = abc[0] * abc[2] - abc[1] * abc[1];
    if (det == 0.0) continue;
    double di = 1.0 / det;
    double conic[3] = {abc[2] * di, -abc[1] * di, abc[0] * di};
    double mid = 0.5 * (abc[0] + abc[2]);
    double l1 = mid + sqrt(fmax(F01, mid * mid - det)), l2 = mid - sqrt(fmax(F01, mid * mid - det));
    double rad = ceil(3.0 * sqrt(fmax(l1, l2)));
    double px = ((proj[0] + 1.0) * W - 1.0) * 0.5, py = ((proj[1] + 1.0) * H - 1.0) * 0.5;
    int x0, y0, x1, y1;
    rect_of(px, py, (int)rad, gx, gy, &x0, &y0, &x1, &y1);
    if ((x1 - x0) * (y1 - y0) == 0) continue;
    if (in->colors_precomp) {
      for (int c = 0; c < 3; c++) g->rgb[3 * (size_t)i + c] = in->colors_precomp[3 * (size_t)i + c];
    } else if (!in->depth_only) {
      double d[3] = {p[0] - in->campos[0], p[1] - in->campos[1], p[2] - in->campos[2]};
      double len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      double bas[16];
      int nk = sh_basis(in->D, d[0] / len, d[1] / len, d[2] / len, bas);
      const float* sh = in->shs + (size_t)i * in->M * 3;
      for (int c = 0; c < 3; c++) {
        double r = 0;
        for (int k = 0; k < nk; k++) r += bas[k] * sh[3 * k + c];
        r += 0.5;
        g->clamped[3 * (size_t)i + c] = r < 0;
        g->rgb[3 * (size_t)i + c] = r < 0 ? 0 : r;
      }
    }
    g->depths[i] = (float)pv[2];
    g->radii[i] = (int)rad;
    g->means2D[2 * (size_t)i] = px;
    g->means2D[2 * (size_t)i + 1] = py;
    double* co = g->conic_o + 4 * (size_t)i;
    co[0] = conic[0]; co[1] = conic[1]; co[2] = conic[2]; co[3] = in->opacities[i];
    g->tiles[i] = (uint32_t)((y1 - y0) * (x1 - x0));
  }
  uint64_t run = 0;
  for (int i = 0; i < P; i++) { run += g->tiles[i]; g->offsets[i] = (uint32_t)run; }
  return (int64_t)run;
}

/* ---- binning ---------------------------------------------------------------------------------------- */
typedef struct { uint64_t key; uint32_t val; uint32_t seq; } KV;
static int kv_cmp(const void* a, const void* b) {
  const KV *x = (const KV*)a, *y = (const KV*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
void o_bin(const OIn* in, const OGeom* g, int64_t R, uint64_t* keys_unsorted, uint32_t* vals_unsorted,
           uint64_t* keys, uint32_t* point_list, uint32_t* ranges /* [T,2] zeroed */) {
  const int gx = (in->W + TILE - 1) / TILE, gy = (in->H + TILE - 1) / TILE;
  KV* kv = (KV*)malloc(sizeof(KV) * (size_t)(R > 0 ? R : 1));
  for (int i = 0; i < in->P; i++) {
    if (g->radii[i] <= 0) continue;
    uint32_t off = i == 0 ? 0 : g->offsets[i - 1];
    int x0, y0, x1, y1;
    rect_of(g->means2D[2 * (size_t)i], g->means2D[2 * (size_t)i + 1], g->radii[i], gx, gy, &x0, &y0, &x1, &y1);
    uint32_t bits;
    memcpy(&bits, &g->depths[i], 4);
    for (int y = y0; y < y1; y++)
      for (int x = x0; x < x1; x++) {
        uint64_t key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | bits;
        keys_unsorted[off] = key;
        vals_unsorted[off] = (uint32_t)i;
        kv[off].key = key; kv[off].val = (uint32_t)i; kv[off].seq = off;
        off++;
      }
  }
  qsort(kv, (size_t)R, sizeof(KV), kv_cmp); /* (key, emission order) == stable sort by key */
  for (int64_t k = 0; k < R; k++) { keys[k] = kv[k].key; point_list[k] = kv[k].val; }
  free(kv);
  for (int64_t k = 0; k < R; k++) {
    uint32_t cur = (uint32_t)(keys[k] >> 32);
    if (k == 0) ranges[2 * cur] = 0;
    else {
      uint32_t prev = (uint32_t)(keys[k - 1] >> 32);
      if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)k; ranges[2 * cur] = (uint32_t)k; }
    }
    if (k == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
  }
}

/* ---- texture emulation ------------------------------------------------------------------------------ */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static int g_tex_mode = 0; /* experiment knob for the weight quantisation, see o_set_tex_mode */
void o_set_tex_mode(int m) { g_tex_mode = m; }
static inline double q8(double f) {
  if (g_tex_mode == 1) return floor(f * 256.0) / 256.0;       /* truncate */
  if (g_tex_mode == 2) return f;                               /* exact (no quantisation) */
  return floor(f * 256.0 + 0.5) / 256.0;                       /* round to nearest (default) */
}
/* one linear-filter tap of plane `img` (W x H), unnormalised coords, clamp addressing */
static double tex_plane(const float* img, int W, int H, double x, double y) {
  if (g_tex_mode == 3) { x = floor(x * 256.0 + 0.5) / 256.0; y = floor(y * 256.0 + 0.5) / 256.0; }
  if (g_tex_mode == 4) { x = (double)(float)x; y = (double)(float)y; }
  double xb = x - 0.5, yb = y - 0.5;
  double fi = floor(xb), fj = floor(yb);
  double a = q8(xb - fi), b = q8(yb - fj);
  int i0 = clampi((int)fi, 0, W - 1), i1 = clampi((int)fi + 1, 0, W - 1);
  int j0 = clampi((int)fj, 0, H - 1), j1 = clampi((int)fj + 1, 0, H - 1);
  double t00 = img[(size_t)j0 * W + i0], t10 = img[(size_t)j0 * W + i1], t01 = img[(size_t)j1 * W + i0],
         t11 = img[(size_t)j1 * W + i1];
  return (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
}
static void tex_color(const OIn* in, int s, double x, double y, double* rgb) {
  size_t plane = (size_t)in->W * in->H;
  for (int c = 0; c < 3; c++) rgb[c] = tex_plane(in->src_images + ((size_t)s * 3 + c) * plane, in->W, in->H, x, y);
}

/* ---- forward: per-pixel ---------------------------------------------------------------------------- */
void o_render(const OIn* in, const OGeom* g, const uint32_t* point_list, const uint32_t* ranges, OImg* im,
              OOut* out) {
  const int W = in->W, H = in->H, BL = in->BL, ns = in->nb_src;
  const size_t N = (size_t)W * H;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const double fy = (double)(float)(H / (2.0f * (float)in->tanfovy)), fx = (double)(float)(W / (2.0f * (float)in->tanfovx));
  const double cx = W * 0.5, cy = H * 0.5;
  const int before_cap = (BL + 1) / 2, below_cap = BL - before_cap;
  const int geo = in->render_geo, donly = in->depth_only;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int ty = 0; ty < gy; ty++)
    for (int tx = 0; tx < gx; tx++) {
      const uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
      const int n = (int)(r1 - r0);
      for (int ly = 0; ly < TILE; ly++)
        for (int lx = 0; lx < TILE; lx++) {
          const int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
          if (pxi >= W || pyi >= H) continue;
          const size_t pid = (size_t)pyi * W + pxi;
          const double pxf = pxi, pyf = pyi;
          const double rayx = (pxf - cx) / fx, rayy = (pyf - cy) / fy;
          double T = 1.0, C[3] = {0, 0, 0}, Nn[3] = {0, 0, 0};
          double zb[MAXBL] = {0}, wb[MAXBL] = {0};
          uint32_t cb[MAXBL] = {0};
          uint32_t contributor = 0, last = 0;
          int before_ptr = 0, below = 0, done = 0;
          double tw = 0, ws = 0;
          for (int base = 0; base < n && !done; base += 256) {
            const int cnt = n - base < 256 ? n - base : 256;
            for (int j = 0; j < cnt && !done; j++) {
              contributor++;
              const uint32_t id = point_list[r0 + base + j];
              const double* co = g->conic_o + 4 * (size_t)id;
              const double dx = g->means2D[2 * (size_t)id] - pxf, dy = g->means2D[2 * (size_t)id + 1] - pyf;
              const double power = -0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
              if (power > 0) continue;
              const double alpha = fmin(F099, co[3] * exp(power));
              if (alpha < F255) continue;
              const double test_T = T * (1 - alpha);
              if (test_T < F1E4) { done = 1; continue; }
              const double aT = alpha * T;
              if (!donly)
                for (int c = 0; c < 3; c++) C[c] += g->rgb[3 * (size_t)id + c] * aT;
              double z = 0;
              if (geo || donly) {
                const float* am = in->all_map + 5 * (size_t)id;
                z = -(double)am[4] / (am[0] * rayx + am[1] * rayy + am[2] + EPS);
              }
              if (geo) {
                const float* am = in->all_map + 5 * (size_t)id;
                for (int c = 0; c < 3; c++) Nn[c] += am[c] * aT;
                if (z > 0) {
                  if (T > 0.5) {
                    zb[before_ptr] = z; wb[before_ptr] = aT; cb[before_ptr] = contributor;
                    before_ptr = (before_ptr + 1) % before_cap;
                  } else if (below < below_cap) {
                    int k = before_cap + below;
                    zb[k] = z; wb[k] = aT; cb[k] = contributor;
                    below++;
                  }
                }
              }
              if (donly && z > 0) {
                if (T > 0.5) {
                  int k = before_ptr;
                  tw -= wb[k]; ws -= wb[k] * zb[k];
                  zb[k] = z; wb[k] = aT;
                  before_ptr = (before_ptr + 1) % before_cap;
                  tw += aT; ws += aT * z;
                } else if (below < below_cap) {
                  int k = before_cap + below;
                  zb[k] = z; wb[k] = aT;
                  below++;
                  tw += aT; ws += aT * z;
                }
                if (below == below_cap) { T = test_T; last = contributor; break; }
              }
              T = test_T;
              last = contributor;
            }
          }
          im->final_T[pid] = T;
          im->n_contrib[pid] = last;
          if (!donly)
            for (int c = 0; c < 3; c++) out->color[c * N + pid] = (float)(C[c] + T * in->bg[c]);
          if (donly) out->depth[pid] = (float)(ws / (tw + EPS));
          if (!geo) continue;

          /* epilogue, forward.cu:512-663 */
          const double ifx = 1.0 / fx, ify = 1.0 / fy, pdx = pxf - cx, pdy = pyf - cy;
          double med = 0, twl = 0, tws[MAXS] = {0}, wc[MAXS * 3] = {0};
          uint32_t lo = cb[0], hi = cb[0];
          for (int i = 0; i < BL; i++) {
            const double w = wb[i];
            if (w == 0) continue;
            const double zz = zb[i];
            const double ip[3] = {pdx * zz * ifx, pdy * zz * ify, zz};
            for (int s = 0; s < ns; s++) {
              const float* m = in->ref_to_src + 16 * s;
              const double qx = m[0] * ip[0] + m[1] * ip[1] + m[2] * ip[2] + m[3];
              const double qy = m[4] * ip[0] + m[5] * ip[1] + m[6] * ip[2] + m[7];
              const double qz = m[8] * ip[0] + m[9] * ip[1] + m[10] * ip[2] + m[11];
              const double iz = 1.0 / (qz + EPS);
              const double u = qx * fx * iz + cx, v = qy * fy * iz + cy;
              if (u >= 0 && u <= W - 1 && v >= 0 && v <= H - 1) {
                double rgb[3];
                tex_color(in, s, u + 0.5, v + 0.5, rgb);
                for (int c = 0; c < 3; c++) wc[3 * s + c] += w * rgb[c];
                tws[s] += w;
              }
            }
            twl += w;
            med += w * zz;
            if (cb[i] < lo) lo = cb[i];
            if (cb[i] > hi) hi = cb[i];
          }
          im->low[pid] = lo; im->high[pid] = hi; im->sum_w[pid] = twl;
          med /= (twl + EPS);
          const double mp[3] = {pdx * med * ifx, pdy * med * ify, med};
          const float* vm = in->view;
          const double pc[3] = {mp[0] - vm[12], mp[1] - vm[13], mp[2] - vm[14]};
          const double mw[3] = {vm[0] * pc[0] + vm[1] * pc[1] + vm[2] * pc[2], vm[4] * pc[0] + vm[5] * pc[1] + vm[6] * pc[2],
                                vm[8] * pc[0] + vm[9] * pc[1] + vm[10] * pc[2]};
          double rd[3] = {mw[0] - in->campos[0], mw[1] - in->campos[1], mw[2] - in->campos[2]};
          const double rl = sqrt(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]) + EPS;
          for (int c = 0; c < 3; c++) { rd[c] /= rl; out->camera_ray[c * N + pid] = (float)rd[c]; }
          int vc = 0;
          double mind = 1.0;
          for (int s = 0; s < ns; s++) {
            const float* m = in->ref_to_src + 16 * s;
            const double qx = m[0] * mp[0] + m[1] * mp[1] + m[2] * mp[2] + m[3];
            const double qy = m[4] * mp[0] + m[5] * mp[1] + m[6] * mp[2] + m[7];
            const double qz = m[8] * mp[0] + m[9] * mp[1] + m[10] * mp[2] + m[11];
            const double iz = 1.0 / (qz + EPS);
            const double u = qx * fx * iz + cx, v = qy * fy * iz + cy;
            double wd = 0;
            if (u >= 0 && u <= W - 1 && v >= 0 && v <= H - 1)
              wd = tex_plane(in->src_depths + (size_t)s * N, W, H, u + 0.5, v + 0.5);
            const double err = fabs(wd - qz) * iz;
            if (wd > 0 && err < in->thr) {
              const double iw = 1.0 / (tws[s] + EPS);
              for (int c = 0; c < 3; c++) {
                wc[3 * s + c] *= iw;
                out->cam_feat[((size_t)vc * 4 + c) * N + pid] = (float)(in->campos[c] - in->src_cam_pos[3 * s + c]);
                out->warped[((size_t)vc * 3 + c) * N + pid] = (float)wc[3 * s + c];
              }
              double sd[3] = {mw[0] - in->src_cam_pos[3 * s], mw[1] - in->src_cam_pos[3 * s + 1], mw[2] - in->src_cam_pos[3 * s + 2]};
              const double sl = sqrt(sd[0] * sd[0] + sd[1] * sd[1] + sd[2] * sd[2]) + EPS;
              out->cam_feat[((size_t)vc * 4 + 3) * N + pid] = (float)((sd[0] * rd[0] + sd[1] * rd[1] + sd[2] * rd[2]) / sl);
              if (s == 0) out->mask[pid] = 1;
              im->valid_idx[(size_t)vc * N + pid] = s;
              im->valid_w[(size_t)vc * N + pid] = tws[s];
              vc++;
              if (err < mind) mind = err;
            }
          }
          if (vc <= MAXS - 1) im->valid_idx[(size_t)vc * N + pid] = -1;
          out->min_depth_diff[pid] = (float)mind;
          out->depth[pid] = (float)med;
          for (int c = 0; c < 3; c++) out->normal[c * N + pid] = (float)Nn[c];
        }
    }
}

/* ---- backward: per-pixel -------------------------------------------------------------------------- */
static inline void addd(double* p, double v) {
#pragma omp atomic
  *p += v;
}
void o_render_backward(const OIn* in, const OGeom* g, const uint32_t* point_list, const uint32_t* ranges,
                       const OImg* im, const float* depth_pix, const float* warped_pix, const float* dL_dcolor,
                       const float* dL_dnormal, const float* dL_ddepth, const float* dL_dwarped, OGrad* gr) {
  const int W = in->W, H = in->H;
  const size_t N = (size_t)W * H;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const double fy = (double)(float)(H / (2.0f * (float)in->tanfovy)), fx = (double)(float)(W / (2.0f * (float)in->tanfovx));
  const double cx = W * 0.5, cy = H * 0.5;
  const int geo = in->render_geo;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int ty = 0; ty < gy; ty++)
    for (int tx = 0; tx < gx; tx++) {
      const uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
      const int n = (int)(r1 - r0);
      for (int ly = 0; ly < TILE; ly++)
        for (int lx = 0; lx < TILE; lx++) {
          const int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
          if (pxi >= W || pyi >= H) continue;
          const size_t pid = (size_t)pyi * W + pxi;
          const double pxf = pxi, pyf = pyi;
          const double rayx = (pxf - cx) / fx, rayy = (pyf - cy) / fy;
          const double T_final = im->final_T[pid];
          double T = T_final;
          const uint32_t last_contributor = im->n_contrib[pid];
          const int lo = geo ? (int)im->low[pid] : 0, hi = geo ? (int)im->high[pid] : 0;
          double acc[3] = {0, 0, 0}, accn[3] = {0, 0, 0}, lastc[3] = {0, 0, 0}, lastn[3] = {0, 0, 0}, last_alpha = 0;
          double dpx[3], dn[3] = {0, 0, 0}, dd = 0;
          for (int c = 0; c < 3; c++) dpx[c] = dL_dcolor[c * N + pid];
          if (geo) {
            for (int c = 0; c < 3; c++) dn[c] = dL_dnormal[c * N + pid];
            dd = dL_ddepth[pid];
          }
          double bgdot = 0;
          for (int c = 0; c < 3; c++) bgdot += in->bg[c] * dpx[c];
          for (int k = n - 1; k >= 0; k--) {
            const uint32_t contributor = (uint32_t)k;
            if (contributor >= last_contributor) continue;
            const uint32_t id = point_list[r0 + k];
            const double* co = g->conic_o + 4 * (size_t)id;
            const double dx = g->means2D[2 * (size_t)id] - pxf, dy = g->means2D[2 * (size_t)id + 1] - pyf;
            const double power = -0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
            if (power > 0) continue;
            const double G = exp(power);
            const double alpha = fmin(F099, co[3] * G);
            if (alpha < F255) continue;
            T = T / (1 - alpha);
            const double w = alpha * T;
            double dalpha = 0;
            for (int c = 0; c < 3; c++) {
              const double col = g->rgb[3 * (size_t)id + c];
              acc[c] = last_alpha * lastc[c] + (1 - last_alpha) * acc[c];
              lastc[c] = col;
              dalpha += (col - acc[c]) * dpx[c];
              addd(&gr->colors[3 * (size_t)id + c], w * dpx[c]);
            }
            if (geo) {
              const float* am = in->all_map + 5 * (size_t)id;
              double tmp[5] = {0, 0, 0, 0, 0};
              for (int c = 0; c < 3; c++) {
                accn[c] = last_alpha * lastn[c] + (1 - last_alpha) * accn[c];
                lastn[c] = am[c];
                dalpha += (am[c] - accn[c]) * dn[c];
                tmp[c] += w * dn[c];
              }
              /* unsigned comparison against (int - 1), as in backward.cu:693 */
              if (contributor >= (uint32_t)(lo - 1) && contributor <= (uint32_t)(hi - 1)) {
                const double tg = am[0] * rayx + am[1] * rayy + am[2] + 1.0e-8;
                const double tg2 = am[4] / (tg * tg);
                const double z = -(double)am[4] / tg;
                if (z > 0) {
                  const double ip[3] = {(pxf - cx) * z / fx, (pyf - cy) * z / fy, z};
                  const double sw = im->sum_w[pid];
                  double dz = dd * w / sw;
                  dalpha += dd * (z - depth_pix[pid]) / sw;
                  for (int m = 0; m < MAXS; m++) {
                    const int s = im->valid_idx[(size_t)m * N + pid];
                    if (s == -1) break;
                    const float* r = in->ref_to_src + 16 * s;
                    const double qx = r[0] * ip[0] + r[1] * ip[1] + r[2] * ip[2] + r[3];
                    const double qy = r[4] * ip[0] + r[5] * ip[1] + r[6] * ip[2] + r[7];
                    const double qz = r[8] * ip[0] + r[9] * ip[1] + r[10] * ip[2] + r[11];
                    const double u = qx * fx / qz + cx, v = qy * fy / qz + cy;
                    if (!(u >= 0 && u <= W - 1 && v >= 0 && v <= H - 1)) continue;
                    double rgb[3], dLc[3];
                    tex_color(in, s, u + 0.5, v + 0.5, rgb);
                    const double vw = im->valid_w[(size_t)m * N + pid];
                    for (int c = 0; c < 3; c++) {
                      const double dLw = dL_dwarped[((size_t)m * 3 + c) * N + pid];
                      dLc[c] = dLw * w / vw;
                      dalpha += dLw * (rgb[c] - warped_pix[((size_t)m * 3 + c) * N + pid]) / vw;
                    }
                    const double A = (pxf - cx) / fx, B = (pyf - cy) / fy;
                    const double U = r[0] * A + r[1] * B + r[2], V = r[4] * A + r[5] * B + r[6], Wc = r[8] * A + r[9] * B + r[10];
                    const double den = Wc * z + r[11];
                    const double dpxd = fx * (U * r[11] - Wc * r[3]) / (den * den);
                    const double dpyd = fy * (V * r[11] - Wc * r[7]) / (den * den);
                    /* bilinearInterpolateBackward, backward.cu:55-109: four LINEAR-FILTER taps at integer coords */
                    const double uu = u + 0.5, vv = v + 0.5;
                    const double u0 = floor(uu), v0 = floor(vv), fu = uu - u0, fv = vv - v0;
                    double c00[3], c01[3], c10[3], c11[3];
                    tex_color(in, s, u0, v0, c00);
                    tex_color(in, s, u0 + 1, v0, c01);
                    tex_color(in, s, u0, v0 + 1, c10);
                    tex_color(in, s, u0 + 1, v0 + 1, c11);
                    double du = 0, dv = 0;
                    for (int c = 0; c < 3; c++) {
                      du += dLc[c] * (-(1 - fv) * c00[c] + (1 - fv) * c01[c] - fv * c10[c] + fv * c11[c]);
                      dv += dLc[c] * (-(1 - fu) * c00[c] - fu * c01[c] + (1 - fu) * c10[c] + fu * c11[c]);
                    }
                    dz += du * dpxd + dv * dpyd;
                    /* accumulated INSIDE the view loop (backward.cu:757-763) */
                    tmp[4] += -dz / tg;
                    tmp[0] += dz * tg2 * rayx;
                    tmp[1] += dz * tg2 * rayy;
                    tmp[2] += dz * tg2;
                  }
                }
              }
              for (int c = 0; c < 5; c++)
                if (tmp[c] != 0) addd(&gr->all_map[5 * (size_t)id + c], tmp[c]);
            }
            dalpha *= T;
            last_alpha = alpha;
            dalpha += (-T_final / (1 - alpha)) * bgdot;
            const double dG = co[3] * dalpha;
            const double gdx = G * dx, gdy = G * dy;
            const double dGx = -gdx * co[0] - gdy * co[1], dGy = -gdy * co[2] - gdx * co[1];
            const double m2x = dG * dGx * (0.5 * W), m2y = dG * dGy * (0.5 * H);
            addd(&gr->means2D[3 * (size_t)id], m2x);
            addd(&gr->means2D[3 * (size_t)id + 1], m2y);
            addd(&gr->means2D_abs[3 * (size_t)id], fabs(m2x));
            addd(&gr->means2D_abs[3 * (size_t)id + 1], fabs(m2y));
            addd(&gr->conic[4 * (size_t)id], -0.5 * gdx * dx * dG);
            addd(&gr->conic[4 * (size_t)id + 1], -0.5 * gdx * dy * dG);
            addd(&gr->conic[4 * (size_t)id + 3], -0.5 * gdy * dy * dG);
            addd(&gr->opacity[id], G * dalpha);
          }
        }
    }
}

/* ---- backward: per-Gaussian ---------------------------------------------------------------------- */
void o_preprocess_backward(const OIn* in, const OGeom* g, OGrad* gr) {
  const int P = in->P, W = in->W, H = in->H;
  const double hy = (double)(float)(H / (2.0f * (float)in->tanfovy)), hx = (double)(float)(W / (2.0f * (float)in->tanfovx));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (!(g->radii[i] > 0)) continue;
    const double mean[3] = {in->means3D[3 * (size_t)i], in->means3D[3 * (size_t)i + 1], in->means3D[3 * (size_t)i + 2]};
    double cov3[6];
    M3 R, Mx;
    cov3d_of(in, i, cov3, &R, &Mx);
    double t[3], txtz, tytz, abc[3];
    M3 T, Vrk, Wm;
    cov2d_of(in, mean, cov3, hx, hy, t, &txtz, &tytz, &T, &Vrk, &Wm, abc);
    const double limx = F13 * in->tanfovx, limy = F13 * in->tanfovy;
    const double xg = (txtz < -limx || txtz > limx) ? 0 : 1, yg = (tytz < -limy || tytz > limy) ? 0 : 1;
    const double a = abc[0], b = abc[1], c = abc[2];
    const double dcx = gr->conic[4 * (size_t)i], dcy = gr->conic[4 * (size_t)i + 1], dcz = gr->conic[4 * (size_t)i + 3];
    const double denom = a * c - b * b;
    const double d2i = 1.0 / (denom * denom + F1E7);
    double da = 0, db = 0, dc = 0, dcov[6] = {0, 0, 0, 0, 0, 0};
    if (d2i != 0) {
      da = d2i * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
      dc = d2i * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
      db = d2i * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
#define TT(cc, rr) T.m[cc][rr]
      dcov[0] = TT(0, 0) * TT(0, 0) * da + TT(0, 0) * TT(1, 0) * db + TT(1, 0) * TT(1, 0) * dc;
      dcov[3] = TT(0, 1) * TT(0, 1) * da + TT(0, 1) * TT(1, 1) * db + TT(1, 1) * TT(1, 1) * dc;
      dcov[5] = TT(0, 2) * TT(0, 2) * da + TT(0, 2) * TT(1, 2) * db + TT(1, 2) * TT(1, 2) * dc;
      dcov[1] = 2 * TT(0, 0) * TT(0, 1) * da + (TT(0, 0) * TT(1, 1) + TT(0, 1) * TT(1, 0)) * db + 2 * TT(1, 0) * TT(1, 1) * dc;
      dcov[2] = 2 * TT(0, 0) * TT(0, 2) * da + (TT(0, 0) * TT(1, 2) + TT(0, 2) * TT(1, 0)) * db + 2 * TT(1, 0) * TT(1, 2) * dc;
      dcov[4] = 2 * TT(0, 2) * TT(0, 1) * da + (TT(0, 1) * TT(1, 2) + TT(0, 2) * TT(1, 1)) * db + 2 * TT(1, 1) * TT(1, 2) * dc;
    }
    for (int k = 0; k < 6; k++) gr->cov3D[6 * (size_t)i + k] = dcov[k];
    double dT[2][3];
    for (int k = 0; k < 3; k++) {
      const double s0 = TT(0, 0) * Vrk.m[k][0] + TT(0, 1) * Vrk.m[k][1] + TT(0, 2) * Vrk.m[k][2];
      const double s1 = TT(1, 0) * Vrk.m[k][0] + TT(1, 1) * Vrk.m[k][1] + TT(1, 2) * Vrk.m[k][2];
      dT[0][k] = 2 * s0 * da + s1 * db;
      dT[1][k] = 2 * s1 * dc + s0 * db;
    }
    const double dJ00 = Wm.m[0][0] * dT[0][0] + Wm.m[0][1] * dT[0][1] + Wm.m[0][2] * dT[0][2];
    const double dJ02 = Wm.m[2][0] * dT[0][0] + Wm.m[2][1] * dT[0][1] + Wm.m[2][2] * dT[0][2];
    const double dJ11 = Wm.m[1][0] * dT[1][0] + Wm.m[1][1] * dT[1][1] + Wm.m[1][2] * dT[1][2];
    const double dJ12 = Wm.m[2][0] * dT[1][0] + Wm.m[2][1] * dT[1][1] + Wm.m[2][2] * dT[1][2];
    const double tz = 1.0 / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const double dtx = xg * -hx * tz2 * dJ02, dty = yg * -hy * tz2 * dJ12;
    const double dtz = -hx * tz2 * dJ00 - hy * tz2 * dJ11 + (2 * hx * t[0]) * tz3 * dJ02 + (2 * hy * t[1]) * tz3 * dJ12;
    const float* v = in->view;
    double dm[3] = {v[0] * dtx + v[1] * dty + v[2] * dtz, v[4] * dtx + v[5] * dty + v[6] * dtz, v[8] * dtx + v[9] * dty + v[10] * dtz};
    /* projection part, backward.cu:467-484 */
    const float* pr = in->proj;
    double mh[4];
    xf44(mean, pr, mh);
    const double mw = 1.0 / (mh[3] + F1E7);
    const double mul1 = (pr[0] * mean[0] + pr[4] * mean[1] + pr[8] * mean[2] + pr[12]) * mw * mw;
    const double mul2 = (pr[1] * mean[0] + pr[5] * mean[1] + pr[9] * mean[2] + pr[13]) * mw * mw;
    const double g2x = gr->means2D[3 * (size_t)i], g2y = gr->means2D[3 * (size_t)i + 1];
    dm[0] += (pr[0] * mw - pr[3] * mul1) * g2x + (pr[1] * mw - pr[3] * mul2) * g2y;
    dm[1] += (pr[4] * mw - pr[7] * mul1) * g2x + (pr[5] * mw - pr[7] * mul2) * g2y;
    dm[2] += (pr[8] * mw - pr[11] * mul1) * g2x + (pr[9] * mw - pr[11] * mul2) * g2y;
    /* SH part, backward.cu:116-235 (basis derivatives by central finite difference-free analytic forms) */
    if (in->shs) {
      const double d0[3] = {mean[0] - in->campos[0], mean[1] - in->campos[1], mean[2] - in->campos[2]};
      const double len = sqrt(d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2]);
      const double x = d0[0] / len, y = d0[1] / len, z = d0[2] / len;
      double bas[16];
      const int nk = sh_basis(in->D, x, y, z, bas);
      const float* sh = in->shs + (size_t)i * in->M * 3;
      double dRGB[3];
      for (int c2 = 0; c2 < 3; c2++) dRGB[c2] = g->clamped[3 * (size_t)i + c2] ? 0 : gr->colors[3 * (size_t)i + c2];
      for (int k = 0; k < nk; k++)
        for (int c2 = 0; c2 < 3; c2++) gr->sh[((size_t)i * in->M + k) * 3 + c2] = bas[k] * dRGB[c2];
      /* d basis / d (x,y,z) */
      double bx[16] = {0}, byy[16] = {0}, bz[16] = {0};
      if (in->D > 0) { bx[3] = -C1; byy[1] = -C1; bz[2] = C1; }
      if (in->D > 1) {
        bx[4] = C2[0] * y; bx[6] = C2[2] * 2 * -x; bx[7] = C2[3] * z; bx[8] = C2[4] * 2 * x;
        byy[4] = C2[0] * x; byy[5] = C2[1] * z; byy[6] = C2[2] * 2 * -y; byy[8] = C2[4] * 2 * -y;
        bz[5] = C2[1] * y; bz[6] = C2[2] * 4 * z; bz[7] = C2[3] * x;
      }
      if (in->D > 2) {
        const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        bx[9] = C3[0] * 6 * xy; bx[10] = C3[1] * yz; bx[11] = C3[2] * -2 * xy; bx[12] = C3[3] * -6 * xz;
        bx[13] = C3[4] * (-3 * xx + 4 * zz - yy); bx[14] = C3[5] * 2 * xz; bx[15] = C3[6] * 3 * (xx - yy);
        byy[9] = C3[0] * 3 * (xx - yy); byy[10] = C3[1] * xz; byy[11] = C3[2] * (-3 * yy + 4 * zz - xx);
        byy[12] = C3[3] * -6 * yz; byy[13] = C3[4] * -2 * xy; byy[14] = C3[5] * -2 * yz; byy[15] = C3[6] * -6 * xy;
        bz[10] = C3[1] * xy; bz[11] = C3[2] * 8 * yz; bz[12] = C3[3] * 3 * (2 * zz - xx - yy);
        bz[13] = C3[4] * 8 * xz; bz[14] = C3[5] * (xx - yy);
      }
      double dd[3] = {0, 0, 0};
      for (int k = 0; k < nk; k++)
        for (int c2 = 0; c2 < 3; c2++) {
          dd[0] += bx[k] * sh[3 * k + c2] * dRGB[c2];
          dd[1] += byy[k] * sh[3 * k + c2] * dRGB[c2];
          dd[2] += bz[k] * sh[3 * k + c2] * dRGB[c2];
        }
      /* dnormvdv, auxiliary.h:111-121 */
      const double s2 = d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2], inv = 1.0 / sqrt(s2 * s2 * s2);
      dm[0] += ((s2 - d0[0] * d0[0]) * dd[0] - d0[1] * d0[0] * dd[1] - d0[2] * d0[0] * dd[2]) * inv;
      dm[1] += (-d0[0] * d0[1] * dd[0] + (s2 - d0[1] * d0[1]) * dd[1] - d0[2] * d0[1] * dd[2]) * inv;
      dm[2] += (-d0[0] * d0[2] * dd[0] - d0[1] * d0[2] * dd[1] + (s2 - d0[2] * d0[2]) * dd[2]) * inv;
    }
    for (int k = 0; k < 3; k++) gr->means3D[3 * (size_t)i + k] = dm[k];
    /* cov3D -> scale / rotation, backward.cu:375-438 */
    if (in->scales && !in->cov3D_precomp) {
      const double q[4] = {in->rotations[4 * (size_t)i], in->rotations[4 * (size_t)i + 1], in->rotations[4 * (size_t)i + 2],
                           in->rotations[4 * (size_t)i + 3]};
      const double r = q[0], x = q[1], y = q[2], z = q[3];
      const double s[3] = {in->scale_modifier * in->scales[3 * (size_t)i], in->scale_modifier * in->scales[3 * (size_t)i + 1],
                           in->scale_modifier * in->scales[3 * (size_t)i + 2]};
      M3 dS, M2, dM, Rt, dMt;
      const double ds[9] = {dcov[0], 0.5 * dcov[1], 0.5 * dcov[2], 0.5 * dcov[1], dcov[3], 0.5 * dcov[4], 0.5 * dcov[2], 0.5 * dcov[4], dcov[5]};
      for (int cc = 0; cc < 3; cc++)
        for (int w = 0; w < 3; w++) { dS.m[cc][w] = ds[3 * cc + w]; M2.m[cc][w] = 2.0 * Mx.m[cc][w]; }
      dM = mm(&M2, &dS);
      Rt = mt(&R);
      dMt = mt(&dM);
      for (int k = 0; k < 3; k++)
        gr->scales[3 * (size_t)i + k] = Rt.m[k][0] * dMt.m[k][0] + Rt.m[k][1] * dMt.m[k][1] + Rt.m[k][2] * dMt.m[k][2];
      for (int k = 0; k < 3; k++)
        for (int w = 0; w < 3; w++) dMt.m[k][w] *= s[k];
#define D(cc, rr) dMt.m[cc][rr]
      gr->rots[4 * (size_t)i + 0] = 2 * z * (D(0, 1) - D(1, 0)) + 2 * y * (D(2, 0) - D(0, 2)) + 2 * x * (D(1, 2) - D(2, 1));
      gr->rots[4 * (size_t)i + 1] = 2 * y * (D(1, 0) + D(0, 1)) + 2 * z * (D(2, 0) + D(0, 2)) + 2 * r * (D(1, 2) - D(2, 1)) - 4 * x * (D(2, 2) + D(1, 1));
      gr->rots[4 * (size_t)i + 2] = 2 * x * (D(1, 0) + D(0, 1)) + 2 * r * (D(2, 0) - D(0, 2)) + 2 * z * (D(1, 2) + D(2, 1)) - 4 * y * (D(2, 2) + D(0, 0));
      gr->rots[4 * (size_t)i + 3] = 2 * r * (D(0, 1) - D(1, 0)) + 2 * x * (D(2, 0) + D(0, 2)) + 2 * y * (D(1, 2) + D(2, 1)) - 4 * z * (D(1, 1) + D(0, 0));
    }
  }
}

/* ---- distCUDA2 as exact brute force ------------------------------------------------------------------- */
void o_dist2(int P, const float* pts, float* out) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    float best[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    const float px = pts[3 * (size_t)i], py = pts[3 * (size_t)i + 1], pz = pts[3 * (size_t)i + 2];
    for (int j = 0; j < P; j++) {
      if (j == i) continue;
      const float dx = pts[3 * (size_t)j] - px, dy = pts[3 * (size_t)j + 1] - py, dz = pts[3 * (size_t)j + 2] - pz;
      float d = dx * dx + dy * dy + dz * dz;
      for (int k = 0; k < 3; k++)
        if (best[k] > d) { float tt = best[k]; best[k] = d; d = tt; }
    }
    out[i] = (best[0] + best[1] + best[2]) / 3.0f;
  }
}

int o_mark_visible(int P, const float* means3D, const float* view, uint8_t* present) {
  for (int i = 0; i < P; i++) {
    double p[3] = {means3D[3 * (size_t)i], means3D[3 * (size_t)i + 1], means3D[3 * (size_t)i + 2]}, pv[3];
    xf43(p, view, pv);
    present[i] = !((float)pv[2] <= 0.2f);
  }
  return 0;
}
