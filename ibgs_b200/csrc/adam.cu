// adam.cu -- one-launch Adam step over a flat parameter arena.
//
// Reference behaviour: gaussians.optimizer.step() (train.py:422) = torch.optim.Adam(l, lr=0.0, eps=1e-15) over eight
// parameter groups with their own learning rates (scene/gaussian_model.py:227-240; xyz / offset rates rescheduled every
// iteration, :251-262), followed by zero_grad (train.py:424).  torch's default CUDA path runs ~10 multi-tensor kernels
// per step, each a full pass over parameters / gradients / moments.
//
// Here parameters, gradients (the same flat arena the data-parallel all-reduce uses, ibgs_b200/parallel.py) and both
// moments live in four flat float arrays with one layout; a step is ONE launch that reads p, g, m, v once and writes
// p, m, v once (28 bytes per parameter; 32 with the fused zero_grad), which is the HBM floor of the update.
// Arithmetic follows torch/optim/adam.py (_single_tensor_adam, amsgrad=False, weight_decay=0, maximize=False):
//   m += (g - m)(1 - beta1);  v = beta2 v + (1 - beta2) g g;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
#include "common.cuh"

namespace {

struct AdamKArgs {
  float* p;
  float* g;
  float* m;
  float* v;
  int num_groups;
  long long begin[IBGS_ADAM_MAX_GROUPS];
  long long end[IBGS_ADAM_MAX_GROUPS];
  float step_size[IBGS_ADAM_MAX_GROUPS];   // lr / bias_correction1
  long long lo, hi;                         // covered index range
  float one_m_beta1, beta2, one_m_beta2, inv_bc2_sqrt, eps, grad_scale;
  int zero_grads;
};

constexpr int ADAM_UNROLL = 4;

__global__ void __launch_bounds__(256) adam_step_kernel(const AdamKArgs a) {
  const long long base = a.lo + ((long long)blockIdx.x * ADAM_UNROLL) * blockDim.x + threadIdx.x;
  float p[ADAM_UNROLL], g[ADAM_UNROLL], m[ADAM_UNROLL], v[ADAM_UNROLL], ss[ADAM_UNROLL];
  bool on[ADAM_UNROLL];
  long long g_begin = 0, g_end = 0;
  float g_ss = 0.f;
#pragma unroll
  for (int u = 0; u < ADAM_UNROLL; u++) {
    const long long i = base + (long long)u * blockDim.x;
    on[u] = false;
    ss[u] = 0.f;
    if (i < a.hi) {
      if (i >= g_begin && i < g_end) {   // same group as the previous element of this thread: the common case
        on[u] = true;
        ss[u] = g_ss;
      } else {
        for (int k = 0; k < a.num_groups; k++)
          if (i >= a.begin[k] && i < a.end[k]) {
            on[u] = true;
            ss[u] = g_ss = a.step_size[k];
            g_begin = a.begin[k];
            g_end = a.end[k];
          }
      }
    }
    if (on[u]) { p[u] = a.p[i]; g[u] = a.g[i]; m[u] = a.m[i]; v[u] = a.v[i]; }
  }
#pragma unroll
  for (int u = 0; u < ADAM_UNROLL; u++) {
    if (!on[u]) continue;
    const long long i = base + (long long)u * blockDim.x;
    const float gr = g[u] * a.grad_scale;
    const float mn = m[u] + (gr - m[u]) * a.one_m_beta1;          // exp_avg.lerp_(grad, 1 - beta1)
    const float vn = v[u] * a.beta2 + a.one_m_beta2 * gr * gr;     // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vn) * a.inv_bc2_sqrt + a.eps;        // (sqrt(v) / bias_correction2_sqrt).add_(eps)
    a.p[i] = p[u] - ss[u] * (mn / denom);                          // addcdiv_(exp_avg, denom, value=-step_size)
    a.m[i] = mn;
    a.v[i] = vn;
    if (a.zero_grads) a.g[i] = 0.f;
  }
}

}  // namespace

extern "C" int ibgs_adam_step(const IbgsAdamArgs* f, void* stream_v) {
  cudaStream_t s = (cudaStream_t)stream_v;
  if (!f) { ibgs_set_error("args is NULL"); return IBGS_EINVAL; }
  if (f->num_groups < 0 || f->num_groups > IBGS_ADAM_MAX_GROUPS) {
    ibgs_set_error("num_groups must be in [0,%d], got %d", IBGS_ADAM_MAX_GROUPS, f->num_groups);
    return IBGS_EINVAL;
  }
  if (f->step < 1) { ibgs_set_error("step must be >= 1 (the 1-based count of this update), got %lld", (long long)f->step); return IBGS_EINVAL; }
  if (f->num_groups == 0) return IBGS_OK;
  if (!f->params || !f->grads || !f->exp_avg || !f->exp_avg_sq) {
    ibgs_set_error("params / grads / exp_avg / exp_avg_sq must not be NULL");
    return IBGS_EINVAL;
  }
  AdamKArgs a;
  a.p = f->params; a.g = f->grads; a.m = f->exp_avg; a.v = f->exp_avg_sq;
  a.num_groups = f->num_groups;
  // bias corrections in double like torch's Python scalars, then rounded once
  const double bc1 = 1.0 - pow((double)f->beta1, (double)f->step);
  const double bc2 = 1.0 - pow((double)f->beta2, (double)f->step);
  long long lo = -1, hi = -1;
  for (int k = 0; k < f->num_groups; k++) {
    const IbgsAdamGroup& gk = f->groups[k];
    if (gk.offset < 0 || gk.count < 0) { ibgs_set_error("group %d: negative offset / count", k); return IBGS_EINVAL; }
    a.begin[k] = gk.offset;
    a.end[k] = gk.offset + gk.count;
    a.step_size[k] = (float)((double)gk.lr / bc1);
    if (gk.count > 0) {
      if (lo < 0 || gk.offset < lo) lo = gk.offset;
      if (a.end[k] > hi) hi = a.end[k];
    }
  }
  if (lo < 0) return IBGS_OK;
  a.lo = lo; a.hi = hi;
  a.one_m_beta1 = (float)(1.0 - (double)f->beta1);
  a.beta2 = f->beta2;
  a.one_m_beta2 = (float)(1.0 - (double)f->beta2);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  a.eps = f->eps;
  a.grad_scale = f->grad_scale;
  a.zero_grads = f->zero_grads;
  const long long n = hi - lo;
  const long long per_block = 256LL * ADAM_UNROLL;
  const long long blocks = (n + per_block - 1) / per_block;
  if (blocks > 0x7fffffffLL) { ibgs_set_error("arena too large for one launch"); return IBGS_ELIMIT; }
  adam_step_kernel<<<(unsigned)blocks, 256, 0, s>>>(a);
  KERNEL_CHECK(0, s);
  return IBGS_OK;
}
