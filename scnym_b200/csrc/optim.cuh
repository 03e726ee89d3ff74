// Optimizer update rules shared by the elementwise arena kernel (optim_ops.cu) and the fused
// first-layer gradient+update kernel (csr_ops.cu).  torch.optim single-tensor semantics
// (SURVEY.md Appendix A7); reference call sites scnym/main.py:36-41, 374-396.
#pragma once
#include "common.cuh"

enum { OPT_NONE = -1, OPT_ADADELTA = 0, OPT_ADAM = 1, OPT_ADAMW = 2, OPT_SGD = 3 };

struct OptimConsts {
  float lr, wd, b1, b2, eps, rho, mom;
  float step_size;     // lr / (1 - b1^t)
  float bc2_sqrt;      // sqrt(1 - b2^t)
  float inv_bc2_sqrt;  // 1 / sqrt(1 - b2^t)
  int first;           // t <= 1 (SGD momentum buffer initialisation)
};

// The update is a pure HBM stream (28 bytes per element), but IEEE sqrtf / division expand to ~15-instruction
// sequences each and made the fused first-layer kernels issue-bound (ncu: 16 µs of a 38 µs kernel with loads and
// stores removed).  The special-function-unit forms below are single instructions with <= 2 ulp error -- five orders
// of magnitude inside the 1e-4 parity budget against torch.optim.
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// hp = [lr, weight_decay, beta1, beta2, eps, rho, momentum, 0]; st[0] = t (1-based at this point)
__device__ __forceinline__ OptimConsts optim_consts(const float* __restrict__ hp, const int64_t* __restrict__ st) {
  OptimConsts c;
  c.lr = hp[0]; c.wd = hp[1]; c.b1 = hp[2]; c.b2 = hp[3]; c.eps = hp[4]; c.rho = hp[5]; c.mom = hp[6];
  const double t = (double)st[0];
  const double bc1 = 1.0 - pow((double)c.b1, t), bc2 = 1.0 - pow((double)c.b2, t);
  c.step_size = (float)((double)c.lr / bc1);
  c.bc2_sqrt = (float)sqrt(bc2);
  c.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  c.first = st[0] <= 1;
  return c;
}

// One element.  p: parameter, g: raw gradient, a/b: the two state slots
// (Adadelta: square_avg, acc_delta; Adam/AdamW: exp_avg, exp_avg_sq; SGD: momentum buffer, unused).
template <int OPT>
__device__ __forceinline__ void optim_update(float& p, float g, float& a, float& b, const OptimConsts& c) {
  if (OPT == OPT_ADADELTA) {
    g = fmaf(c.wd, p, g);
    const float sq = c.rho * a + (1.f - c.rho) * g * g;
    const float delta = fast_sqrt(b + c.eps) * fast_rsqrt(sq + c.eps) * g;
    b = c.rho * b + (1.f - c.rho) * delta * delta;
    a = sq;
    p = p - c.lr * delta;
  } else if (OPT == OPT_ADAM || OPT == OPT_ADAMW) {
    if (OPT == OPT_ADAMW) p = p * (1.f - c.lr * c.wd);
    else g = fmaf(c.wd, p, g);
    const float m = a + (1.f - c.b1) * (g - a);
    const float v = c.b2 * b + (1.f - c.b2) * g * g;
    const float denom = fmaf(fast_sqrt(v), c.inv_bc2_sqrt, c.eps);
    a = m;
    b = v;
    p = fmaf(-c.step_size, m * fast_rcp(denom), p);
  } else if (OPT == OPT_SGD) {  // momentum, dampening 0, no nesterov
    g = fmaf(c.wd, p, g);
    const float buf = c.first ? g : c.mom * a + g;
    a = buf;
    p = p - c.lr * buf;
  }
}

// Data-parallel ownership of the flat arena: rank r owns float4 elements [r * S, min(n4, (r + 1) * S)), S = this many
// float4 -- the even share rounded up to 64 float4 (256 floats), so that a row of W1T (H0 floats, H0 | 256) never
// straddles two owners (the dW1 kernel then sends each gradient row to exactly one rank).
__host__ __device__ __forceinline__ long long dp_slice_f4(long long total_floats, int world) {
  const long long n4 = total_floats >> 2;
  const long long per = (n4 + world - 1) / world;
  return (per + 63) / 64 * 64;
}
