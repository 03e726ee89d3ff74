// Fused optimizer / teacher kernels over the flat parameter arena (sm_100a).
//   scnb_step_begin   advance the device-side step counters (optimizer t, RNG counter, MixMatch step)
//   scnb_optim_step   Adadelta / Adam / AdamW / SGD-momentum, one elementwise pass over the arena
//                     (torch.optim single-tensor rules; reference call sites main.py:36-41,374-396)
//   scnb_ema_update   mean-teacher EMA t <- a t + (1-a) p, a = min(1 - 1/(step+1), decay)   (losses.py:382-406)
// Hyper-parameters live in a small device buffer so that a captured CUDA graph can be replayed
// while lr / weights change:  hp = [lr, weight_decay, beta1, beta2, eps, rho, momentum, 0]
// state (int64): st = [t (optimizer steps taken), rng_seed, rng_counter, mm_step]
#include "optim.cuh"

__global__ void k_step_begin(int64_t* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st[0] += 1;  // optimizer step index t (1-based when read by optim_step)
    st[2] += 1;  // RNG counter
  }
}

extern "C" int scnb_step_begin(int64_t* state4, void* stream) {
  SCNB_CHECK_ARG(state4, "step_begin: null state");
  k_step_begin<<<1, 32, 0, (cudaStream_t)stream>>>(state4);
  SCNB_CHECK_LAUNCH("step_begin");
  return SCNB_OK;
}

__global__ void k_mm_step_inc(int64_t* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) st[3] += 1;
}

template <int OPT>
__global__ void __launch_bounds__(256) k_optim(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ s1,
                                               float* __restrict__ s2, size_t n, const float* __restrict__ hp,
                                               const int64_t* __restrict__ st) {
  PDL_ENTRY();
  __shared__ OptimConsts cs;
  if (threadIdx.x == 0) cs = optim_consts(hp, st);
  __syncthreads();
  const OptimConsts c = cs;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float pi = p[i], a = s1[i], b = (OPT == OPT_SGD) ? 0.f : s2[i];
    optim_update<OPT>(pi, g[i], a, b, c);
    s1[i] = a;
    if (OPT != OPT_SGD) s2[i] = b;
    p[i] = pi;
  }
}

extern "C" int scnb_optim_step(int opt, float* params, const float* grads, float* state1, float* state2, long long n,
                               const float* hp8, const int64_t* state4, void* stream) {
  SCNB_CHECK_ARG(params && grads && state1 && hp8 && state4 && n >= 0, "optim_step: bad arguments");
  SCNB_CHECK_ARG(opt == OPT_SGD || state2, "optim_step: second state buffer required");
  if (n == 0) return SCNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  switch (opt) {
    case OPT_ADADELTA: scnb_launch_pdl(k_optim<OPT_ADADELTA>, dim3(blocks), dim3(256), 0, st, params, grads, state1, state2, (size_t)n, hp8, state4); break;
    case OPT_ADAM: scnb_launch_pdl(k_optim<OPT_ADAM>, dim3(blocks), dim3(256), 0, st, params, grads, state1, state2, (size_t)n, hp8, state4); break;
    case OPT_ADAMW: scnb_launch_pdl(k_optim<OPT_ADAMW>, dim3(blocks), dim3(256), 0, st, params, grads, state1, state2, (size_t)n, hp8, state4); break;
    case OPT_SGD: scnb_launch_pdl(k_optim<OPT_SGD>, dim3(blocks), dim3(256), 0, st, params, grads, state1, state2, (size_t)n, hp8, state4); break;
    default: scnb_set_error("optim_step: unknown optimizer %d", opt); return SCNB_ERR_ARG;
  }
  SCNB_CHECK_LAUNCH("optim_step");
  return SCNB_OK;
}

__global__ void __launch_bounds__(256) k_ema(float* __restrict__ t, const float* __restrict__ p, size_t n, float decay,
                                             const int64_t* __restrict__ st) {
  const double step = (double)st[3];
  const float a = fminf((float)(1.0 - 1.0 / (step + 1.0)), decay);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    t[i] = t[i] * a + (1.f - a) * p[i];
}

extern "C" int scnb_ema_update(float* teacher, const float* params, long long n, float decay, const int64_t* state4,
                               void* stream) {
  SCNB_CHECK_ARG(teacher && params && state4 && n >= 0, "ema_update: bad arguments");
  if (n == 0) return SCNB_OK;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  k_ema<<<blocks, 256, 0, (cudaStream_t)stream>>>(teacher, params, (size_t)n, decay, state4);
  SCNB_CHECK_LAUNCH("ema_update");
  return SCNB_OK;
}

extern "C" int scnb_mm_step_inc(int64_t* state4, void* stream) {
  SCNB_CHECK_ARG(state4, "mm_step_inc: null state");
  k_mm_step_inc<<<1, 32, 0, (cudaStream_t)stream>>>(state4);
  SCNB_CHECK_LAUNCH("mm_step_inc");
  return SCNB_OK;
}

// Per-step bookkeeping without host syncs: loss scalars of step t go to hist[(t-1) % hist_cap],
// pseudolabel confidences to conf_ring[(t-1) % ring_cap] (replaces the per-step .item()/.cpu()
// reads of scnym/trainer.py:666-681 and losses.py:714-725; read back once per epoch).
__global__ void k_record_step(const float* __restrict__ loss8, float* __restrict__ hist, int hist_cap,
                              const float* __restrict__ conf, int U, float* __restrict__ conf_ring, int ring_cap,
                              const int64_t* __restrict__ st) {
  const int64_t t = st[0] - 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (hist && i < 8) hist[(t % hist_cap) * 8 + i] = loss8[i];
  if (conf_ring && i < U) conf_ring[(t % ring_cap) * U + i] = conf[i];
}

extern "C" int scnb_record_step(const float* loss8, float* hist, int hist_cap, const float* conf, int U, float* conf_ring,
                                int ring_cap, const int64_t* state4, void* stream) {
  SCNB_CHECK_ARG(loss8 && state4 && (!hist || hist_cap > 0) && (!conf_ring || (conf && ring_cap > 0)), "record_step: bad arguments");
  const int n = U > 8 ? U : 8;
  k_record_step<<<scnb_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(loss8, hist, hist_cap, conf, U, conf_ring, ring_cap, state4);
  SCNB_CHECK_LAUNCH("record_step");
  return SCNB_OK;
}
