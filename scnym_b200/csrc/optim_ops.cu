// Fused optimizer / teacher kernels over the flat parameter arena (sm_100a).
//   scnb_step_begin   advance the device-side step counters (optimizer t, RNG counter, MixMatch step)
//   scnb_optim_step   Adadelta / Adam / AdamW / SGD-momentum, one elementwise pass over the arena
//                     (torch.optim single-tensor rules; reference call sites main.py:36-41,374-396)
//   scnb_ema_update   mean-teacher EMA t <- a t + (1-a) p, a = min(1 - 1/(step+1), decay)   (losses.py:382-406)
// Hyper-parameters live in a small device buffer so that a captured CUDA graph can be replayed
// while lr / weights change:  hp = [lr, weight_decay, beta1, beta2, eps, rho, momentum, 0]
// state (int64): st = [t (optimizer steps taken), rng_seed, rng_counter, mm_step]
#include <stdlib.h>

#include "optim.cuh"
#include "tc_common.cuh"

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

// Zero-fill of up to two fp32 buffers in one launch (16-byte aligned; the step's accumulation targets: the split-K first
// layer output and the gradient arena behind the first layer) -- keeps every launch of the captured step inside this library.
__global__ void __launch_bounds__(256) k_zero2(float4* __restrict__ a, long long na4, float4* __restrict__ b, long long nb4) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = i0; i < na4; i += stride) a[i] = z;
  for (long long i = i0; i < nb4; i += stride) b[i] = z;
}

extern "C" int scnb_zero2(float* a, long long na, float* b, long long nb, void* stream) {
  SCNB_CHECK_ARG(na >= 0 && nb >= 0 && (na == 0 || a) && (nb == 0 || b), "zero2: bad arguments");
  SCNB_CHECK_ARG(na % 4 == 0 && nb % 4 == 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0), "zero2: buffers must be 16-byte aligned multiples of 4 floats");
  if (na + nb == 0) return SCNB_OK;
  const long long n4 = (na > nb ? na : nb) / 4;
  const int grid = (int)(n4 / 256 + 1 < 148 * 4 ? n4 / 256 + 1 : 148 * 4);
  k_zero2<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)a, na / 4, (float4*)b, nb / 4);
  SCNB_CHECK_LAUNCH("zero2");
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

// ---------------------------------------------------------------------------------------------
// Data-parallel gradient exchange fused with the optimizer, over peer memory (one process per GPU, one node):
//   reduce-scatter  : rank r owns the r-th slice of the flat arena; it LOADS that slice of every rank's gradient arena
//                     over NVLink (fixed rank order -> the same sum wherever it is computed) and averages;
//   optimizer       : Adadelta / Adam / AdamW / SGD on the owned slice only (optimizer state is touched by its owner
//                     alone: 1/world of the single-GPU HBM traffic);
//   all-gather      : the updated parameters are STORED straight into every rank's parameter arena.
// Replaces ncclAllReduce(grad arena) + scnb_optim_step(whole arena): no intermediate averaged gradient, no second pass.
// Each rank's exchange buffer holds [... | flags | gradient arena | parameter arena]; bufs[p] + off_* address rank p's.
// Protocol (flags carry the step number, so nothing is ever reset):
//   ready[r] : rank r's gradients of this step are complete (written by r into every rank's flag block at kernel start)
//   done[r]  : rank r has read all gradients it needs and its parameter stores are visible (written by its last CTA)
// A rank's kernel returns only when every done[] flag has arrived: its parameter arena is then complete and its
// gradient arena may be overwritten by the next step.
// ---------------------------------------------------------------------------------------------
struct DpX {
  const unsigned long long* bufs;
  const long long* epoch;
  long long off_flags, off_grad, off_flat;   // byte offsets inside every rank's exchange buffer
  int rank, world, spin_limit;
  long long off_inbox;   // >= 0: gradients of arena elements [0, pushed_f4) were PUSHED by their producers into the owners' inboxes
  long long pushed_f4;   //        (scnb_csr_w1_bwd_push_tc): region q of this rank's inbox = rank q's gradient of this rank's slice
  int probe;   // timing experiments only (SCNB_DP_PROBE): bit 0 = no stores to peers, bit 1 = no loads from peers (results are wrong)
};

__device__ __forceinline__ void dp_wait_flag(const unsigned* f, unsigned e, int spin_limit) {
  unsigned got;
  int spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(f) : "memory");
    if (got != e && ++spins > spin_limit) __trap();
  } while (got != e);
}

// Bulk-copy form.  A peer load over NVLink has ~2-3 us of latency: with one 16-byte load per thread and peer in flight
// (the first version of this kernel: 75 us for the 31 MB arena of config 3 at two ranks, ~200 GB/s per direction) the link
// idles most of the time.  Here one thread per CTA keeps `stages` chunks of DPB_F4 float4 elements in flight with
// cp.async.bulk -- the slice of every rank's gradient arena (peer memory) and the CTA's own p / state rows land in shared
// memory -- so a persistent grid of one CTA per SM has several MB outstanding; the 256 threads then sum in rank order,
// apply the optimizer and store the new parameters to every rank (posted stores, nothing waits for them until the
// fence in front of the done flag).
#define DPB_T 256
#define DPB_F4 512    // float4 elements per chunk and source: 8 KB

template <int OPT>
__global__ void __launch_bounds__(DPB_T, 1) k_dp_reduce_optim_bcast(DpX dx, float* __restrict__ s1, float* __restrict__ s2,
                                                                   long long total, const float* __restrict__ hp,
                                                                   const int64_t* __restrict__ st, int stages) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ OptimConsts cs;
  __shared__ int is_last;
  __shared__ __align__(8) unsigned long long bars[8];
  const int t = threadIdx.x, W = dx.world;
  constexpr bool TWO = (OPT != OPT_SGD);
  const int nsrc = W + 2 + (TWO ? 1 : 0);                       // gradient of every rank, p, state1 (, state2)
  const uint32_t stage_bytes = (uint32_t)nsrc * DPB_F4 * 16u;
  const unsigned e = (unsigned)dx.epoch[0];
  // flag block layout (u32 words): ready[r] at 2r, done[r] at 64 + 2r, CTA counter at 128
  unsigned* my_flags = reinterpret_cast<unsigned*>(dx.bufs[dx.rank] + dx.off_flags);
  const long long n4 = total >> 2;                                   // arena length is a multiple of 4 floats
  const long long per = dp_slice_f4(total, W);                       // float4 elements per rank
  const long long lo = per * dx.rank, hi = min(n4, lo + per);
  // two runs of chunks: [lo, mid) whose gradients sit in this rank's inbox, [mid, hi) pulled from the peers' arenas
  const long long mid = (dx.off_inbox >= 0) ? max(lo, min(hi, dx.pushed_f4)) : lo;
  const long long nchunk_a = mid > lo ? (mid - lo + DPB_F4 - 1) / DPB_F4 : 0;
  const long long nchunk = nchunk_a + (hi > mid ? (hi - mid + DPB_F4 - 1) / DPB_F4 : 0);
  const long long mine = nchunk > blockIdx.x ? (nchunk - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // chunks of this CTA
  auto chunk_range = [&](long long k, long long& base, long long& lim) -> bool {   // true: inbox run
    if (k < nchunk_a) { base = lo + k * DPB_F4; lim = mid; return true; }
    base = mid + (k - nchunk_a) * DPB_F4; lim = hi; return false;
  };
  float4* s1v = reinterpret_cast<float4*>(s1);
  float4* s2v = reinterpret_cast<float4*>(s2);
  const float4* my_flat = reinterpret_cast<const float4*>(dx.bufs[dx.rank] + dx.off_flat);
  // thread 0: every source of the CTA's j-th chunk into stage j % stages.  which: 1 = this rank's own rows (p, state,
  // gradient: nothing to wait for), 2 = the peers' gradients (after their ready flags), 3 = both
  auto issue = [&](long long j, int which) {
    const int s = (int)(j % stages);
    long long base, lim;
    const bool inbox = chunk_range(blockIdx.x + j * gridDim.x, base, lim);
    const uint32_t bytes = (uint32_t)min((long long)DPB_F4, lim - base) * 16u;
    const uint32_t bar = smem_u32(&bars[s]);
    const uint32_t dst = smem_u32(dsm) + (uint32_t)s * stage_bytes;
    if (which & 1) {
      mbar_arrive_expect_tx(bar, bytes * (uint32_t)nsrc);
      bulk_g2s(dst + (uint32_t)W * DPB_F4 * 16u, my_flat + base, bytes, bar);
      bulk_g2s(dst + (uint32_t)(W + 1) * DPB_F4 * 16u, s1v + base, bytes, bar);
      if (TWO) bulk_g2s(dst + (uint32_t)(W + 2) * DPB_F4 * 16u, s2v + base, bytes, bar);
    }
    for (int q = 0; q < W; ++q) {
      if (inbox) {
        // every rank's gradient of this slice is in LOCAL memory (region q of the inbox), but it was written by rank q:
        // all of them wait for the ready flags (the own region too -- one rule)
        if (which & 2)
          bulk_g2s(dst + (uint32_t)q * DPB_F4 * 16u,
                   reinterpret_cast<const float4*>(dx.bufs[dx.rank] + dx.off_inbox) + (long long)q * per + (base - lo), bytes, bar);
      } else if (which & (q == dx.rank ? 1 : 2)) {
        bulk_g2s(dst + (uint32_t)q * DPB_F4 * 16u,
                 reinterpret_cast<const float4*>(dx.bufs[(dx.probe & 2) ? dx.rank : q] + dx.off_grad) + base, bytes, bar);
      }
    }
  };
  if (t == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // this rank's gradients were written through the generic proxy by the kernels in front: order the bulk (async-proxy)
    // reads behind them, and start on the local rows before anything is known about the peers
    asm volatile("fence.proxy.async;" ::: "memory");
    for (long long j = 0; j < mine && j < stages; ++j) issue(j, 1);
  }
  if (t == 32) cs = optim_consts(hp, st);     // double-precision pow / sqrt: off the loader's path
  if (blockIdx.x == 0 && t >= 64 && t < 64 + W) {
    __threadfence_system();
    unsigned* f = reinterpret_cast<unsigned*>(dx.bufs[t - 64] + dx.off_flags) + 2 * dx.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(e) : "memory");
  }
  if (t >= 96 && t < 96 + W) dp_wait_flag(my_flags + 2 * (t - 96), e, dx.spin_limit);
  __syncthreads();
  if (t == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");   // the peers' gradients, seen through the flags acquired above
    for (long long j = 0; j < mine && j < stages; ++j) issue(j, 2);
  }
  const OptimConsts c = cs;
  const float inv_w = 1.0f / (float)W;
  for (long long j = 0; j < mine; ++j) {
    const int s = (int)(j % stages);
    long long base, lim;
    chunk_range(blockIdx.x + j * gridDim.x, base, lim);
    const int n = (int)min((long long)DPB_F4, lim - base);
    mbar_wait(smem_u32(&bars[s]), (uint32_t)((j / stages) & 1));
    const float4* sv = reinterpret_cast<const float4*>(dsm + (size_t)s * stage_bytes);
    float4 g[2], pi[2], a[2], b[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = t + h * DPB_T;
      g[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n) {
        for (int q = 0; q < W; ++q) {        // rank order: the same sum wherever it is computed
          const float4 v = sv[q * DPB_F4 + i];
          g[h].x += v.x; g[h].y += v.y; g[h].z += v.z; g[h].w += v.w;
        }
        pi[h] = sv[W * DPB_F4 + i];
        a[h] = sv[(W + 1) * DPB_F4 + i];
        b[h] = TWO ? sv[(W + 2) * DPB_F4 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();                          // the stage has been read: refill it
    if (t == 0 && j + stages < mine) issue(j + stages, 3);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = t + h * DPB_T;
      if (i < n) {
        float4 gg = g[h];
        gg.x *= inv_w; gg.y *= inv_w; gg.z *= inv_w; gg.w *= inv_w;
        if (!(dx.probe & 8)) {
        optim_update<OPT>(pi[h].x, gg.x, a[h].x, b[h].x, c);
        optim_update<OPT>(pi[h].y, gg.y, a[h].y, b[h].y, c);
        optim_update<OPT>(pi[h].z, gg.z, a[h].z, b[h].z, c);
        optim_update<OPT>(pi[h].w, gg.w, a[h].w, b[h].w, c);
        } else { pi[h].x += gg.x; }
        if ((dx.probe & 4) && pi[h].x != 123.456f) continue;
        s1v[base + i] = a[h];
        if (TWO) s2v[base + i] = b[h];
        for (int q = 0; q < W; ++q)   // weak store; made visible by the system-scope fence in front of the done flag
          if (!(dx.probe & 1) || q == dx.rank) __stcg(reinterpret_cast<float4*>(dx.bufs[q] + dx.off_flat) + base + i, pi[h]);
      }
    }
  }
  // completion: every CTA's stores are fenced, the last CTA to finish publishes done[rank] and waits for all ranks
  __syncthreads();
  if (t == 0) {
    __threadfence_system();
    const unsigned ticket = atomicAdd(my_flags + 128, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    if (t == 0) my_flags[128] = 0u;
    if (t < W) {
      __threadfence_system();
      unsigned* f = reinterpret_cast<unsigned*>(dx.bufs[t] + dx.off_flags) + 64 + 2 * dx.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(e) : "memory");
      dp_wait_flag(my_flags + 64 + 2 * t, e, dx.spin_limit);
    }
  }
}

extern "C" int scnb_dp_slice_floats(long long total, int world, long long* slice_floats_out) {
  SCNB_CHECK_ARG(total >= 0 && total % 4 == 0 && world >= 1 && slice_floats_out, "dp_slice_floats: bad arguments");
  *slice_floats_out = 4 * dp_slice_f4(total, world);
  return SCNB_OK;
}

extern "C" int scnb_dp_reduce_optim_bcast2(int opt, const uint64_t* peer_bufs, int rank, int world, long long off_flags,
                                           long long off_grad, long long off_flat, long long off_inbox, long long pushed_floats,
                                           long long total, float* state1, float* state2, const float* hp8,
                                           const int64_t* state4, const int64_t* epoch, void* stream);

extern "C" int scnb_dp_reduce_optim_bcast(int opt, const uint64_t* peer_bufs, int rank, int world, long long off_flags,
                                          long long off_grad, long long off_flat, long long total, float* state1, float* state2,
                                          const float* hp8, const int64_t* state4, const int64_t* epoch, void* stream) {
  return scnb_dp_reduce_optim_bcast2(opt, peer_bufs, rank, world, off_flags, off_grad, off_flat, -1, 0, total, state1, state2, hp8,
                                     state4, epoch, stream);
}

extern "C" int scnb_dp_reduce_optim_bcast2(int opt, const uint64_t* peer_bufs, int rank, int world, long long off_flags,
                                           long long off_grad, long long off_flat, long long off_inbox, long long pushed_floats,
                                           long long total, float* state1, float* state2, const float* hp8,
                                           const int64_t* state4, const int64_t* epoch, void* stream) {
  SCNB_CHECK_ARG(peer_bufs && state1 && hp8 && state4 && epoch && total >= 0 && total % 4 == 0, "dp_reduce_optim_bcast: bad arguments");
  SCNB_CHECK_ARG(off_inbox < 0 || (off_inbox % 16 == 0 && pushed_floats >= 0 && pushed_floats <= total && pushed_floats % 4 == 0),
                 "dp_reduce_optim_bcast: bad inbox arguments");
  SCNB_CHECK_ARG(world >= 1 && world <= 8 && rank >= 0 && rank < world, "dp_reduce_optim_bcast: bad rank/world (one NVSwitch node: world <= 8)");
  SCNB_CHECK_ARG(opt == OPT_SGD || state2, "dp_reduce_optim_bcast: second state buffer required");
  SCNB_CHECK_ARG(off_flags % 16 == 0 && off_grad % 16 == 0 && off_flat % 16 == 0, "dp_reduce_optim_bcast: offsets must be 16-byte aligned");
  if (total == 0) return SCNB_OK;
  static int spin = 0;
  if (!spin) {
    const char* ev = getenv("SCNB_PEER_SPIN_LIMIT");
    spin = ev ? atoi(ev) : (1 << 24);
    if (spin <= 0) spin = 1 << 24;
  }
  static int probe = -1;
  if (probe < 0) {
    const char* ev = getenv("SCNB_DP_PROBE");
    probe = ev ? atoi(ev) : 0;
  }
  DpX dx{(const unsigned long long*)peer_bufs, (const long long*)epoch, off_flags, off_grad, off_flat, rank, world, spin, off_inbox,
         pushed_floats >> 2, probe};
  // One CTA per SM, every CTA co-resident with its peers' (a CTA waits only for flags that are already on their way and
  // for the last CTA of the OTHER grids).  Ranks emulated on one GPU (tests) share its SMs: SCNB_DP_GRID caps the grid.
  const long long per = dp_slice_f4(total, world);
  static int grid_cap = 0;
  if (!grid_cap) {
    const char* ev = getenv("SCNB_DP_GRID");
    grid_cap = ev ? atoi(ev) : 148;
    if (grid_cap < 1 || grid_cap > 148) grid_cap = 148;
  }
  const long long want = (per + DPB_F4 - 1) / DPB_F4;
  int blocks = (int)(want < grid_cap ? want : grid_cap);
  if (blocks < 1) blocks = 1;
  const int nsrc = world + 2 + (opt != OPT_SGD ? 1 : 0);
  int stages = (200 * 1024) / (nsrc * DPB_F4 * 16);
  if (stages > 6) stages = 6;
  static int stage_cap = -1;
  if (stage_cap < 0) {
    const char* ev = getenv("SCNB_DP_STAGES");
    stage_cap = ev ? atoi(ev) : 0;
  }
  if (stage_cap > 0 && stages > stage_cap) stages = stage_cap;
  if ((long long)stages > (want + blocks - 1) / blocks) stages = (int)((want + blocks - 1) / blocks);
  if (stages < 1) stages = 1;
  const size_t smem = (size_t)stages * nsrc * DPB_F4 * 16;
  cudaStream_t st = (cudaStream_t)stream;
#define DP_GO(O)                                                                                                         \
  do {                                                                                                                   \
    static bool attr = false;                                                                                            \
    if (!attr) {                                                                                                         \
      cudaError_t ce = cudaFuncSetAttribute(k_dp_reduce_optim_bcast<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (ce != cudaSuccess) {                                                                                           \
        scnb_set_error("dp_reduce_optim_bcast: cannot reserve shared memory: %s", cudaGetErrorString(ce));               \
        return SCNB_ERR_CUDA;                                                                                            \
      }                                                                                                                  \
      attr = true;                                                                                                       \
    }                                                                                                                    \
    k_dp_reduce_optim_bcast<O><<<blocks, DPB_T, smem, st>>>(dx, state1, state2, total, hp8, state4, stages);          \
  } while (0)
  switch (opt) {
    case OPT_ADADELTA: DP_GO(OPT_ADADELTA); break;
    case OPT_ADAM: DP_GO(OPT_ADAM); break;
    case OPT_ADAMW: DP_GO(OPT_ADAMW); break;
    case OPT_SGD: DP_GO(OPT_SGD); break;
    default: scnb_set_error("dp_reduce_optim_bcast: unknown optimizer %d", opt); return SCNB_ERR_ARG;
  }
#undef DP_GO
  SCNB_CHECK_LAUNCH("dp_reduce_optim_bcast");
  return SCNB_OK;
}
