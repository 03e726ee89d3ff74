// Shared device/host helpers for libscnym_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define SCNB_OK 0
#define SCNB_ERR_ARG -1
#define SCNB_ERR_CUDA -2
#define SCNB_ERR_UNSUPPORTED -3

void scnb_set_error(const char* fmt, ...);

#define SCNB_CHECK_ARG(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      scnb_set_error(__VA_ARGS__);         \
      return SCNB_ERR_ARG;                 \
    }                                      \
  } while (0)

// Launch errors only (no sync: every entry point is asynchronous and graph-capturable).
#define SCNB_CHECK_LAUNCH(name)                                              \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      scnb_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return SCNB_ERR_CUDA;                                                  \
    }                                                                        \
  } while (0)

static inline int scnb_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch (PDL): a kernel launched with scnb_launch_pdl may be scheduled while
// its predecessor in the stream is still draining, which hides its launch latency -- the step is a
// chain of ~25 short dependent kernels.  Such a kernel starts with PDL_ENTRY(): wait until the
// predecessor grid has completed and its writes are visible, then allow its own successor to be
// scheduled.  Both instructions are no-ops for a kernel launched the ordinary way.
int scnb_pdl_enabled(void);   // SCNB_PDL=1 enables (api.cu; off by default: no measurable gain inside CUDA graphs)

#ifdef __CUDACC__
// Split form: PDL_LAUNCH() first thing (lets the successor start its own prologue as soon as this grid's CTAs are placed),
// then everything that does not read the predecessor's output (barrier setup, weight panels, BatchNorm parameters), then
// PDL_WAIT() in front of the first access to the predecessor's output.
#define PDL_LAUNCH() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define PDL_ENTRY()                                             \
  do {                                                          \
    asm volatile("griddepcontrol.wait;" ::: "memory");          \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)

template <typename... P, typename... A>
static inline cudaError_t scnb_launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = scnb_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Philox4x32-10 counter RNG (Salmon et al. 2011), used for in-kernel dropout masks when no
// mask pointer is injected.  ctr = (idx_lo, idx_hi, stream, 0), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// Uniform in [0,1) for element `idx` of logical stream `stream` (one Philox call per 4 elements).
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t stream, uint64_t idx) {
  uint64_t blk = idx >> 2;
  uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t v = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return (float)(v >> 8) * (1.0f / 16777216.0f);
}
// Hidden-layer dropout stream (shared by every BatchNorm -> Dropout -> ReLU kernel, forward and backward, so that all of
// them see the same mask): one Philox call per (row, 8-column group) gives eight 16-bit uniforms, element j keeps iff
// uniform[j] >= thr16 = round(p * 65536); an injected byte mask is read 8 bytes at a time (elem0 = row*H + first column).
// 8-bit keep pattern (bit j = keep column j) of element group `grp` = row*(H/8) + column group
__device__ __forceinline__ uint32_t keep8(const uint8_t* __restrict__ mask, size_t elem0, uint64_t grp, uint64_t seed, uint64_t strm,
                                          uint32_t thr16) {
  uint32_t bits = 0;
  if (mask) {
    const uint2 m = *reinterpret_cast<const uint2*>(mask + elem0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bits |= (((m.x >> (8 * j)) & 0xffu) != 0u ? 1u : 0u) << j;
      bits |= (((m.y >> (8 * j)) & 0xffu) != 0u ? 1u : 0u) << (4 + j);
    }
    return bits;
  }
  const uint4 r = philox4x32_10(make_uint4((uint32_t)grp, (uint32_t)(grp >> 32), (uint32_t)strm, (uint32_t)(strm >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bits |= ((w[j] & 0xffffu) >= thr16 ? 1u : 0u) << (2 * j);
    bits |= ((w[j] >> 16) >= thr16 ? 1u : 0u) << (2 * j + 1);
  }
  return bits;
}
// keep decision for dropout: injected mask wins; else Philox(seed, stream, idx) >= p.
__device__ __forceinline__ bool dropout_keep(const uint8_t* mask, uint64_t idx, uint64_t seed, uint64_t stream, float p) {
  if (mask) return mask[idx] != 0;
  return philox_uniform(seed, stream, idx) >= p;
}
#endif
