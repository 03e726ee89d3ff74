// Dense hidden-stack kernels of the scNym hot path (sm_100a), v1 SIMT fp32.
//   scnb_sgemm            C = op(A) op(B) (+beta C) (+bias) (+relu)           (nn.Linear fwd / dA / dW)
//   scnb_colsum           out[N] (+)= sum_rows A[M,N]                          (bias gradients)
//   scnb_bn_act_fwd/bwd   segmented BatchNorm1d -> Dropout -> ReLU             (model.py:212-217 etc.)
//   scnb_mix_rows / scnb_unmix_rows   hidden-space MixUp and its adjoint        (dataprep.py:665-689)
//   scnb_bn_running_update            running_mean / running_var / num_batches_tracked
// "Segments" are the row ranges of the student super-batch [U-mixed | L-mixed | DAN] that need
// separate batch statistics; seg_off lives in device memory because the DAN segment length is
// data dependent when pseudolabel thresholding is on.  Rows >= seg_off[n_seg] are written as 0.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

#define BN_EPS 1e-5f

// ---------------------------------------------------------------------------------------------
// SGEMM (fp32 SIMT): BMxBNx16 tiles, 256 threads, (BM/16)x(BN/16) register micro-tile,
// register-prefetched double buffer.  Row-major.
//   op(A)[m,k] = TA ? A[k*lda+m] : A[m*lda+k];  op(B)[k,n] = TB ? B[n*ldb+k] : B[k*ldb+n].
// The hidden GEMMs of this path are tiny (<= 1024 x 256 x 256): the launcher picks 32-wide tiles
// when 64x64 tiles would leave most of the 148 SMs idle, and splits K over gridDim.z (atomic
// accumulation into C, which then must hold the beta*C term already) for the weight-gradient
// shape M=N=256, K=rows.
// ---------------------------------------------------------------------------------------------
#define GBK 16
#define GPAD 4

template <int BM, bool TA, bool VEC>
__device__ __forceinline__ void load_a_frag(const float* __restrict__ A, int lda, int M, int K, int m0, int k0, int k1, int tid,
                                            float (&r)[4]) {
  r[0] = r[1] = r[2] = r[3] = 0.f;
  if (tid >= BM * 4) return;
  if (!TA) {  // contiguous along k: thread -> (m = tid/4, k = (tid%4)*4 .. +3)
    const int m = m0 + (tid >> 2), k = k0 + ((tid & 3) << 2);
    if (VEC) {
      if (m < M && k < k1) {
        const float4 v = *reinterpret_cast<const float4*>(A + (size_t)m * lda + k);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = (m < M && k + i < k1) ? A[(size_t)m * lda + k + i] : 0.f;
    }
  } else {  // contiguous along m: thread -> (k = tid/(BM/4), m = (tid%(BM/4))*4 .. +3)
    const int k = k0 + tid / (BM / 4), m = m0 + ((tid % (BM / 4)) << 2);
    if (VEC) {
      if (k < k1 && m < M) {
        const float4 v = *reinterpret_cast<const float4*>(A + (size_t)k * lda + m);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = (k < k1 && m + i < M) ? A[(size_t)k * lda + m + i] : 0.f;
    }
  }
}
template <int BM, bool TA>
__device__ __forceinline__ void store_a_frag(float (*As)[BM + GPAD], int tid, const float (&r)[4]) {
  if (tid >= BM * 4) return;
  if (!TA) {
    const int m = tid >> 2, k = (tid & 3) << 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) As[k + i][m] = r[i];
  } else {
    const int k = tid / (BM / 4), m = (tid % (BM / 4)) << 2;
    *reinterpret_cast<float4*>(&As[k][m]) = make_float4(r[0], r[1], r[2], r[3]);
  }
}
template <int BN, bool TB, bool VEC>
__device__ __forceinline__ void load_b_frag(const float* __restrict__ B, int ldb, int N, int K, int n0, int k0, int k1, int tid,
                                            float (&r)[4]) {
  r[0] = r[1] = r[2] = r[3] = 0.f;
  if (tid >= BN * 4) return;
  if (!TB) {  // B stored [K,N], contiguous along n: thread -> (k = tid/(BN/4), n = (tid%(BN/4))*4)
    const int k = k0 + tid / (BN / 4), n = n0 + ((tid % (BN / 4)) << 2);
    if (VEC) {
      if (k < k1 && n < N) {
        const float4 v = *reinterpret_cast<const float4*>(B + (size_t)k * ldb + n);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = (k < k1 && n + i < N) ? B[(size_t)k * ldb + n + i] : 0.f;
    }
  } else {  // B stored [N,K], contiguous along k: thread -> (n = tid/4, k = (tid%4)*4)
    const int n = n0 + (tid >> 2), k = k0 + ((tid & 3) << 2);
    if (VEC) {
      if (n < N && k < k1) {
        const float4 v = *reinterpret_cast<const float4*>(B + (size_t)n * ldb + k);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = (n < N && k + i < k1) ? B[(size_t)n * ldb + k + i] : 0.f;
    }
  }
}
template <int BN, bool TB>
__device__ __forceinline__ void store_b_frag(float (*Bs)[BN + GPAD], int tid, const float (&r)[4]) {
  if (tid >= BN * 4) return;
  if (!TB) {
    const int k = tid / (BN / 4), n = (tid % (BN / 4)) << 2;
    *reinterpret_cast<float4*>(&Bs[k][n]) = make_float4(r[0], r[1], r[2], r[3]);
  } else {
    const int n = tid >> 2, k = (tid & 3) << 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[k + i][n] = r[i];
  }
}

template <int BM, int BN, bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(256) k_sgemm(int M, int N, int K, const float* __restrict__ A, int lda,
                                               const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                                               float alpha, float beta, const float* __restrict__ bias, int relu,
                                               const float* __restrict__ gate, int k_chunk) {
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ __align__(16) float As[2][GBK][BM + GPAD];
  __shared__ __align__(16) float Bs[2][GBK][BN + GPAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * k_chunk, ke = min(K, kb + k_chunk);  // this CTA's K range (split-K)
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float ra[4], rb[4];
  load_a_frag<BM, TA, VEC>(A, lda, M, K, m0, kb, ke, tid, ra);
  load_b_frag<BN, TB, VEC>(B, ldb, N, K, n0, kb, ke, tid, rb);
  store_a_frag<BM, TA>(As[0], tid, ra);
  store_b_frag<BN, TB>(Bs[0], tid, rb);
  __syncthreads();
  const int nk = (ke - kb + GBK - 1) / GBK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_a_frag<BM, TA, VEC>(A, lda, M, K, m0, kb + (kt + 1) * GBK, ke, tid, ra);
      load_b_frag<BN, TB, VEC>(B, ldb, N, K, n0, kb + (kt + 1) * GBK, ke, tid, rb);
    }
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[cur][k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[cur][k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_a_frag<BM, TA>(As[cur ^ 1], tid, ra);
      store_b_frag<BN, TB>(Bs[cur ^ 1], tid, rb);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (split) {  // C already holds beta*C (+bias); partial products are accumulated atomically
        atomicAdd(C + (size_t)m * ldc + n, v);
        continue;
      }
      if (beta != 0.f) v += beta * C[(size_t)m * ldc + n];
      if (bias) v += bias[n];
      if (relu) v = fmaxf(v, 0.f);
      if (gate && !(gate[(size_t)m * ldc + n] > 0.f)) v = 0.f;  // ReLU backward gate (same layout as C)
      C[(size_t)m * ldc + n] = v;
    }
  }
}

template <int BM, int BN>
static void launch_sgemm_tile(bool tA, bool tB, bool vec, dim3 grid, cudaStream_t st, int M, int N, int K, const float* A, int lda,
                              const float* B, int ldb, float* C, int ldc, float alpha, float beta, const float* bias, int relu,
                              const float* gate, int k_chunk) {
#define SCNB_LAUNCH_SGEMM(TA, TB, V) \
  k_sgemm<BM, BN, TA, TB, V><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, k_chunk)
  if (!tA && !tB) { if (vec) SCNB_LAUNCH_SGEMM(false, false, true); else SCNB_LAUNCH_SGEMM(false, false, false); }
  else if (!tA && tB) { if (vec) SCNB_LAUNCH_SGEMM(false, true, true); else SCNB_LAUNCH_SGEMM(false, true, false); }
  else if (tA && !tB) { if (vec) SCNB_LAUNCH_SGEMM(true, false, true); else SCNB_LAUNCH_SGEMM(true, false, false); }
  else { if (vec) SCNB_LAUNCH_SGEMM(true, true, true); else SCNB_LAUNCH_SGEMM(true, true, false); }
#undef SCNB_LAUNCH_SGEMM
}

int scnb_tgemm_try(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                   float alpha, float beta, const float* bias, int relu, const float* gate, cudaStream_t st);
// 1 (default): 3xTF32 mma.sync kernel for the hidden GEMMs; 0: fp32 SIMT panel kernel.  SCNB_GEMM=simt selects 0.
static int scnb_gemm_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("SCNB_GEMM");
    mode = (e && e[0] == 's') ? 0 : 1;
  }
  return mode;
}
int scnb_sgemm_panel_try(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                         int ldc, float alpha, float beta, const float* bias, int relu, const float* gate, cudaStream_t st);
int scnb_gemm_tc_try(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                     float alpha, float beta, const float* bias, int relu, const float* gate, cudaStream_t st);

extern "C" int scnb_sgemm(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                          float* C, int ldc, float alpha, float beta, const float* bias, int relu, const float* gate,
                          void* stream) {
  SCNB_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && A && B && C, "sgemm: bad arguments");
  if (M == 0 || N == 0) return SCNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (K > 0) {  // 3xTF32 tensor-core kernel, else the fp32 cp.async panel kernel (gemm_ops.cu), when operands allow 16-byte chunks
    // large products: tcgen05 bf16x3 kernel (tc_gemm.cu); small ones: the latency-oriented 3xTF32 mma.sync kernel
    int rc = scnb_gemm_tc_try(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, st);
    if (rc == 0 && scnb_gemm_mode() == 1) rc = scnb_tgemm_try(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, st);
    if (rc == 0) rc = scnb_sgemm_panel_try(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, st);
    if (rc < 0) return rc;
    if (rc == 1) {
      SCNB_CHECK_LAUNCH("sgemm/panel");
      return SCNB_OK;
    }
  }
  const bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && (((uintptr_t)A | (uintptr_t)B) % 16 == 0) &&
                   (transA ? (M % 4 == 0) : (K % 4 == 0)) && (transB ? (K % 4 == 0) : (N % 4 == 0));
  // split-K first (keeps the efficient 64x64 tile): only for plain accumulating products (the dW
  // shape M,N <= 256, K = rows): no bias / relu / gate, beta == 1
  int bm = 64, bn = 64;
  auto ctas = [&](int a, int b) { return (long long)scnb_cdiv(M, a) * scnb_cdiv(N, b); };
  int splits = 1;
  const bool splittable = !bias && !relu && !gate && (beta == 1.f) && K >= 128;
  if (splittable) {
    const long long c = ctas(bm, bn);
    while (splits < 32 && c * splits < 2 * 148 && K / (splits * 2) >= 32) splits *= 2;
  } else {
    // otherwise fall to 32-wide tiles while the grid would leave most of the 148 SMs idle
    if (ctas(bm, bn) < 148 && N > 32) bn = 32;
    if (ctas(bm, bn) < 74 && M > 32) bm = 32;
  }
  int k_chunk = scnb_cdiv(scnb_cdiv(K, splits), GBK) * GBK;
  splits = scnb_cdiv(K, k_chunk);
  dim3 grid(scnb_cdiv(N, bn), scnb_cdiv(M, bm), splits);
#define SCNB_TILE(BM_, BN_) \
  launch_sgemm_tile<BM_, BN_>(transA, transB, vec, grid, st, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, k_chunk)
  if (bm == 64 && bn == 64) SCNB_TILE(64, 64);
  else if (bm == 64 && bn == 32) SCNB_TILE(64, 32);
  else if (bm == 32 && bn == 64) SCNB_TILE(32, 64);
  else SCNB_TILE(32, 32);
#undef SCNB_TILE
  SCNB_CHECK_LAUNCH("sgemm");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Column sums (bias gradients).  Block = 32 columns x 32 row lanes; deterministic tree.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_colsum(const float* __restrict__ A, int M, int N, int lda, float* __restrict__ out,
                                                 int accumulate) {
  PDL_ENTRY();
  __shared__ float red[32][33];
  __shared__ float part[32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  // tall matrices (config 5: 16 384 rows): the rows are split over the gridDim.y CTAs of a thread-block cluster and the
  // partial sums meet in rank order through distributed shared memory (deterministic; one launch)
  const int CL = (int)gridDim.y;
  const int chunk = (M + CL - 1) / CL;
  const int rb = min(M, (int)blockIdx.y * chunk), re = min(M, rb + chunk);
  float s = 0.f;
  if (c < N)
    for (int r = rb + ty; r < re; r += 32) s += A[(size_t)r * lda + c];
  red[ty][tx] = s;
  __syncthreads();
  float t = 0.f;
  if (ty == 0) {
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][tx];
  }
  if (CL > 1) {
    if (ty == 0) part[tx] = t;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (ty == 0 && blockIdx.y == 0) {
      const uint32_t mine = (uint32_t)__cvta_generic_to_shared(part + tx);
      t = 0.f;
      for (int p = 0; p < CL; ++p) {
        uint32_t ra;
        float v;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(mine), "r"(p));
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
        t += v;
      }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (ty == 0 && blockIdx.y == 0 && c < N) out[c] = accumulate ? out[c] + t : t;
}

extern "C" int scnb_colsum(const float* A, int M, int N, int lda, float* out, int accumulate, void* stream) {
  SCNB_CHECK_ARG(A && out && M >= 0 && N > 0, "colsum: bad arguments");
  if (M >= 4096) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(scnb_cdiv(N, 32), 8);
    cfg.blockDim = dim3(1024);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = 8;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_colsum, A, M, N, lda, out, accumulate);
    SCNB_CHECK_LAUNCH("colsum/cluster");
    return SCNB_OK;
  }
  scnb_launch_pdl(k_colsum, dim3(scnb_cdiv(N, 32)), dim3(1024), 0, (cudaStream_t)stream, A, M, N, lda, out, accumulate);
  SCNB_CHECK_LAUNCH("colsum");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Segmented BatchNorm1d -> Dropout -> ReLU.
// Block = 8 columns x 128 row lanes (a warp reads 4 rows x 32 B); grid = ceil(H/8); the block walks
// the segments in order, so gamma/beta gradients are accumulated deterministically.
// stats layout: [n_seg][3][H] = mean, biased var, invstd.
// ---------------------------------------------------------------------------------------------
#define BN_CW 8
#define BN_RL 128

// Hidden-layer dropout stream: one Philox call per (row, 8-column group) gives eight 16-bit uniforms;
// element (r, c) keeps iff uniform[(c & 7)] >= p * 65536.  The scalar and the vector kernels share this
// definition, so forward and backward always see the same mask whichever variant runs.
__device__ __forceinline__ bool dropout_keep_h(const uint8_t* mask, int r, int c, int H, uint64_t seed, uint64_t strm, float p) {
  if (mask) return mask[(size_t)r * H + c] != 0;
  const uint64_t grp = (uint64_t)r * (uint64_t)((H + 7) >> 3) + (uint64_t)(c >> 3);
  const uint4 q = philox4x32_10(make_uint4((uint32_t)grp, (uint32_t)(grp >> 32), (uint32_t)strm, (uint32_t)(strm >> 32)),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const int j = c & 7;
  const uint32_t w = (j >> 1) == 0 ? q.x : (j >> 1) == 1 ? q.y : (j >> 1) == 2 ? q.z : q.w;
  const uint32_t u = (j & 1) ? (w >> 16) : (w & 0xffffu);
  return u >= (uint32_t)(p * 65536.0f + 0.5f);
}

// Column sums over the 128 row lanes of each of the 8 columns of a 1024-thread block
// (tid = ry*8 + cx).  A warp holds 4 row lanes x 8 columns: two xor-shuffles fold the row lanes,
// 32 per-warp partials per column meet in shared memory and every thread adds them in the same
// order (deterministic).  `red` is [BN_CW][32]; use a different array for each call between
// two __syncthreads-separated phases.
__device__ __forceinline__ float bn_block_colsum(float v, float (*red)[32], int cx) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane < BN_CW) red[lane][wid] = v;
  __syncthreads();
  const float4* r4 = reinterpret_cast<const float4*>(red[cx]);
  float out = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = r4[i];
    out += (t.x + t.y) + (t.z + t.w);
  }
  return out;
}

// Grid = (ceil(H/8), n_seg): one CTA per (8-column group, segment).
__global__ void __launch_bounds__(1024) k_bn_act_fwd(const float* __restrict__ Z, float* __restrict__ Aout,
                                                     float* __restrict__ pre_out, int H, int n_seg,
                                                     const int32_t* __restrict__ seg_off, int max_rows, int train,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     const float* __restrict__ rmean, const float* __restrict__ rvar,
                                                     float* __restrict__ stats, const uint8_t* __restrict__ keep_mask,
                                                     const uint64_t* __restrict__ rng, uint32_t stream_id, float p_drop,
                                                     int relu, float* __restrict__ partial_out,
                                                     const float* __restrict__ ext_sums, const float* __restrict__ ext_cnt) {
  // Data-parallel (cross-rank BatchNorm statistics) protocol:
  //   pass 1  partial_out != NULL : write local [seg][0]=sum z, [seg][1]=sum z^2, [n_seg*2*H + seg]=row count; no apply
  //   (all-reduce SUM of that buffer over ranks)
  //   pass 2  ext_sums != NULL    : statistics from the reduced sums / counts (ext_cnt), then apply
  __shared__ __align__(16) float red1[BN_CW][32];
  __shared__ __align__(16) float red2[BN_CW][32];
  const int cx = threadIdx.x & (BN_CW - 1), ry = threadIdx.x / BN_CW;
  const int c = blockIdx.x * BN_CW + cx;
  const int s = blockIdx.y;
  const bool cok = c < H;
  const float g = cok ? gamma[c] : 1.f, b = cok ? beta[c] : 0.f;
  uint64_t seed = 0, strm = 0;
  if (train && p_drop > 0.f && !keep_mask) {
    seed = rng[0];
    strm = rng[1] * 64ull + stream_id;
  }
  const float drop_scale = 1.0f / (1.0f - p_drop);
  const int r0 = min(seg_off[s], max_rows), r1 = min(seg_off[s + 1], max_rows);
  const int n = r1 - r0;
  float mean, invstd;
  if (train && partial_out) {
    float a1 = 0.f, a2 = 0.f;
    if (cok)
      for (int r = r0 + ry; r < r1; r += BN_RL) {
        const float z = Z[(size_t)r * H + c];
        a1 += z;
        a2 = fmaf(z, z, a2);
      }
    a1 = bn_block_colsum(a1, red1, cx);
    a2 = bn_block_colsum(a2, red2, cx);
    if (cok && ry == 0) {
      partial_out[((size_t)s * 2 + 0) * H + c] = a1;
      partial_out[((size_t)s * 2 + 1) * H + c] = a2;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) partial_out[(size_t)n_seg * 2 * H + s] = (float)n;
    return;
  }
  if (train) {
    float var;
    if (ext_sums) {
      const float ng = fmaxf(ext_cnt[s], 1.f);
      mean = cok ? ext_sums[((size_t)s * 2 + 0) * H + c] / ng : 0.f;
      var = cok ? fmaxf(ext_sums[((size_t)s * 2 + 1) * H + c] / ng - mean * mean, 0.f) : 1.f;
    } else {
      float acc = 0.f;
      if (cok)
        for (int r = r0 + ry; r < r1; r += BN_RL) acc += Z[(size_t)r * H + c];
      mean = bn_block_colsum(acc, red1, cx) / (float)max(n, 1);
      acc = 0.f;
      if (cok)
        for (int r = r0 + ry; r < r1; r += BN_RL) {
          const float d = Z[(size_t)r * H + c] - mean;
          acc = fmaf(d, d, acc);
        }
      var = bn_block_colsum(acc, red2, cx) / (float)max(n, 1);
    }
    invstd = 1.0f / sqrtf(var + BN_EPS);
    if (cok && ry == 0 && stats) {
      stats[((size_t)s * 3 + 0) * H + c] = mean;
      stats[((size_t)s * 3 + 1) * H + c] = var;
      stats[((size_t)s * 3 + 2) * H + c] = invstd;
    }
  } else {
    mean = cok ? rmean[c] : 0.f;
    invstd = cok ? 1.0f / sqrtf(rvar[c] + BN_EPS) : 1.f;
  }
  if (!cok) return;
  for (int r = r0 + ry; r < r1; r += BN_RL) {
    const size_t o = (size_t)r * H + c;
    float y = (Z[o] - mean) * invstd * g + b;
    if (train && p_drop > 0.f) y = dropout_keep_h(keep_mask, r, c, H, seed, strm, p_drop) ? y * drop_scale : 0.f;
    if (pre_out) pre_out[o] = y;
    Aout[o] = relu ? fmaxf(y, 0.f) : y;
  }
  // inactive tail rows (data-dependent DAN segment length): the last segment's CTAs clear them
  if (s == n_seg - 1)
    for (int r = r1 + ry; r < max_rows; r += BN_RL) {
      Aout[(size_t)r * H + c] = 0.f;
      if (pre_out) pre_out[(size_t)r * H + c] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// Vector form of the segmented BatchNorm -> Dropout -> ReLU kernels (H % 8 == 0, 16-byte aligned rows).
// Grid = (H/8, n_seg), 256 threads: thread t owns rows r0+t, r0+t+256, ... of its segment and, per row,
// the CTA's 8 columns (two float4 = one 32-byte sector).  One Philox call per (row, 8-column group)
// yields eight 16-bit uniforms (`keep = u16 >= p*65536`); an injected byte mask is read 8 bytes at a
// time.  Column sums: butterfly over the 32 lanes, then 8 per-warp partials per column in shared
// memory that every thread adds in the same order (broadcast reads, deterministic).
// ---------------------------------------------------------------------------------------------
#define BNV_T 256
#define BNV_W (BNV_T / 32)

struct F8 { float v[8]; };

__device__ __forceinline__ F8 ld8(const float* __restrict__ p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  F8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st8(float* __restrict__ p, const F8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
// W column groups per CTA: thread t owns the 8 columns of group (t % W) and the rows (t / W) + k * (BNV_T / W); consecutive
// lanes read consecutive 32-byte pieces of ONE row (W = 4: a full 128-byte line per 4 lanes).  W = 1 is the small-batch
// form (a CTA = one column group, every thread a different row).
// Sums over all rows of the CTA for each of the 8 columns of the thread's group; result in every thread.
template <int W>
__device__ __forceinline__ void bnv_colsum8(F8& x, float (*red)[W][8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = x.v[j];
#pragma unroll
    for (int o = 16; o >= W; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // lanes with the same (lane % W)
    x.v[j] = v;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane < W) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[wid][lane][j] = x.v[j];
  }
  __syncthreads();
  const int cg = threadIdx.x % W;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < BNV_W; ++w) t += red[w][cg][j];
    x.v[j] = t;
  }
}

// Large batches (config 5: 16 384 rows): the rows of a segment are split over the CTAs of a thread-block CLUSTER
// (cluster dims (1, 1, CL), cluster rank = blockIdx.z), so that a launch has CL times as many CTAs in flight on the same
// 32-byte column group; the per-CTA column sums meet through distributed shared memory and are added in rank order
// (deterministic).  One launch, no workspace.  With gridDim.z == 1 nothing changes.
__device__ __forceinline__ void bnv_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// a, b: this CTA's 8 + 8 block sums of the thread's column group (in every thread) -> the cluster's sums.
// xs: shared [32 * W].
template <int W>
__device__ __forceinline__ void bnv_cluster_sum16(F8& a, F8& b, float* xs) {
  const int CL = (int)gridDim.z;
  if (CL == 1) return;
  const int t = threadIdx.x;
  if (t < W) {                                   // thread t holds the sums of column group t
#pragma unroll
    for (int j = 0; j < 8; ++j) { xs[t * 16 + j] = a.v[j]; xs[t * 16 + 8 + j] = b.v[j]; }
  }
  bnv_cluster_sync();
  if (t < 16 * W) {
    const uint32_t mine_addr = (uint32_t)__cvta_generic_to_shared(xs + t);
    float acc = 0.f;
    for (int p = 0; p < CL; ++p) {
      uint32_t ra;
      float v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(mine_addr), "r"(p));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
      acc += v;
    }
    xs[16 * W + t] = acc;
  }
  bnv_cluster_sync();            // nobody overwrites (or leaves with) xs[0 .. 16 W) while a peer still reads it
  const int cg = t % W;
#pragma unroll
  for (int j = 0; j < 8; ++j) { a.v[j] = xs[16 * W + cg * 16 + j]; b.v[j] = xs[16 * W + cg * 16 + 8 + j]; }
  __syncthreads();
}
// this CTA's share [rb, re) of the rows [r0, r1) of a segment
__device__ __forceinline__ void bnv_row_share(int r0, int r1, int& rb, int& re) {
  const int CL = (int)gridDim.z;
  const int chunk = (r1 - r0 + CL - 1) / CL;
  rb = min(r1, r0 + (int)blockIdx.z * chunk);
  re = min(r1, rb + chunk);
}

// ---------------------------------------------------------------------------------------------
// Cross-rank BatchNorm statistics through peer memory (data-parallel training, one process per GPU).
// Every rank owns an exchange buffer that all ranks can address (CUDA IPC over NVLink / NVSwitch; bufs[r] is rank
// r's buffer as seen from this process).  A BatchNorm CTA owns 8 columns of one segment on every rank, so the
// all-reduce decomposes per CTA: it stores its 16 local column sums + its row count into EVERY rank's buffer and adds
// what the other ranks stored into its own, in rank order -- bit-identical statistics on all ranks, no collective
// launch, one kernel.  Each value travels as one 8-byte word {value, step number}: the word is written atomically, so
// the step number doubles as its ready flag and no fence or separate flag round-trip is needed (the "LL" protocol);
// the cost of the exchange is one NVLink store latency.  Packets are double-buffered by step parity; the gradient
// exchange that ends a step separates reuses.
//   packet(slot, parity, src, cta) : 32 words of 8 bytes = [sum_a[8] | sum_b[8] | rows | unused]
// ---------------------------------------------------------------------------------------------
// Operand planes of the tcgen05 hidden GEMMs (tc_gemm.cu, planes form) emitted by the PRODUCER of an activation / gradient
// tensor: bf16 hi / lo planes tiled [rows / 256][cols / 32], every tile a SWIZZLE_64B K-major image of 256 x 32.  A thread
// of the BatchNorm kernels owns 8 consecutive columns of a row = one 16-byte chunk of a plane row.  The engine binds plane
// buffers to the tensors (scnb_gemm_tc_bind_planes); the GEMM that reads a bound tensor skips its conversion pass.
struct PlaneSink {
  uint16_t* hi;
  uint16_t* lo;
  int n_kt;        // cols / 32
  int rows_pad;    // rows of the plane buffer (a multiple of 256): rows past the tensor's are written as zeros
};
bool scnb_planes_lookup(const float* X, int cols, int min_rows, uint16_t** hi, uint16_t** lo, int* rows_pad);   // tc_gemm.cu
__device__ __forceinline__ void plane_store8(const PlaneSink& sk, int r, int cgi, const float (&x)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_bf16x2(x[2 * j], x[2 * j + 1], h[j], l[j]);
  const uint32_t rr = (uint32_t)(r & 255);
  const size_t off = ((size_t)(r >> 8) * sk.n_kt + (size_t)(cgi >> 2)) * (256 * 32) + (size_t)rr * 32 +
                     (size_t)(swz_chunk<32>((uint32_t)(cgi & 3), rr) >> 1);
  *reinterpret_cast<uint4*>(sk.hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(sk.lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}
static PlaneSink plane_sink_of(const float* X, int cols, int rows) {
  PlaneSink sk{nullptr, nullptr, 0, 0};
  uint16_t *hi = nullptr, *lo = nullptr;
  int rp = 0;
  if (cols % 32 == 0 && scnb_planes_lookup(X, cols, rows, &hi, &lo, &rp)) sk = PlaneSink{hi, lo, cols / 32, rp};
  return sk;
}

struct PeerX {
  const unsigned long long* bufs;   // device array [world] of exchange-buffer base addresses (NULL: no exchange)
  const long long* epoch;           // device scalar: step number (> 0, changes every step)
  int rank, world, slot, spin_limit;
  int cta_stride;                   // packets reserved per (slot, parity, source rank): >= CTAs of any launch using the buffer
};

#define PEER_PKT_WORDS 32           // 8-byte words per packet

__device__ __forceinline__ void peer_allreduce17(const PeerX& px, F8& a, F8& b, float& rows, float* xs /* shared [32] */) {
  const int t = threadIdx.x;
  const unsigned e = (unsigned)px.epoch[0];
  const size_t n_cta = (size_t)px.cta_stride, cta = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  const size_t pkt0 = ((size_t)(px.slot * 2 + (int)(e & 1u)) * px.world) * n_cta + cta;   // + src * n_cta
  if (t < 17) {
    float mine = rows;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (t == j) mine = a.v[j];
      if (t == 8 + j) mine = b.v[j];
    }
    const size_t my_word = (pkt0 + (size_t)px.rank * n_cta) * PEER_PKT_WORDS + t;
    for (int p = 0; p < px.world; ++p) {
      uint2* dst = reinterpret_cast<uint2*>(px.bufs[p]) + my_word;
      asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(mine)), "r"(e) : "memory");
    }
    float acc = 0.f;
    for (int p = 0; p < px.world; ++p) {   // rank order: the same sum on every rank
      const uint2* src = reinterpret_cast<const uint2*>(px.bufs[px.rank]) + (pkt0 + (size_t)p * n_cta) * PEER_PKT_WORDS + t;
      uint32_t v, f;
      int spins = 0;
      do {
        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(src) : "memory");
        if (f != e && ++spins > px.spin_limit) __trap();   // a peer never arrived: fail loudly instead of hanging the GPU
      } while (f != e);
      acc += __uint_as_float(v);
    }
    xs[t] = acc;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) { a.v[j] = xs[j]; b.v[j] = xs[8 + j]; }
  rows = xs[16];
  __syncthreads();
}

template <int W>
__global__ void __launch_bounds__(BNV_T) k_bn_act_fwd_v8(const float* __restrict__ Z, float* __restrict__ Aout,
                                                         float* __restrict__ pre_out, int H, int n_seg,
                                                         const int32_t* __restrict__ seg_off, int max_rows,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float* __restrict__ stats, const uint8_t* __restrict__ keep_mask,
                                                         const uint64_t* __restrict__ rng, uint32_t stream_id, float p_drop,
                                                         int relu, float* __restrict__ partial_out,
                                                         const float* __restrict__ ext_sums, const float* __restrict__ ext_cnt,
                                                         PeerX px, float* __restrict__ glob_out, PlaneSink sink) {
  PDL_ENTRY();
  __shared__ __align__(16) float red1[BNV_W][W][8];
  __shared__ __align__(16) float red2[BNV_W][W][8];
  __shared__ float xs[32 * W];
  constexpr int RS = BNV_T / W;                  // rows per pass of the CTA
  const int t = threadIdx.x;
  const int tr = t / W;                          // row lane of this thread
  const int cgi = blockIdx.x * W + (t % W);      // its column group
  const int c0 = cgi * 8, s = blockIdx.y;
  const int Hg = H >> 3;
  const int r0 = min(seg_off[s], max_rows), r1 = min(seg_off[s + 1], max_rows);
  const int n = r1 - r0;
  const F8 g = ld8(gamma + c0), b = ld8(beta + c0);
  uint64_t seed = 0, strm = 0;
  if (p_drop > 0.f && !keep_mask) {
    seed = rng[0];
    strm = rng[1] * 64ull + stream_id;
  }
  const uint32_t thr16 = (uint32_t)(p_drop * 65536.0f + 0.5f);
  const float drop_scale = 1.0f / (1.0f - p_drop);
  F8 mean, invstd;
  const bool cached = (W == 1) && (n <= 2 * BNV_T) && !partial_out && !ext_sums && !px.bufs && gridDim.z == 1;
  const bool has0 = t < n, has1 = t + BNV_T < n;
  int rb, re;                      // this CTA's rows of the segment (all of them unless the launch is a cluster launch)
  bnv_row_share(r0, r1, rb, re);
  F8 zc0, zc1;
  if (partial_out) {  // data-parallel pass 1: local sums only
    F8 a1, a2;
#pragma unroll
    for (int j = 0; j < 8; ++j) a1.v[j] = a2.v[j] = 0.f;
    for (int r = r0 + t; r < r1; r += BNV_T) {
      const F8 z = ld8(Z + (size_t)r * H + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) { a1.v[j] += z.v[j]; a2.v[j] = fmaf(z.v[j], z.v[j], a2.v[j]); }
    }
    bnv_colsum8<W>(a1, red1);
    bnv_colsum8<W>(a2, red2);
    if (t < 8) {
      partial_out[((size_t)s * 2 + 0) * H + c0 + t] = a1.v[t];
      partial_out[((size_t)s * 2 + 1) * H + c0 + t] = a2.v[t];
    }
    if (blockIdx.x == 0 && t == 0) partial_out[(size_t)n_seg * 2 * H + s] = (float)n;
    return;
  }
  F8 var;
  if (px.bufs) {  // data parallel, one launch: local sums -> peer-memory all-reduce -> global statistics
    F8 a1, a2;
#pragma unroll
    for (int j = 0; j < 8; ++j) a1.v[j] = a2.v[j] = 0.f;
    for (int r = r0 + t; r < r1; r += BNV_T) {
      const F8 z = ld8(Z + (size_t)r * H + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) { a1.v[j] += z.v[j]; a2.v[j] = fmaf(z.v[j], z.v[j], a2.v[j]); }
    }
    bnv_colsum8<W>(a1, red1);
    bnv_colsum8<W>(a2, red2);
    float rows = (float)n;
    peer_allreduce17(px, a1, a2, rows, xs);
    if (glob_out) {   // [seg][2][H] global sums + [seg] global rows: what the running-statistics update reads
      if (t < 8) {
        float u = 0.f, w = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j == t) { u = a1.v[j]; w = a2.v[j]; }
        glob_out[((size_t)s * 2 + 0) * H + c0 + t] = u;
        glob_out[((size_t)s * 2 + 1) * H + c0 + t] = w;
      }
      if (blockIdx.x == 0 && t == 0) glob_out[(size_t)n_seg * 2 * H + s] = rows;
    }
    const float ng = fmaxf(rows, 1.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean.v[j] = a1.v[j] / ng;
      var.v[j] = fmaxf(a2.v[j] / ng - mean.v[j] * mean.v[j], 0.f);
    }
  } else if (ext_sums) {
    const float ng = fmaxf(ext_cnt[s], 1.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean.v[j] = ext_sums[((size_t)s * 2 + 0) * H + c0 + j] / ng;
      var.v[j] = fmaxf(ext_sums[((size_t)s * 2 + 1) * H + c0 + j] / ng - mean.v[j] * mean.v[j], 0.f);
    }
  } else if (W > 1 || gridDim.z > 1) {
    // Large batches: ONE statistics pass over Z (sums of d = z - z0 and d^2, z0 = the segment's first row: a sample of the
    // column, so E[d]^2 is of the order of the variance and the subtraction below does not cancel) instead of the two
    // passes of the centred form -- Z is read twice per launch, not three times.
    const float inv_n = 1.0f / (float)max(n, 1);
    F8 z0, a1, a2;
#pragma unroll
    for (int j = 0; j < 8; ++j) z0.v[j] = a1.v[j] = a2.v[j] = 0.f;
    if (n > 0) z0 = ld8(Z + (size_t)r0 * H + c0);
    {
      // four rows per trip, all four loads issued before the first use: with one 32-byte load in flight per thread the
      // pass was latency-bound at 2.3 TB/s (ncu: DRAM 28 %, long scoreboard); the order of the additions is unchanged
      int r = rb + tr;
      for (; r + 3 * RS < re; r += 4 * RS) {
        F8 zq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) zq[u] = ld8(Z + (size_t)(r + u * RS) * H + c0);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = zq[u].v[j] - z0.v[j]; a1.v[j] += d; a2.v[j] = fmaf(d, d, a2.v[j]); }
        }
      }
      for (; r < re; r += RS) {
        const F8 z = ld8(Z + (size_t)r * H + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = z.v[j] - z0.v[j]; a1.v[j] += d; a2.v[j] = fmaf(d, d, a2.v[j]); }
      }
    }
    bnv_colsum8<W>(a1, red1);
    bnv_colsum8<W>(a2, red2);
    bnv_cluster_sum16<W>(a1, a2, xs);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m = a1.v[j] * inv_n;
      mean.v[j] = z0.v[j] + m;
      var.v[j] = fmaxf(a2.v[j] * inv_n - m * m, 0.f);
    }
  } else {
    const float inv_n = 1.0f / (float)max(n, 1);
#pragma unroll
    for (int j = 0; j < 8; ++j) mean.v[j] = 0.f;
    if (cached) {   // <= 2 rows per thread: Z is read from global memory once and stays in registers for all passes
      if (has0) zc0 = ld8(Z + (size_t)(r0 + t) * H + c0);
      if (has1) zc1 = ld8(Z + (size_t)(r0 + t + BNV_T) * H + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) mean.v[j] = (has0 ? zc0.v[j] : 0.f) + (has1 ? zc1.v[j] : 0.f);
    } else {
      for (int r = rb + tr; r < re; r += RS) {
        const F8 z = ld8(Z + (size_t)r * H + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) mean.v[j] += z.v[j];
      }
    }
    bnv_colsum8<W>(mean, red1);
    {
      F8 dummy;
#pragma unroll
      for (int j = 0; j < 8; ++j) dummy.v[j] = 0.f;
      bnv_cluster_sum16<W>(mean, dummy, xs);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { mean.v[j] *= inv_n; var.v[j] = 0.f; }
    if (cached) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d0 = has0 ? zc0.v[j] - mean.v[j] : 0.f, d1 = has1 ? zc1.v[j] - mean.v[j] : 0.f;
        var.v[j] = fmaf(d0, d0, d1 * d1);
      }
    } else {
      for (int r = rb + tr; r < re; r += RS) {
        const F8 z = ld8(Z + (size_t)r * H + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = z.v[j] - mean.v[j]; var.v[j] = fmaf(d, d, var.v[j]); }
      }
    }
    bnv_colsum8<W>(var, red2);
    {
      F8 dummy;
#pragma unroll
      for (int j = 0; j < 8; ++j) dummy.v[j] = 0.f;
      bnv_cluster_sum16<W>(var, dummy, xs);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) var.v[j] *= inv_n;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) invstd.v[j] = 1.0f / sqrtf(var.v[j] + BN_EPS);
  if (stats && t < 8 * W && blockIdx.z == 0) {
    // thread (group t % W, column t / W of the group); dynamic register indexing would spill: compile-time unrolled select
    float m = 0.f, v = 0.f, is = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) if (j == tr) { m = mean.v[j]; v = var.v[j]; is = invstd.v[j]; }
    stats[((size_t)s * 3 + 0) * H + c0 + tr] = m;
    stats[((size_t)s * 3 + 1) * H + c0 + tr] = v;
    stats[((size_t)s * 3 + 2) * H + c0 + tr] = is;
  }
  auto apply_row = [&](int r, const F8& z) {
    const size_t o = (size_t)r * H + c0;
    uint32_t kb = 0xffu;
    if (p_drop > 0.f) kb = keep8(keep_mask, o, (uint64_t)r * Hg + cgi, seed, strm, thr16);
    F8 y;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float yy = (z.v[j] - mean.v[j]) * invstd.v[j] * g.v[j] + b.v[j];
      if (p_drop > 0.f) yy = ((kb >> j) & 1u) ? yy * drop_scale : 0.f;
      y.v[j] = yy;
    }
    if (pre_out) st8(pre_out + o, y);
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y.v[j] = fmaxf(y.v[j], 0.f);
    }
    st8(Aout + o, y);
    if (sink.hi) plane_store8(sink, r, cgi, y.v);
  };
  if (cached) {
    for (int r = rb + tr; r < re; r += RS) apply_row(r, r == r0 + t ? zc0 : zc1);
  } else {
    int r = rb + tr;
    for (; r + RS < re; r += 2 * RS) {       // two rows per trip, both loads in flight before the first use
      const F8 za = ld8(Z + (size_t)r * H + c0), zb = ld8(Z + (size_t)(r + RS) * H + c0);
      apply_row(r, za);
      apply_row(r + RS, zb);
    }
    if (r < re) apply_row(r, ld8(Z + (size_t)r * H + c0));
  }
  if (sink.hi && s == n_seg - 1) {
    const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = r1 + tr + (int)blockIdx.z * RS; r < sink.rows_pad; r += RS * (int)gridDim.z) plane_store8(sink, r, cgi, zero8);
  }
  if (s == n_seg - 1) {  // inactive tail rows (data-dependent DAN segment length)
    F8 zero;
#pragma unroll
    for (int j = 0; j < 8; ++j) zero.v[j] = 0.f;
    for (int r = r1 + tr + (int)blockIdx.z * RS; r < max_rows; r += RS * (int)gridDim.z) {
      st8(Aout + (size_t)r * H + c0, zero);
      if (pre_out) st8(pre_out + (size_t)r * H + c0, zero);
    }
  }
}

template <int W>
__global__ void __launch_bounds__(BNV_T) k_bn_act_bwd_v8(const float* __restrict__ dA, const float* __restrict__ Z,
                                                         const float* __restrict__ Aact, int H, int n_seg,
                                                         const int32_t* __restrict__ seg_off, int max_rows,
                                                         const float* __restrict__ gamma, const float* __restrict__ stats,
                                                         const uint8_t* __restrict__ keep_mask, const uint64_t* __restrict__ rng,
                                                         uint32_t stream_id, float p_drop, int relu, float* __restrict__ dZ,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                         float* __restrict__ partial_out, const float* __restrict__ ext_sums,
                                                         const float* __restrict__ ext_cnt, PeerX px, PlaneSink sink) {
  PDL_ENTRY();
  __shared__ __align__(16) float red1[BNV_W][W][8];
  __shared__ __align__(16) float red2[BNV_W][W][8];
  __shared__ float xs[32 * W];
  constexpr int RS = BNV_T / W;                  // rows per pass of the CTA
  const int t = threadIdx.x;
  const int tr = t / W;                          // row lane of this thread
  const int cgi = blockIdx.x * W + (t % W);      // its column group
  const int c0 = cgi * 8, s = blockIdx.y;
  const int Hg = H >> 3;
  const int r0 = min(seg_off[s], max_rows), r1 = min(seg_off[s + 1], max_rows);
  const int n = r1 - r0;
  const F8 g = ld8(gamma + c0);
  const F8 mean = ld8(stats + ((size_t)s * 3 + 0) * H + c0), invstd = ld8(stats + ((size_t)s * 3 + 2) * H + c0);
  uint64_t seed = 0, strm = 0;
  if (p_drop > 0.f && !keep_mask) {
    seed = rng[0];
    strm = rng[1] * 64ull + stream_id;
  }
  const uint32_t thr16 = (uint32_t)(p_drop * 65536.0f + 0.5f);
  const float drop_scale = 1.0f / (1.0f - p_drop);
  F8 s1, s2;
#pragma unroll
  for (int j = 0; j < 8; ++j) s1.v[j] = s2.v[j] = 0.f;
  float inv_n = 1.0f / (float)max(n, 1);
  const bool cached = (W == 1) && (n <= 2 * BNV_T) && !ext_sums && !partial_out && gridDim.z == 1;
  int rb, re;                      // this CTA's rows of the segment (all of them unless the launch is a cluster launch)
  bnv_row_share(r0, r1, rb, re);
  F8 dyc0, zhc0, dyc1, zhc1;
  if (ext_sums) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1.v[j] = ext_sums[((size_t)s * 2 + 0) * H + c0 + j];
      s2.v[j] = ext_sums[((size_t)s * 2 + 1) * H + c0 + j];
    }
    inv_n = 1.0f / fmaxf(ext_cnt[s], 1.f);
  } else {
    int slot = 0;
    for (int r = rb + tr; r < re; r += RS, ++slot) {
      const size_t o = (size_t)r * H + c0;
      const F8 da = ld8(dA + o), z = ld8(Z + o);
      F8 av;
      if (relu) av = ld8(Aact + o);
      uint32_t kb = 0xffu;
      if (p_drop > 0.f) kb = keep8(keep_mask, o, (uint64_t)r * Hg + cgi, seed, strm, thr16);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dy = da.v[j];
        if (relu && !(av.v[j] > 0.f)) dy = 0.f;
        if (p_drop > 0.f) dy = ((kb >> j) & 1u) ? dy * drop_scale : 0.f;
        const float zh = (z.v[j] - mean.v[j]) * invstd.v[j];
        s1.v[j] += dy;
        s2.v[j] = fmaf(dy, zh, s2.v[j]);
        if (cached) {   // <= 2 rows per thread: dy and zhat stay in registers for the second pass
          if (slot == 0) { dyc0.v[j] = dy; zhc0.v[j] = zh; } else { dyc1.v[j] = dy; zhc1.v[j] = zh; }
        }
      }
    }
    bnv_colsum8<W>(s1, red1);
    bnv_colsum8<W>(s2, red2);
    bnv_cluster_sum16<W>(s1, s2, xs);
    if (t < 8 * W && blockIdx.z == 0) {         // thread (group t % W, column t / W of the group)
      float a = 0.f, b2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j == tr) { a = s1.v[j]; b2 = s2.v[j]; }
      if (n > 0) {
        atomicAdd(dgamma + c0 + tr, b2);
        atomicAdd(dbeta + c0 + tr, a);
      }
      if (partial_out) {
        partial_out[((size_t)s * 2 + 0) * H + c0 + tr] = a;
        partial_out[((size_t)s * 2 + 1) * H + c0 + tr] = b2;
      }
    }
    if (partial_out) {
      if (blockIdx.x == 0 && t == 0) partial_out[(size_t)n_seg * 2 * H + s] = (float)n;
      return;
    }
    if (px.bufs) {   // data parallel, one launch: the two column sums and the row count become global here
      float rows = (float)n;
      peer_allreduce17(px, s1, s2, rows, xs);
      inv_n = 1.0f / fmaxf(rows, 1.f);
    }
  }
  if (cached) {
    int slot = 0;
    for (int r = r0 + t; r < r1; r += BNV_T, ++slot) {
      F8 out;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dy = slot == 0 ? dyc0.v[j] : dyc1.v[j], zh = slot == 0 ? zhc0.v[j] : zhc1.v[j];
        out.v[j] = g.v[j] * invstd.v[j] * (dy - s1.v[j] * inv_n - zh * (s2.v[j] * inv_n));
      }
      st8(dZ + (size_t)r * H + c0, out);
      if (sink.hi) plane_store8(sink, r, cgi, out.v);
    }
  } else {
    for (int r = rb + tr; r < re; r += RS) {
      const size_t o = (size_t)r * H + c0;
      const F8 da = ld8(dA + o), z = ld8(Z + o);
      F8 av;
      if (relu) av = ld8(Aact + o);
      uint32_t kb = 0xffu;
      if (p_drop > 0.f) kb = keep8(keep_mask, o, (uint64_t)r * Hg + cgi, seed, strm, thr16);
      F8 out;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dy = da.v[j];
        if (relu && !(av.v[j] > 0.f)) dy = 0.f;
        if (p_drop > 0.f) dy = ((kb >> j) & 1u) ? dy * drop_scale : 0.f;
        const float zh = (z.v[j] - mean.v[j]) * invstd.v[j];
        out.v[j] = g.v[j] * invstd.v[j] * (dy - s1.v[j] * inv_n - zh * (s2.v[j] * inv_n));
      }
      st8(dZ + o, out);
      if (sink.hi) plane_store8(sink, r, cgi, out.v);
    }
  }
  if (sink.hi && s == n_seg - 1) {
    const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = r1 + tr + (int)blockIdx.z * RS; r < sink.rows_pad; r += RS * (int)gridDim.z) plane_store8(sink, r, cgi, zero8);
  }
  if (s == n_seg - 1) {
    F8 zero;
#pragma unroll
    for (int j = 0; j < 8; ++j) zero.v[j] = 0.f;
    for (int r = r1 + tr + (int)blockIdx.z * RS; r < max_rows; r += RS * (int)gridDim.z) st8(dZ + (size_t)r * H + c0, zero);
  }
}

// The vector kernels need whole 32-byte column groups and 16-byte aligned rows (the dropout stream is
// defined per 8-column group there, so forward and backward must take the same decision: it depends
// only on H and on pointer alignment, which the caller keeps identical between the two calls).
// cluster launch with dims (1, 1, grid.z)
template <typename... P, typename... A>
static inline cudaError_t bnv_launch_cluster(void (*kern)(P...), dim3 grid, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(BNV_T);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = grid.z;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}
// rows of a segment are split over a cluster from this many rows on (SCNB_BN_CLUSTER_ROWS overrides; 0 = never)
static int bnv_cluster_size(int max_rows) {
  static int thr = -1;
  if (thr < 0) {
    const char* e = getenv("SCNB_BN_CLUSTER_ROWS");
    thr = e ? atoi(e) : 4096;
  }
  if (thr <= 0 || max_rows < thr) return 1;
  return max_rows >= 2 * thr ? 8 : 4;
}

static bool bn_vec8_ok(int H, const void* a, const void* b, const void* c, const void* d, const void* e, const void* f,
                       const void* mask) {
  const uintptr_t bits = (uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d | (uintptr_t)e | (uintptr_t)f;
  return (H % 8 == 0) && (bits % 16 == 0) && ((uintptr_t)mask % 8 == 0);
}

// Eval mode (running statistics, no dropout) is purely elementwise: a flat grid-stride kernel over
// rows [0, seg_off[n_seg]) so that predict-sized inputs (65 536 rows) use the whole chip.
template <bool VEC4>
__global__ void __launch_bounds__(256) k_bn_eval_fwd(const float* __restrict__ Z, float* __restrict__ Aout,
                                                     float* __restrict__ pre_out, int H, const int32_t* __restrict__ seg_off,
                                                     int n_seg, int max_rows, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, const float* __restrict__ rmean,
                                                     const float* __restrict__ rvar, int relu) {
  PDL_ENTRY();
  const int rows = min(seg_off[n_seg], max_rows);
  if (VEC4) {
    const int Hv = H >> 2;
    const size_t total = (size_t)max_rows * Hv, active = (size_t)rows * Hv;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < active) {
        const int cv = (int)(i % Hv);
        const float4 z = reinterpret_cast<const float4*>(Z)[i];
        const float4 m = reinterpret_cast<const float4*>(rmean)[cv], v = reinterpret_cast<const float4*>(rvar)[cv];
        const float4 g = reinterpret_cast<const float4*>(gamma)[cv], b = reinterpret_cast<const float4*>(beta)[cv];
        y.x = (z.x - m.x) * (1.0f / sqrtf(v.x + BN_EPS)) * g.x + b.x;
        y.y = (z.y - m.y) * (1.0f / sqrtf(v.y + BN_EPS)) * g.y + b.y;
        y.z = (z.z - m.z) * (1.0f / sqrtf(v.z + BN_EPS)) * g.z + b.z;
        y.w = (z.w - m.w) * (1.0f / sqrtf(v.w + BN_EPS)) * g.w + b.w;
      }
      if (pre_out) reinterpret_cast<float4*>(pre_out)[i] = y;
      if (relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
      reinterpret_cast<float4*>(Aout)[i] = y;
    }
  } else {
    const size_t total = (size_t)max_rows * H, active = (size_t)rows * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
      float y = 0.f;
      if (i < active) {
        const int c = (int)(i % H);
        y = (Z[i] - rmean[c]) * (1.0f / sqrtf(rvar[c] + BN_EPS)) * gamma[c] + beta[c];
      }
      if (pre_out) pre_out[i] = y;
      Aout[i] = relu ? fmaxf(y, 0.f) : y;
    }
  }
}

extern "C" int scnb_bn_act_fwd(const float* Z, float* A, float* pre_out, int H, int n_seg, const int32_t* seg_off_dev,
                               int max_rows, int train, const float* gamma, const float* beta, const float* running_mean,
                               const float* running_var, float* stats, const uint8_t* keep_mask, const uint64_t* rng,
                               uint32_t stream_id, float p_drop, int relu, float* dp_partial_out, const float* dp_sums,
                               void* stream) {
  SCNB_CHECK_ARG(Z && A && gamma && beta && seg_off_dev && H > 0 && n_seg > 0 && max_rows >= 0, "bn_act_fwd: bad arguments");
  SCNB_CHECK_ARG(train || (running_mean && running_var), "bn_act_fwd: eval mode needs running statistics");
  SCNB_CHECK_ARG(!(train && p_drop > 0.f) || keep_mask || rng, "bn_act_fwd: dropout needs keep_mask or rng state");
  SCNB_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "bn_act_fwd: p_drop out of range");
  SCNB_CHECK_ARG(!(dp_partial_out && dp_sums), "bn_act_fwd: partial and reduced statistics are exclusive");
  const float* ext_cnt = dp_sums ? dp_sums + (size_t)n_seg * 2 * H : nullptr;
  const PlaneSink sink = plane_sink_of(A, H, max_rows);
#define PLANE_SINK_ARG sink
  SCNB_CHECK_ARG(!sink.hi || (train && bn_vec8_ok(H, Z, A, pre_out, gamma, beta, stats, keep_mask)),
                 "bn_act_fwd: operand planes are bound to this output, but this path (eval mode / unaligned) does not emit them");
  if (!train) {
    if (max_rows == 0) return SCNB_OK;
    const bool v4 = (H % 4 == 0) && ((((uintptr_t)Z | (uintptr_t)A | (uintptr_t)pre_out | (uintptr_t)gamma | (uintptr_t)beta |
                                       (uintptr_t)running_mean | (uintptr_t)running_var) % 16) == 0);
    const size_t work = (size_t)max_rows * (v4 ? H / 4 : H);
    const int blocks = (int)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
    if (v4)
      scnb_launch_pdl(k_bn_eval_fwd<true>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, Z, A, pre_out, H, seg_off_dev, n_seg, max_rows, gamma, beta,
                                                                    running_mean, running_var, relu);
    else
      k_bn_eval_fwd<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(Z, A, pre_out, H, seg_off_dev, n_seg, max_rows, gamma, beta,
                                                                     running_mean, running_var, relu);
    SCNB_CHECK_LAUNCH("bn_act_fwd/eval");
    return SCNB_OK;
  }
  if (bn_vec8_ok(H, Z, A, pre_out, gamma, beta, stats, keep_mask)) {
    dim3 gridv(H / 8, n_seg);
    const int cl = (dp_partial_out || dp_sums) ? 1 : bnv_cluster_size(max_rows);
    if (cl > 1) {
      gridv.z = cl;
      if (H % 32 == 0) {   // four column groups per CTA: every 4 lanes read a full 128-byte line of one row
        gridv.x = H / 32;
        bnv_launch_cluster(k_bn_act_fwd_v8<4>, gridv, (cudaStream_t)stream, Z, A, pre_out, H, n_seg, seg_off_dev, max_rows, gamma, beta, stats,
                           keep_mask, rng, stream_id, p_drop, relu, dp_partial_out, dp_sums, ext_cnt,
                           PeerX{nullptr, nullptr, 0, 1, 0, 0, 0}, (float*)nullptr, PLANE_SINK_ARG);
      } else
      bnv_launch_cluster(k_bn_act_fwd_v8<1>, gridv, (cudaStream_t)stream, Z, A, pre_out, H, n_seg, seg_off_dev, max_rows, gamma, beta, stats,
                         keep_mask, rng, stream_id, p_drop, relu, dp_partial_out, dp_sums, ext_cnt,
                         PeerX{nullptr, nullptr, 0, 1, 0, 0, 0}, (float*)nullptr, PLANE_SINK_ARG);
      SCNB_CHECK_LAUNCH("bn_act_fwd/v8 cluster");
      return SCNB_OK;
    }
    scnb_launch_pdl(k_bn_act_fwd_v8<1>, gridv, dim3(BNV_T), 0, (cudaStream_t)stream, Z, A, pre_out, H, n_seg, seg_off_dev, max_rows, gamma, beta, stats,
                                                               keep_mask, rng, stream_id, p_drop, relu, dp_partial_out, dp_sums,
                                                               ext_cnt, PeerX{nullptr, nullptr, 0, 1, 0, 0, 0}, (float*)nullptr, PLANE_SINK_ARG);
    SCNB_CHECK_LAUNCH("bn_act_fwd/v8");
    return SCNB_OK;
  }
  dim3 grid(scnb_cdiv(H, BN_CW), n_seg);
  k_bn_act_fwd<<<grid, 1024, 0, (cudaStream_t)stream>>>(Z, A, pre_out, H, n_seg, seg_off_dev, max_rows, train, gamma, beta,
                                                         running_mean, running_var, stats, keep_mask, rng, stream_id, p_drop,
                                                         relu, dp_partial_out, dp_sums, ext_cnt);
  SCNB_CHECK_LAUNCH("bn_act_fwd");
  return SCNB_OK;
#undef PLANE_SINK_ARG
}

// Backward (train mode).  dA is the gradient w.r.t. the post-ReLU activation A.
//   dd = dA*[A>0]; dy = dd*keep/(1-p); dbeta += sum dy; dgamma += sum dy*zhat;
//   dz = gamma*invstd*(dy - mean(dy) - zhat*mean(dy*zhat))
// Grid = (ceil(H/8), n_seg); dgamma/dbeta receive one atomic add per segment.
__global__ void __launch_bounds__(1024) k_bn_act_bwd(const float* __restrict__ dA, const float* __restrict__ Z,
                                                     const float* __restrict__ Aact, int H, int n_seg,
                                                     const int32_t* __restrict__ seg_off, int max_rows,
                                                     const float* __restrict__ gamma, const float* __restrict__ stats,
                                                     const uint8_t* __restrict__ keep_mask, const uint64_t* __restrict__ rng,
                                                     uint32_t stream_id, float p_drop, int relu, float* __restrict__ dZ,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     float* __restrict__ partial_out, const float* __restrict__ ext_sums,
                                                     const float* __restrict__ ext_cnt) {
  // Data-parallel protocol as in the forward: pass 1 (partial_out) writes the local column sums
  // [seg][0]=sum dy, [seg][1]=sum dy*zhat and accumulates the local dgamma/dbeta; after the
  // all-reduce, pass 2 (ext_sums) forms dZ from the global sums and global row counts.
  __shared__ __align__(16) float red1[BN_CW][32];
  __shared__ __align__(16) float red2[BN_CW][32];
  const int cx = threadIdx.x & (BN_CW - 1), ry = threadIdx.x / BN_CW;
  const int c = blockIdx.x * BN_CW + cx;
  const int s = blockIdx.y;
  const bool cok = c < H;
  const float g = cok ? gamma[c] : 1.f;
  uint64_t seed = 0, strm = 0;
  if (p_drop > 0.f && !keep_mask) {
    seed = rng[0];
    strm = rng[1] * 64ull + stream_id;
  }
  const float drop_scale = 1.0f / (1.0f - p_drop);
  const int r0 = min(seg_off[s], max_rows), r1 = min(seg_off[s + 1], max_rows);
  const int n = r1 - r0;
  const float mean = cok ? stats[((size_t)s * 3 + 0) * H + c] : 0.f;
  const float invstd = cok ? stats[((size_t)s * 3 + 2) * H + c] : 1.f;
  float s1 = 0.f, s2 = 0.f;
  float inv_n = 1.0f / (float)max(n, 1);
  if (ext_sums) {
    s1 = cok ? ext_sums[((size_t)s * 2 + 0) * H + c] : 0.f;
    s2 = cok ? ext_sums[((size_t)s * 2 + 1) * H + c] : 0.f;
    inv_n = 1.0f / fmaxf(ext_cnt[s], 1.f);
  } else {
    if (cok)
      for (int r = r0 + ry; r < r1; r += BN_RL) {
        const size_t o = (size_t)r * H + c;
        float dy = dA[o];
        if (relu && !(Aact[o] > 0.f)) dy = 0.f;
        if (p_drop > 0.f) dy = dropout_keep_h(keep_mask, r, c, H, seed, strm, p_drop) ? dy * drop_scale : 0.f;
        s1 += dy;
        s2 = fmaf(dy, (Z[o] - mean) * invstd, s2);
      }
    s1 = bn_block_colsum(s1, red1, cx);
    s2 = bn_block_colsum(s2, red2, cx);
    if (cok && ry == 0 && n > 0) {
      atomicAdd(dgamma + c, s2);
      atomicAdd(dbeta + c, s1);
    }
    if (partial_out) {
      if (cok && ry == 0) {
        partial_out[((size_t)s * 2 + 0) * H + c] = s1;
        partial_out[((size_t)s * 2 + 1) * H + c] = s2;
      }
      if (blockIdx.x == 0 && threadIdx.x == 0) partial_out[(size_t)n_seg * 2 * H + s] = (float)n;
      return;
    }
  }
  if (!cok) return;
  const float m1 = s1 * inv_n, m2 = s2 * inv_n, k = g * invstd;
  for (int r = r0 + ry; r < r1; r += BN_RL) {
    const size_t o = (size_t)r * H + c;
    float dy = dA[o];
    if (relu && !(Aact[o] > 0.f)) dy = 0.f;
    if (p_drop > 0.f) dy = dropout_keep_h(keep_mask, r, c, H, seed, strm, p_drop) ? dy * drop_scale : 0.f;
    const float zh = (Z[o] - mean) * invstd;
    dZ[o] = k * (dy - m1 - zh * m2);
  }
  if (s == n_seg - 1)
    for (int r = r1 + ry; r < max_rows; r += BN_RL) dZ[(size_t)r * H + c] = 0.f;
}

extern "C" int scnb_bn_act_bwd(const float* dA, const float* Z, const float* A, int H, int n_seg, const int32_t* seg_off_dev,
                               int max_rows, const float* gamma, const float* stats, const uint8_t* keep_mask,
                               const uint64_t* rng, uint32_t stream_id, float p_drop, int relu, float* dZ, float* dgamma,
                               float* dbeta, float* dp_partial_out, const float* dp_sums, void* stream) {
  SCNB_CHECK_ARG(dA && Z && A && gamma && stats && dZ && dgamma && dbeta && seg_off_dev && H > 0 && n_seg > 0,
                 "bn_act_bwd: bad arguments");
  SCNB_CHECK_ARG(!(p_drop > 0.f) || keep_mask || rng, "bn_act_bwd: dropout needs keep_mask or rng state");
  SCNB_CHECK_ARG(!(dp_partial_out && dp_sums), "bn_act_bwd: partial and reduced sums are exclusive");
  const float* ext_cnt = dp_sums ? dp_sums + (size_t)n_seg * 2 * H : nullptr;
  const PlaneSink sink = plane_sink_of(dZ, H, max_rows);
#define PLANE_SINK_ARG sink
  SCNB_CHECK_ARG(!sink.hi || bn_vec8_ok(H, dA, Z, A, gamma, dZ, stats, keep_mask),
                 "bn_act_bwd: operand planes are bound to this output, but the unaligned path does not emit them");
  if (bn_vec8_ok(H, dA, Z, A, gamma, dZ, stats, keep_mask)) {
    dim3 gridv(H / 8, n_seg);
    const int cl = (dp_partial_out || dp_sums) ? 1 : bnv_cluster_size(max_rows);
    if (cl > 1) {
      gridv.z = cl;
      if (H % 32 == 0) {
        gridv.x = H / 32;
        bnv_launch_cluster(k_bn_act_bwd_v8<4>, gridv, (cudaStream_t)stream, dA, Z, A, H, n_seg, seg_off_dev, max_rows, gamma, stats, keep_mask,
                           rng, stream_id, p_drop, relu, dZ, dgamma, dbeta, dp_partial_out, dp_sums, ext_cnt,
                           PeerX{nullptr, nullptr, 0, 1, 0, 0, 0}, PLANE_SINK_ARG);
      } else
      bnv_launch_cluster(k_bn_act_bwd_v8<1>, gridv, (cudaStream_t)stream, dA, Z, A, H, n_seg, seg_off_dev, max_rows, gamma, stats, keep_mask,
                         rng, stream_id, p_drop, relu, dZ, dgamma, dbeta, dp_partial_out, dp_sums, ext_cnt,
                         PeerX{nullptr, nullptr, 0, 1, 0, 0, 0}, PLANE_SINK_ARG);
      SCNB_CHECK_LAUNCH("bn_act_bwd/v8 cluster");
      return SCNB_OK;
    }
    scnb_launch_pdl(k_bn_act_bwd_v8<1>, gridv, dim3(BNV_T), 0, (cudaStream_t)stream, dA, Z, A, H, n_seg, seg_off_dev, max_rows, gamma, stats, keep_mask,
                                                               rng, stream_id, p_drop, relu, dZ, dgamma, dbeta, dp_partial_out,
                                                               dp_sums, ext_cnt, PeerX{nullptr, nullptr, 0, 1, 0, 0, 0}, PLANE_SINK_ARG);
    SCNB_CHECK_LAUNCH("bn_act_bwd/v8");
    return SCNB_OK;
  }
  dim3 grid(scnb_cdiv(H, BN_CW), n_seg);
  k_bn_act_bwd<<<grid, 1024, 0, (cudaStream_t)stream>>>(dA, Z, A, H, n_seg, seg_off_dev, max_rows, gamma, stats, keep_mask, rng,
                                                         stream_id, p_drop, relu, dZ, dgamma, dbeta, dp_partial_out, dp_sums,
                                                         ext_cnt);
  SCNB_CHECK_LAUNCH("bn_act_bwd");
  return SCNB_OK;
#undef PLANE_SINK_ARG
}

// ---- data-parallel one-launch forms: statistics all-reduced through peer memory inside the kernel ----
static int peer_check(const char* who, const unsigned long long* peer_bufs, int rank, int world, int slot, int n_slots,
                      const long long* epoch, long long buf_bytes, int H, int n_seg, int* cta_stride) {
  SCNB_CHECK_ARG(peer_bufs && epoch && world >= 1 && world <= 32 && rank >= 0 && rank < world && slot >= 0 && slot < n_slots,
                 "%s: bad peer arguments", who);
  // the buffer is cut into n_slots * 2 (step parity) * world (source rank) equal runs of packets
  *cta_stride = (int)(buf_bytes / ((long long)n_slots * 2 * world * PEER_PKT_WORDS * 8));
  SCNB_CHECK_ARG(*cta_stride >= (H / 8) * n_seg, "%s: exchange buffer too small (%lld bytes: %d packets per run, %d needed)", who,
                 buf_bytes, *cta_stride, (H / 8) * n_seg);
  return SCNB_OK;
}

static int peer_spin_limit() {
  static int v = 0;
  if (!v) {
    const char* ev = getenv("SCNB_PEER_SPIN_LIMIT");
    v = ev ? atoi(ev) : (1 << 24);   // ~10 s of polling before the kernel traps
    if (v <= 0) v = 1 << 24;
  }
  return v;
}

extern "C" int scnb_bn_peer_bytes(int H, int n_seg, int n_slots, int world) {
  return (int)((size_t)n_slots * 2 * world * (size_t)((H + 7) / 8) * n_seg * PEER_PKT_WORDS * 8);
}

extern "C" int scnb_bn_act_fwd_peer(const float* Z, float* A, float* pre_out, int H, int n_seg, const int32_t* seg_off_dev,
                                    int max_rows, const float* gamma, const float* beta, float* stats, const uint8_t* keep_mask,
                                    const uint64_t* rng, uint32_t stream_id, float p_drop, int relu, const uint64_t* peer_bufs,
                                    int rank, int world, int slot, int n_slots, const int64_t* epoch, long long buf_bytes,
                                    float* global_sums_out, void* stream) {
  SCNB_CHECK_ARG(Z && A && gamma && beta && seg_off_dev && stats && H > 0 && n_seg > 0 && max_rows >= 0, "bn_act_fwd_peer: bad arguments");
  SCNB_CHECK_ARG(!(p_drop > 0.f) || keep_mask || rng, "bn_act_fwd_peer: dropout needs keep_mask or rng state");
  SCNB_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "bn_act_fwd_peer: p_drop out of range");
  SCNB_CHECK_ARG(bn_vec8_ok(H, Z, A, pre_out, gamma, beta, stats, keep_mask),
                 "bn_act_fwd_peer: needs H %% 8 == 0 and 16-byte aligned buffers");
  int stride = 0;
  int rc = peer_check("bn_act_fwd_peer", (const unsigned long long*)peer_bufs, rank, world, slot, n_slots, (const long long*)epoch,
                      buf_bytes, H, n_seg, &stride);
  if (rc != SCNB_OK) return rc;
  SCNB_CHECK_ARG(!plane_sink_of(A, H, max_rows).hi, "bn_act_fwd_peer: operand planes are bound to this output; the peer form does not emit them");
  PeerX px{(const unsigned long long*)peer_bufs, (const long long*)epoch, rank, world, slot, peer_spin_limit(), stride};
  k_bn_act_fwd_v8<1><<<dim3(H / 8, n_seg), BNV_T, 0, (cudaStream_t)stream>>>(Z, A, pre_out, H, n_seg, seg_off_dev, max_rows, gamma, beta, stats,
                                                                         keep_mask, rng, stream_id, p_drop, relu, nullptr, nullptr,
                                                                         nullptr, px, global_sums_out, PlaneSink{nullptr, nullptr, 0, 0});
  SCNB_CHECK_LAUNCH("bn_act_fwd_peer");
  return SCNB_OK;
}

extern "C" int scnb_bn_act_bwd_peer(const float* dA, const float* Z, const float* A, int H, int n_seg, const int32_t* seg_off_dev,
                                    int max_rows, const float* gamma, const float* stats, const uint8_t* keep_mask,
                                    const uint64_t* rng, uint32_t stream_id, float p_drop, int relu, float* dZ, float* dgamma,
                                    float* dbeta, const uint64_t* peer_bufs, int rank, int world, int slot, int n_slots,
                                    const int64_t* epoch, long long buf_bytes, void* stream) {
  SCNB_CHECK_ARG(dA && Z && A && gamma && stats && dZ && dgamma && dbeta && seg_off_dev && H > 0 && n_seg > 0,
                 "bn_act_bwd_peer: bad arguments");
  SCNB_CHECK_ARG(!(p_drop > 0.f) || keep_mask || rng, "bn_act_bwd_peer: dropout needs keep_mask or rng state");
  SCNB_CHECK_ARG(bn_vec8_ok(H, dA, Z, A, gamma, dZ, stats, keep_mask), "bn_act_bwd_peer: needs H %% 8 == 0 and 16-byte aligned buffers");
  int stride = 0;
  int rc = peer_check("bn_act_bwd_peer", (const unsigned long long*)peer_bufs, rank, world, slot, n_slots, (const long long*)epoch,
                      buf_bytes, H, n_seg, &stride);
  if (rc != SCNB_OK) return rc;
  SCNB_CHECK_ARG(!plane_sink_of(dZ, H, max_rows).hi, "bn_act_bwd_peer: operand planes are bound to this output; the peer form does not emit them");
  PeerX px{(const unsigned long long*)peer_bufs, (const long long*)epoch, rank, world, slot, peer_spin_limit(), stride};
  k_bn_act_bwd_v8<1><<<dim3(H / 8, n_seg), BNV_T, 0, (cudaStream_t)stream>>>(dA, Z, A, H, n_seg, seg_off_dev, max_rows, gamma, stats, keep_mask, rng,
                                                                         stream_id, p_drop, relu, dZ, dgamma, dbeta, nullptr, nullptr,
                                                                         nullptr, px, PlaneSink{nullptr, nullptr, 0, 0});
  SCNB_CHECK_LAUNCH("bn_act_bwd_peer");
  return SCNB_OK;
}

// running statistics: one update per (segment, application), segment-major -- the order in which
// the reference's three forwards touch a BatchNorm module; a block shared by several applications
// (n_layers >= 3, model.py:207) passes one stats pointer per application.  SURVEY A2: momentum 0.1,
// unbiased variance in the running estimate.  nbt += number of updates applied.
struct StatsList { const float* p[4]; };
__global__ void k_bn_running_update(float* __restrict__ rmean, float* __restrict__ rvar, int64_t* __restrict__ nbt,
                                    StatsList sl, int n_stats, const int32_t* __restrict__ seg_off, int n_seg, int H,
                                    float momentum, const float* __restrict__ dp_cnt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  int applied = 0;
  float rm = 0.f, rv = 0.f;
  if (c < H) {
    rm = rmean[c];
    rv = rvar[c];
  }
  for (int s = 0; s < n_seg; ++s) {
    const int n = dp_cnt ? (int)dp_cnt[s] : seg_off[s + 1] - seg_off[s];  // global rows under data parallelism
    if (n <= 0) continue;
    for (int a = 0; a < n_stats; ++a) {
      ++applied;
      if (c < H) {
        const float mean = sl.p[a][((size_t)s * 3 + 0) * H + c];
        const float var = sl.p[a][((size_t)s * 3 + 1) * H + c];
        const float unb = var * ((float)n / (float)max(n - 1, 1));
        rm = (1.f - momentum) * rm + momentum * mean;
        rv = (1.f - momentum) * rv + momentum * unb;
      }
    }
  }
  if (c < H) {
    rmean[c] = rm;
    rvar[c] = rv;
  }
  if (c == 0 && nbt) *nbt += applied;
}

extern "C" int scnb_bn_running_update(float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                      const float* stats0, const float* stats1, const float* stats2, const float* stats3,
                                      const int32_t* seg_off_dev, int n_seg, int H, float momentum, const float* dp_cnt,
                                      void* stream) {
  SCNB_CHECK_ARG(running_mean && running_var && stats0 && seg_off_dev && n_seg > 0 && H > 0, "bn_running_update: bad arguments");
  StatsList sl;
  sl.p[0] = stats0; sl.p[1] = stats1; sl.p[2] = stats2; sl.p[3] = stats3;
  int n_stats = 1;
  while (n_stats < 4 && sl.p[n_stats]) ++n_stats;
  k_bn_running_update<<<scnb_cdiv(H, 256), 256, 0, (cudaStream_t)stream>>>(running_mean, running_var, num_batches_tracked, sl,
                                                                            n_stats, seg_off_dev, n_seg, H, momentum, dp_cnt);
  SCNB_CHECK_LAUNCH("bn_running_update");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Hidden-space MixUp: out[i,:] = gam[i]*Z[srcA[i],:] + (1-gam[i])*Z[srcB[i],:], i < *n_active;
// zero rows after.  Adjoint: dZ[r,:] = sum_{k<3} coef[k][r] * dOut[inv[k][r],:]  (inv = -1 -> none).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mix_rows(const float* __restrict__ Z, int H, const int32_t* __restrict__ srcA,
                                                  const int32_t* __restrict__ srcB, const float* __restrict__ gam,
                                                  const int32_t* __restrict__ n_active, int max_rows, float* __restrict__ out) {
  PDL_ENTRY();
  const int i = blockIdx.x;
  const int na = n_active ? min(*n_active, max_rows) : max_rows;
  float* o = out + (size_t)i * H;
  if (i >= na) {
    for (int h = threadIdx.x; h < H; h += blockDim.x) o[h] = 0.f;
    return;
  }
  const int a = srcA[i], b = srcB ? srcB[i] : a;
  const float g = gam ? gam[i] : 1.f;
  const float* za = Z + (size_t)a * H;
  const float* zb = Z + (size_t)b * H;
  if (a == b || g == 1.f) {
    for (int h = threadIdx.x; h < H; h += blockDim.x) o[h] = za[h];
  } else {
    const float g1 = 1.f - g;
    for (int h = threadIdx.x; h < H; h += blockDim.x) o[h] = g * za[h] + g1 * zb[h];
  }
}

// Tall super-batches (config 5: 16 384 rows): one WARP per row, 16 bytes per lane, the row's indices and gamma loaded by every
// lane (broadcast) -- 8 rows per CTA instead of one: an eighth of the CTAs, each with its dependent index -> row loads of 8
// rows in flight at once.  Same expressions as k_mix_rows (bit-identical output).
__global__ void __launch_bounds__(256) k_mix_rows_w(const float* __restrict__ Z, int H, const int32_t* __restrict__ srcA,
                                                    const int32_t* __restrict__ srcB, const float* __restrict__ gam,
                                                    const int32_t* __restrict__ n_active, int max_rows, float* __restrict__ out) {
  PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= max_rows) return;
  const int na = n_active ? min(*n_active, max_rows) : max_rows;
  float4* o = reinterpret_cast<float4*>(out + (size_t)i * H);
  const int H4 = H >> 2;
  if (i >= na) {
    for (int h = lane; h < H4; h += 32) o[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int a = srcA[i], b = srcB ? srcB[i] : a;
  const float g = gam ? gam[i] : 1.f;
  const float4* za = reinterpret_cast<const float4*>(Z + (size_t)a * H);
  const float4* zb = reinterpret_cast<const float4*>(Z + (size_t)b * H);
  if (a == b || g == 1.f) {
    for (int h = lane; h < H4; h += 32) o[h] = za[h];
  } else {
    const float g1 = 1.f - g;
    for (int h = lane; h < H4; h += 32) {
      const float4 x = za[h], y = zb[h];
      o[h] = make_float4(g * x.x + g1 * y.x, g * x.y + g1 * y.y, g * x.z + g1 * y.z, g * x.w + g1 * y.w);
    }
  }
}

extern "C" int scnb_mix_rows(const float* Z, int H, const int32_t* srcA, const int32_t* srcB, const float* gam,
                             const int32_t* n_active_dev, int max_rows, float* out, void* stream) {
  SCNB_CHECK_ARG(Z && srcA && out && H > 0 && max_rows >= 0, "mix_rows: bad arguments");
  if (max_rows == 0) return SCNB_OK;
  if (max_rows >= 4096 && H % 4 == 0 && (((uintptr_t)Z | (uintptr_t)out) % 16) == 0) {
    scnb_launch_pdl(k_mix_rows_w, dim3(scnb_cdiv(max_rows, 8)), dim3(256), 0, (cudaStream_t)stream, Z, H, srcA, srcB, gam, n_active_dev, max_rows,
                    out);
    SCNB_CHECK_LAUNCH("mix_rows/warp");
    return SCNB_OK;
  }
  scnb_launch_pdl(k_mix_rows, dim3(max_rows), dim3(256), 0, (cudaStream_t)stream, Z, H, srcA, srcB, gam, n_active_dev, max_rows, out);
  SCNB_CHECK_LAUNCH("mix_rows");
  return SCNB_OK;
}

__global__ void __launch_bounds__(256) k_unmix_rows(const float* __restrict__ dOut, int H, const int32_t* __restrict__ inv,
                                                    const float* __restrict__ coef, int n_base, float* __restrict__ dZ) {
  PDL_ENTRY();
  const int r = blockIdx.x;
  int i0 = inv[r], i1 = inv[n_base + r], i2 = inv[2 * n_base + r];
  float c0 = coef[r], c1 = coef[n_base + r], c2 = coef[2 * n_base + r];
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float v = 0.f;
    if (i0 >= 0) v = fmaf(c0, dOut[(size_t)i0 * H + h], v);
    if (i1 >= 0) v = fmaf(c1, dOut[(size_t)i1 * H + h], v);
    if (i2 >= 0) v = fmaf(c2, dOut[(size_t)i2 * H + h], v);
    dZ[(size_t)r * H + h] = v;
  }
}

extern "C" int scnb_unmix_rows(const float* dOut, int H, const int32_t* inv3, const float* coef3, int n_base, float* dZ,
                               void* stream) {
  SCNB_CHECK_ARG(dOut && inv3 && coef3 && dZ && H > 0 && n_base >= 0, "unmix_rows: bad arguments");
  if (n_base == 0) return SCNB_OK;
  scnb_launch_pdl(k_unmix_rows, dim3(n_base), dim3(256), 0, (cudaStream_t)stream, dOut, H, inv3, coef3, n_base, dZ);
  SCNB_CHECK_LAUNCH("unmix_rows");
  return SCNB_OK;
}

// out = g * [y > 0]   (backward of a ReLU fused into a Linear epilogue)
__global__ void k_relu_bwd(const float* __restrict__ g, const float* __restrict__ y, float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = y[i] > 0.f ? g[i] : 0.f;
}

extern "C" int scnb_relu_bwd(const float* g, const float* y, float* out, long long n, void* stream) {
  SCNB_CHECK_ARG(g && y && out && n >= 0, "relu_bwd: bad arguments");
  if (n == 0) return SCNB_OK;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  k_relu_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, y, out, (size_t)n);
  SCNB_CHECK_LAUNCH("relu_bwd");
  return SCNB_OK;
}

// Eval-mode BatchNorm backward w.r.t. its input: dz = dy * gamma / sqrt(running_var + eps) (optionally ReLU-gated by A).
__global__ void k_bn_eval_bwd(const float* __restrict__ dA, const float* __restrict__ Aact, int relu, size_t n, int H,
                              const float* __restrict__ gamma, const float* __restrict__ rvar, float* __restrict__ dZ) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % H);
    float dy = dA[i];
    if (relu && !(Aact[i] > 0.f)) dy = 0.f;
    dZ[i] = dy * gamma[c] / sqrtf(rvar[c] + BN_EPS);
  }
}

extern "C" int scnb_bn_eval_bwd(const float* dA, const float* A, int relu, int rows, int H, const float* gamma,
                                const float* running_var, float* dZ, void* stream) {
  SCNB_CHECK_ARG(dA && A && gamma && running_var && dZ && rows >= 0 && H > 0, "bn_eval_bwd: bad arguments");
  const size_t n = (size_t)rows * H;
  if (n == 0) return SCNB_OK;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  k_bn_eval_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(dA, A, relu, n, H, gamma, running_var, dZ);
  SCNB_CHECK_LAUNCH("bn_eval_bwd");
  return SCNB_OK;
}
