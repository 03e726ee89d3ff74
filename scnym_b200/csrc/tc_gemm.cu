// General dense GEMM on the 5th-generation tensor cores for the large hidden-layer products (config 5: 16 384 x 1024 x 1024):
//   C[M, N] = alpha * op(A) op(B) (+ beta C) (+ bias[n]) (+ relu) (+ gate)          fp32 in, fp32 out
// for the three operand layouts of nn.Linear's forward, input gradient and weight gradient (scnym/model.py:218, 229):
//   LAY 0   A[m][k], B[n][k]     Z  = A W^T          (k contiguous in both)
//   LAY 1   A[m][k], B[k][n]     dA = dZ W           (B stored by k rows)
//   LAY 2   A[k][m], B[k][n]     dW = dZ^T A         (both stored by k rows)
// A CTA owns a 128 x 256 tile of C (one fp32 accumulator: 128 TMEM lanes x 256 columns) and walks K in tiles of 32:
//   * eight loader warps read the fp32 operand tiles from global memory (coalesced along whichever axis is contiguous),
//     split every value into bf16 hi = bf16(x), lo = bf16(x - hi) and write the two planes of each operand into a
//     shared-memory stage as SWIZZLE_64B K-major tiles -- no bf16 copy of activations or weights ever exists in HBM;
//   * one thread issues tcgen05.mma (kind::f16, M = 128, N <= 256, K = 16) x 2 K-slices x 3 products
//       acc += A_lo B_hi + A_hi B_lo + A_hi B_hi          (~16 mantissa bits per product, fp32 accumulation in TMEM)
//     and commits the stage back to the loaders through an mbarrier (4 stages);
//   * the loader warps then read the accumulator with tcgen05.ld and apply the epilogue; with a K split (weight
//     gradients: few output tiles, K = rows of the batch) the partial tiles are added atomically (beta must be 1).
#include <stdlib.h>

#include "tc_common.cuh"

#define TG5_THREADS 320   // warp 0: idle, warp 1: MMA issuer / TMEM owner, warps 2-9: operand staging, then epilogue
#define TG5_LOAD 256
#define TG5_BM 128
#define TG5_BN 256
#define TG5_KT 32

__device__ __forceinline__ void tg5_split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_bf16x2(x[2 * j], x[2 * j + 1], h[j], l[j]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool ok) {
  const int n = ok ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float ld_shared_f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// Operand staging, two phases per k-tile, both on the SAME shared-memory bytes of a stage:
//   (1) tg5_fetch: cp.async (16 bytes per request, zero-filled outside the matrix) brings the fp32 tile in, two k-tiles
//       ahead of its use, with no registers held across the global-memory latency;
//   (2) tg5_read / tg5_write: every thread reads its fp32 chunks, all 256 threads meet at a named barrier, then the bf16
//       hi / lo planes are written over the fp32 image (a tile of `rows` x 32 fp32 and its two planes are both rows*128 B).
// fp32 image: KMAJOR (X[row][k]): [rows][32] floats;  else (X[k][row]): [32][rows] floats.
// A thread's task = (row r, 8-k chunk c): rows of 64 bytes per plane, chunk c of row r at r*64 + swizzle(c, r).
// A thread's NP 16-byte pieces of an operand tile: global address of the piece in the FIRST k-tile, its k offset inside a
// k-tile and its validity are fixed for the whole K walk, so they are computed once; per k-tile a piece costs one cp.async
// and one pointer increment (the address arithmetic per piece had made the staging warps issue-bound: ncu, 5 900
// warp instructions per k-tile).
template <int NP>
struct Tg5Pieces {
  const float* p[NP];
  int kk[NP];        // k offset of the piece inside a k-tile; -1: the piece does not exist (outside the tile / matrix rows)
  size_t step;       // elements to advance per k-tile
};
template <bool KMAJOR, int NP>
__device__ __forceinline__ void tg5_pieces(Tg5Pieces<NP>& pc, const float* __restrict__ X, int ld, int row0, int n_rows, int tile_rows,
                                           int kb, int t) {
  pc.step = KMAJOR ? (size_t)TG5_KT : (size_t)TG5_KT * ld;
  const int pr = tile_rows >> 2;
#pragma unroll
  for (int u = 0; u < NP; ++u) {
    const int q = t + u * TG5_LOAD;
    int r, kk;
    if (KMAJOR) { r = q >> 3; kk = 4 * (q & 7); } else { kk = q / pr; r = 4 * (q - kk * pr); }
    const bool ok = (KMAJOR ? q < tile_rows * 8 : q < 32 * pr) && (row0 + r < n_rows);
    pc.kk[u] = ok ? kk : -1;
    pc.p[u] = ok ? (KMAJOR ? X + (size_t)(row0 + r) * ld + kb + kk : X + (size_t)(kb + kk) * ld + row0 + r) : X;
  }
}
template <int NP>
__device__ __forceinline__ void tg5_fetch(Tg5Pieces<NP>& pc, int k_left /* K elements left from this k-tile's start */, uint32_t smem,
                                          int t, int n_pieces) {
#pragma unroll
  for (int u = 0; u < NP; ++u) {
    const int q = t + u * TG5_LOAD;
    if (q < n_pieces) {                                            // pieces of the tile image (zero-filled when invalid)
      const bool ok = pc.kk[u] >= 0 && pc.kk[u] < k_left;
      cp_async16_zfill(smem + (uint32_t)q * 16u, pc.p[u], ok);
      pc.p[u] += pc.step;
    }
  }
}
template <bool KMAJOR, int NT>
__device__ __forceinline__ void tg5_read(uint32_t smem, int tile_rows, int t, float (&x)[NT][8]) {
#pragma unroll
  for (int u = 0; u < NT; ++u) {
    const int task = t + u * TG5_LOAD;
    if (task < tile_rows * 4) {
      if (KMAJOR) {
        const int r = task >> 2, c = task & 3;
        const float4 a = ld_shared_f4(smem + (uint32_t)r * 128u + (uint32_t)c * 32u);
        const float4 b = ld_shared_f4(smem + (uint32_t)r * 128u + (uint32_t)c * 32u + 16u);
        x[u][0] = a.x; x[u][1] = a.y; x[u][2] = a.z; x[u][3] = a.w; x[u][4] = b.x; x[u][5] = b.y; x[u][6] = b.z; x[u][7] = b.w;
      } else {
        const int r = task % tile_rows, c = task / tile_rows;
#pragma unroll
        for (int j = 0; j < 8; ++j) x[u][j] = ld_shared_f(smem + ((uint32_t)(8 * c + j) * (uint32_t)tile_rows + (uint32_t)r) * 4u);
      }
    }
  }
}
template <bool KMAJOR, int NT>
__device__ __forceinline__ void tg5_write(uint32_t smem_hi, uint32_t smem_lo, int tile_rows, int t, const float (&x)[NT][8]) {
#pragma unroll
  for (int u = 0; u < NT; ++u) {
    const int task = t + u * TG5_LOAD;
    if (task < tile_rows * 4) {
      const int r = KMAJOR ? (task >> 2) : (task % tile_rows), c = KMAJOR ? (task & 3) : (task / tile_rows);
      uint4 hi, lo;
      tg5_split8(x[u], hi, lo);
      const uint32_t off = (uint32_t)r * 64u + swz_chunk<TG5_KT>((uint32_t)c, (uint32_t)r);
      st_shared_v4(smem_hi + off, hi);
      st_shared_v4(smem_lo + off, lo);
    }
  }
}

template <int LAY>
__global__ void __launch_bounds__(TG5_THREADS, 1)
k_gemm_tc(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
          float alpha, float beta, const float* __restrict__ bias, int relu, const float* __restrict__ gate, int k_chunk, int stages,
          int bn /* columns of C per CTA: a multiple of 16, <= TG5_BN (small products use narrow tiles to cover more SMs) */) {
  constexpr uint32_t A_PLANE = TG5_BM * 64, B_PLANE = TG5_BN * 64;
  constexpr uint32_t STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;   // A_hi | A_lo | B_hi | B_lo = 48 KB
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = tc_smem_raw + (base - raw);
  const uint32_t bars = base + (uint32_t)stages * STAGE_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8u * stages, bar_acc = bars + 16u * stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (size_t)stages * STAGE_BYTES + 16 * stages + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TG5_BM, n0 = blockIdx.x * bn;
  const int kb = blockIdx.z * k_chunk, ke = min(K, kb + k_chunk);
  const int n_it = (ke - kb + TG5_KT - 1) / TG5_KT;
  // rows of the B tile that exist, rounded up to the MMA's N granularity (16)
  const int n_tile = min(bn, (N - n0 + 15) / 16 * 16);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8u * s, TG5_LOAD);
      mbar_init(bar_empty + 8u * s, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
      for (int i = 0; i < n_it; ++i) {
        const int s = i % stages;
        const uint32_t ph = (uint32_t)(i / stages) & 1u;
        mbar_wait(bar_full + 8u * s, ph);
        tc_fence_after();
        const uint32_t stage = base + (uint32_t)s * STAGE_BYTES;
        const uint64_t a_hi = umma_desc_kmajor<TG5_KT>(stage), a_lo = umma_desc_kmajor<TG5_KT>(stage + A_PLANE);
        const uint64_t b_hi = umma_desc_kmajor<TG5_KT>(stage + 2u * A_PLANE), b_lo = umma_desc_kmajor<TG5_KT>(stage + 2u * A_PLANE + B_PLANE);
#pragma unroll
        for (int k = 0; k < TG5_KT / 16; ++k) {
          const uint64_t ko = (uint64_t)(k * 2);
          tc_mma_bf16(tmem_base, a_lo + ko, b_hi + ko, idesc, (i > 0 || k > 0) ? 1u : 0u);
          tc_mma_bf16(tmem_base, a_hi + ko, b_lo + ko, idesc, 1u);
          tc_mma_bf16(tmem_base, a_hi + ko, b_hi + ko, idesc, 1u);
        }
        tc_commit(bar_empty + 8u * s);
      }
      if (n_it > 0) tc_commit(bar_acc);
    }
  } else if (warp >= 2) {
    // ===== operand staging: fp32 tiles -> bf16 hi / lo planes in shared memory =====
    const int t = tid - 64;
    constexpr int PF = 3;                       // k-tiles fetched ahead (stages >= PF + 1)
    Tg5Pieces<4> pa;
    Tg5Pieces<8> pb;
    tg5_pieces<LAY != 2, 4>(pa, A, lda, m0, M, TG5_BM, kb, t);
    tg5_pieces<LAY == 0, 8>(pb, B, ldb, n0, N, n_tile, kb, t);
    const int np_a = TG5_BM * 8, np_b = n_tile * 8;                // 16-byte pieces of the two fp32 tile images
    auto fetch = [&](int i) {   // k-tiles are fetched in order: the piece pointers advance by one k-tile per call
      const int s = i % stages;
      mbar_wait(bar_empty + 8u * s, ((uint32_t)(i / stages) & 1u) ^ 1u);      // the MMAs that read this stage are done
      const uint32_t stage = base + (uint32_t)s * STAGE_BYTES;
      const int k_left = ke - (kb + i * TG5_KT);
      tg5_fetch<4>(pa, k_left, stage, t, np_a);
      tg5_fetch<8>(pb, k_left, stage + 2u * A_PLANE, t, np_b);
    };
    for (int j = 0; j < PF; ++j) {
      if (j < n_it) fetch(j);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int i = 0; i < n_it; ++i) {
      asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");       // this thread's pieces of tile i have landed
      asm volatile("bar.sync 1, 256;" ::: "memory");                           // ... and everybody else's
      const uint32_t stage = base + (uint32_t)(i % stages) * STAGE_BYTES;
      float xa[2][8], xb[4][8];
      tg5_read<LAY != 2, 2>(stage, TG5_BM, t, xa);
      tg5_read<LAY == 0, 4>(stage + 2u * A_PLANE, n_tile, t, xb);
      asm volatile("bar.sync 1, 256;" ::: "memory");                           // all fp32 reads before the planes overwrite them
      tg5_write<LAY != 2, 2>(stage, stage + A_PLANE, TG5_BM, t, xa);
      tg5_write<LAY == 0, 4>(stage + 2u * A_PLANE, stage + 2u * A_PLANE + B_PLANE, n_tile, t, xb);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_full + 8u * (i % stages));
      if (i + PF < n_it) fetch(i + PF);         // refill the stage the tensor core released longest ago
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ===== epilogue: TMEM lane = row of C; warps 2-5 take columns [0, 128), warps 6-9 columns [128, 256) =====
    if (n_it > 0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      const int q = warp & 3, half = (warp - 2) >> 2;
      const int m = m0 + q * 32 + lane;
      const bool split = gridDim.z > 1;
      for (int cb = half * 128; cb < half * 128 + 128 && cb < n_tile; cb += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < M) {
          float* out = C + (size_t)m * ldc + n0 + cb;
          const float* gt = gate ? gate + (size_t)m * ldc + n0 + cb : nullptr;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + cb + j;
            if (n < N) {
              float x = alpha * __uint_as_float(v[j]);
              if (split) {
                atomicAdd(out + j, x);
              } else {
                if (beta != 0.f) x += beta * out[j];
                if (bias) x += bias[n];
                if (relu) x = fmaxf(x, 0.f);
                if (gt && !(gt[j] > 0.f)) x = 0.f;
                out[j] = x;
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// Minimum size at which the tensor-core kernel takes over from the latency-oriented mma.sync kernel (SCNB_GEMM_TC_MIN_FLOPS
// overrides; 0 disables).  The 1024 x 256 x 256 products of config 2 stay below it.
static double g_tg5_min_flops = -1.0;
static double tg5_min_flops() {
  if (g_tg5_min_flops < 0.0) {
    const char* e = getenv("SCNB_GEMM_TC_MIN_FLOPS");
    g_tg5_min_flops = e ? atof(e) : 2.0e9;
  }
  return g_tg5_min_flops;
}
// Override the take-over size at run time (tests; tuning): products of at least `flops` = 2 M N K go to tcgen05.
extern "C" int scnb_gemm_tc_min_flops(long long flops) {
  SCNB_CHECK_ARG(flops >= 0, "gemm_tc_min_flops: negative size");
  g_tg5_min_flops = (double)flops;
  return SCNB_OK;
}

// Returns 1 if launched, 0 if not applicable, <0 on error.
int scnb_gemm_tc_try(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                     float alpha, float beta, const float* bias, int relu, const float* gate, cudaStream_t st) {
  if (transA && transB) return 0;
  const int lay = transA ? 2 : (transB ? 0 : 1);
  const double flops = 2.0 * M * (double)N * K;
  if (tg5_min_flops() <= 0.0 || flops < tg5_min_flops()) return 0;
  if (M < TG5_BM || N < 64 || K < 64) return 0;
  // tile width: 256 columns; SCNB_GEMM_TC_BN = 32 / 64 / 128 narrows the tiles (tools/bench_gemm_small.py).  Measured on the
  // 1024 x 256 x 256 products of config 2 the narrow-tile form needs 16.7 us against 10.4 us of the mma.sync kernel (TMEM
  // allocation, a 4-stage pipeline and the fp32 -> bf16 staging are fixed costs per CTA), so those stay where they are.
  static int bn_env = -1;
  if (bn_env < 0) { const char* e = getenv("SCNB_GEMM_TC_BN"); bn_env = e ? atoi(e) : 0; }
  int bn = TG5_BN;
  if (bn_env >= 32 && bn_env <= TG5_BN && bn_env % 32 == 0) bn = bn_env;
  // 16-byte cp.async pieces: leading dimensions and bases aligned; pieces never straddle a matrix edge
  if ((lda % 4) || (ldb % 4) || (((uintptr_t)A | (uintptr_t)B) % 16)) return 0;
  if (lay == 0 && (K % 4)) return 0;                 // A[m][k], B[n][k]
  if (lay == 1 && ((K % 4) || (N % 4))) return 0;    // A[m][k], B[k][n]
  if (lay == 2 && ((M % 4) || (N % 4))) return 0;    // A[k][m], B[k][n]
  const long long tiles = (long long)scnb_cdiv(M, TG5_BM) * scnb_cdiv(N, bn);
  int splits = 1;
  const bool splittable = !bias && !relu && !gate && (beta == 1.f);
  if (splittable)
    while (splits < 64 && tiles * splits < 148 && K / (splits * 2) >= 512) splits *= 2;
  int k_chunk = scnb_cdiv(scnb_cdiv(K, splits), TG5_KT) * TG5_KT;
  splits = scnb_cdiv(K, k_chunk);
  const int stages = 4;
  const size_t smem = 1024 + (size_t)stages * (2 * TG5_BM * 64 + 2 * TG5_BN * 64) + 16 * stages + 16 + 16;
  dim3 grid(scnb_cdiv(N, bn), scnb_cdiv(M, TG5_BM), splits);
  cudaError_t e;
#define TG5_GO(L)                                                                                                          \
  e = cudaFuncSetAttribute(k_gemm_tc<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                          \
  if (e == cudaSuccess)                                                                                                    \
    k_gemm_tc<L><<<grid, TG5_THREADS, smem, st>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, k_chunk, stages, bn)
  if (lay == 0) { TG5_GO(0); } else if (lay == 1) { TG5_GO(1); } else { TG5_GO(2); }
#undef TG5_GO
  if (e != cudaSuccess) {
    scnb_set_error("sgemm/tcgen05: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  return 1;
}
