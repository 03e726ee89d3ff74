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

// =====================================================================================================================
// Planes form (used when the caller has registered a workspace for the stream, scnb_gemm_tc_workspace): the fp32 -> bf16
// hi / lo conversion is taken OFF the tensor pipe's critical resource.  In k_gemm_tc above eight warps convert 12 288
// floats per k-tile between two named barriers while one thread issues 6 MMAs (768 cycles): the staging warps, not the
// tensor cores, set the pace (ncu: tensor pipe 17 % active, profiles/r01_ncu_full_v11_gemm_tc.txt).  Here
//   (1) k_gemm_operand_planes writes each operand ONCE as two bf16 planes in the layout the MMA wants: tiles
//       [row tile of 256][k tile of 32], every tile a ready-made SWIZZLE_64B K-major shared-memory image of 16 KB
//       (HBM-bound streaming pass; the transposed operands of the input- and weight-gradient products go through a
//       shared-memory transposition so that both the reads and the writes are full lines);
//   (2) k_gemm_planes owns a 256 x 256 tile of C (two 128-lane fp32 accumulators = all 512 TMEM columns): one thread
//       brings a stage in with four 16 KB bulk copies (A_hi, A_lo, B_hi, B_lo), one thread issues 12 MMAs per stage
//       (M = 128, N = 256, K = 16; acc += A_lo B_hi + A_hi B_lo + A_hi B_hi), eight warps only run the epilogue.
//       64 KB per 1536 MMA cycles per SM = 12 TB/s of L2 -> SM traffic over the chip when every SM runs at the tensor
//       peak -- the same operand bytes per flop as fp32 tiles, but no instruction issue between L2 and the tensor core.
// =====================================================================================================================
#define GP_BM 256
#define GP_BN 256
#define GP_KT 32
#define GP_TILE_BYTES (256u * 64u)        // one plane of one 256 x 32 tile
#define GP_THREADS 320                    // warp 0: loader, warp 1: TMEM owner + MMA issuer, warps 2-9: epilogue

// P[r][k] (r: the operand's row on the MMA's M or N axis, k: the reduction index) as tiled planes.
//   TRANS == false: P[r][k] = X[r * ld + k]   (k contiguous)      one thread per (row, 8-k chunk)
//   TRANS == true : P[r][k] = X[k * ld + r]   (r contiguous)      one CTA per tile, transposed through shared memory
// Rows >= R and k >= K are written as zeros (the GEMM always multiplies full tiles).
template <bool TRANS>
__global__ void __launch_bounds__(256) k_gemm_operand_planes(const float* __restrict__ X, int ld, int R, int K, int n_rt, int n_kt,
                                                             uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  if (!TRANS) {
    const long long total = (long long)n_rt * 256 * n_kt * 4;
    const bool vec = (ld % 4 == 0) && ((uintptr_t)X % 16 == 0);
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
      const int cpr = n_kt * 4;
      const int r = (int)(e / cpr), cg = (int)(e - (long long)r * cpr);
      const int k0 = cg * 8;
      float x[8];
      if (r < R && vec && k0 + 8 <= K) {
        const float4 a = *reinterpret_cast<const float4*>(X + (size_t)r * ld + k0);
        const float4 b = *reinterpret_cast<const float4*>(X + (size_t)r * ld + k0 + 4);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (r < R && k0 + j < K) ? X[(size_t)r * ld + k0 + j] : 0.f;
      }
      uint4 h, l;
      tg5_split8(x, h, l);
      const uint32_t rr = (uint32_t)(r & 255);
      const size_t off = ((size_t)(r >> 8) * n_kt + (size_t)(cg >> 2)) * (256 * 32) + (size_t)rr * 32 +
                         (size_t)(swz_chunk<32>((uint32_t)(cg & 3), rr) >> 1);
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(lo + off) = l;
    }
  } else {
    __shared__ float wt[32][257];
    const int kt = blockIdx.x, rt = blockIdx.y, t = threadIdx.x;
    const int r = rt * 256 + t;
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const int k = kt * 32 + kk;
      wt[kk][t] = (r < R && k < K) ? X[(size_t)k * ld + r] : 0.f;
    }
    __syncthreads();
    const size_t tile = ((size_t)rt * n_kt + kt) * (256 * 32);
    // thread (row rr, chunk c): consecutive threads take the four chunks of one row, so a warp writes 8 full 64-byte rows
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int task = t + 256 * u;
      const uint32_t rr = (uint32_t)(task >> 2), c = (uint32_t)(task & 3);
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = wt[8 * c + j][rr];
      uint4 h, l;
      tg5_split8(x, h, l);
      const size_t off = tile + (size_t)rr * 32 + (size_t)(swz_chunk<32>(c, rr) >> 1);
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(lo + off) = l;
    }
  }
}

__global__ void __launch_bounds__(GP_THREADS, 1)
k_gemm_planes(int M, int N, int n_kt, const uint8_t* __restrict__ a_hi, const uint8_t* __restrict__ a_lo,
              const uint8_t* __restrict__ b_hi, const uint8_t* __restrict__ b_lo, float* __restrict__ C, int ldc, float alpha, float beta,
              const float* __restrict__ bias, int relu, const float* __restrict__ gate, int kt_chunk, int stages) {
  constexpr uint32_t STAGE_BYTES = 4u * GP_TILE_BYTES;   // A_hi | A_lo | B_hi | B_lo = 64 KB
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = tc_smem_raw + (base - raw);
  const uint32_t bars = base + (uint32_t)stages * STAGE_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8u * stages, bar_acc = bars + 16u * stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (size_t)stages * STAGE_BYTES + 16 * stages + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x, nt = blockIdx.y;
  const int kt0 = blockIdx.z * kt_chunk, kt1 = min(n_kt, kt0 + kt_chunk);
  const int n_it = max(kt1 - kt0, 0);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== loader: four 16 KB bulk copies per stage =====
      const size_t a_off0 = ((size_t)mt * n_kt + kt0) * GP_TILE_BYTES, b_off0 = ((size_t)nt * n_kt + kt0) * GP_TILE_BYTES;
      for (int i = 0; i < n_it; ++i) {
        const int s = i % stages;
        mbar_wait(bar_empty + 8u * s, ((uint32_t)(i / stages) & 1u) ^ 1u);
        const uint32_t stage = base + (uint32_t)s * STAGE_BYTES;
        const uint32_t bar = bar_full + 8u * s;
        mbar_arrive_expect_tx(bar, STAGE_BYTES);
        const size_t ao = a_off0 + (size_t)i * GP_TILE_BYTES, bo = b_off0 + (size_t)i * GP_TILE_BYTES;
        bulk_g2s(stage, a_hi + ao, GP_TILE_BYTES, bar);
        bulk_g2s(stage + GP_TILE_BYTES, a_lo + ao, GP_TILE_BYTES, bar);
        bulk_g2s(stage + 2u * GP_TILE_BYTES, b_hi + bo, GP_TILE_BYTES, bar);
        bulk_g2s(stage + 3u * GP_TILE_BYTES, b_lo + bo, GP_TILE_BYTES, bar);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GP_BN >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
      for (int i = 0; i < n_it; ++i) {
        const int s = i % stages;
        mbar_wait(bar_full + 8u * s, (uint32_t)(i / stages) & 1u);
        tc_fence_after();
        const uint32_t stage = base + (uint32_t)s * STAGE_BYTES;
        const uint64_t bh = umma_desc_kmajor<GP_KT>(stage + 2u * GP_TILE_BYTES), bl = umma_desc_kmajor<GP_KT>(stage + 3u * GP_TILE_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t ah = umma_desc_kmajor<GP_KT>(stage + (uint32_t)h * (TC_M * 64u));
          const uint64_t al = umma_desc_kmajor<GP_KT>(stage + GP_TILE_BYTES + (uint32_t)h * (TC_M * 64u));
          const uint32_t d = tmem_base + (uint32_t)(h * GP_BN);
#pragma unroll
          for (int k = 0; k < GP_KT / 16; ++k) {
            const uint64_t ko = (uint64_t)(k * 2);
            tc_mma_bf16(d, al + ko, bh + ko, idesc, (i > 0 || k > 0) ? 1u : 0u);
            tc_mma_bf16(d, ah + ko, bl + ko, idesc, 1u);
            tc_mma_bf16(d, ah + ko, bh + ko, idesc, 1u);
          }
        }
        tc_commit(bar_empty + 8u * s);
      }
      if (n_it > 0) tc_commit(bar_acc);
    }
  } else if (n_it > 0) {
    // ===== epilogue: accumulator `half` holds rows [128 half, 128 half + 128) of the tile; TMEM lane = row =====
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int m = mt * GP_BM + half * TC_M + q * 32 + lane;
    const int n0 = nt * GP_BN;
    const bool split = gridDim.z > 1;
    for (int cb = 0; cb < GP_BN && n0 + cb < N; cb += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * GP_BN + cb);
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
        if (split && n0 + cb + 32 <= N) {    // K split: 16-byte vector atomics (a quarter of the atomic operations)
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            atomicAdd(reinterpret_cast<float4*>(out + j), make_float4(alpha * __uint_as_float(v[j]), alpha * __uint_as_float(v[j + 1]),
                                                                      alpha * __uint_as_float(v[j + 2]), alpha * __uint_as_float(v[j + 3])));
        } else if (split) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + cb + j < N) atomicAdd(out + j, alpha * __uint_as_float(v[j]));
        } else if (n0 + cb + 32 <= N) {      // full 32-column group: 16-byte accesses (ldc % 4 == 0, C 16-byte aligned)
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(alpha * __uint_as_float(v[j]), alpha * __uint_as_float(v[j + 1]), alpha * __uint_as_float(v[j + 2]),
                                   alpha * __uint_as_float(v[j + 3]));
            if (beta != 0.f) {
              const float4 c0 = *reinterpret_cast<const float4*>(out + j);
              o.x += beta * c0.x; o.y += beta * c0.y; o.z += beta * c0.z; o.w += beta * c0.w;
            }
            if (bias) {
              const float4 bb = *reinterpret_cast<const float4*>(bias + n0 + cb + j);
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (gt) {
              const float4 g4 = *reinterpret_cast<const float4*>(gt + j);
              if (!(g4.x > 0.f)) o.x = 0.f;
              if (!(g4.y > 0.f)) o.y = 0.f;
              if (!(g4.z > 0.f)) o.z = 0.f;
              if (!(g4.w > 0.f)) o.w = 0.f;
            }
            *reinterpret_cast<float4*>(out + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + cb + j;
            if (n < N) {
              float x = alpha * __uint_as_float(v[j]);
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
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Weight gradient C[M, N] (+)= alpha * sum_r P[r][m] Q[r][n] from the ROW-major planes of P [R][M] and Q [R][N] (the planes
// their producers wrote for the forward / input-gradient products): both operands are MN-major for this product.  A
// row-major tile (256 rows x 32 columns, 64-byte rows, SWIZZLE_64B by address) read with k = row, mn = column IS the
// canonical MN-major SWIZZLE_64B layout ((8,4,m),(8,k)):((1,8,LBO),(32,SBO)) elements: 32 contiguous mn, the next 32 mn at LBO,
// 8 k rows of 64 bytes, the next 8 at SBO = 512 bytes.  A stage holds 32 rows (k) of 8 + 8 column tiles (256 m, 256 n):
// 32-row slabs of 2 KB per tile and plane (the swizzle repeats every 512 bytes, so a slab is a valid image), LBO = 2 KB.
// CTA tile 256 x 256 (two accumulators), K split over gridDim.z with atomics.
#define GPM_SLAB 2048u                       // 32 rows x 64 bytes of one tile
#define GPM_OP_BYTES (8u * GPM_SLAB)         // one plane of one operand in a stage: 8 column tiles
__device__ __forceinline__ uint64_t umma_desc_mnmajor64(uint32_t smem_addr, uint32_t lbo_bytes) {
  const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
  const uint32_t hi = (512u >> 4) | (1u << 14) | (4u << 29);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
__global__ void __launch_bounds__(GP_THREADS, 1)
k_gemm_planes_mn(int M, int N, int n_rt /* 256-row tiles of the operands */, int p_nkt, int q_nkt, const uint8_t* __restrict__ p_hi,
                 const uint8_t* __restrict__ p_lo, const uint8_t* __restrict__ q_hi, const uint8_t* __restrict__ q_lo,
                 float* __restrict__ C, int ldc, float alpha, int slabs_per_cta, int stages) {
  constexpr uint32_t STAGE_BYTES = 4u * GPM_OP_BYTES;   // P_hi | P_lo | Q_hi | Q_lo = 64 KB
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = tc_smem_raw + (base - raw);
  const uint32_t bars = base + (uint32_t)stages * STAGE_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8u * stages, bar_acc = bars + 16u * stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (size_t)stages * STAGE_BYTES + 16 * stages + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x, nt = blockIdx.y;                    // 256-column blocks of P and of Q
  const int n_slabs = n_rt * 8;                                  // 32-row slabs over all rows
  const int s0 = blockIdx.z * slabs_per_cta, s1 = min(n_slabs, s0 + slabs_per_cta);
  const int n_it = max(s1 - s0, 0);
  // column tiles of this CTA that exist (M, N multiples of 32; a partial 256-block multiplies zero-filled stage bytes)
  const int p_tiles = min(8, p_nkt - mt * 8), q_tiles = min(8, q_nkt - nt * 8);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (p_tiles < 8 || q_tiles < 8) {     // column tiles that do not exist stay zero in every stage
    for (uint32_t o = (uint32_t)tid * 16u; o < (uint32_t)stages * STAGE_BYTES; o += GP_THREADS * 16u) st_shared_zero16(base + o);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== loader: a 2 KB slab per column tile and plane = up to 32 bulk copies per stage.  ONE per lane: a single thread
    // issuing all 32 took longer than the 12 MMAs of a stage (the product ran at 40 % of the tensor peak) =====
    // lane = 16 * operand + 8 * plane + column tile
    const int op = lane >> 4, plane = (lane >> 3) & 1, j = lane & 7;
    const bool mine = j < (op ? q_tiles : p_tiles);
    const uint8_t* src = op ? (plane ? q_lo : q_hi) : (plane ? p_lo : p_hi);
    const int nkt = op ? q_nkt : p_nkt, ct = (op ? nt : mt) * 8 + j;
    const uint32_t dst_off = (uint32_t)(2 * op + plane) * GPM_OP_BYTES + (uint32_t)j * GPM_SLAB;
    for (int i = 0; i < n_it; ++i) {
      const int s = i % stages;
      const uint32_t bar = bar_full + 8u * s;
      if (lane == 0) {
        mbar_wait(bar_empty + 8u * s, ((uint32_t)(i / stages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar, 2u * GPM_SLAB * (uint32_t)(p_tiles + q_tiles));
      }
      __syncwarp();
      if (mine) {
        const int slab = s0 + i;
        const size_t off = ((size_t)(slab >> 3) * nkt + (size_t)ct) * GP_TILE_BYTES + (size_t)(slab & 7) * GPM_SLAB;
        bulk_g2s(base + (uint32_t)s * STAGE_BYTES + dst_off, src + off, GPM_SLAB, bar);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer: both operands MN-major (idesc bits 15, 16) =====
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(GP_BN >> 3) << 17) |
                             ((uint32_t)(TC_M >> 4) << 24);
      for (int i = 0; i < n_it; ++i) {
        const int s = i % stages;
        mbar_wait(bar_full + 8u * s, (uint32_t)(i / stages) & 1u);
        tc_fence_after();
        const uint32_t stage = base + (uint32_t)s * STAGE_BYTES;
#pragma unroll
        for (int h = 0; h < 2; ++h) {          // 128 columns of P (4 column tiles) per accumulator
          const uint32_t d = tmem_base + (uint32_t)(h * GP_BN);
#pragma unroll
          for (int k = 0; k < 2; ++k) {        // 16 rows per MMA: 1 KB further into every slab
            const uint32_t ko = (uint32_t)k * 1024u;
            const uint64_t ph = umma_desc_mnmajor64(stage + (uint32_t)h * 4u * GPM_SLAB + ko, GPM_SLAB);
            const uint64_t pl = umma_desc_mnmajor64(stage + GPM_OP_BYTES + (uint32_t)h * 4u * GPM_SLAB + ko, GPM_SLAB);
            const uint64_t qh = umma_desc_mnmajor64(stage + 2u * GPM_OP_BYTES + ko, GPM_SLAB);
            const uint64_t ql = umma_desc_mnmajor64(stage + 3u * GPM_OP_BYTES + ko, GPM_SLAB);
            tc_mma_bf16(d, pl, qh, idesc, (i > 0 || k > 0) ? 1u : 0u);
            tc_mma_bf16(d, ph, ql, idesc, 1u);
            tc_mma_bf16(d, ph, qh, idesc, 1u);
          }
        }
        tc_commit(bar_empty + 8u * s);
      }
      if (n_it > 0) tc_commit(bar_acc);
    }
  } else if (n_it > 0) {
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int m = mt * GP_BM + half * TC_M + q * 32 + lane;
    const int n0 = nt * GP_BN;
    for (int cb = 0; cb < GP_BN && n0 + cb < N; cb += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * GP_BN + cb);
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
        if (n0 + cb + 32 <= N) {               // 16-byte vector atomics (ldc % 4 == 0, C 16-byte aligned: checked by the launcher)
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            atomicAdd(reinterpret_cast<float4*>(out + j), make_float4(alpha * __uint_as_float(v[j]), alpha * __uint_as_float(v[j + 1]),
                                                                      alpha * __uint_as_float(v[j + 2]), alpha * __uint_as_float(v[j + 3])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + cb + j < N) atomicAdd(out + j, alpha * __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Workspaces for the operand planes, one per stream that issues large products (registered by the engine before any
// capture: the library itself allocates nothing).  A product on a stream without a (large enough) workspace keeps the
// in-kernel staging form above.
struct Tg5Workspace { cudaStream_t st; uint8_t* buf; long long bytes; };
#define TG5_MAX_WS 64
static Tg5Workspace g_tg5_ws[TG5_MAX_WS];
static int g_tg5_n_ws = 0;
extern "C" int scnb_gemm_tc_workspace(void* stream, void* buf, long long bytes) {
  SCNB_CHECK_ARG(bytes >= 0 && ((uintptr_t)buf % 1024) == 0, "gemm_tc_workspace: the buffer must be 1024-byte aligned");
  int free_slot = -1;
  for (int i = 0; i < g_tg5_n_ws; ++i) {
    if (g_tg5_ws[i].buf && g_tg5_ws[i].st == (cudaStream_t)stream) {
      g_tg5_ws[i].buf = (uint8_t*)buf;
      g_tg5_ws[i].bytes = buf ? bytes : 0;
      return SCNB_OK;
    }
    if (!g_tg5_ws[i].buf && free_slot < 0) free_slot = i;
  }
  if (!buf) return SCNB_OK;                      // nothing registered for this stream
  if (free_slot < 0) {
    SCNB_CHECK_ARG(g_tg5_n_ws < TG5_MAX_WS, "gemm_tc_workspace: more than %d streams with a workspace", TG5_MAX_WS);
    free_slot = g_tg5_n_ws++;
  }
  g_tg5_ws[free_slot] = Tg5Workspace{(cudaStream_t)stream, (uint8_t*)buf, bytes};
  return SCNB_OK;
}
// Plane buffers bound to fp32 tensors by the engine: the producer of the tensor (scnb_bn_act_fwd / scnb_bn_act_bwd) writes
// the planes next to it, and a product that reads the tensor as a row-major operand [rows][cols] (lda == cols) takes them
// as they are.  A weight-gradient product whose two operands are both bound reads the SAME planes as MN-major operands
// (k_gemm_planes_mn) -- no transposed copies.  rows_pad x cols bf16 per plane, rows_pad a multiple of 256, cols of 32.
struct Tg5Binding { const float* X; int rows, cols, rows_pad; uint16_t* hi; uint16_t* lo; };
#define TG5_MAX_BIND 64
static Tg5Binding g_tg5_bind[TG5_MAX_BIND];
static int g_tg5_n_bind = 0;
extern "C" int scnb_gemm_tc_bind_planes(const float* X, int rows, int cols, void* hi, void* lo) {
  SCNB_CHECK_ARG(X && rows > 0 && cols > 0, "gemm_tc_bind_planes: bad arguments");
  int free_slot = -1;
  for (int i = 0; i < g_tg5_n_bind; ++i) {
    if (g_tg5_bind[i].X == X) { free_slot = i; break; }
    if (!g_tg5_bind[i].X && free_slot < 0) free_slot = i;
  }
  if (!hi || !lo) {                                           // unbind
    if (free_slot >= 0 && g_tg5_bind[free_slot].X == X) g_tg5_bind[free_slot] = Tg5Binding{nullptr, 0, 0, 0, nullptr, nullptr};
    return SCNB_OK;
  }
  SCNB_CHECK_ARG(cols % 32 == 0, "gemm_tc_bind_planes: cols must be a multiple of 32 (got %d)", cols);
  SCNB_CHECK_ARG((((uintptr_t)hi | (uintptr_t)lo) % 1024) == 0, "gemm_tc_bind_planes: planes must be 1024-byte aligned");
  if (free_slot < 0) {
    SCNB_CHECK_ARG(g_tg5_n_bind < TG5_MAX_BIND, "gemm_tc_bind_planes: more than %d bound tensors", TG5_MAX_BIND);
    free_slot = g_tg5_n_bind++;
  }
  g_tg5_bind[free_slot] = Tg5Binding{X, rows, cols, (rows + 255) / 256 * 256, (uint16_t*)hi, (uint16_t*)lo};
  return SCNB_OK;
}
// X may also start at a later row tile of a bound tensor (a row offset that is a multiple of 256: the planes of rows
// [256 t, ...) are the tiles from t on) -- the adversary branch reads the rows behind the MixMatch segments of As[last].
bool scnb_planes_lookup(const float* X, int cols, int min_rows, uint16_t** hi, uint16_t** lo, int* rows_pad) {
  for (int i = 0; i < g_tg5_n_bind; ++i) {
    const Tg5Binding& b = g_tg5_bind[i];
    if (!b.X || b.cols != cols || X < b.X) continue;
    const size_t d = (size_t)(X - b.X);
    if (d % ((size_t)256 * cols) != 0) continue;
    const long long row0 = (long long)(d / cols);
    if (row0 >= b.rows || b.rows - row0 < min_rows) continue;
    const size_t off = (size_t)row0 * cols;        // bf16 elements: whole tiles
    *hi = b.hi + off; *lo = b.lo + off; *rows_pad = b.rows_pad - (int)row0;
    return true;
  }
  return false;
}

// bytes of workspace a product C[M, N] = op(A) op(B) over K needs
static long long tg5_workspace_bytes(int M, int N, int K) {
  const long long n_kt = (K + GP_KT - 1) / GP_KT;
  return 2ll * GP_TILE_BYTES * n_kt * ((M + GP_BM - 1) / GP_BM + (N + GP_BN - 1) / GP_BN);
}
extern "C" int scnb_gemm_tc_workspace_bytes(int M, int N, int K, long long* bytes_out) {
  SCNB_CHECK_ARG(M > 0 && N > 0 && K > 0 && bytes_out, "gemm_tc_workspace_bytes: bad arguments");
  *bytes_out = tg5_workspace_bytes(M, N, K);
  return SCNB_OK;
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

static long long g_tg5_planes_launches = 0;
static long long g_tg5_bound_uses = 0;      // operands taken from planes their producer wrote (no conversion pass)
extern "C" int scnb_gemm_tc_bound_count(long long* count_out) {
  SCNB_CHECK_ARG(count_out, "gemm_tc_bound_count: null pointer");
  *count_out = g_tg5_bound_uses;
  return SCNB_OK;
}
// number of products issued in the planes form so far (tests assert that the path they mean to cover really ran)
extern "C" int scnb_gemm_tc_planes_count(long long* count_out) {
  SCNB_CHECK_ARG(count_out, "gemm_tc_planes_count: null pointer");
  *count_out = g_tg5_planes_launches;
  return SCNB_OK;
}
extern "C" int scnb_gemm_tc_min_flops_get(long long* flops_out) {
  SCNB_CHECK_ARG(flops_out, "gemm_tc_min_flops_get: null pointer");
  *flops_out = (long long)tg5_min_flops();
  return SCNB_OK;
}

// The planes form.  Returns 1 if launched, 0 if not applicable (no workspace on this stream, operands not aligned).
static int tg5_planes_try(int lay, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, float alpha,
                          float beta, const float* bias, int relu, const float* gate, cudaStream_t st) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("SCNB_GEMM_TC_PLANES"); off = (e && e[0] == '0') ? 1 : 0; }
  if (off) return 0;
  static int bind_off = -1;
  if (bind_off < 0) { const char* e = getenv("SCNB_GEMM_TC_BOUND"); bind_off = (e && e[0] == '0') ? 1 : 0; }
  // weight gradient with both operands bound: the producers' row-major planes as MN-major operands, nothing to convert
  if (lay == 2 && !bind_off && beta == 1.f && !bias && !relu && !gate && (M % 32 == 0) && (N % 32 == 0) && lda == M && ldb == N &&
      (ldc % 4 == 0) && ((uintptr_t)C % 16 == 0)) {
    uint16_t *ph, *pl, *qh, *ql;
    int rp_a = 0, rp_b = 0;
    if (scnb_planes_lookup(A, M, K, &ph, &pl, &rp_a) && scnb_planes_lookup(B, N, K, &qh, &ql, &rp_b)) {
      const int n_rt = scnb_cdiv(K, 256);               // (rows past K are zeros in both operands' planes)
      if (n_rt * 256 <= rp_a && n_rt * 256 <= rp_b) {
        const int n_mt = scnb_cdiv(M, GP_BM), n_nt = scnb_cdiv(N, GP_BN), n_slabs = n_rt * 8;
        int splits = 1;
        while (splits < 64 && (long long)n_mt * n_nt * splits < 148 && n_slabs / (splits * 2) >= 16) splits *= 2;
        const int per = scnb_cdiv(n_slabs, splits);
        splits = scnb_cdiv(n_slabs, per);
        const int stages = 3;
        const size_t smem = 1024 + (size_t)stages * 4 * GPM_OP_BYTES + 16 * stages + 16 + 16;
        cudaError_t e = cudaFuncSetAttribute(k_gemm_planes_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
          scnb_set_error("sgemm/tcgen05 planes (MN-major): cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
          return SCNB_ERR_CUDA;
        }
        k_gemm_planes_mn<<<dim3(n_mt, n_nt, splits), GP_THREADS, smem, st>>>(M, N, n_rt, M / 32, N / 32, (const uint8_t*)ph, (const uint8_t*)pl,
                                                                            (const uint8_t*)qh, (const uint8_t*)ql, C, ldc, alpha, per, stages);
        ++g_tg5_planes_launches;
        ++g_tg5_bound_uses;
        return 1;
      }
    }
  }
  const Tg5Workspace* ws = nullptr;
  for (int i = 0; i < g_tg5_n_ws; ++i)
    if (g_tg5_ws[i].st == st && g_tg5_ws[i].buf) ws = &g_tg5_ws[i];
  if (!ws || tg5_workspace_bytes(M, N, K) > ws->bytes) return 0;
  if ((ldc % 4) || ((uintptr_t)C % 16) || (bias && ((uintptr_t)bias % 16)) || (gate && ((uintptr_t)gate % 16))) return 0;
  const int n_kt = scnb_cdiv(K, GP_KT), n_mt = scnb_cdiv(M, GP_BM), n_nt = scnb_cdiv(N, GP_BN);
  const size_t a_plane = (size_t)n_mt * n_kt * GP_TILE_BYTES, b_plane = (size_t)n_nt * n_kt * GP_TILE_BYTES;
  uint8_t* a_hi = ws->buf;
  uint8_t* a_lo = a_hi + a_plane;
  uint8_t* b_hi = a_lo + a_plane;
  uint8_t* b_lo = b_hi + b_plane;
  // a row-major A operand whose producer has already written its planes (scnb_gemm_tc_bind_planes)
  bool a_bound = false;
  if (lay != 2 && !bind_off && lda == K && K % 32 == 0) {
    uint16_t *bh, *bl;
    int rp = 0;
    if (scnb_planes_lookup(A, K, M, &bh, &bl, &rp) && n_mt * GP_BM <= rp) {
      a_hi = (uint8_t*)bh;
      a_lo = (uint8_t*)bl;
      a_bound = true;
      ++g_tg5_bound_uses;
    }
  }
  auto planes = [&](bool trans, const float* X, int ld, int R, int n_rt, uint8_t* hi, uint8_t* lo) {
    if (trans) {
      k_gemm_operand_planes<true><<<dim3(n_kt, n_rt), 256, 0, st>>>(X, ld, R, K, n_rt, n_kt, (uint16_t*)hi, (uint16_t*)lo);
    } else {
      const long long total = (long long)n_rt * 256 * n_kt * 4;
      const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
      k_gemm_operand_planes<false><<<blocks, 256, 0, st>>>(X, ld, R, K, n_rt, n_kt, (uint16_t*)hi, (uint16_t*)lo);
    }
  };
  if (!a_bound) planes(lay == 2, A, lda, M, n_mt, a_hi, a_lo);
  planes(lay != 0, B, ldb, N, n_nt, b_hi, b_lo);
  int splits = 1;
  const bool splittable = !bias && !relu && !gate && (beta == 1.f);
  if (splittable)
    while (splits < 64 && (long long)n_mt * n_nt * splits < 148 && n_kt / (splits * 2) >= 16) splits *= 2;
  const int kt_chunk = scnb_cdiv(n_kt, splits);
  splits = scnb_cdiv(n_kt, kt_chunk);
  const int stages = 3;
  const size_t smem = 1024 + (size_t)stages * 4 * GP_TILE_BYTES + 16 * stages + 16 + 16;
  cudaError_t e = cudaFuncSetAttribute(k_gemm_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    scnb_set_error("sgemm/tcgen05 planes: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  k_gemm_planes<<<dim3(n_mt, n_nt, splits), GP_THREADS, smem, st>>>(M, N, n_kt, a_hi, a_lo, b_hi, b_lo, C, ldc, alpha, beta, bias, relu, gate,
                                                                    kt_chunk, stages);
  ++g_tg5_planes_launches;
  return 1;
}

// Returns 1 if launched, 0 if not applicable, <0 on error.
int scnb_gemm_tc_try(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                     float alpha, float beta, const float* bias, int relu, const float* gate, cudaStream_t st) {
  if (transA && transB) return 0;
  const int lay = transA ? 2 : (transB ? 0 : 1);
  const double flops = 2.0 * M * (double)N * K;
  if (tg5_min_flops() <= 0.0 || flops < tg5_min_flops()) return 0;
  if (M < TG5_BM || N < 64 || K < 64) return 0;
  {
    const int rc = tg5_planes_try(lay, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, st);
    if (rc != 0) return rc;
  }
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
