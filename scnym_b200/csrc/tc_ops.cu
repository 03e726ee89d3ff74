// Dense (tensor-core) form of the first layer for sm_100a: densify-then-tcgen05 GEMM.
//   Z[n, H0] = X_csr[n, G] * W1T[G, H0] + b1        (scnym/model.py:211 on CSR input)
// The warp-per-row SpMM (csr_ops.cu) gathers 1 KB of W1T per stored entry through L2 (1 MB per cell at
// G=20 000, 5 %), which caps it at the L2->SM bandwidth.  Here a CTA owns a 128-row tile of X and walks
// the gene axis in tiles of 64 genes:
//   * W1T is pre-split into two bf16 planes (hi = bf16(w), lo = bf16(w - hi)) stored as ready-made
//     128-byte-swizzled K-major shared-memory images, one per gene tile, so one elected thread brings
//     a tile in with a single 1-D bulk copy (cp.async.bulk + mbarrier complete_tx), no tensor map;
//   * 128 threads (one per row) densify the row's entries of that gene tile into two swizzled bf16
//     A tiles (hi, lo) -- everything else in the tile is zero;
//   * one thread issues tcgen05.mma (kind::f16, M=128, N=H0, K=16) x 4 K-slices x 3 products
//       acc += A_lo B_hi + A_hi B_lo + A_hi B_hi        (fp32 accumulate in TMEM)
//     which keeps ~16 mantissa bits per product (relative error ~2^-16 per term, well inside the 1e-4
//     parity budget; a single bf16 or tf32 pass is not: SURVEY.md §7);
//   * tcgen05.commit frees the stage; after the last tile the four densify warps read the accumulator
//     back with tcgen05.ld and write (or, when the gene axis is split over CTAs, atomically add) Z.
// Pipeline: `stages` shared-memory stages (A_hi, A_lo, B_hi, B_lo), mbarriers full_b / full_a / empty.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

#include "tc_common.cuh"

// ---------------------------------------------------------------------------------------------
// W1T fp32 [G, H0]  ->  two bf16 planes, tile kt = genes [KT kt, KT kt + KT): image[n][k] at byte
// n*2KT + swizzle(chunk k>>3, row n) + (k&7)*2   (K-major rows of 2KT bytes, 16-byte chunks XOR-swizzled by row).
// ---------------------------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(256) k_w1_planes(const float* __restrict__ W1T, int G, int H0, long long ld_k, long long ld_n,
                                                   uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  extern __shared__ float wt[];  // [KT][H0 + 1]
  constexpr int CPR = KT / 8;    // 16-byte chunks per row
  const int kt = blockIdx.x, tid = threadIdx.x;
  const int S = H0 + 1;
  if (ld_n == 1) {            // [K, N] storage (first-layer W1T): consecutive threads walk n
    for (int e = tid; e < KT * H0; e += 256) {
      const int k = e / H0, n = e - k * H0;
      const int g = kt * KT + k;
      wt[k * S + n] = g < G ? W1T[(size_t)g * ld_k + n] : 0.f;
    }
  } else {                    // [N, K] storage (nn.Linear weight): consecutive threads walk k
    for (int e = tid; e < KT * H0; e += 256) {
      const int n = e / KT, k = e - n * KT;
      const int g = kt * KT + k;
      wt[k * S + n] = g < G ? W1T[(size_t)g * ld_k + (size_t)n * ld_n] : 0.f;
    }
  }
  __syncthreads();
  const size_t tile = (size_t)kt * H0 * KT;  // elements per plane tile
  for (int e = tid; e < H0 * CPR; e += 256) {
    const int n = e / CPR, c = e % CPR;
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x0 = wt[(8 * c + 2 * j) * S + n], x1 = wt[(8 * c + 2 * j + 1) * S + n];
      split_bf16x2(x0, x1, h[j], l[j]);
    }
    const size_t off = tile + (size_t)n * KT + (size_t)(swz_chunk<KT>((uint32_t)c, (uint32_t)n) >> 1);  // in bf16 elements
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// Tile geometry used for a given first-layer width (shared by the plane builder and the GEMM launcher).
static inline int tc_tile_genes(int H0) { return 32; }

extern "C" int scnb_w1_planes(const float* W1T, int G, int H0, void* planes_hi, void* planes_lo, void* stream) {
  SCNB_CHECK_ARG(W1T && planes_hi && planes_lo && G > 0 && H0 > 0 && H0 % 8 == 0, "w1_planes: bad arguments");
  SCNB_CHECK_ARG((((uintptr_t)planes_hi | (uintptr_t)planes_lo) % 16) == 0, "w1_planes: planes must be 16-byte aligned");
  const int KT = tc_tile_genes(H0);
  const int n_kt = scnb_cdiv(G, KT);
  const size_t smem = (size_t)KT * (H0 + 1) * sizeof(float);
  if (KT == 64) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_w1_planes<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_w1_planes<64><<<n_kt, 256, smem, (cudaStream_t)stream>>>(W1T, G, H0, (long long)H0, 1ll, (uint16_t*)planes_hi, (uint16_t*)planes_lo);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_w1_planes<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_w1_planes<32><<<n_kt, 256, smem, (cudaStream_t)stream>>>(W1T, G, H0, (long long)H0, 1ll, (uint16_t*)planes_hi, (uint16_t*)planes_lo);
  }
  SCNB_CHECK_LAUNCH("w1_planes");
  return SCNB_OK;
}

// Same planes for an nn.Linear weight W [N = out_features, K = in_features] (row-major): the B operand of
// Y = A W^T on the tensor-core path (scnym/model.py:218 hidden layers).
extern "C" int scnb_linear_planes(const float* W, int N, int K, void* planes_hi, void* planes_lo, void* stream) {
  SCNB_CHECK_ARG(W && planes_hi && planes_lo && N > 0 && K > 0 && N % 8 == 0, "linear_planes: bad arguments");
  SCNB_CHECK_ARG((((uintptr_t)planes_hi | (uintptr_t)planes_lo) % 16) == 0, "linear_planes: planes must be 16-byte aligned");
  const int n_kt = scnb_cdiv(K, 32);
  const size_t smem = (size_t)32 * (N + 1) * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_w1_planes<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_w1_planes<32><<<n_kt, 256, smem, (cudaStream_t)stream>>>(W, K, N, 1ll, (long long)K, (uint16_t*)planes_hi, (uint16_t*)planes_lo);
  SCNB_CHECK_LAUNCH("linear_planes");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Eval-mode BatchNorm -> ReLU that ALSO emits its output as the A operand of the next tensor-core GEMM:
// two bf16 planes tiled [row tile of 256][k tile of 32] in the SWIZZLE_64B shared-memory image layout
// (so the GEMM brings a tile in with one bulk copy).  Rows >= n_rows of the last tile are written as zeros.
// One thread per (row, 8-column group).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bn_eval_planes(const float* __restrict__ Z, int n_rows, int H,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const float* __restrict__ rmean, const float* __restrict__ rvar,
                                                        int relu, float* __restrict__ A_out, float* __restrict__ pre_out,
                                                        uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  // per-column constants once per CTA: the IEEE sqrt + division per ELEMENT made this pass issue-bound (0.9 TB/s)
  extern __shared__ float bnp_smem[];
  float* s_mean = bnp_smem;
  float* s_inv = s_mean + H;
  float* s_gam = s_inv + H;
  float* s_bet = s_gam + H;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    s_mean[c] = rmean[c];
    s_inv[c] = 1.0f / sqrtf(rvar[c] + 1e-5f);
    s_gam[c] = gamma[c];
    s_bet[c] = beta[c];
  }
  __syncthreads();
  const int Hg = H >> 3;
  const long long rows_pad = ((long long)n_rows + 255) / 256 * 256;
  const long long total = rows_pad * Hg;
  const int n_kt = H >> 5;
  // (256 % Hg == 0: a thread keeps its column group and walks rows -- no 64-bit division per element)
  const bool fixed_cg = (256 % Hg) == 0;
  const int rows_per_it = fixed_cg ? 256 / Hg : 0;
  const long long e0 = fixed_cg ? (long long)blockIdx.x * rows_per_it + threadIdx.x / Hg : (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long e_end = fixed_cg ? rows_pad : total;
  const long long e_step = fixed_cg ? (long long)gridDim.x * rows_per_it : (long long)gridDim.x * blockDim.x;
  const int cg_fixed = fixed_cg ? (int)(threadIdx.x % Hg) : 0;
  for (long long e = e0; e < e_end; e += e_step) {
    const long long r = fixed_cg ? e : e / Hg;
    const int cg = fixed_cg ? cg_fixed : (int)(e - r * Hg), c0 = cg * 8;
    float y[8];
    if (r < n_rows) {
      const float4 z0 = *reinterpret_cast<const float4*>(Z + r * H + c0), z1 = *reinterpret_cast<const float4*>(Z + r * H + c0 + 4);
      const float z[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        y[j] = (z[j] - s_mean[c]) * s_inv[c] * s_gam[c] + s_bet[c];
      }
      if (pre_out) {
        *reinterpret_cast<float4*>(pre_out + r * H + c0) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(pre_out + r * H + c0 + 4) = make_float4(y[4], y[5], y[6], y[7]);
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.f);
      }
      if (A_out) {
        *reinterpret_cast<float4*>(A_out + r * H + c0) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(A_out + r * H + c0 + 4) = make_float4(y[4], y[5], y[6], y[7]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = 0.f;
    }
    if (hi) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_bf16x2(y[2 * j], y[2 * j + 1], h[j], l[j]);
      const long long rt = r >> 8;
      const uint32_t rr = (uint32_t)(r & 255), rl = rr & 127u;
      const int kt = cg >> 2;
      const size_t tile = ((size_t)rt * n_kt + kt) * (256 * 32);             // bf16 elements per (row tile, k tile)
      const size_t off = tile + (size_t)rr * 32 + (size_t)(swz_chunk<32>((uint32_t)(cg & 3), rl) >> 1);
      *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

extern "C" int scnb_bn_eval_planes(const float* Z, int n_rows, int H, const float* gamma, const float* beta,
                                   const float* running_mean, const float* running_var, int relu, float* A_out, float* pre_out,
                                   void* planes_hi, void* planes_lo, void* stream) {
  SCNB_CHECK_ARG(Z && gamma && beta && running_mean && running_var && n_rows >= 0 && H > 0 && H % 32 == 0,
                 "bn_eval_planes: bad arguments (H must be a multiple of 32)");
  SCNB_CHECK_ARG((planes_hi == nullptr) == (planes_lo == nullptr), "bn_eval_planes: give both planes or neither");
  SCNB_CHECK_ARG((((uintptr_t)Z | (uintptr_t)A_out | (uintptr_t)pre_out | (uintptr_t)planes_hi | (uintptr_t)planes_lo) % 16) == 0,
                 "bn_eval_planes: buffers must be 16-byte aligned");
  if (n_rows == 0) return SCNB_OK;
  const long long total = ((long long)n_rows + 255) / 256 * 256 * (H / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  SCNB_CHECK_ARG(H <= 2048, "bn_eval_planes: H must be <= 2048 (got %d)", H);
  k_bn_eval_planes<<<blocks, 256, 4 * (size_t)H * sizeof(float), (cudaStream_t)stream>>>(Z, n_rows, H, gamma, beta, running_mean, running_var, relu, A_out,
                                                            pre_out, (uint16_t*)planes_hi, (uint16_t*)planes_lo);
  SCNB_CHECK_LAUNCH("bn_eval_planes");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Roles: warp 0 = W1T tile loader, warp 1 = TMEM allocation + MMA issuer, warps 2..2+4*HALVES-1 = densify
// (one thread per row of the CTA's tile) and, afterwards, epilogue (TMEM -> registers -> Z).
// ---------------------------------------------------------------------------------------------
// DENSE_A: the A operand is not densified from a CSR but arrives as pre-tiled bf16 planes (a_hi / a_lo, written by
// scnb_bn_eval_planes) that the loader thread bulk-copies next to the W tiles; warps 2.. then only run the epilogue.
// TRN (split-K training form, H0 a multiple of 128): the accumulator is kept TRANSPOSED, D^T[H0, rows] = W1T^T X^T -- the
// W1T plane tile is the MMA's A operand (M = 128 columns of Z per accumulator), the densified X tile its B operand
// (N = 256 rows).  TMEM lanes are then columns of Z, so a warp of the epilogue adds 32 CONSECUTIVE floats of one row of Z
// per instruction (one 128-byte line per atomic request) instead of 32 rows x 16 bytes: the 148 partial products of the
// gene-axis split meet in L2 with a quarter of the atomic transactions.
// Optional epilogue of the predict form: eval-mode BatchNorm -> ReLU of the first layer's output written straight as the bf16
// hi / lo operand planes of the next Linear (the layout of scnb_bn_eval_planes), instead of fp32 Z: the 77 MB per chunk of Z
// are neither written nor read back, and the separate BatchNorm launch disappears.  Same arithmetic, operation for operation,
// as scnb_csr_linear_fwd_tc_i64 followed by scnb_bn_eval_planes (bit-identical planes).
struct FwdBnPlanes {
  const float* gamma;
  const float* beta;
  const float* rmean;
  const float* rvar;
  uint16_t* hi;
  uint16_t* lo;
  int relu;
};

// ... or, for the LAST hidden layer, straight into the classifier: BatchNorm -> ReLU -> logits[r, :] (+)= A[r, :] Wc^T + bc
// (scnym/model.py:229), the pre-ReLU BatchNorm output optionally kept as the embedding (scnym/api.py:700).  A thread owns a
// row: it walks its accumulator 32 columns at a time and carries up to FWD_HEAD_CMAX running dot products; Wc sits in the
// freed pipeline stages.  Neither Z nor A of the last layer ever exist in HBM.
#define FWD_HEAD_CMAX 16
struct FwdHead {
  const float* Wc;      // [C, N]
  const float* bc;      // [C] or NULL
  float* logits;        // [n, C]
  float* pre_out;       // [n, N] or NULL
  int C;
  int accumulate;       // logits += (ensembles: scnym/predict.py:178-192 sums the models' outputs)
};

template <typename IdxT, int HALVES, int KT, bool DENSE_A, bool TRN>
__global__ void __launch_bounds__(64 + 128 * HALVES, 1)
k_csr_dense_fwd_tc(const IdxT* __restrict__ indptr, const int32_t* __restrict__ idx, const float* __restrict__ val, int n_rows,
                   const uint8_t* __restrict__ planes_hi, const uint8_t* __restrict__ planes_lo, int N, int n_ktiles,
                   const float* __restrict__ bias, float* __restrict__ Z, int stages, int tmem_cols,
                   const uint8_t* __restrict__ a_hi_planes, const uint8_t* __restrict__ a_lo_planes,
                   const int32_t* __restrict__ fwd_splits, long long* __restrict__ dbg, FwdBnPlanes bnp, FwdHead head) {
#define FWDTC_GSTAMP(k) do { if (dbg) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[4 * (blockIdx.y * gridDim.x + blockIdx.x) + (k)] = (long long)t_; } } while (0)
  constexpr uint32_t ROWB = 2 * KT;                       // bytes per operand row (bf16)
  constexpr uint32_t A_HALF = TC_M * ROWB;                // one 128-row A block of one plane
  constexpr uint32_t A_PLANE = HALVES * A_HALF;
  constexpr int CPR = KT / 8;                             // 16-byte chunks per row
  constexpr int ROWS = TC_M * HALVES;
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;           // swizzled tiles need (at most) 1024-byte alignment
  uint8_t* sm = tc_smem_raw + (base - raw);
  const uint32_t B_BYTES = (uint32_t)N * ROWB;
  const uint32_t STAGE_BYTES = 2u * A_PLANE + 2u * B_BYTES;   // A_hi | A_lo | B_hi | B_lo
  const uint32_t bars = base + (uint32_t)stages * STAGE_BYTES;  // full_b[S] | full_a[S] | empty[S] | acc | tmem slot
  const uint32_t bar_full_b = bars, bar_full_a = bars + 8u * stages, bar_empty = bars + 16u * stages,
                 bar_acc = bars + 24u * stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (size_t)stages * STAGE_BYTES + 24 * stages + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // this CTA's share of the gene tiles
  const int per = (n_ktiles + gridDim.y - 1) / gridDim.y;
  const int kt0 = blockIdx.y * per, kt1 = min(n_ktiles, kt0 + per);
  const int n_it = max(kt1 - kt0, 0);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full_b + 8u * s, 1);
      mbar_init(bar_full_a + 8u * s, ROWS);
      mbar_init(bar_empty + 8u * s, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== W1T loader: one bulk copy per plane per gene tile =====
    if (lane == 0) {
      for (int i = 0; i < n_it; ++i) {
        const int s = i % stages;
        const uint32_t ph = (uint32_t)(i / stages) & 1u;
        mbar_wait(bar_empty + 8u * s, ph ^ 1u);
        const uint32_t stage = base + (uint32_t)s * STAGE_BYTES;
        mbar_arrive_expect_tx(bar_full_b + 8u * s, 2u * B_BYTES + (DENSE_A ? 2u * A_PLANE : 0u));
        const size_t off = (size_t)(kt0 + i) * B_BYTES;
        bulk_g2s(stage + 2u * A_PLANE, planes_hi + off, B_BYTES, bar_full_b + 8u * s);
        bulk_g2s(stage + 2u * A_PLANE + B_BYTES, planes_lo + off, B_BYTES, bar_full_b + 8u * s);
        if (DENSE_A) {
          const size_t aoff = ((size_t)blockIdx.x * n_ktiles + (size_t)(kt0 + i)) * A_PLANE;
          bulk_g2s(stage, a_hi_planes + aoff, A_PLANE, bar_full_b + 8u * s);
          bulk_g2s(stage + A_PLANE, a_lo_planes + aoff, A_PLANE, bar_full_b + 8u * s);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((TRN ? ROWS : N) >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
      for (int i = 0; i < n_it; ++i) {
        const int s = i % stages;
        const uint32_t ph = (uint32_t)(i / stages) & 1u;
        mbar_wait(bar_full_b + 8u * s, ph);
        if (!DENSE_A) mbar_wait(bar_full_a + 8u * s, ph);
        tc_fence_after();
        const uint32_t stage = base + (uint32_t)s * STAGE_BYTES;
        const uint64_t b_hi = umma_desc_kmajor<KT>(stage + 2u * A_PLANE), b_lo = umma_desc_kmajor<KT>(stage + 2u * A_PLANE + B_BYTES);
        if (TRN) {
          const uint64_t x_hi = umma_desc_kmajor<KT>(stage), x_lo = umma_desc_kmajor<KT>(stage + A_PLANE);
          for (int h = 0; h < (N >> 7); ++h) {       // 128 columns of Z per accumulator: TMEM columns [h*ROWS, h*ROWS + ROWS)
            const uint64_t w_hi = umma_desc_kmajor<KT>(stage + 2u * A_PLANE + (uint32_t)h * A_HALF);
            const uint64_t w_lo = umma_desc_kmajor<KT>(stage + 2u * A_PLANE + B_BYTES + (uint32_t)h * A_HALF);
            const uint32_t d = tmem_base + (uint32_t)(h * ROWS);
#pragma unroll
            for (int k = 0; k < KT / 16; ++k) {
              const uint64_t ko = (uint64_t)(k * 2);
              tc_mma_bf16(d, w_lo + ko, x_hi + ko, idesc, (i > 0 || k > 0) ? 1u : 0u);
              tc_mma_bf16(d, w_hi + ko, x_lo + ko, idesc, 1u);
              tc_mma_bf16(d, w_hi + ko, x_hi + ko, idesc, 1u);
            }
          }
          tc_commit(bar_empty + 8u * s);
          continue;
        }
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          const uint64_t a_hi = umma_desc_kmajor<KT>(stage + (uint32_t)h * A_HALF);
          const uint64_t a_lo = umma_desc_kmajor<KT>(stage + A_PLANE + (uint32_t)h * A_HALF);
          const uint32_t d = tmem_base + (uint32_t)(h * N);   // accumulator of this row block: columns [h*N, h*N + N)
#pragma unroll
          for (int k = 0; k < KT / 16; ++k) {
            const uint64_t ko = (uint64_t)(k * 2);  // 32 bytes per K=16 slice, in 16-byte units
            tc_mma_bf16(d, a_lo + ko, b_hi + ko, idesc, (i > 0 || k > 0) ? 1u : 0u);
            tc_mma_bf16(d, a_hi + ko, b_lo + ko, idesc, 1u);
            tc_mma_bf16(d, a_hi + ko, b_hi + ko, idesc, 1u);
          }
        }
        tc_commit(bar_empty + 8u * s);   // stage reusable once these MMAs have read it
      }
      if (n_it > 0) tc_commit(bar_acc);  // accumulators complete
    }
  } else {
    // ===== densify (one thread per row of the tile), then epilogue =====
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;             // which 128-row block
    const int rl = q * 32 + lane;                 // row within the block == TMEM lane
    const int r = half * TC_M + rl;               // row within the CTA tile
    const long long grow = (long long)blockIdx.x * ROWS + r;
    long long p = 0, pend = 0;
    if (tid == 64) FWDTC_GSTAMP(0);
    if (!DENSE_A && fwd_splits && grow < n_rows) {
      // split points of this CTA's gene range, found off the critical path (scnb_row_fwd_splits): the ten dependent loads
      // of the binary search below were ~5 us at the head of every step
      p = (long long)fwd_splits[(size_t)blockIdx.y * n_rows + grow];
      pend = (long long)fwd_splits[(size_t)(blockIdx.y + 1) * n_rows + grow];
    } else if (!DENSE_A && grow < n_rows) {
      p = (long long)indptr[grow];
      pend = (long long)indptr[grow + 1];
    }
    if (!fwd_splits && kt0 > 0 && p < pend) {  // first stored gene >= kt0*KT (rows are sorted by gene)
      const int target = kt0 * KT;
      long long lo_p = p, hi_p = pend;
      while (lo_p < hi_p) {
        const long long mid = (lo_p + hi_p) >> 1;
        if (idx[mid] < target) lo_p = mid + 1; else hi_p = mid;
      }
      p = lo_p;
    }
    // Entries are consumed from a register window of TC_W (gene, value) pairs; the next window is already
    // in flight while the current one is consumed, so the per-lane memory latency of this divergent walk
    // is paid once per window and off the critical path (a warp otherwise stalls on its slowest lane at
    // every entry).
    const int NONE = 0x7fffffff;
    // (Two windows in flight behind the current one were measured: no gain at 20 % density -- config 5's first layer stays at
    // 375 us -- and 5.6 % slower at 5 % density: predict's first layer 1.71 -> 1.81 ms per chunk.  One window ahead it is.)
    int cw[TC_W], cn[TC_W];
    float vw[TC_W], vn[TC_W];
#pragma unroll
    for (int w = 0; w < TC_W; ++w) {
      cw[w] = p + w < pend ? idx[p + w] : NONE;
      vw[w] = p + w < pend ? val[p + w] : 0.f;
    }
#pragma unroll
    for (int w = 0; w < TC_W; ++w) {
      cn[w] = p + TC_W + w < pend ? idx[p + TC_W + w] : NONE;
      vn[w] = p + TC_W + w < pend ? val[p + TC_W + w] : 0.f;
    }
    // The head of the window is always cw[0] (consumed entries are shifted out), so every lane that still has an
    // entry for this tile executes the SAME store instruction: one shared-memory wavefront group per step of the
    // loop instead of one per lane (the MMA reads its operands from the same shared memory; stores that
    // serialise per lane take the bandwidth it needs).
    int cnt = TC_W;                                // entries of the current window not yet consumed (incl. NONE padding)
    if (tid == 64) FWDTC_GSTAMP(1);
    for (int i = 0; i < (DENSE_A ? 0 : n_it); ++i) {
      const int s = i % stages;
      const uint32_t ph = (uint32_t)(i / stages) & 1u;
      mbar_wait(bar_empty + 8u * s, ph ^ 1u);
      const uint32_t row_hi = base + (uint32_t)s * STAGE_BYTES + (uint32_t)half * A_HALF + (uint32_t)rl * ROWB;
      const uint32_t row_lo = row_hi + A_PLANE;
#pragma unroll
      for (int j = 0; j < CPR; ++j) {              // clear the row (lanes stagger the chunk: no bank conflicts)
        const uint32_t ch = (uint32_t)((j + lane) & (CPR - 1)) << 4;
        st_shared_zero16(row_hi + ch);
        st_shared_zero16(row_lo + ch);
      }
      const int k0 = (kt0 + i) * KT, kend = k0 + KT;
      while (true) {
        while (cw[0] < kend) {
          const uint32_t g = (uint32_t)(cw[0] - k0);
          const __nv_bfloat16 h = __float2bfloat16_rn(vw[0]);
          const __nv_bfloat16 l = __float2bfloat16_rn(vw[0] - __bfloat162float(h));
          const uint32_t off = swz_chunk<KT>(g >> 3, (uint32_t)rl) + ((g & 7u) << 1);
          st_shared_u16(row_hi + off, __bfloat16_as_ushort(h));
          st_shared_u16(row_lo + off, __bfloat16_as_ushort(l));
#pragma unroll
          for (int w = 0; w + 1 < TC_W; ++w) {     // shift the window: the next entry becomes the head
            cw[w] = cw[w + 1];
            vw[w] = vw[w + 1];
          }
          cw[TC_W - 1] = NONE;
          --cnt;
        }
        if (cnt > 0) break;                        // the head belongs to a later tile (or the row is finished)
        // window exhausted: the prefetched one becomes current, the one after it is requested now
        p += TC_W;
#pragma unroll
        for (int w = 0; w < TC_W; ++w) {
          cw[w] = cn[w];
          vw[w] = vn[w];
        }
#pragma unroll
        for (int w = 0; w < TC_W; ++w) {
          cn[w] = p + TC_W + w < pend ? idx[p + TC_W + w] : NONE;
          vn[w] = p + TC_W + w < pend ? val[p + TC_W + w] : 0.f;
        }
        cnt = TC_W;
        if (cw[0] >= kend) break;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
      mbar_arrive(bar_full_a + 8u * s);
    }
    if (TRN && n_it > 0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      if (tid == 64) FWDTC_GSTAMP(2);
      // accumulator hh holds columns [128 hh, 128 hh + 128) of Z; a CTA has HALVES warp groups for N / 128 accumulators
      for (int hh = half; hh * TC_M < N; hh += HALVES) {
        const int col = hh * TC_M + rl;                 // column of Z held by this TMEM lane
        const float bc = (bias && blockIdx.y == 0) ? bias[col] : 0.f;
        // every gene split adds into the same rows of Z: the splits start at different 32-row groups, so that at any moment
        // the CTAs of a row tile are spread over different lines (the L2 atomic unit serialises per address)
        for (int rr = 0; rr < ROWS; rr += 32) {
          const int r0 = (rr + 32 * (int)blockIdx.y) & (ROWS - 1);
          if ((long long)blockIdx.x * ROWS + r0 >= n_rows) continue;
          uint32_t v[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hh * ROWS + r0);
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
          float* out = Z + ((size_t)blockIdx.x * ROWS + r0) * N + col;
          const int n_here = (int)min((long long)32, (long long)n_rows - ((long long)blockIdx.x * ROWS + r0));
          // full 32-row groups at the default width: immediate row offsets, no predicates (a third of the epilogue's
          // instructions were 64-bit address arithmetic, as in the dW1 kernel)
          if (n_here == 32 && N == 256) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(out + j * 256, __uint_as_float(v[j]) + bc);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < n_here) atomicAdd(out + (size_t)j * N, __uint_as_float(v[j]) + bc);
          }
        }
      }
    } else if (n_it > 0 && head.logits) {
      // ===== predict form, last hidden layer: Z + b -> BatchNorm(eval) [-> embedding] -> ReLU -> classifier logits =====
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      float* cst = reinterpret_cast<float*>(sm);          // [mean | 1/sqrt(var + eps) | gamma | beta | bias], N floats each
      float* Ws = cst + 5 * N;                            // [C][N]
      for (int c = tid - 64; c < N; c += 128 * HALVES) {
        cst[c] = bnp.rmean[c];
        cst[N + c] = 1.0f / sqrtf(bnp.rvar[c] + 1e-5f);
        cst[2 * N + c] = bnp.gamma[c];
        cst[3 * N + c] = bnp.beta[c];
        cst[4 * N + c] = bias ? bias[c] : 0.f;
      }
      for (int e = tid - 64; e < head.C * N; e += 128 * HALVES) Ws[e] = head.Wc[e];
      asm volatile("bar.sync 1, %0;" ::"n"(128 * HALVES) : "memory");
      float lg[FWD_HEAD_CMAX];
#pragma unroll
      for (int c = 0; c < FWD_HEAD_CMAX; ++c) lg[c] = 0.f;
      for (int cb = 0; cb < N; cb += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * N + cb);
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
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          float a4[4], y4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = cb + j4 + u;
            float z = __uint_as_float(v[j4 + u]);
            z += cst[4 * N + c];
            y4[u] = (z - cst[c]) * cst[N + c] * cst[2 * N + c] + cst[3 * N + c];
            a4[u] = bnp.relu ? fmaxf(y4[u], 0.f) : y4[u];
          }
          if (head.pre_out && grow < n_rows)
            *reinterpret_cast<float4*>(head.pre_out + (size_t)grow * N + cb + j4) = make_float4(y4[0], y4[1], y4[2], y4[3]);
#pragma unroll
          for (int c = 0; c < FWD_HEAD_CMAX; ++c) {
            if (c < head.C) {
              const float4 w = *reinterpret_cast<const float4*>(Ws + (size_t)c * N + cb + j4);
              lg[c] = fmaf(a4[0], w.x, lg[c]);
              lg[c] = fmaf(a4[1], w.y, lg[c]);
              lg[c] = fmaf(a4[2], w.z, lg[c]);
              lg[c] = fmaf(a4[3], w.w, lg[c]);
            }
          }
        }
      }
      if (grow < n_rows) {
#pragma unroll
        for (int c = 0; c < FWD_HEAD_CMAX; ++c) {
          if (c < head.C) {
            float o = lg[c] + (head.bc ? head.bc[c] : 0.f);
            float* dst = head.logits + (size_t)grow * head.C + c;
            if (head.accumulate) o += *dst;
            *dst = o;
          }
        }
      }
    } else if (n_it > 0 && bnp.hi) {
      // ===== predict form, fused: Z + b -> BatchNorm(eval) -> ReLU -> operand planes of the next layer (no split over genes) =====
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      // per-column constants in the (now free) first stage: [mean | 1/sqrt(var + eps) | gamma | beta | bias], N floats each
      float* cst = reinterpret_cast<float*>(sm);
      for (int c = tid - 64; c < N; c += 128 * HALVES) {
        cst[c] = bnp.rmean[c];
        cst[N + c] = 1.0f / sqrtf(bnp.rvar[c] + 1e-5f);
        cst[2 * N + c] = bnp.gamma[c];
        cst[3 * N + c] = bnp.beta[c];
        cst[4 * N + c] = bias ? bias[c] : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * HALVES) : "memory");
      const int n_kt_out = N >> 5;
      const uint32_t rr = (uint32_t)(grow & 255);
      const size_t tile_row = (size_t)(grow >> 8) * n_kt_out;
      for (int cb = 0; cb < N; cb += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * N + cb);
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
        const size_t tile = (tile_row + (size_t)(cb >> 5)) * (256 * 32) + (size_t)rr * 32;
#pragma unroll
        for (int jc = 0; jc < 4; ++jc) {
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = cb + 8 * jc + j;
            float z = __uint_as_float(v[8 * jc + j]);
            z += cst[4 * N + c];
            float yy = (z - cst[c]) * cst[N + c] * cst[2 * N + c] + cst[3 * N + c];
            if (bnp.relu) yy = fmaxf(yy, 0.f);
            y[j] = grow < n_rows ? yy : 0.f;       // rows past the chunk: zeros, as scnb_bn_eval_planes writes them
          }
          uint32_t h[4], l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) split_bf16x2(y[2 * j], y[2 * j + 1], h[j], l[j]);
          const size_t off = tile + (size_t)(swz_chunk<32>((uint32_t)jc, rr & 127u) >> 1);
          *reinterpret_cast<uint4*>(bnp.hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(bnp.lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
    } else if (n_it > 0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      const bool split = gridDim.y > 1;
      const bool add_bias = bias && blockIdx.y == 0;
      for (int cb = 0; cb < N; cb += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * N + cb);
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
        if (grow < n_rows) {
          float* out = Z + (size_t)grow * N + cb;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]));
            if (add_bias) {
              o.x += bias[cb + j]; o.y += bias[cb + j + 1]; o.z += bias[cb + j + 2]; o.w += bias[cb + j + 3];
            }
            if (split) atomicAdd(reinterpret_cast<float4*>(out + j), o);
            else *reinterpret_cast<float4*>(out + j) = o;
          }
        }
      }
    }
  }
  if (tid == 64) FWDTC_GSTAMP(3);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
#undef FWDTC_GSTAMP
}

// splits[b * nb + r] = first position p of row r (absolute index into b_idx) with b_idx[p] >= b * GT, b = 0..n_blk;
// optionally a second table with its own block size in the same launch (the split points of the tensor-core forward).
__global__ void __launch_bounds__(128) k_row_gene_splits(const int32_t* __restrict__ b_indptr, const int32_t* __restrict__ b_idx,
                                                         int nb, int GT, int n_blk, int32_t* __restrict__ splits, int GT2 = 0,
                                                         int n_blk2 = 0, int32_t* __restrict__ splits2 = nullptr) {
  const int r = blockIdx.x;
  const int s = b_indptr[r], e = b_indptr[r + 1];
  const int n1 = splits ? n_blk + 1 : 0, n2 = splits2 ? n_blk2 + 1 : 0;
  for (int b = threadIdx.x; b < n1 + n2; b += blockDim.x) {
    const bool second = b >= n1;
    const int bb = second ? b - n1 : b;
    const int target = bb * (second ? GT2 : GT);
    int lo = s, hi = e;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (b_idx[mid] < target) lo = mid + 1; else hi = mid;
    }
    (second ? splits2 : splits)[(size_t)bb * nb + r] = lo;
  }
}

// gene-axis splits of the training forward: until the grid covers the 148 SMs
static inline int fwd_tc_auto_splits(int n, int G, int rows_per_cta, int KT) {
  const int row_tiles = scnb_cdiv(n, rows_per_cta), n_kt = scnb_cdiv(G, KT);
  int ks = 148 / row_tiles;
  if (ks < 1) ks = 1;
  return ks > n_kt ? n_kt : ks;
}

template <typename IdxT, int HALVES, int KT, bool DENSE_A, bool TRN = false>
static int launch_dense_fwd_tc(const IdxT* indptr, const int32_t* idx, const float* val, int n, int G, int H0, const void* hi,
                               const void* lo, const float* b1, float* Z, int ksplits, cudaStream_t st, const void* a_hi = nullptr,
                               const void* a_lo = nullptr, const int32_t* fwd_splits = nullptr, FwdBnPlanes bnp = FwdBnPlanes{},
                               FwdHead head = FwdHead{}) {
  constexpr int ROWS = TC_M * HALVES;
  const int n_kt = scnb_cdiv(G, KT);
  const int row_tiles = scnb_cdiv(n, ROWS);
  const bool prezeroed = ksplits < 0;   // the caller has already cleared Z (off its critical path)
  if (ksplits <= 0) ksplits = fwd_tc_auto_splits(n, G, ROWS, KT);
  if (ksplits > n_kt) ksplits = n_kt;
  const size_t stage_bytes = 2 * (size_t)ROWS * 2 * KT + 2 * (size_t)H0 * 2 * KT;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > 6) stages = 6;
  if (stages < 2) {
    scnb_set_error("csr_linear_fwd_tc: H0=%d needs more shared memory than one SM has", H0);
    return SCNB_ERR_UNSUPPORTED;
  }
  const size_t smem = 1024 + stages * stage_bytes + 24 * stages + 16 + 16;
  if (TRN && (ksplits <= 1 || H0 % 128 != 0))   // the transposed accumulator only pays for (and is only built for) split-K sums
    return launch_dense_fwd_tc<IdxT, HALVES, KT, DENSE_A, false>(indptr, idx, val, n, G, H0, hi, lo, b1, Z, prezeroed ? -1 : ksplits, st,
                                                                 a_hi, a_lo, fwd_splits, bnp, head);
  int tmem_cols = 32;
  while (tmem_cols < (TRN ? (H0 / 128) * ROWS : HALVES * H0)) tmem_cols *= 2;
  cudaError_t e = cudaFuncSetAttribute(k_csr_dense_fwd_tc<IdxT, HALVES, KT, DENSE_A, TRN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    scnb_set_error("csr_linear_fwd_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  if (ksplits > 1 && !prezeroed) {  // partial products of the gene-axis splits are added atomically
    e = cudaMemsetAsync(Z, 0, (size_t)n * H0 * sizeof(float), st);
    if (e != cudaSuccess) {
      scnb_set_error("csr_linear_fwd_tc: memset failed: %s", cudaGetErrorString(e));
      return SCNB_ERR_CUDA;
    }
  }
  dim3 grid(row_tiles, ksplits);
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* ev = getenv("SCNB_FWDTC_DBG");
    dbg_on = (ev && ev[0] == '1') ? 1 : 0;
    if (dbg_on) cudaMalloc(&dbg, 4 * 4096 * sizeof(long long));
  }
  const int n_cta = row_tiles * ksplits;
  const bool stamp = dbg_on && !DENSE_A && n_cta <= 4096;
  k_csr_dense_fwd_tc<IdxT, HALVES, KT, DENSE_A, TRN><<<grid, 64 + 128 * HALVES, smem, st>>>(
      indptr, idx, val, n, (const uint8_t*)hi, (const uint8_t*)lo, H0, n_kt, b1, Z, stages, tmem_cols, (const uint8_t*)a_hi,
      (const uint8_t*)a_lo, fwd_splits, stamp ? dbg : nullptr, bnp, head);
  SCNB_CHECK_LAUNCH("csr_linear_fwd_tc");
  if (stamp) {  // debugging aid only (synchronises): per-CTA wall-clock stamps (ns)
    static long long hc[4 * 4096];
    cudaStreamSynchronize(st);
    cudaMemcpy(hc, dbg, sizeof(long long) * 4 * n_cta, cudaMemcpyDeviceToHost);
    long long t0 = 1LL << 62;
    for (int b = 0; b < n_cta; ++b) if (hc[4 * b] && hc[4 * b] < t0) t0 = hc[4 * b];
    const char* nm[4] = {"set up", "row windows loaded", "accumulator ready", "end"};
    for (int k = 0; k < 4; ++k) {
      long long mn = 1LL << 62, mx = 0, sum = 0;
      int cnt = 0;
      for (int b = 0; b < n_cta; ++b) {
        if (!hc[4 * b + k] || !hc[4 * b + 2]) continue;   // (CTAs without gene tiles never reach stamp 2)
        long long v = hc[4 * b + k] - t0; if (v < mn) mn = v; if (v > mx) mx = v; sum += v; ++cnt;
      }
      fprintf(stderr, "[fwdtc] %s: min %lld mean %lld max %lld ns (%d of %d CTAs, %d stages, splits table %s)\n", nm[k], mn,
              cnt ? sum / cnt : 0, mx, cnt, n_cta, stages, fwd_splits ? "yes" : "no");
    }
    cudaMemset(dbg, 0, sizeof(long long) * 4 * n_cta);
  }
  return SCNB_OK;
}

static int tc_check(int n, int G, int H0, const void* hi, const void* lo, const float* Z) {
  SCNB_CHECK_ARG(n >= 0 && G > 0 && hi && lo && Z, "csr_linear_fwd_tc: bad arguments");
  SCNB_CHECK_ARG(H0 >= 32 && H0 <= 256 && H0 % 32 == 0, "csr_linear_fwd_tc: H0 must be a multiple of 32 in [32, 256] (got %d)", H0);
  SCNB_CHECK_ARG((((uintptr_t)hi | (uintptr_t)lo | (uintptr_t)Z) % 16) == 0, "csr_linear_fwd_tc: buffers must be 16-byte aligned");
  return SCNB_OK;
}

extern "C" int scnb_csr_linear_fwd_tc(const int32_t* b_indptr, const int32_t* b_idx, const float* b_val, int n, int G, int H0,
                                      const void* planes_hi, const void* planes_lo, const float* b1, float* Z, int ksplits,
                                      void* stream) {
  SCNB_CHECK_ARG(b_indptr && b_idx && b_val, "csr_linear_fwd_tc: null CSR");
  int rc = tc_check(n, G, H0, planes_hi, planes_lo, Z);
  if (rc != SCNB_OK) return rc;
  if (n == 0) return SCNB_OK;
  // SCNB_FWD_ROWS=128: 128-row CTAs (twice the row tiles, half the gene splits: half the split-K atomics, twice the
  // operand-plane reads out of L2); default 256
  static int rows128 = -1;
  if (rows128 < 0) {
    const char* ev = getenv("SCNB_FWD_ROWS");
    rows128 = (ev && atoi(ev) == 128) ? 1 : 0;
  }
  if (rows128)
    return launch_dense_fwd_tc<int32_t, 1, 32, false, true>(b_indptr, b_idx, b_val, n, G, H0, planes_hi, planes_lo, b1, Z, ksplits,
                                                            (cudaStream_t)stream);
  return launch_dense_fwd_tc<int32_t, 2, 32, false, true>(b_indptr, b_idx, b_val, n, G, H0, planes_hi, planes_lo, b1, Z, ksplits,
                                             (cudaStream_t)stream);
}

// Geometry of scnb_csr_linear_fwd_tc with automatic splits: number of gene-axis splits and genes per split.
extern "C" int scnb_csr_linear_fwd_tc_geometry(int n, int G, int H0, int* ksplits, int* genes_per_split) {
  SCNB_CHECK_ARG(n > 0 && G > 0 && ksplits && genes_per_split, "csr_linear_fwd_tc_geometry: bad arguments");
  (void)H0;
  const int KT = 32, ROWS = 256;
  const int ks = fwd_tc_auto_splits(n, G, ROWS, KT);
  *ksplits = ks;
  *genes_per_split = scnb_cdiv(scnb_cdiv(G, KT), ks) * KT;
  return SCNB_OK;
}

// fwd_splits[s * n + r] = first position p of batch row r with b_idx[p] >= s * genes_per_split, s = 0..ksplits: where
// each CTA of the split forward starts (and ends) in every row.  Depends only on the batch, so the engine computes it
// on the prologue branch, one step ahead.
extern "C" int scnb_row_fwd_splits(const int32_t* b_indptr, const int32_t* b_idx, int n, int G, int H0, int32_t* fwd_splits,
                                   void* stream) {
  SCNB_CHECK_ARG(b_indptr && b_idx && fwd_splits && n >= 0 && G > 0, "row_fwd_splits: bad arguments");
  if (n == 0) return SCNB_OK;
  int ks = 0, gp = 0;
  scnb_csr_linear_fwd_tc_geometry(n, G, H0, &ks, &gp);
  k_row_gene_splits<<<n, 128, 0, (cudaStream_t)stream>>>(b_indptr, b_idx, n, gp, ks, fwd_splits);
  SCNB_CHECK_LAUNCH("row_fwd_splits");
  return SCNB_OK;
}

// scnb_csr_linear_fwd_tc with the split points of scnb_row_fwd_splits (automatic splits only: ksplits 0, or < 0 when Z is
// already zeroed).
extern "C" int scnb_csr_linear_fwd_tc_splits(const int32_t* b_indptr, const int32_t* b_idx, const float* b_val, int n, int G, int H0,
                                             const void* planes_hi, const void* planes_lo, const float* b1, float* Z, int ksplits,
                                             const int32_t* fwd_splits, void* stream) {
  SCNB_CHECK_ARG(b_indptr && b_idx && b_val && fwd_splits, "csr_linear_fwd_tc_splits: null CSR or split table");
  SCNB_CHECK_ARG(ksplits <= 0, "csr_linear_fwd_tc_splits: the split table is laid out for the automatic geometry (ksplits <= 0)");
  int rc = tc_check(n, G, H0, planes_hi, planes_lo, Z);
  if (rc != SCNB_OK) return rc;
  if (n == 0) return SCNB_OK;
  return launch_dense_fwd_tc<int32_t, 2, 32, false, true>(b_indptr, b_idx, b_val, n, G, H0, planes_hi, planes_lo, b1, Z, ksplits,
                                                          (cudaStream_t)stream, nullptr, nullptr, fwd_splits);
}

extern "C" int scnb_csr_linear_fwd_tc_i64(const int64_t* indptr, const int32_t* indices, const float* data, int n, int G, int H0,
                                          const void* planes_hi, const void* planes_lo, const float* b1, float* Z, int ksplits,
                                          void* stream) {
  SCNB_CHECK_ARG(indptr && indices && data, "csr_linear_fwd_tc_i64: null CSR");
  int rc = tc_check(n, G, H0, planes_hi, planes_lo, Z);
  if (rc != SCNB_OK) return rc;
  if (n == 0) return SCNB_OK;
  return launch_dense_fwd_tc<int64_t, 2, 32, false>(indptr, indices, data, n, G, H0, planes_hi, planes_lo, b1, Z, ksplits,
                                             (cudaStream_t)stream);
}

// scnb_csr_linear_fwd_tc_i64 + scnb_bn_eval_planes in one launch (FwdBnPlanes above): the first layer of batched predict
// (scnym/model.py:211-216 in eval mode on CSR input) ending in the operand planes of the next Linear.  a_hi_out / a_lo_out:
// ceil(n / 256) * 256 x H0 bf16 each.  Never splits the gene axis (the BatchNorm needs the complete sum).
extern "C" int scnb_csr_linear_fwd_tc_bn_planes_i64(const int64_t* indptr, const int32_t* indices, const float* data, int n, int G, int H0,
                                                    const void* planes_hi, const void* planes_lo, const float* b1, const float* gamma,
                                                    const float* beta, const float* running_mean, const float* running_var, int relu,
                                                    void* a_hi_out, void* a_lo_out, void* stream) {
  SCNB_CHECK_ARG(indptr && indices && data && gamma && beta && running_mean && running_var && a_hi_out && a_lo_out,
                 "csr_linear_fwd_tc_bn_planes_i64: null argument");
  SCNB_CHECK_ARG(n >= 0 && G > 0 && planes_hi && planes_lo, "csr_linear_fwd_tc_bn_planes_i64: bad arguments");
  SCNB_CHECK_ARG(H0 >= 32 && H0 <= 256 && H0 % 32 == 0, "csr_linear_fwd_tc_bn_planes_i64: H0 must be a multiple of 32 in [32, 256] (got %d)", H0);
  SCNB_CHECK_ARG((((uintptr_t)planes_hi | (uintptr_t)planes_lo | (uintptr_t)a_hi_out | (uintptr_t)a_lo_out) % 16) == 0,
                 "csr_linear_fwd_tc_bn_planes_i64: buffers must be 16-byte aligned");
  if (n == 0) return SCNB_OK;
  FwdBnPlanes bnp{gamma, beta, running_mean, running_var, (uint16_t*)a_hi_out, (uint16_t*)a_lo_out, relu};
  // (Z is not written: the kernel needs a non-null pointer only for its unfused epilogue)
  return launch_dense_fwd_tc<int64_t, 2, 32, false>(indptr, indices, data, n, G, H0, planes_hi, planes_lo, b1, (float*)a_hi_out, 1,
                                                    (cudaStream_t)stream, nullptr, nullptr, nullptr, bnp);
}

// Y[n, N] = A[n, K] W[N, K]^T + b with both operands given as pre-tiled bf16 hi/lo planes
// (A: scnb_bn_eval_planes, W: scnb_linear_planes): the hidden nn.Linear layers of batched predict
// (scnym/model.py:218; scnym/predict.py:178-192) on tcgen05.
extern "C" int scnb_linear_fwd_tc_planes(const void* a_hi, const void* a_lo, int n, int K, int N, const void* w_hi, const void* w_lo,
                                         const float* bias, float* Y, void* stream) {
  SCNB_CHECK_ARG(a_hi && a_lo && w_hi && w_lo && Y && n >= 0 && K > 0 && K % 32 == 0, "linear_fwd_tc_planes: bad arguments");
  SCNB_CHECK_ARG(N >= 32 && N <= 256 && N % 32 == 0, "linear_fwd_tc_planes: N must be a multiple of 32 in [32, 256] (got %d)", N);
  SCNB_CHECK_ARG((((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)Y) % 16) == 0,
                 "linear_fwd_tc_planes: buffers must be 16-byte aligned");
  if (n == 0) return SCNB_OK;
  return launch_dense_fwd_tc<int32_t, 2, 32, true>(nullptr, nullptr, nullptr, n, K, N, w_hi, w_lo, bias, Y, 1, (cudaStream_t)stream,
                                                   a_hi, a_lo);
}

// scnb_linear_fwd_tc_planes + scnb_bn_eval_planes in one launch: a hidden Linear of batched predict whose output goes through
// the eval-mode BatchNorm -> ReLU straight into the operand planes of the NEXT Linear (FwdBnPlanes; bit-identical planes).
extern "C" int scnb_linear_fwd_tc_planes_bn_planes(const void* a_hi, const void* a_lo, int n, int K, int N, const void* w_hi,
                                                   const void* w_lo, const float* bias, const float* gamma, const float* beta,
                                                   const float* running_mean, const float* running_var, int relu, void* out_hi,
                                                   void* out_lo, void* stream) {
  SCNB_CHECK_ARG(a_hi && a_lo && w_hi && w_lo && out_hi && out_lo && gamma && beta && running_mean && running_var && n >= 0 && K > 0 &&
                     K % 32 == 0, "linear_fwd_tc_planes_bn_planes: bad arguments");
  SCNB_CHECK_ARG(N >= 32 && N <= 256 && N % 32 == 0, "linear_fwd_tc_planes_bn_planes: N must be a multiple of 32 in [32, 256] (got %d)", N);
  SCNB_CHECK_ARG((((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)out_hi | (uintptr_t)out_lo) % 16) == 0,
                 "linear_fwd_tc_planes_bn_planes: buffers must be 16-byte aligned");
  SCNB_CHECK_ARG(out_hi != a_hi && out_lo != a_lo, "linear_fwd_tc_planes_bn_planes: the output planes must not alias the input planes");
  if (n == 0) return SCNB_OK;
  FwdBnPlanes bnp{gamma, beta, running_mean, running_var, (uint16_t*)out_hi, (uint16_t*)out_lo, relu};
  return launch_dense_fwd_tc<int32_t, 2, 32, true>(nullptr, nullptr, nullptr, n, K, N, w_hi, w_lo, bias, (float*)out_hi, 1,
                                                   (cudaStream_t)stream, a_hi, a_lo, nullptr, bnp);
}

// The LAST hidden Linear of batched predict, its eval-mode BatchNorm -> ReLU and the classifier in one launch (FwdHead above):
// planes in, logits (and optionally the embedding) out.  C <= FWD_HEAD_CMAX.
extern "C" int scnb_linear_fwd_tc_planes_bn_head(const void* a_hi, const void* a_lo, int n, int K, int N, const void* w_hi,
                                                 const void* w_lo, const float* bias, const float* gamma, const float* beta,
                                                 const float* running_mean, const float* running_var, int relu, const float* Wc,
                                                 const float* bc, int C, float* logits, int accumulate, float* pre_out, void* stream) {
  SCNB_CHECK_ARG(a_hi && a_lo && w_hi && w_lo && gamma && beta && running_mean && running_var && Wc && logits && n >= 0 && K > 0 &&
                     K % 32 == 0, "linear_fwd_tc_planes_bn_head: bad arguments");
  SCNB_CHECK_ARG(N >= 32 && N <= 256 && N % 32 == 0, "linear_fwd_tc_planes_bn_head: N must be a multiple of 32 in [32, 256] (got %d)", N);
  SCNB_CHECK_ARG(C >= 1 && C <= FWD_HEAD_CMAX, "linear_fwd_tc_planes_bn_head: C must be in [1, %d] (got %d)", FWD_HEAD_CMAX, C);
  SCNB_CHECK_ARG((((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)Wc | (uintptr_t)pre_out) % 16) == 0,
                 "linear_fwd_tc_planes_bn_head: buffers must be 16-byte aligned");
  if (n == 0) return SCNB_OK;
  FwdBnPlanes bnp{gamma, beta, running_mean, running_var, nullptr, nullptr, relu};
  FwdHead head{Wc, bc, logits, pre_out, C, accumulate};
  return launch_dense_fwd_tc<int32_t, 2, 32, true>(nullptr, nullptr, nullptr, n, K, N, w_hi, w_lo, bias, logits, 1, (cudaStream_t)stream,
                                                   a_hi, a_lo, nullptr, bnp, head);
}

// =============================================================================================
// Tensor-core form of the fused first-layer gradient + optimizer kernel (K8+K9):
//   dW1T[g, :] = sum_r X[r, g] * dZ1[r, :]      then   optimizer update of W1T[g, :], state1[g, :], state2[g, :]
// as a tcgen05 GEMM  D[genes, H0] = X^T[genes, rows] * dZ1[rows, H0]  with the optimizer in the epilogue.
//   * a CTA owns GT <= 128 consecutive genes (one M=128 accumulator of H0 columns in TMEM); two CTAs share an SM,
//     so the epilogue of one (pure HBM streaming of p / m / v) overlaps the MMAs of the other;
//   * the A operand (X^T tile: genes x 32 batch rows, bf16 hi/lo) is densified straight from the batch CSR: rows
//     are sorted by gene, so row r's entries of this gene block are the contiguous range
//     [splits[c][r], splits[c+1][r]) computed by scnb_row_gene_splits (no gene-major copy of the batch is needed);
//   * the B operand (dZ1 tile: H0 x 32 rows) comes from bf16 planes of dZ1 (scnb_w1_planes on dZ1) by bulk copy;
//   * acc += A_lo B_hi + A_hi B_lo + A_hi B_hi in fp32 (TMEM); the epilogue reads it back 32 columns at a time and
//     applies Adadelta / Adam / AdamW / SGD to 128-byte segments of the gene's parameter and state rows.
// =============================================================================================
#include <type_traits>

#include "optim.cuh"


#define W1TC_THREADS 320   // warp 0: dZ1 tile loader + L2 prefetch, warp 1: MMA issuer, warps 2-9: densify, then epilogue
#define W1TC_DENSE 256     // densify / epilogue threads
#define W1TC_PRE 8         // batch entries per densify thread held in registers one k-tile ahead
#define W1TC_SINK_BYTES 16384u

// D^T[c, g] = sum_r dZ1[r, c] * X[r, g]: the MMA's M axis is the H0 axis (two 128-lane accumulators for H0 = 256), its
// N axis the CTA's genes (any multiple of 16 up to 256, so a gene block of ceil(G / 148) genes is not padded to 128),
// K the batch rows.  One CTA per SM.
//   A operand: bf16 hi/lo planes of dZ1 (scnb_w1_planes: tile i = batch rows [32 i, 32 i + 32), image [c][r]) by bulk copy;
//   B operand: X^T tile [gene][r], densified from the batch CSR by 256 threads (8 per batch row of the tile);
//   acc += dz_lo x_hi + dz_hi x_lo + dz_hi x_hi  (fp32 in TMEM).
// Epilogue: TMEM lane = column c of W1T, TMEM column = gene.  tcgen05.ld hands a warp 32 consecutive c of 16 genes, so
// for every gene the warp reads and writes 128 contiguous bytes of the parameter / state rows -- coalesced without a
// transposition through shared memory; the 48 loads of a 16-gene chunk are in flight before the first update.
// EXP selects how the epilogue gets the parameter / optimizer-state rows (tools/exp_w1tc.py times the forms side by side):
//   0  register form: every epilogue thread loads the rows of the next 16-gene chunk while it updates the current one;
//   6  as 0, and the rows are READ once during the tensor phase (bulk copies into a 16 KB shared-memory sink, a few per
//      k-tile) so that they sit in L2 when the epilogue asks for them.  (cp.async.bulk.prefetch.L2 and prefetch.global.L2
//      were measured to have no effect here; real reads do.)
//   8  ring form: after the last MMA the freed operand stages become a ring of 16-gene chunks [p rows | m rows | v rows]
//      filled by bulk copies (up to the whole ring, ~190 KB, in flight per SM instead of 48 KB of register loads that
//      queue behind the epilogue's own stores); the epilogue warps read their columns from shared memory;
//   9  8 + the sink reads of 6.
// HC: first-layer width known at compile time (256) or 0 = use the run-time H0.  The epilogue addresses 64 rows of four
// arrays per 16-gene chunk; with a run-time row pitch that is ~390 of its 900 warp instructions per chunk (ncu source
// view, profiles/r01_ncu_full_v12_w1tc_ring.txt), with HC every row is an immediate offset from one pointer per array.
template <int OPT, int EXP = 0, int HC = 0>
__global__ void __launch_bounds__(W1TC_THREADS, 1)
k_w1_bwd_optim_tc(const int32_t* __restrict__ splits, const int32_t* __restrict__ b_idx, const float* __restrict__ b_val, int nb,
                  int GT, int NT, int G, int H0, const uint8_t* __restrict__ dz_hi, const uint8_t* __restrict__ dz_lo,
                  float* __restrict__ W, float* __restrict__ grad_out, float* __restrict__ S1, float* __restrict__ S2,
                  const float* __restrict__ hp, const int64_t* __restrict__ st, uint8_t* __restrict__ pl_hi,
                  uint8_t* __restrict__ pl_lo, int stages, int tmem_cols, long long* __restrict__ dbg,
                  const unsigned long long* __restrict__ push_bufs, long long off_inbox, long long per_f4, int push_rank) {
  constexpr int KT = 32;
  constexpr uint32_t ROWB = 2 * KT;
  if (HC) H0 = HC;                                   // lets the compiler fold the row pitch
#define W1TC_GSTAMP(k) do { if (dbg) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[4 * blockIdx.x + (k)] = (long long)t_; } } while (0)
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ OptimConsts cs;
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = tc_smem_raw + (base - raw);
  const uint32_t A_PLANE = (uint32_t)H0 * ROWB, B_PLANE = (uint32_t)NT * ROWB;
  const uint32_t STAGE_BYTES = 2u * A_PLANE + 2u * B_PLANE;        // [dz_hi][dz_lo][x_hi][x_lo]
  const uint32_t bars = base + (uint32_t)stages * STAGE_BYTES;
  const uint32_t bar_full_dz = bars, bar_full_x = bars + 8u * stages, bar_empty = bars + 16u * stages, bar_acc = bars + 24u * stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + (size_t)stages * STAGE_BYTES + 24 * stages + 8);
  constexpr bool SINK = (EXP == 6 || EXP == 9), RING = (EXP == 8 || EXP == 9);
  // behind everything else: [16 KB sink (SINK only)][sink mbarrier][ring full x 8][ring empty x 8]
  const uint32_t sink = (bars + 24u * (uint32_t)stages + 16u + 127u) & ~127u;
  const uint32_t bar_sink = sink + (SINK ? W1TC_SINK_BYTES : 0u);
  const uint32_t ring_full = bar_sink + 16u, ring_empty = ring_full + 64u;
  const int n_arr = (OPT == OPT_SGD) ? 2 : 3;                       // parameter, state1[, state2]
  const uint32_t arr_stride = 16u * (uint32_t)H0 * 4u;              // 16 gene rows of one array
  const uint32_t slot_bytes = (uint32_t)n_arr * arr_stride;
  const int n_slots = min(8, (int)(((uint32_t)stages * STAGE_BYTES) / slot_bytes));   // >= 2 (stages >= 3)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x;
  const int g0 = c * GT, g1 = min(G, g0 + GT);
  const int n_it = (nb + KT - 1) / KT;
  const int halves = (H0 + 127) >> 7;
  if (tid == 32) W1TC_GSTAMP(3);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full_dz + 8u * s, 1);
      mbar_init(bar_full_x + 8u * s, W1TC_DENSE / 2);
      mbar_init(bar_empty + 8u * s, 1);
    }
    mbar_init(bar_acc, 1);
    if (SINK) mbar_init(bar_sink, 1);
    if (RING) {
      for (int k = 0; k < 8; ++k) {
        mbar_init(ring_full + 8u * k, 1);
        mbar_init(ring_empty + 8u * k, (uint32_t)(H0 >> 5));       // one arrival per epilogue warp that owns columns
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (OPT != OPT_NONE) cs = optim_consts(hp, st);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 1 && OPT != OPT_NONE && EXP == 0) {
      // Ask L2 for this CTA's parameter and optimizer-state rows (contiguous: W1T is gene-major) while HBM is idle.
      // Measured: no effect on B200 (tools/exp_w1tc.py); kept in the register form only.
      const size_t off = (size_t)g0 * H0;
      const uint32_t bytes = (uint32_t)(g1 - g0) * (uint32_t)H0 * 4u;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(W + off), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(S1 + off), "r"(bytes) : "memory");
      if (OPT != OPT_SGD) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(S2 + off), "r"(bytes) : "memory");
    }
    if (lane == 0) {  // ===== dZ1 tile loader =====
      // SINK: this CTA's parameter / state rows, 16 KB at a time behind each dZ1 tile
      const uint32_t row_bytes = (uint32_t)(g1 - g0) * (uint32_t)H0 * 4u;
      const int n_rd = (OPT == OPT_NONE || !SINK) ? 0 : n_arr;
      const int per_arr = (int)((row_bytes + W1TC_SINK_BYTES - 1u) / W1TC_SINK_BYTES);
      const int n_sink = n_rd * per_arr;
      const int per_it = (n_sink + n_it - 1) / n_it;
      int k_sink = 0;
      if (n_sink > 0) mbar_arrive_expect_tx(bar_sink, (uint32_t)n_rd * row_bytes);
      for (int i = 0; i < n_it; ++i) {
        const int s = i % stages;
        const uint32_t ph = (uint32_t)(i / stages) & 1u;
        mbar_wait(bar_empty + 8u * s, ph ^ 1u);
        const uint32_t stage = base + (uint32_t)s * STAGE_BYTES;
        mbar_arrive_expect_tx(bar_full_dz + 8u * s, 2u * A_PLANE);
        const size_t off = (size_t)i * A_PLANE;
        bulk_g2s(stage, dz_hi + off, A_PLANE, bar_full_dz + 8u * s);
        bulk_g2s(stage + A_PLANE, dz_lo + off, A_PLANE, bar_full_dz + 8u * s);
        for (int u = 0; u < per_it && k_sink < n_sink; ++u, ++k_sink) {
          const int a = k_sink / per_arr, j = k_sink % per_arr;
          const float* arr = a == 0 ? W : (a == 1 ? S1 : S2);
          const uint32_t o = (uint32_t)j * W1TC_SINK_BYTES;
          const uint32_t len = min(W1TC_SINK_BYTES, row_bytes - o);
          bulk_g2s(sink, reinterpret_cast<const uint8_t*>(arr + (size_t)g0 * H0) + o, len, bar_sink);
        }
      }
      if (n_sink > 0) mbar_wait(bar_sink, 0);   // no bulk copy may be in flight when the CTA exits
      if (RING && OPT != OPT_NONE) {
        // ===== ring producer: the stages are free once every MMA has read them =====
        mbar_wait(bar_acc, 0);
        const int n_g = g1 - g0;
        for (int gb = 0, k = 0; gb < n_g; gb += 16, ++k) {
          const int slot = k % n_slots;
          const uint32_t use = (uint32_t)(k / n_slots);
          if (use > 0) mbar_wait(ring_empty + 8u * slot, (use - 1u) & 1u);
          const uint32_t bytes = (uint32_t)min(16, n_g - gb) * (uint32_t)H0 * 4u;
          const uint32_t dst = base + (uint32_t)slot * slot_bytes;
          const size_t off = (size_t)(g0 + gb) * H0;
          mbar_arrive_expect_tx(ring_full + 8u * slot, (uint32_t)n_arr * bytes);
          bulk_g2s(dst, W + off, bytes, ring_full + 8u * slot);
          bulk_g2s(dst + arr_stride, S1 + off, bytes, ring_full + 8u * slot);
          if (OPT != OPT_SGD) bulk_g2s(dst + 2u * arr_stride, S2 + off, bytes, ring_full + 8u * slot);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
      for (int i = 0; i < n_it; ++i) {
        const int s = i % stages;
        const uint32_t ph = (uint32_t)(i / stages) & 1u;
        mbar_wait(bar_full_dz + 8u * s, ph);
        mbar_wait(bar_full_x + 8u * s, ph);
        tc_fence_after();
        const uint32_t stage = base + (uint32_t)s * STAGE_BYTES;
        const uint64_t x_hi = umma_desc_kmajor<KT>(stage + 2u * A_PLANE), x_lo = umma_desc_kmajor<KT>(stage + 2u * A_PLANE + B_PLANE);
        for (int h = 0; h < halves; ++h) {
          const uint64_t d_hi = umma_desc_kmajor<KT>(stage + (uint32_t)h * TC_M * ROWB);
          const uint64_t d_lo = umma_desc_kmajor<KT>(stage + A_PLANE + (uint32_t)h * TC_M * ROWB);
          const uint32_t acc = tmem_base + (uint32_t)(h * NT);
#pragma unroll
          for (int k = 0; k < KT / 16; ++k) {
            const uint64_t ko = (uint64_t)(k * 2);
            tc_mma_bf16(acc, d_lo + ko, x_hi + ko, idesc, (i > 0 || k > 0) ? 1u : 0u);
            tc_mma_bf16(acc, d_hi + ko, x_lo + ko, idesc, 1u);
            tc_mma_bf16(acc, d_hi + ko, x_hi + ko, idesc, 1u);
          }
        }
        tc_commit(bar_empty + 8u * s);
      }
      tc_commit(bar_acc);
    }
  } else {
    // ===== densify X^T tiles from the batch CSR.  The refill of a stage is a chain of latencies (entry loads, clear,
    // barrier, scatter, proxy fence, arrive), so the eight warps work as two groups of four on alternate k-tiles: two
    // refills are in flight.  Within a group 4 threads share a batch row of the tile. =====
    const int t = tid - 64;                 // 0..255
    const int grp = t >> 7, tg = t & 127;   // densify group, thread within the group
    const int rl = tg >> 2, sub = tg & 3;   // batch row within the k-tile, entry phase
    const int32_t* sp0 = splits + (size_t)c * nb;
    const int32_t* sp1 = splits + (size_t)(c + 1) * nb;
    // software pipeline over this group's k-tiles: split points are fetched two tiles ahead and the tile's first entries
    // one tile ahead, so no global-memory latency sits between the release of a stage and its refill
    int lo = 0, hi = 0, lo_n = 0, hi_n = 0;
    if (grp < n_it && grp * KT + rl < nb) { lo = sp0[grp * KT + rl]; hi = sp1[grp * KT + rl]; }
    if (grp + 2 < n_it && (grp + 2) * KT + rl < nb) { lo_n = sp0[(grp + 2) * KT + rl]; hi_n = sp1[(grp + 2) * KT + rl]; }
    // W1TC_PRE entries per thread (4 x W1TC_PRE per batch row and gene block) are in registers before the stage is released;
    // rows with more take the dependent loop below.  8 covers 20 % density (27 +- 5 entries per row and block of 136 genes:
    // with 4 every k-tile of config 5 paid three dependent global round trips inside the refill)
    int gj[W1TC_PRE];
    float vj[W1TC_PRE];
#pragma unroll
    for (int j = 0; j < W1TC_PRE; ++j) {
      const int pj = lo + sub + 4 * j;
      gj[j] = pj < hi ? b_idx[pj] : -1;
      vj[j] = pj < hi ? b_val[pj] : 0.f;
    }
    if (t == 0) W1TC_GSTAMP(0);
    const uint32_t n_chunks = 2u * B_PLANE / 16u;
    for (int i = grp; i < n_it; i += 2) {
      const int s = i % stages;
      const uint32_t ph = (uint32_t)(i / stages) & 1u;
      const int rn2 = (i + 4) * KT + rl;
      int lo_n2 = 0, hi_n2 = 0;
      if (i + 4 < n_it && rn2 < nb) { lo_n2 = sp0[rn2]; hi_n2 = sp1[rn2]; }
      int gn[W1TC_PRE];
      float vn[W1TC_PRE];
#pragma unroll
      for (int j = 0; j < W1TC_PRE; ++j) {         // first entries of this group's next tile
        const int pj = lo_n + sub + 4 * j;
        gn[j] = pj < hi_n ? b_idx[pj] : -1;
        vn[j] = pj < hi_n ? b_val[pj] : 0.f;
      }
      mbar_wait(bar_empty + 8u * s, ph ^ 1u);
      const uint32_t x_hi_base = base + (uint32_t)s * STAGE_BYTES + 2u * A_PLANE, x_lo_base = x_hi_base + B_PLANE;
      for (uint32_t ch = (uint32_t)tg; ch < n_chunks; ch += 128u) st_shared_zero16(x_hi_base + ch * 16u);
      // all clears before any scatter: named barrier of this group's four warps
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      const uint32_t kcol = (uint32_t)rl;
#pragma unroll
      for (int j = 0; j < W1TC_PRE; ++j) {
        if (gj[j] >= 0 && vj[j] != 0.f) {
          const uint32_t g = (uint32_t)(gj[j] - g0);
          const __nv_bfloat16 h = __float2bfloat16_rn(vj[j]);
          const __nv_bfloat16 l = __float2bfloat16_rn(vj[j] - __bfloat162float(h));
          const uint32_t off = g * ROWB + swz_chunk<KT>(kcol >> 3, g) + ((kcol & 7u) << 1);
          st_shared_u16(x_hi_base + off, __bfloat16_as_ushort(h));
          st_shared_u16(x_lo_base + off, __bfloat16_as_ushort(l));
        }
      }
      for (int pj = lo + sub + 4 * W1TC_PRE; pj < hi; pj += 4) {   // rows with more than 4 * W1TC_PRE entries in this gene block
        const float v = b_val[pj];
        if (v != 0.f) {
          const uint32_t g = (uint32_t)(b_idx[pj] - g0);
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
          const uint32_t off = g * ROWB + swz_chunk<KT>(kcol >> 3, g) + ((kcol & 7u) << 1);
          st_shared_u16(x_hi_base + off, __bfloat16_as_ushort(h));
          st_shared_u16(x_lo_base + off, __bfloat16_as_ushort(l));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_full_x + 8u * s);
      lo = lo_n; hi = hi_n;
      lo_n = lo_n2; hi_n = hi_n2;
#pragma unroll
      for (int j = 0; j < W1TC_PRE; ++j) { gj[j] = gn[j]; vj[j] = vn[j]; }
    }
    // ===== epilogue: warp (h, q) owns columns c = 128 h + 32 q + lane of W1T for all genes of the CTA =====
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int h = (warp - 2) >> 2;
    const int col = h * 128 + q * 32 + lane;
    if (h * 128 + q * 32 < H0) {
      OptimConsts oc;
      if (OPT != OPT_NONE) oc = cs;
      const int n_g = g1 - g0;
      float p[16], a[16], b[16];
      auto load_chunk = [&](int gb, float(&pp)[16], float(&aa)[16], float(&bb)[16]) {
        if (OPT == OPT_NONE) return;
        const int n_here = n_g - gb;
        const size_t o0 = (size_t)(g0 + gb) * H0 + col;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j < n_here) {
            pp[j] = W[o0 + (size_t)j * H0];
            aa[j] = S1[o0 + (size_t)j * H0];
            bb[j] = (OPT != OPT_SGD) ? S2[o0 + (size_t)j * H0] : 0.f;
          }
        }
      };
      // gradient of 16 genes from TMEM (asynchronous; completed by tmem_wait)
      auto tmem_load16 = [&](int gb, uint32_t(&v)[16]) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * NT + gb);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr)
            : "memory");
      };
      // optimizer update of the chunk's rows (this thread's column), stores, operand planes of the updated rows.
      // `full` (a std::integral_constant): all 16 rows exist -- no per-row predicates.
      auto finish_chunk = [&](int gb, const uint32_t(&v)[16], float(&pp)[16], float(&aa)[16], float(&bb)[16], auto full) {
        constexpr bool FULL = decltype(full)::value;
        const int n_here = FULL ? 16 : n_g - gb;
        const size_t o0 = (size_t)(g0 + gb) * H0 + col;
        float* const pw = W + o0;
        float* const p1 = S1 + o0;
        float* const p2 = S2 + o0;
        float* const pg = grad_out ? grad_out + o0 : nullptr;
        // data parallel, push form: gradient row g goes straight into the inbox of the rank that owns row g of the arena
        // (region `push_rank` of that rank's inbox; rows never straddle owners: per_f4 is a multiple of a row) -- posted
        // NVLink stores that overlap this kernel instead of round-trip loads in the exchange kernel behind it
        int own = 0;
        long long own_next = 0, rows_per = 0;
        float* pushp = nullptr;
        if (OPT == OPT_NONE && push_bufs) {
          rows_per = per_f4 / (H0 >> 2);
          own = (int)((long long)(g0 + gb) / rows_per);
          own_next = (long long)(own + 1) * rows_per;
          pushp = reinterpret_cast<float*>(push_bufs[own] + off_inbox) + ((long long)push_rank - own) * per_f4 * 4 + o0;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (FULL || j < n_here) {
            const float gr = __uint_as_float(v[j]);
            if (OPT == OPT_NONE && push_bufs) {
              while ((long long)(g0 + gb + j) >= own_next) {
                ++own;
                own_next += rows_per;
                pushp = reinterpret_cast<float*>(push_bufs[own] + off_inbox) + ((long long)push_rank - own) * per_f4 * 4 + o0;
              }
              pushp[j * H0] = gr;
            }
            if (pg) pg[j * H0] = gr;
            if (OPT != OPT_NONE) {
              optim_update<OPT>(pp[j], gr, aa[j], bb[j], oc);
              pw[j * H0] = pp[j];
              p1[j * H0] = aa[j];
              if (OPT != OPT_SGD) p2[j * H0] = bb[j];
            }
          }
        }
        if (OPT != OPT_NONE && pl_hi) {
          // bf16 hi/lo operand planes of the UPDATED rows for the next step's tensor-core forward (same image as
          // k_w1_planes<32>: tile = 32 genes, row = column of W1T, 16-byte chunk = 8 consecutive genes, XOR-swizzled)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (FULL || 8 * u < n_here) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x0 = (FULL || (8 * u + 2 * j) < n_here) ? pp[8 * u + 2 * j] : 0.f;
                const float x1 = (FULL || (8 * u + 2 * j + 1) < n_here) ? pp[8 * u + 2 * j + 1] : 0.f;
                split_bf16x2(x0, x1, hw[j], lw[j]);
              }
              const uint32_t gg = (uint32_t)(g0 + gb + 8 * u);
              const size_t off = ((size_t)(gg >> 5) * H0 * KT) * 2 + (size_t)col * ROWB + swz_chunk<KT>((gg & 31u) >> 3, (uint32_t)col);
              *reinterpret_cast<uint4*>(pl_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(pl_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        }
      };
      auto finish = [&](int gb, const uint32_t(&v)[16], float(&pp)[16], float(&aa)[16], float(&bb)[16]) {
        if (n_g - gb >= 16) finish_chunk(gb, v, pp, aa, bb, std::true_type{});
        else finish_chunk(gb, v, pp, aa, bb, std::false_type{});
      };
      if (RING) {
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        if (t == 0) W1TC_GSTAMP(1);
        for (int gb = 0, k = 0; gb < n_g; gb += 16, ++k) {
          uint32_t v[16];
          tmem_load16(gb, v);
          if (OPT != OPT_NONE) {
            const int slot = k % n_slots;
            const uint32_t use = (uint32_t)(k / n_slots);
            mbar_wait(ring_full + 8u * slot, use & 1u);
            const float* sp = reinterpret_cast<const float*>(sm + (size_t)slot * slot_bytes) + col;
            const int n_here = n_g - gb;
            if (n_here >= 16) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                p[j] = sp[j * H0];
                a[j] = sp[16 * H0 + j * H0];
                b[j] = (OPT != OPT_SGD) ? sp[32 * H0 + j * H0] : 0.f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (j < n_here) {
                  p[j] = sp[j * H0];
                  a[j] = sp[16 * H0 + j * H0];
                  b[j] = (OPT != OPT_SGD) ? sp[32 * H0 + j * H0] : 0.f;
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(ring_empty + 8u * slot);      // this warp has its columns of the slot in registers
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          finish(gb, v, p, a, b);
        }
      } else {
        load_chunk(0, p, a, b);          // in flight while the last MMAs drain
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        if (t == 0) W1TC_GSTAMP(1);
        for (int gb = 0; gb < n_g; gb += 16) {
          uint32_t v[16];
          tmem_load16(gb, v);
          float pn[16], an[16], bn[16];
          if (gb + 16 < n_g) load_chunk(gb + 16, pn, an, bn);   // next chunk's rows, in flight during this chunk's update
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          finish(gb, v, p, a, b);
#pragma unroll
          for (int j = 0; j < 16; ++j) { p[j] = pn[j]; a[j] = an[j]; b[j] = bn[j]; }
        }
      }
    } else {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
    }
    if (t == 0) W1TC_GSTAMP(2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
#undef W1TC_GSTAMP
}

// MixUp adjoint (scnb_unmix_rows: dZ1[r] = sum_k coef[k][r] * dOut[inv[k][r]]) that also emits dZ1 as the bf16 hi/lo
// operand planes of the tensor-core dW1 kernel (tile = 32 batch rows, plane row = column of dZ1, 16-byte chunk = 8
// consecutive batch rows): a CTA owns 8 batch rows -- one chunk of every plane row -- and thread t column t, so the
// separate scnb_w1_planes pass over dZ1 (and its launch on the critical chain) disappears.
__global__ void __launch_bounds__(256) k_unmix_rows_planes(const float* __restrict__ dOut, int H, const int32_t* __restrict__ inv,
                                                           const float* __restrict__ coef, int n_base, float* __restrict__ dZ,
                                                           uint8_t* __restrict__ pl_hi, uint8_t* __restrict__ pl_lo) {
  __shared__ int s_inv[8][3];
  __shared__ float s_coef[8][3];
  const int r0 = blockIdx.x * 8;
  if (threadIdx.x < 24) {
    const int j = threadIdx.x / 3, k = threadIdx.x % 3, r = r0 + j;
    s_inv[j][k] = r < n_base ? inv[k * n_base + r] : -1;
    s_coef[j][k] = r < n_base ? coef[k * n_base + r] : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int i = s_inv[j][k];
        if (i >= 0) x = fmaf(s_coef[j][k], dOut[(size_t)i * H + c], x);
      }
      v[j] = x;
      if (r0 + j < n_base) dZ[(size_t)(r0 + j) * H + c] = x;
    }
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      split_bf16x2(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
    }
    const size_t off = ((size_t)(r0 >> 5) * H * 32) * 2 + (size_t)c * 64 + swz_chunk<32>((uint32_t)(r0 & 31) >> 3, (uint32_t)c);
    *reinterpret_cast<uint4*>(pl_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(pl_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

extern "C" int scnb_unmix_rows_planes(const float* dOut, int H, const int32_t* inv3, const float* coef3, int n_base, float* dZ,
                                      void* planes_hi, void* planes_lo, void* stream) {
  SCNB_CHECK_ARG(dOut && inv3 && coef3 && dZ && planes_hi && planes_lo && H > 0 && n_base >= 0, "unmix_rows_planes: bad arguments");
  SCNB_CHECK_ARG(H % 32 == 0 && (((uintptr_t)planes_hi | (uintptr_t)planes_lo) % 16) == 0, "unmix_rows_planes: H %% 32 and 16-byte aligned planes");
  if (n_base == 0) return SCNB_OK;
  k_unmix_rows_planes<<<scnb_cdiv(n_base, 8), H < 256 ? H : 256, 0, (cudaStream_t)stream>>>(dOut, H, inv3, coef3, n_base, dZ, (uint8_t*)planes_hi,
                                                                                         (uint8_t*)planes_lo);
  SCNB_CHECK_LAUNCH("unmix_rows_planes");
  return SCNB_OK;
}

// gene block of a CTA of the tensor-core dW1 kernel: one CTA per SM when the block fits one MMA (N <= 256)
static inline int w1tc_gene_block(int G) {
  int gt = ((G + 147) / 148 + 7) / 8 * 8;   // multiple of 8: a 16-byte chunk of the operand planes never straddles two CTAs
  if (gt > 256) gt = 256;
  if (gt < 16) gt = 16;
  return gt;
}

extern "C" int scnb_w1tc_gene_block(int G) { return w1tc_gene_block(G); }

// both tables in one launch (the prologue branch of the step): dW1 gene blocks and the forward's gene splits
extern "C" int scnb_row_gene_splits2(const int32_t* b_indptr, const int32_t* b_idx, int nb, int G, int H0, int32_t* splits,
                                     int32_t* fwd_splits, void* stream) {
  SCNB_CHECK_ARG(b_indptr && b_idx && splits && fwd_splits && nb >= 0 && G > 0, "row_gene_splits2: bad arguments");
  if (nb == 0) return SCNB_OK;
  const int GT = w1tc_gene_block(G);
  const int n_blk = scnb_cdiv(G, GT);
  int ks = 0, gp = 0;
  scnb_csr_linear_fwd_tc_geometry(nb, G, H0, &ks, &gp);
  k_row_gene_splits<<<nb, 128, 0, (cudaStream_t)stream>>>(b_indptr, b_idx, nb, GT, n_blk, splits, gp, ks, fwd_splits);
  SCNB_CHECK_LAUNCH("row_gene_splits2");
  return SCNB_OK;
}

extern "C" int scnb_row_gene_splits(const int32_t* b_indptr, const int32_t* b_idx, int nb, int G, int32_t* splits, void* stream) {
  SCNB_CHECK_ARG(b_indptr && b_idx && splits && nb >= 0 && G > 0, "row_gene_splits: bad arguments");
  if (nb == 0) return SCNB_OK;
  const int GT = w1tc_gene_block(G);
  const int n_blk = scnb_cdiv(G, GT);
  k_row_gene_splits<<<nb, 128, 0, (cudaStream_t)stream>>>(b_indptr, b_idx, nb, GT, n_blk, splits);
  SCNB_CHECK_LAUNCH("row_gene_splits");
  return SCNB_OK;
}

template <int OPT, int EXP = 0, int HC = 0>
static int launch_w1_bwd_optim_tc_any(const int32_t* splits, const int32_t* b_idx, const float* b_val, int nb, int H0, int G,
                                      const void* dz_hi, const void* dz_lo, float* W1T, float* grad_out, float* s1, float* s2,
                                      const float* hp, const int64_t* st4, void* pl_hi, void* pl_lo, cudaStream_t st,
                                      const unsigned long long* push_bufs = nullptr, long long off_inbox = 0, long long per_f4 = 0,
                                      int push_rank = 0) {
  const int GT = w1tc_gene_block(G);
  const int NT = (GT + 15) / 16 * 16;
  const int n_blk = scnb_cdiv(G, GT);
  const int halves = (H0 + 127) / 128;
  const size_t stage_bytes = 2 * (size_t)H0 * 64 + 2 * (size_t)NT * 64;
  // the last half-tile of the dZ1 planes may be read past H0 rows by the M = 128 MMA (the lanes it feeds are never
  // read back): keep that read inside the dynamic window
  const size_t slack = (size_t)(halves * 128 - H0) * 64;
  int stages = (int)((200 * 1024 - slack) / stage_bytes);
  if (stages > 6) stages = 6;
  if (stages < 2) stages = 2;
  const size_t smem = 1024 + stages * stage_bytes + slack + 24 * stages + 16 + 16 + ((EXP == 6 || EXP == 9) ? W1TC_SINK_BYTES : 0) + 128 + 16 + 128;
  int tmem_cols = 32;
  while (tmem_cols < halves * NT) tmem_cols *= 2;
  cudaError_t e = cudaFuncSetAttribute(k_w1_bwd_optim_tc<OPT, EXP, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    scnb_set_error("csr_w1_bwd_optim_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  static long long* dbg = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* ev = getenv("SCNB_W1TC_DBG");
    dbg_on = (ev && ev[0] == '1') ? 1 : 0;
    if (dbg_on) cudaMalloc(&dbg, 4 * 4096 * sizeof(long long));
  }
  k_w1_bwd_optim_tc<OPT, EXP, HC><<<n_blk, W1TC_THREADS, smem, st>>>(splits, b_idx, b_val, nb, GT, NT, G, H0, (const uint8_t*)dz_hi,
                                                          (const uint8_t*)dz_lo, W1T, grad_out, s1, s2, hp, st4, (uint8_t*)pl_hi, (uint8_t*)pl_lo,
                                                          stages, tmem_cols, dbg_on ? dbg : nullptr, push_bufs, off_inbox, per_f4, push_rank);
  SCNB_CHECK_LAUNCH("csr_w1_bwd_optim_tc");
  if (dbg_on && n_blk <= 4096) {  // debugging aid only (synchronises): per-CTA wall-clock stamps (ns)
    static long long hc[4 * 4096];
    cudaStreamSynchronize(st);
    cudaMemcpy(hc, dbg, sizeof(long long) * 4 * n_blk, cudaMemcpyDeviceToHost);
    long long t0 = hc[3];
    for (int b = 0; b < n_blk; ++b) if (hc[4 * b + 3] < t0) t0 = hc[4 * b + 3];
    const char* nm[4] = {"start", "accumulator ready", "end", "kernel entry"};
    for (int k = 0; k < 4; ++k) {
      long long mn = 1LL << 62, mx = 0, sum = 0;
      for (int b = 0; b < n_blk; ++b) { long long v = hc[4 * b + k] - t0; if (v < mn) mn = v; if (v > mx) mx = v; sum += v; }
      fprintf(stderr, "[w1tc] %s: min %lld mean %lld max %lld ns (%d CTAs, %d stages, GT %d NT %d)\n", nm[k], mn, sum / n_blk, mx, n_blk,
              stages, GT, NT);
    }
  }
  return SCNB_OK;
}

extern "C" int scnb_csr_w1_bwd_optim_tc(int opt, const int32_t* splits, const int32_t* b_idx, const float* b_val, int nb, int H0, int G,
                                        const void* dz_planes_hi, const void* dz_planes_lo, float* W1T, float* grad_out,
                                        float* state1, float* state2, const float* hp8, const int64_t* state4, void* planes_hi_out,
                                        void* planes_lo_out, void* stream) {
  SCNB_CHECK_ARG(splits && b_idx && b_val && dz_planes_hi && dz_planes_lo && nb > 0 && G > 0, "csr_w1_bwd_optim_tc: bad arguments");
  SCNB_CHECK_ARG(H0 >= 32 && H0 <= 256 && H0 % 32 == 0, "csr_w1_bwd_optim_tc: H0 must be a multiple of 32 in [32, 256] (got %d)", H0);
  SCNB_CHECK_ARG(opt == OPT_NONE ? (grad_out != nullptr) : (W1T && state1 && hp8 && state4 && (opt == OPT_SGD || state2)),
                 "csr_w1_bwd_optim_tc: missing buffers for the requested mode");
  SCNB_CHECK_ARG((((uintptr_t)W1T | (uintptr_t)grad_out | (uintptr_t)state1 | (uintptr_t)state2 | (uintptr_t)dz_planes_hi |
                   (uintptr_t)dz_planes_lo) % 16) == 0, "csr_w1_bwd_optim_tc: buffers must be 16-byte aligned");
  SCNB_CHECK_ARG((planes_hi_out == nullptr) == (planes_lo_out == nullptr) && ((uintptr_t)planes_hi_out | (uintptr_t)planes_lo_out) % 16 == 0,
                 "csr_w1_bwd_optim_tc: operand planes must be given together, 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
#define W1TC_GO(O, E, HCV) return launch_w1_bwd_optim_tc_any<O, E, HCV>(splits, b_idx, b_val, nb, H0, G, dz_planes_hi, dz_planes_lo, W1T, grad_out, state1, state2, hp8, state4, planes_hi_out, planes_lo_out, st)
  // Form of the epilogue (see the kernel): 9 = ring + sink reads is the production form, specialised for the default width
  // H0 = 256.  SCNB_W1TC_EXP=0 selects the register form, 6 / 8 (Adam only) the two halves of 9 on their own, 10 the
  // production form without the H0 specialisation.  Read per launch so that tools/exp_w1tc.py can time every form in one
  // process.
  const char* ev = getenv("SCNB_W1TC_EXP");
  const int form = ev ? atoi(ev) : 9;
  if (opt == OPT_ADAM && form == 6) W1TC_GO(OPT_ADAM, 6, 0);
  if (opt == OPT_ADAM && form == 8) W1TC_GO(OPT_ADAM, 8, 0);
  if (form == 0) {
    switch (opt) {
      case OPT_ADADELTA: W1TC_GO(OPT_ADADELTA, 0, 0);
      case OPT_ADAM: W1TC_GO(OPT_ADAM, 0, 0);
      case OPT_ADAMW: W1TC_GO(OPT_ADAMW, 0, 0);
      case OPT_SGD: W1TC_GO(OPT_SGD, 0, 0);
      default: break;
    }
  }
  if (H0 == 256 && form != 10) {
    switch (opt) {
      case OPT_NONE: W1TC_GO(OPT_NONE, 0, 256);
      case OPT_ADADELTA: W1TC_GO(OPT_ADADELTA, 9, 256);
      case OPT_ADAM: W1TC_GO(OPT_ADAM, 9, 256);
      case OPT_ADAMW: W1TC_GO(OPT_ADAMW, 9, 256);
      case OPT_SGD: W1TC_GO(OPT_SGD, 9, 256);
      default: break;
    }
  }
  switch (opt) {
    case OPT_NONE: W1TC_GO(OPT_NONE, 0, 0);
    case OPT_ADADELTA: W1TC_GO(OPT_ADADELTA, 9, 0);
    case OPT_ADAM: W1TC_GO(OPT_ADAM, 9, 0);
    case OPT_ADAMW: W1TC_GO(OPT_ADAMW, 9, 0);
    case OPT_SGD: W1TC_GO(OPT_SGD, 9, 0);
    default: scnb_set_error("csr_w1_bwd_optim_tc: unknown optimizer %d", opt); return SCNB_ERR_ARG;
  }
#undef W1TC_GO
}

// Data-parallel form of the tensor-core dW1 kernel: gradient only, and every gradient row is STORED straight into the
// inbox of the rank that owns that row of the flat arena (scnb_dp_slice_floats; region `rank` of the owner's inbox at
// off_inbox of its exchange buffer) while the kernel runs -- scnb_dp_reduce_optim_bcast2 then finds all ranks' gradients of
// its slice in local memory.  Needs 256 %% H0 == 0 (a row of W1T never straddles two owners).  grad_out (may be NULL): also
// keep the local gradient [G, H0] (parity tests).
extern "C" int scnb_csr_w1_bwd_push_tc(const int32_t* splits, const int32_t* b_idx, const float* b_val, int nb, int H0, int G,
                                       const void* dz_planes_hi, const void* dz_planes_lo, const uint64_t* peer_bufs, int rank,
                                       int world, long long off_inbox, long long total, float* grad_out, void* stream) {
  SCNB_CHECK_ARG(splits && b_idx && b_val && dz_planes_hi && dz_planes_lo && peer_bufs, "csr_w1_bwd_push_tc: null argument");
  SCNB_CHECK_ARG(nb > 0 && G > 0 && H0 >= 32 && H0 <= 256 && 256 % H0 == 0, "csr_w1_bwd_push_tc: H0 must divide 256 (got %d)", H0);
  SCNB_CHECK_ARG(world >= 1 && world <= 8 && rank >= 0 && rank < world && off_inbox >= 0 && off_inbox % 16 == 0 &&
                     total >= (long long)G * H0 && total % 4 == 0, "csr_w1_bwd_push_tc: bad exchange arguments");
  SCNB_CHECK_ARG((((uintptr_t)dz_planes_hi | (uintptr_t)dz_planes_lo) % 16) == 0, "csr_w1_bwd_push_tc: buffers must be 16-byte aligned");
  const long long per_f4 = dp_slice_f4(total, world);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long* pb = (const unsigned long long*)peer_bufs;
  if (H0 == 256)
    return launch_w1_bwd_optim_tc_any<OPT_NONE, 0, 256>(splits, b_idx, b_val, nb, H0, G, dz_planes_hi, dz_planes_lo, nullptr, grad_out,
                                                        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st, pb, off_inbox, per_f4, rank);
  return launch_w1_bwd_optim_tc_any<OPT_NONE, 0, 0>(splits, b_idx, b_val, nb, H0, G, dz_planes_hi, dz_planes_lo, nullptr, grad_out, nullptr,
                                                    nullptr, nullptr, nullptr, nullptr, nullptr, st, pb, off_inbox, per_f4, rank);
}
