// Tile-GEMM core, BatchNorm statistics combination and staging helpers of the fused hidden stack, shared by hidden_ops.cu
// and head_ops.cu.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

#define BN_EPS 1e-5f
#define HS_BM 32
#define HS_BN 64
#define HS_T 512          // threads of the tile-GEMM kernels: 16 warps = 4 column slices x 4 K quarters
#define HS_TE 256         // threads of the elementwise tile kernels (thread -> row tid/8, 8 columns)
#define HS_KQ 4           // K splits inside a CTA
#define HS_MAXSEG 3

struct HsSeg {
  int n_seg;
  int off[HS_MAXSEG + 1];   // row offsets, multiples of HS_BM
};

__device__ __forceinline__ int hs_seg_of(const HsSeg& sg, int row) {
  int s = 0;
#pragma unroll
  for (int i = 1; i < HS_MAXSEG; ++i)
    if (i < sg.n_seg && row >= sg.off[i]) s = i;
  return s;
}

// ---- small PTX helpers (same forms as gemm_ops.cu) ----
__device__ __forceinline__ void hs_cp16(float* smem_dst, const float* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void hs_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void hs_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
// 3xTF32 operand split.  The tensor core reads a tf32 operand from a 32-bit register and ignores the low 13 mantissa bits,
// so `hi` is x itself (used as trunc(x)) and `lo = x - trunc(x)` is exact in fp32 (then truncated the same way): two
// full-rate instructions per element.  (`cvt.rna.tf32.f32` is emulated on sm_100a -- add / compare / select / mask, and
// again for the residual: nine instructions per element, which made the MMA loop issue-bound.)  |lo| < 2^-10 |x|, so the
// dropped terms (lo*lo and the truncation of lo) are below 2^-20 relative per product.
__device__ __forceinline__ void hs_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x);
  lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
}
__device__ __forceinline__ void hs_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Stage `rows` rows of `cpr` 16-byte chunks each (row pitches: `gld` floats in global memory, `sld` in shared memory).
// When the chunk count per row divides the thread count every thread keeps its chunk column and walks rows with plain
// additions (the generic index arithmetic was a third of the tile kernels' instructions).
__device__ __forceinline__ void hs_stage_rows(float* __restrict__ sdst, int sld, const float* __restrict__ gsrc, size_t gld, int rows,
                                              int cpr) {
  const int tid = threadIdx.x;
  if (HS_T % cpr == 0) {
    const int rpp = HS_T / cpr, r0 = tid / cpr, q = tid - r0 * cpr;
    float* d = sdst + r0 * sld + q * 4;
    const float* g = gsrc + (size_t)r0 * gld + q * 4;
    for (int r = r0; r < rows; r += rpp, d += rpp * sld, g += (size_t)rpp * gld) hs_cp16(d, g);
  } else {
    for (int ch = tid; ch < rows * cpr; ch += HS_T) {
      const int r = ch / cpr, q = ch - r * cpr;
      hs_cp16(sdst + r * sld + q * 4, gsrc + (size_t)r * gld + q * 4);
    }
  }
}

// Per-segment batch statistics of column c from the tile partials part[tile][2][H] (tiles t0 .. t1-1 of 32 rows each),
// combined in tile order (Chan et al.; equal tile sizes).  Up to 16 tiles per segment (the usual case: <= 512 rows) all
// partials are fetched in ONE round trip (independent loads); longer segments go in chunks of 8.
#define HS_CT 16
__device__ __forceinline__ void hs_combine_stats(const float* __restrict__ part, int H, int c, int t0, int t1, float& mean, float& var) {
  const int n = t1 - t0;
  float m = 0.f, M2 = 0.f;
  if (n <= HS_CT) {
    float mt[HS_CT], qt[HS_CT];
#pragma unroll
    for (int i = 0; i < HS_CT; ++i) {
      const bool ok = i < n;
      mt[i] = ok ? part[((size_t)(t0 + i) * 2 + 0) * H + c] : 0.f;
      qt[i] = ok ? part[((size_t)(t0 + i) * 2 + 1) * H + c] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < HS_CT; ++i) m += mt[i];
    m /= (float)n;
#pragma unroll
    for (int i = 0; i < HS_CT; ++i)
      if (i < n) {
        const float d = mt[i] - m;
        M2 += qt[i] + (float)HS_BM * d * d;
      }
  } else {
    for (int t = t0; t < t1; t += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (t + i < t1) ? part[((size_t)(t + i) * 2 + 0) * H + c] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) m += v[i];
    }
    m /= (float)n;
    for (int t = t0; t < t1; t += 8) {
      float v[8], q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool ok = t + i < t1;
        v[i] = ok ? part[((size_t)(t + i) * 2 + 0) * H + c] : m;
        q[i] = ok ? part[((size_t)(t + i) * 2 + 1) * H + c] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[i] - m;
        M2 += q[i] + (float)HS_BM * d * d;
      }
    }
  }
  mean = m;
  var = M2 / (float)(n * HS_BM);
}
// plain sums of the backward partials (sum dy, sum dy*zhat) over a segment's tiles, in tile order
__device__ __forceinline__ void hs_combine_sums(const float* __restrict__ part, int H, int c, int t0, int t1, float& s1, float& s2) {
  float a = 0.f, b = 0.f;
  for (int t = t0; t < t1; t += HS_CT) {
    float u[HS_CT], w[HS_CT];
#pragma unroll
    for (int i = 0; i < HS_CT; ++i) {
      const bool ok = t + i < t1;
      u[i] = ok ? part[((size_t)(t + i) * 2 + 0) * H + c] : 0.f;
      w[i] = ok ? part[((size_t)(t + i) * 2 + 1) * H + c] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < HS_CT; ++i) {
      a += u[i];
      b += w[i];
    }
  }
  s1 = a;
  s2 = b;
}

// Column statistics of a 32 x 64 tile held as v[8] per thread (thread -> row = tid/8, columns (tid%8)*8 .. +7):
// mean and M2 over the 32 rows, written to part[tile][0/1][c0 + ..].  red: shared [32][64 + 1].
__device__ __forceinline__ void hs_tile_stats_rowmajor(const float (&v)[8], float (*red)[HS_BN + 1], float* __restrict__ part, int H,
                                                       int tile, int c0) {
  const int tid = threadIdx.x, r = tid >> 3, cg = (tid & 7) << 3;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[r][cg + j] = v[j];
  __syncthreads();
  float mean = 0.f;
  if (tid < HS_BN) {
#pragma unroll 8
    for (int i = 0; i < HS_BM; ++i) mean += red[i][tid];
    mean *= (1.0f / HS_BM);
    float m2 = 0.f;
#pragma unroll 8
    for (int i = 0; i < HS_BM; ++i) {
      const float d = red[i][tid] - mean;
      m2 = fmaf(d, d, m2);
    }
    part[((size_t)tile * 2 + 0) * H + c0 + tid] = mean;
    part[((size_t)tile * 2 + 1) * H + c0 + tid] = m2;
  }
  __syncthreads();
}
// same for two plain sums (backward): a[8], b[8] per thread
__device__ __forceinline__ void hs_tile_sums_rowmajor(const float (&a)[8], const float (&b)[8], float (*red)[HS_BN + 1],
                                                      float* __restrict__ part, int H, int tile, int c0) {
  const int tid = threadIdx.x, r = tid >> 3, cg = (tid & 7) << 3;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[r][cg + j] = a[j];
  __syncthreads();
  if (tid < HS_BN) {
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < HS_BM; ++i) s += red[i][tid];
    part[((size_t)tile * 2 + 0) * H + c0 + tid] = s;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) red[r][cg + j] = b[j];
  __syncthreads();
  if (tid < HS_BN) {
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < HS_BM; ++i) s += red[i][tid];
    part[((size_t)tile * 2 + 1) * H + c0 + tid] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ void hs_emit_keep(uint8_t* __restrict__ keep_out, int H, int row0, int c0, const uint64_t* __restrict__ rng,
                                             uint32_t stream_id, float p_drop);
__device__ __forceinline__ void hs_ld8(const float* __restrict__ p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void hs_st8(float* __restrict__ p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}


// ---------------------------------------------------------------------------------------------------------------------
// Tile GEMM core: 32 x 64 output tile, K (multiple of 32, <= 512) staged whole in shared memory.
//   sA : [32][K + 4]  (row-major, k contiguous)
//   sB : LAY 0 (NT: weight stored [n][k])  [64][K + 4];  LAY 1 (NN: weight stored [k][n])  [K][64 + 8]
// 16 warps = 4 (16-column slices) x 4 (K quarters); a warp runs m16n8k8 3xTF32 over a 32 x 16 tile of its K quarter.  The
// four partial tiles then meet in shared memory as plain [4][32][72] row-major tiles (sT, which may alias the operand
// panels: callers synchronise before and after), so that the epilogue runs on ALL 512 threads with float4 rows.
// Fragment reads are bank-conflict free: row strides are 4 (k-contiguous) or 8 (k-row panels) modulo 32 words.
// ---------------------------------------------------------------------------------------------------------------------
#define HS_TP (HS_BN + 8)                       // row pitch of a partial tile
#define HS_TILE_FLOATS (HS_KQ * HS_BM * HS_TP)  // floats of the four partial tiles

template <int LAY>
__device__ __forceinline__ void hs_tile_mma(const float* __restrict__ sA, const float* __restrict__ sB, int K, float (&acc)[2][2][4]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wn = warp & 3, kq = warp >> 2;
  const int SA = K + 4, SBr = K + 4, SBc = HS_BN + 8;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
  const int kb = kq * (K / HS_KQ), ke = kb + K / HS_KQ;
#pragma unroll 2
  for (int kk = kb; kk < ke; kk += 8) {
    uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float x0 = sA[(16 * i + g) * SA + kk + t];
      const float x1 = sA[(16 * i + g + 8) * SA + kk + t];
      const float x2 = sA[(16 * i + g) * SA + kk + t + 4];
      const float x3 = sA[(16 * i + g + 8) * SA + kk + t + 4];
      hs_split(x0, ah[i][0], al[i][0]);
      hs_split(x1, ah[i][1], al[i][1]);
      hs_split(x2, ah[i][2], al[i][2]);
      hs_split(x3, ah[i][3], al[i][3]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = wn * 16 + 8 * j + g;
      float y0, y1;
      if (LAY == 0) {
        y0 = sB[n * SBr + kk + t];
        y1 = sB[n * SBr + kk + t + 4];
      } else {
        y0 = sB[(kk + t) * SBc + n];
        y1 = sB[(kk + t + 4) * SBc + n];
      }
      hs_split(y0, bh[j][0], bl[j][0]);
      hs_split(y1, bh[j][1], bl[j][1]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        hs_mma(acc[i][j], al[i], bh[j]);
        hs_mma(acc[i][j], ah[i], bl[j]);
        hs_mma(acc[i][j], ah[i], bh[j]);
      }
  }
}
// this warp's partial accumulators -> sT[kq][row][col] (lane (g, t): rows 16i + g (+8), columns wn*16 + 8j + 2t, +1)
__device__ __forceinline__ void hs_tile_store_partials(float* __restrict__ sT, const float (&acc)[2][2][4]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3, wn = warp & 3, kq = warp >> 2;
  float* base = sT + (size_t)kq * HS_BM * HS_TP;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = wn * 16 + 8 * j + 2 * t;
      *reinterpret_cast<float2*>(base + (16 * i + g) * HS_TP + c) = make_float2(acc[i][j][0], acc[i][j][1]);
      *reinterpret_cast<float2*>(base + (16 * i + g + 8) * HS_TP + c) = make_float2(acc[i][j][2], acc[i][j][3]);
    }
}
// element group of thread `tid` in the epilogue: row tid/16, columns 4 (tid%16) .. +3; sum of the four partial tiles
__device__ __forceinline__ float4 hs_tile_sum4(const float* __restrict__ sT, int row, int c4) {
  float4 v = *reinterpret_cast<const float4*>(sT + row * HS_TP + c4);
#pragma unroll
  for (int q = 1; q < HS_KQ; ++q) {
    const float4 u = *reinterpret_cast<const float4*>(sT + (size_t)q * HS_BM * HS_TP + row * HS_TP + c4);
    v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
  }
  return v;
}
// one bulk copy per row: `rows` rows of `bytes` bytes, global pitch gld floats -> shared pitch sld floats (thread r issues row r)
__device__ __forceinline__ void hs_bulk_rows(float* sdst, int sld, const float* gsrc, size_t gld, int rows, uint32_t bytes, uint32_t bar,
                                             int tid0) {
  for (int r = (int)threadIdx.x - tid0; r >= 0 && r < rows; r += HS_T) bulk_g2s(smem_u32(sdst + (size_t)r * sld), gsrc + (size_t)r * gld, bytes, bar);
}
// 0/1 keep bytes of the four columns c .. c+3 of row rg (mask: [R, H] bytes) as a 4-bit pattern.  Split form for
// prefetching: hs_keep4_raw only loads (the word can sit in a register while other round trips are in flight -- any use of
// it stalls the thread until it has arrived), hs_keep4_bits converts.
__device__ __forceinline__ uint32_t hs_keep4_raw(const uint8_t* __restrict__ mask, size_t elem) {
  return *reinterpret_cast<const uint32_t*>(mask + elem);
}
__device__ __forceinline__ uint32_t hs_keep4_bits(uint32_t m) {
  return ((m & 0xffu) ? 1u : 0u) | ((m & 0xff00u) ? 2u : 0u) | ((m & 0xff0000u) ? 4u : 0u) | ((m & 0xff000000u) ? 8u : 0u);
}
__device__ __forceinline__ uint32_t hs_keep4(const uint8_t* __restrict__ mask, size_t elem) { return hs_keep4_bits(hs_keep4_raw(mask, elem)); }
// Dropout keep bytes of a 32 x 64 tile of application `a`'s BatchNorm output (rows row0.., columns c0..), written by the
// kernel that PRODUCES Z[a] (one Philox call per (row, 8-column group), the stream every BatchNorm -> Dropout kernel of this
// library uses): the consumers, which all stage the same rows, then read bytes instead of each re-running Philox.
__device__ __forceinline__ void hs_emit_keep(uint8_t* __restrict__ keep_out, int H, int row0, int c0, const uint64_t* __restrict__ rng,
                                             uint32_t stream_id, float p_drop) {
  const int tid = threadIdx.x;
  if (tid >= HS_BM * (HS_BN / 8)) return;
  const int r = row0 + (tid >> 3), c = c0 + ((tid & 7) << 3);
  const uint64_t seed = rng[0], strm = rng[1] * 64ull + stream_id;
  const uint32_t thr16 = (uint32_t)(p_drop * 65536.0f + 0.5f);
  const uint32_t kb = keep8(nullptr, 0, (uint64_t)r * (uint64_t)(H >> 3) + (uint64_t)(c >> 3), seed, strm, thr16);
  uint2 o;
  o.x = (kb & 1u) | ((kb & 2u) << 7) | ((kb & 4u) << 14) | ((kb & 8u) << 21);
  o.y = ((kb >> 4) & 1u) | ((kb & 32u) << 3) | ((kb & 64u) << 10) | ((kb & 128u) << 17);
  *reinterpret_cast<uint2*>(keep_out + (size_t)r * H + c) = o;
}


static int hs_fill_seg(HsSeg& sg, int n_seg, const int32_t* seg_off_host, int R, const char* who) {
  SCNB_CHECK_ARG(seg_off_host && n_seg >= 1 && n_seg <= HS_MAXSEG, "%s: bad segments", who);
  sg.n_seg = n_seg;
  for (int i = 0; i <= HS_MAXSEG; ++i) sg.off[i] = seg_off_host[i <= n_seg ? i : n_seg];
  for (int i = 0; i <= n_seg; ++i) SCNB_CHECK_ARG(sg.off[i] % HS_BM == 0, "%s: segment boundaries must be multiples of 32 rows", who);
  SCNB_CHECK_ARG(sg.off[0] == 0 && sg.off[n_seg] == R, "%s: segments must cover all rows", who);
  return SCNB_OK;
}

template <typename KernelT>
static int hs_set_smem(KernelT kern, size_t bytes, const char* who) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      scnb_set_error("%s: cannot reserve %zu bytes of shared memory: %s", who, bytes, cudaGetErrorString(e));
      return SCNB_ERR_CUDA;
    }
  }
  return SCNB_OK;
}

