// Fused hidden stack of the scNym training step (sm_100a): BatchNorm -> Dropout -> ReLU folded into the GEMMs around it.
//
// Reference arithmetic: `[Linear -> BatchNorm1d(train) -> Dropout(0.5) -> ReLU] x n` (scnym/model.py:210-233, SURVEY
// Appendix A2) on the student super-batch [U-mixed | L-mixed | DAN] with per-segment batch statistics, and its adjoint.
//
// The unfused path runs every BatchNorm and every Linear as its own launch (12 launches on the step's critical chain,
// each 5-12 us for <= 1 MB of data).  Here a 32-row x 64-column output tile is the unit of work and the data-dependent
// parts of BatchNorm travel between launches as per-tile partial statistics:
//
//   forward  hs_mix      Z0 = MixUp(Z1) (hidden-space MixUp, dataprep.py:665-689)            + tile statistics of Z0
//            hs_fwd(a)   A[a-1] = relu(drop(bn(Z[a-1]))) applied WHILE the A operand is staged in shared memory,
//                        Z[a] = A[a-1] W_a^T + b_a (3xTF32 mma.sync, fp32 accumulate)          + tile statistics of Z[a]
//            hs_act      A[last] = relu(drop(bn(Z[last])))                                    (input of the heads)
//   backward hs_bwd_act  dY[last] = dA[last] * [A > 0] / (1-p)                                + tile sums (dy, dy*zhat)
//            hs_bwd(a)   dZ[a] = gamma*invstd*(dY - mean(dY) - zhat*mean(dY*zhat)) applied while staging the A operand,
//                        dA[a-1] = dZ[a] W_a;  dY[a-1] = dA[a-1] * [A[a-1] > 0] / (1-p) in the epilogue + tile sums
//            hs_bwd_in   dZ[0] (input of the MixUp adjoint)
//
// Tile statistics are (mean, M2 = sum (z - mean)^2) over the tile's 32 rows, combined per segment with Chan's parallel
// formula in a FIXED order by every consumer: the result is as accurate as the two-pass form of the unfused kernels and
// bit-reproducible run to run (no atomics in the statistics).  Backward tile sums are plain sums combined in tile order.
// Segment boundaries must be multiples of 32 rows and known on the host (pseudolabel thresholding off): the engine
// falls back to the unfused kernels otherwise.
#include "common.cuh"
#include "tc_common.cuh"
#include "hidden_common.cuh"
#include <stdlib.h>

// Phase stamps of CTA (0, 0) of the tile-GEMM kernels (SCNB_HS_DBG=1; tools/hs_phases.py reads them): %globaltimer in ns.
__device__ unsigned long long g_hs_dbg[2][16];
__device__ __forceinline__ void hs_stamp(int on, int kern, int slot) {
  if (on && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_hs_dbg[kern][slot] = t;
  }
}
static int hs_dbg_on() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SCNB_HS_DBG");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}
extern "C" int scnb_hs_debug_read(unsigned long long* out32_host) {
  cudaError_t e = cudaMemcpyFromSymbol(out32_host, g_hs_dbg, sizeof(unsigned long long) * 32);
  if (e != cudaSuccess) {
    scnb_set_error("hs_debug_read: %s", cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// hs_mix: out[i,:] = gam[i]*Z[srcA[i],:] + (1-gam[i])*Z[srcB[i],:] (srcA == NULL: out = Z, statistics only) + tile
// statistics of `out`.  grid = (H/64, R/32), 256 threads: thread -> (row tid/8, 8 columns).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HS_TE) k_hs_mix(const float* __restrict__ Z, int H, const int32_t* __restrict__ srcA,
                                                 const int32_t* __restrict__ srcB, const float* __restrict__ gam, int R,
                                                 float* __restrict__ out, float* __restrict__ part, uint8_t* __restrict__ keep_out,
                                                 const uint64_t* __restrict__ rng, uint32_t stream_id, float p_drop) {
  PDL_ENTRY();
  __shared__ float red[HS_BM][HS_BN + 1];
  const int tid = threadIdx.x, c0 = blockIdx.x * HS_BN, tile = blockIdx.y;
  const int i = tile * HS_BM + (tid >> 3), c = c0 + ((tid & 7) << 3);
  float v[8];
  if (!srcA) {
    hs_ld8(Z + (size_t)i * H + c, v);
  } else {
    const int a = srcA[i], b = srcB ? srcB[i] : a;
    const float g = gam ? gam[i] : 1.f;
    hs_ld8(Z + (size_t)a * H + c, v);
    if (!(a == b || g == 1.f)) {
      float w[8];
      hs_ld8(Z + (size_t)b * H + c, w);
      const float g1 = 1.f - g;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = g * v[j] + g1 * w[j];
    }
  }
  if (out != Z) hs_st8(out + (size_t)i * H + c, v);
  hs_tile_stats_rowmajor(v, red, part, H, tile, c0);
  if (keep_out) hs_emit_keep(keep_out, H, tile * HS_BM, c0, rng, stream_id, p_drop);   // dropout keep bytes of application 0
}

extern "C" int scnb_hs_mix(const float* Z, int H, const int32_t* srcA, const int32_t* srcB, const float* gam, int R, float* out,
                           float* part, uint8_t* keep_out, const uint64_t* rng, uint32_t stream_id, float p_drop, void* stream) {
  SCNB_CHECK_ARG(Z && out && part && H > 0 && H % HS_BN == 0 && R > 0 && R % HS_BM == 0, "hs_mix: needs H %% 64 == 0, rows %% 32 == 0");
  SCNB_CHECK_ARG(!keep_out || (rng && p_drop > 0.f && p_drop < 1.f), "hs_mix: keep_out needs rng state and 0 < p_drop < 1");
  scnb_launch_pdl(k_hs_mix, dim3(H / HS_BN, R / HS_BM), dim3(HS_TE), 0, (cudaStream_t)stream, Z, H, srcA, srcB, gam, R, out, part, keep_out,
                  rng, stream_id, p_drop);
  SCNB_CHECK_LAUNCH("hs_mix");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// hs_act: A = relu(drop(bn(Z))) for the LAST application (the heads read A); writes the segment statistics.
// grid = (H/64, R/32).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HS_TE) k_hs_act(const float* __restrict__ Z, float* __restrict__ A, int H, HsSeg sg,
                                                 const float* __restrict__ part, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, float* __restrict__ stats,
                                                 const uint8_t* __restrict__ keep_mask, const uint64_t* __restrict__ rng,
                                                 uint32_t stream_id, float p_drop) {
  PDL_ENTRY();
  __shared__ float sMean[HS_BN], sInv[HS_BN];
  const int tid = threadIdx.x, c0 = blockIdx.x * HS_BN, tile = blockIdx.y;
  const int row0 = tile * HS_BM, s = hs_seg_of(sg, row0);
  const int t0 = sg.off[s] / HS_BM, t1 = sg.off[s + 1] / HS_BM;
  if (tid < HS_BN) {
    if (!part) {   // data parallel: the exchange kernel has written the GLOBAL segment statistics
      sMean[tid] = stats[((size_t)s * 3 + 0) * H + c0 + tid];
      sInv[tid] = stats[((size_t)s * 3 + 2) * H + c0 + tid];
    } else {
      float mean, var;
      hs_combine_stats(part, H, c0 + tid, t0, t1, mean, var);
      const float inv = 1.0f / sqrtf(var + BN_EPS);
      sMean[tid] = mean;
      sInv[tid] = inv;
      if (tile == t0) {
        stats[((size_t)s * 3 + 0) * H + c0 + tid] = mean;
        stats[((size_t)s * 3 + 1) * H + c0 + tid] = var;
        stats[((size_t)s * 3 + 2) * H + c0 + tid] = inv;
      }
    }
  }
  __syncthreads();
  uint64_t seed = 0, strm = 0;
  if (p_drop > 0.f && !keep_mask) {
    seed = rng[0];
    strm = rng[1] * 64ull + stream_id;
  }
  const uint32_t thr16 = (uint32_t)(p_drop * 65536.0f + 0.5f);
  const float drop_scale = 1.0f / (1.0f - p_drop);
  const int r = row0 + (tid >> 3), cl = (tid & 7) << 3, c = c0 + cl;
  const size_t o = (size_t)r * H + c;
  float z[8], g[8], b[8];
  hs_ld8(Z + o, z);
  hs_ld8(gamma + c, g);
  hs_ld8(beta + c, b);
  uint32_t kb = 0xffu;
  if (p_drop > 0.f) kb = keep8(keep_mask, o, (uint64_t)r * (H >> 3) + (c >> 3), seed, strm, thr16);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float y = (z[j] - sMean[cl + j]) * sInv[cl + j] * g[j] + b[j];
    if (p_drop > 0.f) y = ((kb >> j) & 1u) ? y * drop_scale : 0.f;
    z[j] = fmaxf(y, 0.f);
  }
  hs_st8(A + o, z);
}

extern "C" int scnb_hs_act(const float* Z, float* A, int H, int R, int n_seg, const int32_t* seg_off_host, const float* part,
                           const float* gamma, const float* beta, float* stats, const uint8_t* keep_mask, const uint64_t* rng,
                           uint32_t stream_id, float p_drop, void* stream) {
  SCNB_CHECK_ARG(Z && A && gamma && beta && stats && seg_off_host, "hs_act: null argument");
  SCNB_CHECK_ARG(H % HS_BN == 0 && R % HS_BM == 0 && n_seg >= 1 && n_seg <= HS_MAXSEG, "hs_act: bad shape");
  SCNB_CHECK_ARG(!(p_drop > 0.f) || keep_mask || rng, "hs_act: dropout needs keep_mask or rng state");
  HsSeg sg;
  sg.n_seg = n_seg;
  for (int i = 0; i <= HS_MAXSEG; ++i) sg.off[i] = seg_off_host[i <= n_seg ? i : n_seg];
  for (int i = 0; i <= n_seg; ++i) SCNB_CHECK_ARG(sg.off[i] % HS_BM == 0, "hs_act: segment boundaries must be multiples of 32 rows");
  SCNB_CHECK_ARG(sg.off[n_seg] == R, "hs_act: segments must cover all rows");
  scnb_launch_pdl(k_hs_act, dim3(H / HS_BN, R / HS_BM), dim3(HS_TE), 0, (cudaStream_t)stream, Z, A, H, sg, part, gamma, beta, stats,
                  keep_mask, rng, stream_id, p_drop);
  SCNB_CHECK_LAUNCH("hs_act");
  return SCNB_OK;
}

struct HsFwdArgs {
  const float* Zin;      // [R, K] pre-BatchNorm input of application a-1
  float* Aout;           // [R, K] activation of application a-1 (materialised by the column-tile-0 CTAs; NULL: not kept)
  const float* part_in;  // [R/32][2][K] tile statistics of Zin (train mode)
  float* stats_in;       // [n_seg][3][K] segment statistics of Zin (written here; train mode)
  const float* rmean;    // eval mode (the MixMatch teacher, losses.py:660-737): running statistics instead of batch
  const float* rvar;     //   statistics, no dropout, no statistics of the output
  const float* gamma;    // BatchNorm of application a-1
  const float* beta;
  const uint8_t* keep;   // dropout keep bytes of application a-1 ([R, K]; written by the producer of Zin, or injected)
  float p_drop;
  const float* W;        // [N, K] Linear of application a
  const float* bias;     // [N]
  float* Zout;           // [R, N]
  float* part_out;       // [R/32][2][N]
  uint8_t* keep_out;     // [R, N] dropout keep bytes of application a (NULL: not generated here)
  const uint64_t* rng;
  uint32_t out_stream_id;
  float out_p_drop;
  int K, N, R;
  int dbg;
  HsSeg sg;
};

// grid = (N/64, R/32), 512 threads; dynamic shared memory: sA 32*(K+4) + max(sB 64*(K+4), partial tiles) + 4*K floats.
__global__ void __launch_bounds__(HS_T) k_hs_fwd(const __grid_constant__ HsFwdArgs p) {
  PDL_LAUNCH();
  extern __shared__ __align__(16) float hs_smem[];
  __shared__ __align__(8) unsigned long long bar_mem;
  const int K = p.K, N = p.N;
  const int b_floats = max(HS_BN * (K + 4), HS_TILE_FLOATS);
  float* sA = hs_smem;
  float* sB = sA + HS_BM * (K + 4);
  float* sT = sB;                       // partial tiles alias the weight panel once the MMAs are done
  float* sMean = sB + b_floats;
  float* sInv = sMean + K;
  float* sGam = sInv + K;
  float* sBet = sGam + K;
  const int tid = threadIdx.x, n0 = blockIdx.x * HS_BN, tile = blockIdx.y, row0 = tile * HS_BM;
  const int s = hs_seg_of(p.sg, row0);
  const int t0 = p.sg.off[s] / HS_BM, t1 = p.sg.off[s + 1] / HS_BM;
  const uint32_t bar = smem_u32(&bar_mem);
  hs_stamp(p.dbg, 0, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // stage the raw operands: one bulk copy per row (96 copies per CTA instead of 6144 16-byte cp.async)
  if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)((HS_BM + HS_BN) * K * sizeof(float)));
  // the weight panel and the BatchNorm parameters do not depend on the kernel in front: fetched before the dependency wait
  hs_bulk_rows(sB, K + 4, p.W + (size_t)n0 * K, K, HS_BN, (uint32_t)(K * sizeof(float)), bar, HS_BM);
  PDL_WAIT();
  hs_bulk_rows(sA, K + 4, p.Zin + (size_t)row0 * K, K, HS_BM, (uint32_t)(K * sizeof(float)), bar, 0);
  // (a global load whose value goes to shared memory stalls its thread for the round trip: every copy is issued first and
  // the BatchNorm parameters are requested together with the tile statistics below)
  // dropout keep bytes of the float4 groups this thread transforms below, requested now: the round trip overlaps the operand
  // copies and the statistics instead of sitting between "operands landed" and the MMAs
  const int fpr = K >> 2;
  const bool drop = p.p_drop > 0.f;
  const bool fixed_cols = (HS_T % fpr == 0) && (HS_BM * fpr <= 8 * HS_T);
  uint32_t kbv[8];
  if (fixed_cols && drop) {
    const int rpp = HS_T / fpr, r0 = tid / fpr, c = (tid - r0 * fpr) << 2;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = r0 + u * rpp;
      kbv[u] = r < HS_BM ? hs_keep4_raw(p.keep, (size_t)(row0 + r) * K + c) : 0u;
    }
  }
  hs_stamp(p.dbg, 0, 1);
  // segment statistics of the input while the copies fly
  const bool eval = p.rmean != nullptr;
  for (int c = tid; c < K; c += HS_T) {
    float mean, var, inv;
    const float gam_c = p.gamma[c], bet_c = p.beta[c];
    const bool given = !eval && !p.part_in;   // data parallel: GLOBAL statistics already in stats_in
    if (eval) {
      mean = p.rmean[c];
      var = p.rvar[c];
      inv = 1.0f / sqrtf(var + BN_EPS);
    } else if (given) {
      mean = p.stats_in[((size_t)s * 3 + 0) * K + c];
      var = 0.f;
      inv = p.stats_in[((size_t)s * 3 + 2) * K + c];
    } else {
      hs_combine_stats(p.part_in, K, c, t0, t1, mean, var);
      inv = 1.0f / sqrtf(var + BN_EPS);
    }
    sMean[c] = mean;
    sInv[c] = inv;
    sGam[c] = gam_c;
    sBet[c] = bet_c;
    if (!eval && !given && blockIdx.x == 0 && tile == t0) {
      p.stats_in[((size_t)s * 3 + 0) * K + c] = mean;
      p.stats_in[((size_t)s * 3 + 1) * K + c] = var;
      p.stats_in[((size_t)s * 3 + 2) * K + c] = inv;
    }
  }
  hs_stamp(p.dbg, 0, 2);
  mbar_wait(bar, 0);
  __syncthreads();
  hs_stamp(p.dbg, 0, 3);
  // BatchNorm -> Dropout -> ReLU in place on the A operand (consecutive lanes take consecutive float4: no bank conflicts)
  const float drop_scale = 1.0f / (1.0f - p.p_drop);
  auto transform = [&](int r, int c, uint32_t kb_pre, bool have_kb) {
    float* a = sA + r * (K + 4) + c;
    const int rg = row0 + r;
    const float4 z = *reinterpret_cast<const float4*>(a);
    const float4 mu = *reinterpret_cast<const float4*>(sMean + c), is = *reinterpret_cast<const float4*>(sInv + c);
    const float4 gm = *reinterpret_cast<const float4*>(sGam + c), bt = *reinterpret_cast<const float4*>(sBet + c);
    float4 y;
    y.x = (z.x - mu.x) * is.x * gm.x + bt.x;
    y.y = (z.y - mu.y) * is.y * gm.y + bt.y;
    y.z = (z.z - mu.z) * is.z * gm.z + bt.z;
    y.w = (z.w - mu.w) * is.w * gm.w + bt.w;
    if (drop) {
      const uint32_t kb = have_kb ? hs_keep4_bits(kb_pre) : hs_keep4(p.keep, (size_t)rg * K + c);
      y.x = (kb & 1u) ? y.x * drop_scale : 0.f;
      y.y = (kb & 2u) ? y.y * drop_scale : 0.f;
      y.z = (kb & 4u) ? y.z * drop_scale : 0.f;
      y.w = (kb & 8u) ? y.w * drop_scale : 0.f;
    }
    y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
    *reinterpret_cast<float4*>(a) = y;
    if (blockIdx.x == 0 && p.Aout) *reinterpret_cast<float4*>(p.Aout + (size_t)rg * K + c) = y;
  };
  if (fixed_cols) {        // every thread keeps its float4 column and walks rows (no divisions in the loop)
    const int rpp = HS_T / fpr, r0 = tid / fpr, c = (tid - r0 * fpr) << 2;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = r0 + u * rpp;
      if (r < HS_BM) transform(r, c, kbv[u], true);
    }
  } else if (HS_T % fpr == 0) {
    const int rpp = HS_T / fpr, r0 = tid / fpr, c = (tid - r0 * fpr) << 2;
#pragma unroll 4
    for (int r = r0; r < HS_BM; r += rpp) transform(r, c, 0u, false);
  } else {
    for (int f = tid; f < HS_BM * fpr; f += HS_T) {
      const int r = f / fpr;
      transform(r, (f - r * fpr) << 2, 0u, false);
    }
  }
  __syncthreads();
  hs_stamp(p.dbg, 0, 4);
  float acc[2][2][4];
  hs_tile_mma<0>(sA, sB, K, acc);
  __syncthreads();                       // every warp is done with the weight panel: the partial tiles may overwrite it
  hs_tile_store_partials(sT, acc);
  __syncthreads();
  hs_stamp(p.dbg, 0, 5);
  // epilogue on all threads: thread -> (row tid/16, 4 columns): bias, Z[a] tile, tile statistics
  {
    const int row = tid >> 4, c4 = (tid & 15) << 2;
    float4 v = hs_tile_sum4(sT, row, c4);
    if (p.bias) {
      const float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + c4);
      v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
    }
    *reinterpret_cast<float4*>(p.Zout + (size_t)(row0 + row) * N + n0 + c4) = v;
    *reinterpret_cast<float4*>(sT + row * HS_TP + c4) = v;      // own elements only: no hazard with other threads' reads
  }
  if (p.part_out) {
    __syncthreads();
    if (tid < HS_BN) {
      float mean = 0.f;
#pragma unroll 8
      for (int i = 0; i < HS_BM; ++i) mean += sT[i * HS_TP + tid];
      mean *= (1.0f / HS_BM);
      float m2 = 0.f;
#pragma unroll 8
      for (int i = 0; i < HS_BM; ++i) {
        const float d = sT[i * HS_TP + tid] - mean;
        m2 = fmaf(d, d, m2);
      }
      p.part_out[((size_t)tile * 2 + 0) * N + n0 + tid] = mean;
      p.part_out[((size_t)tile * 2 + 1) * N + n0 + tid] = m2;
    }
  }
  if (p.keep_out) hs_emit_keep(p.keep_out, N, row0, n0, p.rng, p.out_stream_id, p.out_p_drop);
  hs_stamp(p.dbg, 0, 6);
}

extern "C" int scnb_hs_fwd(const float* Zin, float* Aout, const float* part_in, float* stats_in, const float* gamma,
                           const float* beta, const uint8_t* keep_mask, float p_drop, const float* W, const float* bias, float* Zout,
                           float* part_out, uint8_t* keep_out, const uint64_t* rng, uint32_t out_stream_id, float out_p_drop, int K,
                           int N, int R, int n_seg, const int32_t* seg_off_host, const float* running_mean,
                           const float* running_var, void* stream) {
  SCNB_CHECK_ARG(Zin && gamma && beta && W && Zout, "hs_fwd: null argument");
  SCNB_CHECK_ARG((running_mean && running_var && p_drop == 0.f) || (!running_mean && stats_in && Aout && part_out),
                 "hs_fwd: train mode needs stats_in / Aout / part_out (+ part_in unless the statistics are given), eval mode running "
                 "statistics and no dropout");
  SCNB_CHECK_ARG(K % 32 == 0 && K >= 32 && K <= 512 && N % HS_BN == 0 && R % HS_BM == 0, "hs_fwd: needs K %% 32 == 0 (<= 512), N %% 64 == 0, rows %% 32 == 0");
  SCNB_CHECK_ARG(!(p_drop > 0.f) || keep_mask, "hs_fwd: dropout needs the keep bytes of the input application");
  SCNB_CHECK_ARG(!keep_out || (rng && out_p_drop > 0.f && out_p_drop < 1.f), "hs_fwd: keep_out needs rng state and 0 < out_p_drop < 1");
  SCNB_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "hs_fwd: p_drop out of range");
  SCNB_CHECK_ARG((((uintptr_t)Zin | (uintptr_t)W | (uintptr_t)Zout | (uintptr_t)bias | (uintptr_t)keep_mask | (uintptr_t)keep_out) % 16) == 0,
                 "hs_fwd: operands must be 16-byte aligned");
  HsFwdArgs a;
  a.Zin = Zin; a.Aout = Aout; a.part_in = part_in; a.stats_in = stats_in; a.gamma = gamma; a.beta = beta; a.keep = keep_mask;
  a.p_drop = p_drop; a.W = W; a.bias = bias; a.Zout = Zout; a.part_out = part_out; a.keep_out = keep_out; a.rng = rng;
  a.out_stream_id = out_stream_id; a.out_p_drop = out_p_drop;
  a.K = K; a.N = N; a.R = R; a.rmean = running_mean; a.rvar = running_var; a.dbg = hs_dbg_on();
  int rc = hs_fill_seg(a.sg, n_seg, seg_off_host, R, "hs_fwd");
  if (rc != SCNB_OK) return rc;
  const size_t b_floats = (size_t)HS_BN * (K + 4) > (size_t)HS_TILE_FLOATS ? (size_t)HS_BN * (K + 4) : (size_t)HS_TILE_FLOATS;
  const size_t smem = ((size_t)HS_BM * (K + 4) + b_floats + 4 * (size_t)K) * sizeof(float);
  rc = hs_set_smem(k_hs_fwd, smem, "hs_fwd");
  if (rc != SCNB_OK) return rc;
  scnb_launch_pdl(k_hs_fwd, dim3(N / HS_BN, R / HS_BM), dim3(HS_T), smem, (cudaStream_t)stream, a);
  SCNB_CHECK_LAUNCH("hs_fwd");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// hs_bwd_act: dY = dA * [A > 0] / (1-p) for the LAST application + tile sums (sum dY, sum dY*zhat).  grid = (H/64, R/32).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HS_TE) k_hs_bwd_act(const float* __restrict__ dA, const float* __restrict__ Z,
                                                     const float* __restrict__ A, int H, HsSeg sg, const float* __restrict__ stats,
                                                     float p_drop, float* __restrict__ dY, float* __restrict__ part) {
  PDL_ENTRY();
  __shared__ float red[HS_BM][HS_BN + 1];
  const int tid = threadIdx.x, c0 = blockIdx.x * HS_BN, tile = blockIdx.y;
  const int r = tile * HS_BM + (tid >> 3), c = c0 + ((tid & 7) << 3);
  const int s = hs_seg_of(sg, tile * HS_BM);
  const size_t o = (size_t)r * H + c;
  const float drop_scale = 1.0f / (1.0f - p_drop);
  float da[8], z[8], a[8], mean[8], inv[8], dy[8], dyz[8];
  hs_ld8(dA + o, da);
  hs_ld8(Z + o, z);
  hs_ld8(A + o, a);
  hs_ld8(stats + ((size_t)s * 3 + 0) * H + c, mean);
  hs_ld8(stats + ((size_t)s * 3 + 2) * H + c, inv);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dy[j] = a[j] > 0.f ? da[j] * drop_scale : 0.f;      // A > 0 implies the element was kept by the dropout
    dyz[j] = dy[j] * ((z[j] - mean[j]) * inv[j]);
  }
  hs_st8(dY + o, dy);
  hs_tile_sums_rowmajor(dy, dyz, red, part, H, tile, c0);
}

extern "C" int scnb_hs_bwd_act(const float* dA, const float* Z, const float* A, int H, int R, int n_seg, const int32_t* seg_off_host,
                               const float* stats, float p_drop, float* dY, float* part, void* stream) {
  SCNB_CHECK_ARG(dA && Z && A && stats && dY && part, "hs_bwd_act: null argument");
  SCNB_CHECK_ARG(H % HS_BN == 0 && R % HS_BM == 0 && p_drop >= 0.f && p_drop < 1.f, "hs_bwd_act: bad shape");
  HsSeg sg;
  int rc = hs_fill_seg(sg, n_seg, seg_off_host, R, "hs_bwd_act");
  if (rc != SCNB_OK) return rc;
  scnb_launch_pdl(k_hs_bwd_act, dim3(H / HS_BN, R / HS_BM), dim3(HS_TE), 0, (cudaStream_t)stream, dA, Z, A, H, sg, stats, p_drop, dY, part);
  SCNB_CHECK_LAUNCH("hs_bwd_act");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// hs_bwd: dZ[a] = gamma*invstd*(dY - mean(dY) - zhat*mean(dY*zhat)) formed while the A operand is staged (and written
// out by the column-tile-0 CTAs, for the weight gradients), dA[a-1] = dZ[a] W_a, dY[a-1] = dA[a-1]*[A[a-1] > 0]/(1-p)
// in the epilogue + tile sums.  The segment totals of (dY, dY*zhat) are the BatchNorm beta / gamma gradients.
// ---------------------------------------------------------------------------------------------------------------------
struct HsBwdArgs {
  const float* dYin;     // [R, K] gradient w.r.t. the BatchNorm output of application a (before the dropout scale: already applied)
  const float* Zin;      // [R, K] pre-BatchNorm input of application a
  const float* stats_in; // [n_seg][3][K]
  const float* part_in;  // [R/32][2][K] tile sums of (dY, dY*zhat); NULL: bmean_in holds the GLOBAL segment means
  const float* bmean_in; // [n_seg][2][K] mean(dY), mean(dY*zhat) (data parallel)
  const float* gamma;    // [K]
  float* dgamma;         // [K] += over segments
  float* dbeta;
  float* dZout;          // [R, K] materialised dZ[a]
  const float* W;        // [K, N] Linear of application a (out = K, in = N)
  const float* Zprev;    // [R, N] pre-BatchNorm input of application a-1
  const float* Aprev;    // [R, N] activation of application a-1
  const float* stats_prev;  // [n_seg][3][N]
  float p_drop_prev;
  float* dYout;          // [R, N]
  float* part_out;       // [R/32][2][N]
  int K, N, R;
  HsSeg sg;
};

// grid = (N/64, R/32), 512 threads; dynamic shared memory: sA, sZ 32*(K+4) each + max(sB K*(64+8), partial tiles) + 5*K floats.
__global__ void __launch_bounds__(HS_T) k_hs_bwd(const __grid_constant__ HsBwdArgs p) {
  PDL_LAUNCH();
  extern __shared__ __align__(16) float hs_smem[];
  __shared__ __align__(8) unsigned long long bar_mem;
  const int K = p.K, N = p.N;
  const int b_floats = max(K * (HS_BN + 8), HS_TILE_FLOATS);
  float* sA = hs_smem;                       // dY tile, transformed into dZ in place
  float* sZ = sA + HS_BM * (K + 4);          // Z[a] tile [32][K + 4]
  float* sB = sZ + HS_BM * (K + 4);          // W_a panel [K][64 + 8]
  float* sT = sB;                            // partial tiles alias the weight panel after the MMAs
  float* sM1 = sB + b_floats;                // mean(dY) per column
  float* sM2 = sM1 + K;                      // mean(dY*zhat)
  float* sMu = sM2 + K;                      // BatchNorm mean / invstd / gamma of application a
  float* sIs = sMu + K;
  float* sGm = sIs + K;
  const int tid = threadIdx.x, n0 = blockIdx.x * HS_BN, tile = blockIdx.y, row0 = tile * HS_BM;
  const int s = hs_seg_of(p.sg, row0);
  const int t0 = p.sg.off[s] / HS_BM, t1 = p.sg.off[s + 1] / HS_BM;
  const uint32_t bar = smem_u32(&bar_mem);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)((2 * HS_BM * K + K * HS_BN) * sizeof(float)));
  // the weight panel, this application's pre-BatchNorm tile and statistics come from the forward pass: before the wait
  hs_bulk_rows(sB, HS_BN + 8, p.W + n0, N, K, (uint32_t)(HS_BN * sizeof(float)), bar, 2 * HS_BM);
  hs_bulk_rows(sZ, K + 4, p.Zin + (size_t)row0 * K, K, HS_BM, (uint32_t)(K * sizeof(float)), bar, HS_BM);
  const int erow = tid >> 4, ec4 = (tid & 15) << 2;
  const size_t eo = (size_t)(row0 + erow) * N + n0 + ec4;
  const float4 av = *reinterpret_cast<const float4*>(p.Aprev + eo);
  const float4 zv = *reinterpret_cast<const float4*>(p.Zprev + eo);
  const float4 pmu = *reinterpret_cast<const float4*>(p.stats_prev + ((size_t)s * 3 + 0) * N + n0 + ec4);
  const float4 pis = *reinterpret_cast<const float4*>(p.stats_prev + ((size_t)s * 3 + 2) * N + n0 + ec4);
  PDL_WAIT();
  hs_bulk_rows(sA, K + 4, p.dYin + (size_t)row0 * K, K, HS_BM, (uint32_t)(K * sizeof(float)), bar, 0);
  // (loads whose values go to shared memory stall their thread for the round trip: every copy is issued first, and the
  // column constants are requested together with the tile sums -- one round trip for all of them)
  const float inv_n = 1.0f / (float)((t1 - t0) * HS_BM);
  for (int c = tid; c < K; c += HS_T) {
    const float mu_c = p.stats_in[((size_t)s * 3 + 0) * K + c];
    const float is_c = p.stats_in[((size_t)s * 3 + 2) * K + c];
    const float gm_c = p.gamma[c];
    if (!p.part_in) {
      sM1[c] = p.bmean_in[((size_t)s * 2 + 0) * K + c];
      sM2[c] = p.bmean_in[((size_t)s * 2 + 1) * K + c];
    } else {
      float s1, s2;
      hs_combine_sums(p.part_in, K, c, t0, t1, s1, s2);
      sM1[c] = s1 * inv_n;
      sM2[c] = s2 * inv_n;
      if (blockIdx.x == 0 && tile == t0) {   // one CTA per segment: BatchNorm parameter gradients
        atomicAdd(p.dbeta + c, s1);
        atomicAdd(p.dgamma + c, s2);
      }
    }
    sMu[c] = mu_c;
    sIs[c] = is_c;
    sGm[c] = gm_c;
  }
  mbar_wait(bar, 0);
  __syncthreads();
  const int fpr = K >> 2;
  auto transform = [&](int r, int c) {
    float* a = sA + r * (K + 4) + c;
    const float4 dy = *reinterpret_cast<const float4*>(a), z = *reinterpret_cast<const float4*>(sZ + r * (K + 4) + c);
    const float4 mu = *reinterpret_cast<const float4*>(sMu + c), is = *reinterpret_cast<const float4*>(sIs + c);
    const float4 gm = *reinterpret_cast<const float4*>(sGm + c);
    const float4 m1 = *reinterpret_cast<const float4*>(sM1 + c), m2 = *reinterpret_cast<const float4*>(sM2 + c);
    float4 v;
    v.x = gm.x * is.x * (dy.x - m1.x - (z.x - mu.x) * is.x * m2.x);
    v.y = gm.y * is.y * (dy.y - m1.y - (z.y - mu.y) * is.y * m2.y);
    v.z = gm.z * is.z * (dy.z - m1.z - (z.z - mu.z) * is.z * m2.z);
    v.w = gm.w * is.w * (dy.w - m1.w - (z.w - mu.w) * is.w * m2.w);
    *reinterpret_cast<float4*>(a) = v;
    if (blockIdx.x == 0) *reinterpret_cast<float4*>(p.dZout + (size_t)(row0 + r) * K + c) = v;
  };
  if (HS_T % fpr == 0) {
    const int rpp = HS_T / fpr, r0 = tid / fpr, c = (tid - r0 * fpr) << 2;
#pragma unroll 4
    for (int r = r0; r < HS_BM; r += rpp) transform(r, c);
  } else {
    for (int f = tid; f < HS_BM * fpr; f += HS_T) {
      const int r = f / fpr;
      transform(r, (f - r * fpr) << 2);
    }
  }
  __syncthreads();
  float acc[2][2][4];
  hs_tile_mma<1>(sA, sB, K, acc);
  __syncthreads();
  hs_tile_store_partials(sT, acc);
  __syncthreads();
  // epilogue on all threads: dY[a-1] = dA[a-1] * [A[a-1] > 0] / (1-p), tile sums of (dY, dY*zhat)
  {
    const float drop_scale = 1.0f / (1.0f - p.p_drop_prev);
    const float4 da = hs_tile_sum4(sT, erow, ec4);
    float4 d, dz;
    d.x = av.x > 0.f ? da.x * drop_scale : 0.f;       // A > 0 implies the element was kept by the dropout
    d.y = av.y > 0.f ? da.y * drop_scale : 0.f;
    d.z = av.z > 0.f ? da.z * drop_scale : 0.f;
    d.w = av.w > 0.f ? da.w * drop_scale : 0.f;
    dz.x = d.x * ((zv.x - pmu.x) * pis.x);
    dz.y = d.y * ((zv.y - pmu.y) * pis.y);
    dz.z = d.z * ((zv.z - pmu.z) * pis.z);
    dz.w = d.w * ((zv.w - pmu.w) * pis.w);
    *reinterpret_cast<float4*>(p.dYout + eo) = d;
    *reinterpret_cast<float4*>(sT + erow * HS_TP + ec4) = d;                                  // partial tile 0: own elements
    *reinterpret_cast<float4*>(sT + (size_t)HS_BM * HS_TP + erow * HS_TP + ec4) = dz;         // partial tile 1: own elements
  }
  __syncthreads();
  if (tid < 2 * HS_BN) {
    const int which = tid >> 6, c = tid & (HS_BN - 1);
    const float* src = sT + (size_t)which * HS_BM * HS_TP + c;
    float acc1 = 0.f;
#pragma unroll 8
    for (int i = 0; i < HS_BM; ++i) acc1 += src[i * HS_TP];
    p.part_out[((size_t)tile * 2 + which) * N + n0 + c] = acc1;
  }
}

extern "C" int scnb_hs_bwd(const float* dYin, const float* Zin, const float* stats_in, const float* part_in, const float* gamma,
                           float* dgamma, float* dbeta, float* dZout, const float* W, const float* Zprev, const float* Aprev,
                           const float* stats_prev, float p_drop_prev, float* dYout, float* part_out, int K, int N, int R, int n_seg,
                           const int32_t* seg_off_host, const float* bmean_in, void* stream) {
  SCNB_CHECK_ARG(dYin && Zin && stats_in && (part_in || bmean_in) && gamma && dZout && W && Zprev && Aprev && stats_prev &&
                     dYout && part_out && (!part_in || (dgamma && dbeta)), "hs_bwd: null argument");
  SCNB_CHECK_ARG(K % 32 == 0 && K >= 32 && K <= 256 && N % HS_BN == 0 && R % HS_BM == 0, "hs_bwd: needs K %% 32 == 0 (<= 256), N %% 64 == 0, rows %% 32 == 0");
  SCNB_CHECK_ARG(p_drop_prev >= 0.f && p_drop_prev < 1.f, "hs_bwd: p_drop out of range");
  HsBwdArgs a;
  a.dYin = dYin; a.Zin = Zin; a.stats_in = stats_in; a.part_in = part_in; a.gamma = gamma; a.dgamma = dgamma; a.dbeta = dbeta;
  a.dZout = dZout; a.W = W; a.Zprev = Zprev; a.Aprev = Aprev; a.stats_prev = stats_prev; a.p_drop_prev = p_drop_prev; a.dYout = dYout;
  a.part_out = part_out; a.K = K; a.N = N; a.R = R; a.bmean_in = bmean_in;
  int rc = hs_fill_seg(a.sg, n_seg, seg_off_host, R, "hs_bwd");
  if (rc != SCNB_OK) return rc;
  const size_t b_floats = (size_t)K * (HS_BN + 8) > (size_t)HS_TILE_FLOATS ? (size_t)K * (HS_BN + 8) : (size_t)HS_TILE_FLOATS;
  const size_t smem = ((size_t)2 * HS_BM * (K + 4) + b_floats + 5 * (size_t)K) * sizeof(float);
  rc = hs_set_smem(k_hs_bwd, smem, "hs_bwd");
  if (rc != SCNB_OK) return rc;
  scnb_launch_pdl(k_hs_bwd, dim3(N / HS_BN, R / HS_BM), dim3(HS_T), smem, (cudaStream_t)stream, a);
  SCNB_CHECK_LAUNCH("hs_bwd");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// hs_bwd_in: dZ[0] from dY[0] (application 0 has no Linear inside the stack: its input gradient feeds the MixUp adjoint
// and the first-layer weight gradient).  grid = (H/64, R/32).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HS_TE) k_hs_bwd_in(const float* __restrict__ dY, const float* __restrict__ Z, int H, HsSeg sg,
                                                    const float* __restrict__ stats, const float* __restrict__ part,
                                                    const float* __restrict__ gamma, float* __restrict__ dgamma,
                                                    float* __restrict__ dbeta, float* __restrict__ dZ,
                                                    const float* __restrict__ bmean, float* __restrict__ db) {
  PDL_ENTRY();
  __shared__ float sM1[HS_BN], sM2[HS_BN];
  const int tid = threadIdx.x, c0 = blockIdx.x * HS_BN, tile = blockIdx.y, row0 = tile * HS_BM;
  const int s = hs_seg_of(sg, row0);
  const int t0 = sg.off[s] / HS_BM, t1 = sg.off[s + 1] / HS_BM;
  if (tid < HS_BN) {
    if (!part) {
      sM1[tid] = bmean[((size_t)s * 2 + 0) * H + c0 + tid];
      sM2[tid] = bmean[((size_t)s * 2 + 1) * H + c0 + tid];
    } else {
      float s1, s2;
      hs_combine_sums(part, H, c0 + tid, t0, t1, s1, s2);
      const float inv_n = 1.0f / (float)((t1 - t0) * HS_BM);
      sM1[tid] = s1 * inv_n;
      sM2[tid] = s2 * inv_n;
      if (tile == t0) {
        atomicAdd(dbeta + c0 + tid, s1);
        atomicAdd(dgamma + c0 + tid, s2);
      }
    }
  }
  __syncthreads();
  const int r = row0 + (tid >> 3), cl = (tid & 7) << 3, c = c0 + cl;
  const size_t o = (size_t)r * H + c;
  float dy[8], z[8], mean[8], inv[8], g[8];
  hs_ld8(dY + o, dy);
  hs_ld8(Z + o, z);
  hs_ld8(stats + ((size_t)s * 3 + 0) * H + c, mean);
  hs_ld8(stats + ((size_t)s * 3 + 2) * H + c, inv);
  hs_ld8(gamma + c, g);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float zh = (z[j] - mean[j]) * inv[j];
    dy[j] = g[j] * inv[j] * (dy[j] - sM1[cl + j] - zh * sM2[cl + j]);
  }
  hs_st8(dZ + o, dy);
  if (db) {   // bias gradient of the Linear in front of application 0: column sums of dZ[0] over all rows
    __shared__ float red[HS_BM][HS_BN + 1];
    const int rr = tid >> 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rr][cl + j] = dy[j];
    __syncthreads();
    if (tid < HS_BN) {
      float sacc = 0.f;
#pragma unroll 8
      for (int i = 0; i < HS_BM; ++i) sacc += red[i][tid];
      atomicAdd(db + c0 + tid, sacc);
    }
  }
}

extern "C" int scnb_hs_bwd_in(const float* dY, const float* Z, int H, int R, int n_seg, const int32_t* seg_off_host, const float* stats,
                              const float* part, const float* gamma, float* dgamma, float* dbeta, float* dZ, const float* bmean_in,
                              float* db, void* stream) {
  SCNB_CHECK_ARG(dY && Z && stats && (part || bmean_in) && gamma && dZ && (!part || (dgamma && dbeta)), "hs_bwd_in: null argument");
  SCNB_CHECK_ARG(H % HS_BN == 0 && R % HS_BM == 0, "hs_bwd_in: bad shape");
  HsSeg sg;
  int rc = hs_fill_seg(sg, n_seg, seg_off_host, R, "hs_bwd_in");
  if (rc != SCNB_OK) return rc;
  scnb_launch_pdl(k_hs_bwd_in, dim3(H / HS_BN, R / HS_BM), dim3(HS_TE), 0, (cudaStream_t)stream, dY, Z, H, sg, stats, part, gamma, dgamma,
                  dbeta, dZ, bmean_in, db);
  SCNB_CHECK_LAUNCH("hs_bwd_in");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// hs_dw: weight and bias gradient of a hidden Linear from the materialised dZ [R, M] (M = out features) and its input
// activation A [R, N]:  dW[M, N] += dZ^T A,  db[M] += column sums of dZ   (dW = dz^T a, db = sum_n dz).
// One CTA = a 64 x 64 tile of dW over a chunk of 64 rows (grid = (N/64, M/64, ceil(R/64))); partial products are added
// to dW with atomics (the gradient arena is zeroed at the start of the step).  16 warps as 4 x 4 sub-tiles of 16 x 16,
// 3xTF32.  Replaces a split-K GEMM launch plus a column-sum launch per layer on the step's weight-gradient branch.
// ---------------------------------------------------------------------------------------------------------------------
#define HS_DW_ROWS 64    // 37 KB of shared memory per CTA: co-resides with a tile-GEMM CTA of the critical chain on the same SM
__global__ void __launch_bounds__(HS_T) k_hs_dw(const float* __restrict__ dZ, int M, const float* __restrict__ A, int N, int R,
                                                float* __restrict__ dW, float* __restrict__ db) {
  PDL_ENTRY();
  extern __shared__ __align__(16) float hs_smem[];
  constexpr int SP = HS_BN + 8;
  float* sD = hs_smem;                      // [64][72]  dZ[rows][m0 .. m0+63]
  float* sA = sD + HS_DW_ROWS * SP;         // [64][72]  A[rows][n0 .. n0+63]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3, wm = warp >> 2, wn = warp & 3;
  const int n0 = blockIdx.x * HS_BN, m0 = blockIdx.y * HS_BN, r0 = blockIdx.z * HS_DW_ROWS;
  const int rows = min(HS_DW_ROWS, R - r0);
  // 64 rows x 16 chunks per panel; rows beyond R are zero-filled
  for (int ch = tid; ch < HS_DW_ROWS * 16; ch += HS_T) {
    const int r = ch >> 4, q = ch & 15;
    if (r < rows) {
      hs_cp16(sD + r * SP + q * 4, dZ + (size_t)(r0 + r) * M + m0 + q * 4);
      hs_cp16(sA + r * SP + q * 4, A + (size_t)(r0 + r) * N + n0 + q * 4);
    } else {
      *reinterpret_cast<float4*>(sD + r * SP + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(sA + r * SP + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  hs_commit();
  hs_wait_all();
  __syncthreads();
  float acc[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
#pragma unroll 4
  for (int kk = 0; kk < HS_DW_ROWS; kk += 8) {
    uint32_t ah[4], al[4], bh[2][2], bl[2][2];
    const float* pa = sD + (kk + t) * SP + wm * 16 + g;
    hs_split(pa[0], ah[0], al[0]);
    hs_split(pa[8], ah[1], al[1]);
    hs_split(pa[4 * SP], ah[2], al[2]);
    hs_split(pa[4 * SP + 8], ah[3], al[3]);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float* pb = sA + (kk + t) * SP + wn * 16 + 8 * j + g;
      hs_split(pb[0], bh[j][0], bl[j][0]);
      hs_split(pb[4 * SP], bh[j][1], bl[j][1]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      hs_mma(acc[j], al, bh[j]);
      hs_mma(acc[j], ah, bl[j]);
      hs_mma(acc[j], ah, bh[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int m = m0 + wm * 16 + g, n = n0 + wn * 16 + 8 * j + 2 * t;
    atomicAdd(dW + (size_t)m * N + n, acc[j][0]);
    atomicAdd(dW + (size_t)m * N + n + 1, acc[j][1]);
    atomicAdd(dW + (size_t)(m + 8) * N + n, acc[j][2]);
    atomicAdd(dW + (size_t)(m + 8) * N + n + 1, acc[j][3]);
  }
  if (db && blockIdx.x == 0 && tid < HS_BN) {   // bias gradient: one column-tile row of CTAs adds its rows' column sums
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < HS_DW_ROWS; ++r) s += sD[r * SP + tid];
    atomicAdd(db + m0 + tid, s);
  }
}

extern "C" int scnb_hs_dw(const float* dZ, int M, const float* A, int N, int R, float* dW, float* db, void* stream) {
  SCNB_CHECK_ARG(dZ && A && dW && M > 0 && N > 0 && R > 0, "hs_dw: bad arguments");
  SCNB_CHECK_ARG(M % HS_BN == 0 && N % HS_BN == 0 && (((uintptr_t)dZ | (uintptr_t)A) % 16 == 0), "hs_dw: needs M, N %% 64 == 0 and 16-byte aligned operands");
  const size_t smem = (size_t)2 * HS_DW_ROWS * (HS_BN + 8) * sizeof(float);
  int rc = hs_set_smem(k_hs_dw, smem, "hs_dw");
  if (rc != SCNB_OK) return rc;
  scnb_launch_pdl(k_hs_dw, dim3(N / HS_BN, M / HS_BN, scnb_cdiv(R, HS_DW_ROWS)), dim3(HS_T), smem, (cudaStream_t)stream, dZ, M, A, N, R, dW, db);
  SCNB_CHECK_LAUNCH("hs_dw");
  return SCNB_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// hs_xchg: cross-rank exchange of the per-segment BatchNorm statistics (data-parallel training, one process per GPU,
// statistics over the CONCATENATED global batch).  One launch between the kernel that wrote the tile partials and the one
// that consumes them: grid = (H/64, n_seg), 128 threads.  A CTA combines the LOCAL tiles of its (64 columns, segment),
// stores the result into every rank's exchange buffer over NVLink / NVSwitch peer memory and merges what the peers stored,
// in rank order (bit-identical statistics on every rank).  Every value travels as one 8-byte word {value, step number}: the
// step number is its ready flag ("LL" protocol), buffers are double-buffered by step parity.
//   forward  (kind 0): (mean, M2, rows) merged with Chan's formula -> stats[seg][3][H] = mean, biased var, invstd; cnt[seg] = rows
//   backward (kind 1): (sum dY, sum dY*zhat, rows) -> bmean[seg][2][H] = global means; the LOCAL sums are also this rank's
//                      BatchNorm beta / gamma gradients (averaged with all other gradients by the gradient exchange)
// packet (slot, parity, src rank, seg, column block) = 136 words: [v0[64] | v1[64] | rows | pad]
// ---------------------------------------------------------------------------------------------------------------------
#define HS_PKT 136
struct HsPeer {
  const unsigned long long* bufs;   // [world] exchange-buffer base addresses as seen from this process
  const long long* epoch;           // device scalar: step number
  int rank, world, slot, ncb_max, spin_limit;
};

__device__ __forceinline__ void hs_ll_store(unsigned long long base, size_t word, float v, unsigned e) {
  uint2* dst = reinterpret_cast<uint2*>(base) + word;
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(v)), "r"(e) : "memory");
}
__device__ __forceinline__ float hs_ll_load(unsigned long long base, size_t word, unsigned e, int spin_limit) {
  const uint2* src = reinterpret_cast<const uint2*>(base) + word;
  uint32_t v, f;
  int spins = 0;
  do {
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(src) : "memory");
    if (f != e && ++spins > spin_limit) __trap();   // a peer never arrived: fail loudly instead of hanging the GPU
  } while (f != e);
  return __uint_as_float(v);
}

__global__ void __launch_bounds__(128) k_hs_xchg(int kind, const float* __restrict__ part, int H, HsSeg sg, HsPeer px,
                                                 float* __restrict__ out, float* __restrict__ cnt_out, float* __restrict__ dgamma,
                                                 float* __restrict__ dbeta) {
  PDL_ENTRY();
  const int tid = threadIdx.x, cb = blockIdx.x, s = blockIdx.y, c = cb * HS_BN + tid;
  const int t0 = sg.off[s] / HS_BM, t1 = sg.off[s + 1] / HS_BM;
  const unsigned e = (unsigned)px.epoch[0];
  const float n_loc = (float)((t1 - t0) * HS_BM);
  const size_t pkt0 = ((((size_t)(px.slot * 2 + (int)(e & 1u)) * px.world) * sg.n_seg + 0) * px.ncb_max) * HS_PKT;   // + src, seg, cb below
  auto pkt = [&](int src) { return pkt0 + (((size_t)src * sg.n_seg + s) * px.ncb_max + cb) * HS_PKT; };
  if (tid < HS_BN) {
    float v0, v1;
    if (kind == 0) {
      float var;
      hs_combine_stats(part, H, c, t0, t1, v0, var);
      v1 = var * n_loc;                         // M2
    } else {
      hs_combine_sums(part, H, c, t0, t1, v0, v1);
      if (t1 > t0) {
        atomicAdd(dbeta + c, v0);
        atomicAdd(dgamma + c, v1);
      }
    }
    const size_t mine = pkt(px.rank);
    for (int p = 0; p < px.world; ++p) {
      hs_ll_store(px.bufs[p], mine + tid, v0, e);
      hs_ll_store(px.bufs[p], mine + 64 + tid, v1, e);
      if (tid == 0) hs_ll_store(px.bufs[p], mine + 128, n_loc, e);
    }
    const unsigned long long own = px.bufs[px.rank];
    float n = 0.f, a = 0.f, b = 0.f;    // forward: running (rows, mean, M2); backward: (rows, sum, sum)
    // first look at every rank's three words at once (system-scope loads are ~0.7 us each: 24 of them one after the other
    // were 14 us of this kernel at eight ranks), then poll only the words that had not arrived
    constexpr int MAXW = 8;
    uint2 w0[MAXW], w1[MAXW], w2[MAXW];
#pragma unroll
    for (int p = 0; p < MAXW; ++p)
      if (p < px.world) {
        const uint2* src = reinterpret_cast<const uint2*>(own) + pkt(p);
        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(w0[p].x), "=r"(w0[p].y) : "l"(src + tid) : "memory");
        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(w1[p].x), "=r"(w1[p].y) : "l"(src + 64 + tid) : "memory");
        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(w2[p].x), "=r"(w2[p].y) : "l"(src + 128) : "memory");
      }
#pragma unroll
    for (int p = 0; p < 32; ++p) {   // rank order: the same result on every rank
      if (p >= px.world) break;
      const size_t w = pkt(p);
      float u0, u1, np_;
      if (p < MAXW) {
        const int q = p < MAXW ? p : 0;
        u0 = (w0[q].y == e) ? __uint_as_float(w0[q].x) : hs_ll_load(own, w + tid, e, px.spin_limit);
        u1 = (w1[q].y == e) ? __uint_as_float(w1[q].x) : hs_ll_load(own, w + 64 + tid, e, px.spin_limit);
        np_ = (w2[q].y == e) ? __uint_as_float(w2[q].x) : hs_ll_load(own, w + 128, e, px.spin_limit);
      } else {
        u0 = hs_ll_load(own, w + tid, e, px.spin_limit);
        u1 = hs_ll_load(own, w + 64 + tid, e, px.spin_limit);
        np_ = hs_ll_load(own, w + 128, e, px.spin_limit);
      }
      if (kind == 0) {
        if (np_ > 0.f) {
          const float nn = n + np_, d = u0 - a;
          a += d * (np_ / nn);
          b += u1 + d * d * (n * np_ / nn);
          n = nn;
        }
      } else {
        a += u0;
        b += u1;
        n += np_;
      }
    }
    const float ng = fmaxf(n, 1.f);
    if (kind == 0) {
      const float var = b / ng;
      out[((size_t)s * 3 + 0) * H + c] = a;
      out[((size_t)s * 3 + 1) * H + c] = var;
      out[((size_t)s * 3 + 2) * H + c] = 1.0f / sqrtf(var + BN_EPS);
    } else {
      out[((size_t)s * 2 + 0) * H + c] = a / ng;
      out[((size_t)s * 2 + 1) * H + c] = b / ng;
    }
    if (cnt_out && cb == 0 && tid == 0) cnt_out[s] = n;
  }
}

extern "C" int scnb_hs_peer_bytes(int H_max, int n_seg, int n_slots, int world) {
  return (int)((long long)n_slots * 2 * world * n_seg * (long long)((H_max + HS_BN - 1) / HS_BN) * HS_PKT * 8);
}

extern "C" int scnb_hs_xchg(int kind, const float* part, int H, int R, int n_seg, const int32_t* seg_off_host, const uint64_t* peer_bufs,
                            int rank, int world, int slot, int n_slots, int H_max, const int64_t* epoch, long long buf_bytes,
                            float* out, float* cnt_out, float* dgamma, float* dbeta, void* stream) {
  SCNB_CHECK_ARG(part && out && peer_bufs && epoch && (kind == 0 || kind == 1), "hs_xchg: bad arguments");
  SCNB_CHECK_ARG(kind == 0 || (dgamma && dbeta), "hs_xchg: backward exchange needs dgamma / dbeta");
  SCNB_CHECK_ARG(H % HS_BN == 0 && H <= H_max && world >= 1 && world <= 32 && rank >= 0 && rank < world && slot >= 0 && slot < n_slots,
                 "hs_xchg: bad shape / rank / slot");
  SCNB_CHECK_ARG(buf_bytes >= scnb_hs_peer_bytes(H_max, n_seg, n_slots, world), "hs_xchg: exchange buffer too small");
  HsSeg sg;
  int rc = hs_fill_seg(sg, n_seg, seg_off_host, R, "hs_xchg");
  if (rc != SCNB_OK) return rc;
  static int spin = 0;
  if (!spin) {
    const char* ev = getenv("SCNB_PEER_SPIN_LIMIT");
    spin = ev ? atoi(ev) : (1 << 24);
    if (spin <= 0) spin = 1 << 24;
  }
  HsPeer px{(const unsigned long long*)peer_bufs, (const long long*)epoch, rank, world, slot, (H_max + HS_BN - 1) / HS_BN, spin};
  k_hs_xchg<<<dim3(H / HS_BN, n_seg), 128, 0, (cudaStream_t)stream>>>(kind, part, H, sg, px, out, cnt_out, dgamma, dbeta);
  SCNB_CHECK_LAUNCH("hs_xchg");
  return SCNB_OK;
}
