// Heads of the MixMatch + DAN training step fused with the ends of the hidden stack (sm_100a).
//
// Between the last hidden Linear and the backward of the hidden stack the unfused chain was
//     hs_act -> [DAN Linear -> DAN tail -> GRL product]  ||  classifier + losses   -> hs_bwd_act
// i.e. five dependent launches (27 us at R = 1024, H = 256) for work that is local to a row once the segment statistics of
// the last BatchNorm are known.  Here each branch is ONE launch that starts from the pre-BatchNorm tile Z[last] and ends with
// dY[last] and its tile sums (the input of scnb_hs_bwd):
//
//   scnb_hs_classif_head   rows of the MixMatch segments: BatchNorm -> Dropout -> ReLU (scnym/model.py:219-227), classifier
//                          (model.py:229), interpolation-consistency MSE / soft-target CE and dLogits
//                          (scnym/losses.py:880-923), dA = dLogits Wc, dY = dA [A > 0] / (1-p), tile sums.
//                          One CTA per 32-row tile, one warp per row (two rows per warp).
//   scnb_hs_dan_head       rows of the domain-adversary segments: the same activation, Hd = relu(A Wd1^T + b1), domain logits,
//                          CE against the domain label (scnym/model.py:377-383, scnym/losses.py:1209-1243),
//                          dHd = (ddlog Wd2) [Hd > 0], gradient reversal dA = -w dHd Wd1 (model.py:338), dY, tile sums.
//                          A thread-block CLUSTER of H/64 CTAs owns a 32-row tile; CTA j holds rows [64j, 64j + 64) of Wd1
//                          in shared memory ONCE and uses the panel for both products: as the [n][k] operand of
//                          Hd[:, 64j..] = A Wd1[64j.., :]^T and as the [k][n] operand of the partial product
//                          dHd[:, 64j..] Wd1[64j.., :].  The domain logits (partial dot products over each CTA's 64 hidden
//                          columns) and the partial dA tiles meet through distributed shared memory, summed in rank order.
//
// Arithmetic is the hidden stack's: 3xTF32 mma.sync with fp32 accumulation, statistics combined per segment from the tile
// partials in tile order (bit-reproducible), dropout keep bytes written by the producer of Z[last].
#include "common.cuh"
#include "tc_common.cuh"
#include "hidden_common.cuh"
#include "loss_rows.cuh"
#include <stdlib.h>

// Phase stamps of CTA (0, 0) (SCNB_HS_DBG=1; tools/head_phases.py reads them): %globaltimer in ns.
__device__ unsigned long long g_hh_dbg[2][16];
__device__ __forceinline__ void hh_stamp(int on, int kern, int slot) {
  if (on && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_hh_dbg[kern][slot] = t;
  }
}
static int hh_dbg_on() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SCNB_HS_DBG");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}
extern "C" int scnb_hh_debug_read(unsigned long long* out32_host) {
  cudaError_t e = cudaMemcpyFromSymbol(out32_host, g_hh_dbg, sizeof(unsigned long long) * 32);
  if (e != cudaSuccess) {
    scnb_set_error("hh_debug_read: %s", cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  return SCNB_OK;
}

struct HsHeadArgs {
  const float* Z;         // [R, H] pre-BatchNorm input of the last application
  float* A;               // [R, H] its activation (kept for the weight gradients of the heads)
  const float* part_in;   // [R/32][2][H] tile statistics of Z; NULL: `stats` already holds the (global) segment statistics
  float* stats;           // [n_seg][3][H] mean, biased variance, invstd (written here unless given)
  const float* gamma;     // [H]
  const float* beta;      // [H]
  const uint8_t* keep;    // [R, H] dropout keep bytes (p_drop > 0)
  float p_drop;
  float* dY;              // [R, H] gradient w.r.t. the BatchNorm output, dropout scale applied
  float* part_out;        // [R/32][2][H] tile sums of (dY, dY * zhat)
  int H, R;
  int dbg;
  HsSeg sg;
};

// per-column constants of the last BatchNorm for this CTA's segment (all H columns)
__device__ __forceinline__ void hh_segment_consts(const HsHeadArgs& p, int s, int t0, int t1, bool writer, float* sMean, float* sInv,
                                                  float* sGam, float* sBet, int nthreads) {
  const int H = p.H;
  for (int c = threadIdx.x; c < H; c += nthreads) {
    float mean, var, inv;
    const float gam_c = p.gamma[c], bet_c = p.beta[c];     // requested with the partials: one round trip
    if (!p.part_in) {
      mean = p.stats[((size_t)s * 3 + 0) * H + c];
      inv = p.stats[((size_t)s * 3 + 2) * H + c];
    } else {
      hs_combine_stats(p.part_in, H, c, t0, t1, mean, var);
      inv = 1.0f / sqrtf(var + BN_EPS);
      if (writer) {
        p.stats[((size_t)s * 3 + 0) * H + c] = mean;
        p.stats[((size_t)s * 3 + 1) * H + c] = var;
        p.stats[((size_t)s * 3 + 2) * H + c] = inv;
      }
    }
    sMean[c] = mean;
    sInv[c] = inv;
    sGam[c] = gam_c;
    sBet[c] = bet_c;
  }
}

// relu(drop(bn(z))) on four consecutive columns (same expression order as k_hs_act / k_hs_fwd)
__device__ __forceinline__ float4 hh_act4(const float4 z, const float* sMean, const float* sInv, const float* sGam, const float* sBet, int c,
                                          bool drop, uint32_t kb, float drop_scale) {
  const float4 mu = *reinterpret_cast<const float4*>(sMean + c), is = *reinterpret_cast<const float4*>(sInv + c);
  const float4 gm = *reinterpret_cast<const float4*>(sGam + c), bt = *reinterpret_cast<const float4*>(sBet + c);
  float4 y;
  y.x = (z.x - mu.x) * is.x * gm.x + bt.x;
  y.y = (z.y - mu.y) * is.y * gm.y + bt.y;
  y.z = (z.z - mu.z) * is.z * gm.z + bt.z;
  y.w = (z.w - mu.w) * is.w * gm.w + bt.w;
  if (drop) {
    y.x = (kb & 1u) ? y.x * drop_scale : 0.f;
    y.y = (kb & 2u) ? y.y * drop_scale : 0.f;
    y.z = (kb & 4u) ? y.z * drop_scale : 0.f;
    y.w = (kb & 8u) ? y.w * drop_scale : 0.f;
  }
  return make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
}

// =====================================================================================================================
// scnb_hs_classif_head
// =====================================================================================================================
__device__ __forceinline__ void hh_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t hh_map_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float hh_ld_cluster(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 hh_ld_cluster4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

#define HC_T 256               // threads of a classifier-head CTA: one warp per row
#define HC_ROWS (HC_T / 32)    // rows per CTA
#define HC_CS (HS_BM / HC_ROWS)   // CTAs per 32-row tile (one thread-block cluster)

// A cluster of HC_CS CTAs owns a 32-row tile (the unit of the tile sums), a CTA 8 of its rows, a warp one row: the rows'
// work (activation, C dot products of length H, the loss row, C axpys of length H) is issue-bound, so it is spread over
// 4x the SMs a CTA per tile would use; the column sums of the tile meet through distributed shared memory in rank order.
template <int NV>
__global__ void __launch_bounds__(HC_T) k_hs_classif_head(const __grid_constant__ HsHeadArgs p, const float* __restrict__ Wc,
                                                          const float* __restrict__ bc, int C, int U, int L,
                                                          const int32_t* __restrict__ n_conf_dev, const float* __restrict__ Y,
                                                          const int32_t* __restrict__ mapA, const int32_t* __restrict__ mapB,
                                                          const float* __restrict__ gam, const float* __restrict__ class_weight,
                                                          float unsup_weight, float* __restrict__ logits,
                                                          float* __restrict__ dlogits, float* __restrict__ loss8) {
  extern __shared__ __align__(16) float hh_smem[];
  const int H = p.H;
  float* Ws = hh_smem;                                  // [C][H]
  float* scratch = Ws + (size_t)C * H;                  // HC_ROWS x [C][33]: per-class partial dot products of a warp's row
  float* sMean = scratch + (size_t)HC_ROWS * C * 33;
  float* sInv = sMean + H;
  float* sGam = sInv + H;
  float* sBet = sGam + H;
  float* sRed = sBet + H;                               // [HC_ROWS][2][H] (dY, dY zhat) of every row
  float* sPart = sRed + (size_t)HC_ROWS * 2 * H;        // [2][H] column sums over this CTA's rows (read by the cluster)
  float (*sZ)[33] = reinterpret_cast<float (*)[33]>(sPart + 2 * (size_t)H);   // [8][33] logits
  float (*sDz)[33] = sZ + HC_ROWS;                      // dLogits
  float (*sY)[33] = sDz + HC_ROWS;                      // soft label rows (MixUp of the two source rows)
  float (*sYa)[33] = sY + HC_ROWS;                      // dominant source label rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x;                             // rank in the cluster
  const int tile = blockIdx.y, row0 = tile * HS_BM, crow0 = row0 + j * HC_ROWS;
  const int s = hs_seg_of(p.sg, row0);
  const int t0 = p.sg.off[s] / HS_BM, t1 = p.sg.off[s + 1] / HS_BM;
  hh_stamp(p.dbg, 0, 0);
  const bool drop = p.p_drop > 0.f;
  const float drop_scale = 1.0f / (1.0f - p.p_drop);
  const int r = crow0 + warp;                           // this warp's row
  // everything that comes from global memory is requested before anything waits
  float4 z[NV], a[NV];
  uint32_t kb[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const size_t o = (size_t)r * H + (i * 32 + lane) * 4;
    z[i] = *reinterpret_cast<const float4*>(p.Z + o);
    kb[i] = drop ? hs_keep4_raw(p.keep, o) : 0x01010101u;
  }
  const int n_conf_val = (row0 < U && n_conf_dev) ? *n_conf_dev : U;
  // label rows of the CTA's rows: thread -> (row, class); the source rows' indices first (the labels depend on them)
  const int le = tid < HC_ROWS * C ? tid : 0, lrr = le / C, lc = le - lrr * C, lgr = crow0 + lrr;
  int ia = lgr, ib = lgr;
  float lg = 1.f;
  if (mapA) {
    ia = mapA[lgr];
    ib = mapB ? mapB[lgr] : ia;
    lg = gam ? gam[lgr] : 1.f;
  }
  float4 wreg[8];                                       // classifier weights: C * H <= 8192 floats over 256 threads
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int i = (tid + u * HC_T) * 4;
    wreg[u] = i < C * H ? *reinterpret_cast<const float4*>(Wc + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  hh_segment_consts(p, s, t0, t1, j == 0 && tile == t0, sMean, sInv, sGam, sBet, HC_T);
  const float lya = Y[(size_t)ia * C + lc], lyb = Y[(size_t)ib * C + lc];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int i = (tid + u * HC_T) * 4;
    if (i < C * H) *reinterpret_cast<float4*>(Ws + i) = wreg[u];
  }
  if (tid < HC_ROWS * C) {
    sYa[lrr][lc] = lya;
    sY[lrr][lc] = (ia == ib) ? lya : lg * lya + (1.f - lg) * lyb;
  }
  const int n_conf = (row0 < U) ? min(n_conf_val, U) : 0;
  __syncthreads();
  hh_stamp(p.dbg, 0, 1);
  // ---- activation, logits
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    a[i] = hh_act4(z[i], sMean, sInv, sGam, sBet, c, drop, hs_keep4_bits(kb[i]), drop_scale);
    *reinterpret_cast<float4*>(p.A + (size_t)r * H + c) = a[i];
  }
  {
    float* sp = scratch + (size_t)warp * C * 33;
    for (int c = 0; c < C; ++c) {
      float pp = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + (size_t)c * H + (i * 32 + lane) * 4);
        pp = fmaf(a[i].x, w.x, pp); pp = fmaf(a[i].y, w.y, pp); pp = fmaf(a[i].z, w.z, pp); pp = fmaf(a[i].w, w.w, pp);
      }
      sp[c * 33 + lane] = pp;
    }
    __syncwarp();
    if (lane < C) {                                     // lane -> class: sum of the 32 lanes' partials
      const float* q = sp + lane * 33;
      float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        z0 += q[i];
        z1 += q[i + 1];
        z2 += q[i + 2];
        z3 += q[i + 3];
      }
      sZ[warp][lane] = ((z0 + z1) + (z2 + z3)) + (bc ? bc[lane] : 0.f);
    }
  }
  __syncthreads();
  hh_stamp(p.dbg, 0, 2);
  // ---- losses and dLogits (a tile lies in one segment: all MSE rows or all CE rows)
  if (warp == 0) {                                      // four lanes per row: 8 rows = one warp
    float lt = 0.f, hit = 0.f;
    const int lr = lane >> 2, part = lane & 3;
    if (row0 < U) mse_row_group4(true, crow0 + lr < n_conf, part, sZ[lr], C, sY[lr], U, unsup_weight, sDz[lr], lt);
    else ce_row_group4(true, true, part, sZ[lr], C, sY[lr], sYa[lr], class_weight, (float)max(L, 1), 1.0f, 1.0f, sDz[lr], lt, hit);
    lt = warp_sum(lt);
    hit = warp_sum(hit);
    if (lane == 0) {
      if (lt != 0.f) atomicAdd(loss8 + (row0 < U ? 2 : 0), lt);
      if (hit != 0.f) atomicAdd(loss8 + 1, hit);
    }
  }
  __syncthreads();
  hh_stamp(p.dbg, 0, 3);
  if (lane < C) {
    logits[(size_t)r * C + lane] = sZ[warp][lane];
    dlogits[(size_t)r * C + lane] = sDz[warp][lane];
  }
  // ---- dA = dLogits Wc, dY = dA [A > 0] / (1-p)
  {
    float4 d0[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) d0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* dz0 = sDz[warp];
    for (int c = 0; c < C; ++c) {
      const float e0 = dz0[c];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + (size_t)c * H + (i * 32 + lane) * 4);
        d0[i].x = fmaf(e0, w.x, d0[i].x); d0[i].y = fmaf(e0, w.y, d0[i].y); d0[i].z = fmaf(e0, w.z, d0[i].z); d0[i].w = fmaf(e0, w.w, d0[i].w);
      }
    }
    hh_stamp(p.dbg, 0, 4);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 mu = *reinterpret_cast<const float4*>(sMean + c), is = *reinterpret_cast<const float4*>(sInv + c);
      const float4 da = d0[i], av = a[i], zz = z[i];
      float4 d, dz;
      d.x = av.x > 0.f ? da.x * drop_scale : 0.f;       // A > 0 implies the element was kept by the dropout
      d.y = av.y > 0.f ? da.y * drop_scale : 0.f;
      d.z = av.z > 0.f ? da.z * drop_scale : 0.f;
      d.w = av.w > 0.f ? da.w * drop_scale : 0.f;
      dz.x = d.x * ((zz.x - mu.x) * is.x);
      dz.y = d.y * ((zz.y - mu.y) * is.y);
      dz.z = d.z * ((zz.z - mu.z) * is.z);
      dz.w = d.w * ((zz.w - mu.w) * is.w);
      *reinterpret_cast<float4*>(p.dY + (size_t)r * H + c) = d;
      *reinterpret_cast<float4*>(sRed + ((size_t)warp * 2 + 0) * H + c) = d;
      *reinterpret_cast<float4*>(sRed + ((size_t)warp * 2 + 1) * H + c) = dz;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < 2 * H; idx += HC_T) {        // column sums over this CTA's rows, in row order
    const int which = idx >= H ? 1 : 0, c = idx - which * H;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < HC_ROWS; ++w) acc += sRed[((size_t)w * 2 + which) * H + c];
    sPart[idx] = acc;
  }
  hh_stamp(p.dbg, 0, 5);
  hh_cluster_sync();
  {                                                      // tile sums: CTA j adds its quarter of the columns over the cluster, rank order
    const int per = 2 * H / HC_CS;
    const uint32_t base = smem_u32(sPart);
    for (int idx = j * per + tid; idx < (j + 1) * per; idx += HC_T) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < HC_CS; ++c) acc += hh_ld_cluster(hh_map_rank(base + (uint32_t)idx * 4u, (uint32_t)c));
      const int which = idx >= H ? 1 : 0, col = idx - which * H;
      p.part_out[((size_t)tile * 2 + which) * H + col] = acc;
    }
  }
  hh_stamp(p.dbg, 0, 6);
  hh_cluster_sync();   // no CTA leaves while a peer may still read its sums
}

static int hh_fill(HsHeadArgs& a, const float* Z, float* A, const float* part_in, float* stats, const float* gamma, const float* beta,
                   const uint8_t* keep, float p_drop, float* dY, float* part_out, int H, int R, int n_seg, const int32_t* seg_off_host,
                   const char* who) {
  SCNB_CHECK_ARG(Z && A && stats && gamma && beta && dY && part_out, "%s: null argument", who);
  SCNB_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f && (!(p_drop > 0.f) || keep), "%s: dropout needs the keep bytes of the last application", who);
  SCNB_CHECK_ARG((((uintptr_t)Z | (uintptr_t)A | (uintptr_t)dY | (uintptr_t)keep) % 16) == 0, "%s: operands must be 16-byte aligned", who);
  a.Z = Z; a.A = A; a.part_in = part_in; a.stats = stats; a.gamma = gamma; a.beta = beta; a.keep = keep; a.p_drop = p_drop;
  a.dY = dY; a.part_out = part_out; a.H = H; a.R = R; a.dbg = hh_dbg_on();
  return hs_fill_seg(a.sg, n_seg, seg_off_host, R, who);
}

// Returns SCNB_ERR_UNSUPPORTED (-3) for shapes it is not built for (H not 128 or 256, C > 32): the caller keeps the unfused chain.
extern "C" int scnb_hs_classif_head(const float* Z, float* A, const float* part_in, float* stats, const float* gamma,
                                    const float* beta, const uint8_t* keep, float p_drop, int H, int R, int n_seg,
                                    const int32_t* seg_off_host, const float* Wc, const float* bc, int C, int U, int L,
                                    const int32_t* n_conf_dev, const float* Y, const int32_t* mapA, const int32_t* mapB,
                                    const float* gam, const float* class_weight, float unsup_weight, float* logits, float* dlogits,
                                    float* loss8, float* dY, float* part_out, void* stream) {
  SCNB_CHECK_ARG(Wc && Y && mapA && logits && dlogits && loss8 && C > 0 && U >= 0 && L >= 0, "hs_classif_head: bad arguments");
  if ((H != 128 && H != 256) || C > 32) return SCNB_ERR_UNSUPPORTED;
  HsHeadArgs a;
  int rc = hh_fill(a, Z, A, part_in, stats, gamma, beta, keep, p_drop, dY, part_out, H, R, n_seg, seg_off_host, "hs_classif_head");
  if (rc != SCNB_OK) return rc;
  const int nb = U + L;
  bool on_boundary = false;
  for (int i = 1; i <= n_seg; ++i) on_boundary = on_boundary || a.sg.off[i] == nb;
  SCNB_CHECK_ARG(nb > 0 && nb % HS_BM == 0 && on_boundary, "hs_classif_head: the MixMatch rows must end at a segment boundary");
  SCNB_CHECK_ARG(((uintptr_t)Wc % 16) == 0, "hs_classif_head: classifier weight must be 16-byte aligned");
  const size_t smem = ((size_t)C * H + (size_t)HC_ROWS * C * 33 + 4 * (size_t)H + (size_t)HC_ROWS * 2 * H + 2 * (size_t)H + 4 * HC_ROWS * 33) * sizeof(float);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(HC_CS, nb / HS_BM);
  cfg.blockDim = dim3(HC_T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = HC_CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e;
  if (H == 128) {
    rc = hs_set_smem(k_hs_classif_head<1>, smem, "hs_classif_head");
    if (rc != SCNB_OK) return rc;
    e = cudaLaunchKernelEx(&cfg, k_hs_classif_head<1>, a, Wc, bc, C, U, L, n_conf_dev, Y, mapA, mapB, gam, class_weight, unsup_weight, logits,
                           dlogits, loss8);
  } else {
    rc = hs_set_smem(k_hs_classif_head<2>, smem, "hs_classif_head");
    if (rc != SCNB_OK) return rc;
    e = cudaLaunchKernelEx(&cfg, k_hs_classif_head<2>, a, Wc, bc, C, U, L, n_conf_dev, Y, mapA, mapB, gam, class_weight, unsup_weight, logits,
                           dlogits, loss8);
  }
  if (e != cudaSuccess) {
    scnb_set_error("hs_classif_head: launch failed: %s", cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  SCNB_CHECK_LAUNCH("hs_classif_head");
  return SCNB_OK;
}

// =====================================================================================================================
// scnb_hs_dan_head (thread-block cluster of H/64 CTAs per 32-row tile)
// =====================================================================================================================
struct HsDanArgs {
  HsHeadArgs h;
  int row_base;            // first row of the domain-adversary rows in Z (multiple of 32)
  const float* W1;         // [H, H] hidden Linear of the adversary (out, in)
  const float* b1;         // [H]
  const float* W2;         // [D, H]
  const float* b2;         // [D]
  int D;
  const int32_t* n_rows_dev;
  const int32_t* dom_ids;
  const int32_t* src;
  int src_off;
  const float* loss_scale_dev;
  float grl_scale;         // -dan_weight
  float* Hd;               // [Rd, H]
  float* dHd;              // [Rd, H]
  float* dlog;             // [Rd, D]
  float* dlabel;           // [Rd, D]
  float* ddlog;            // [Rd, D]
  float* loss2;
  int Rd;
};

#define HD_DP (HS_BN + 4)      // pitch of the dHd slice [32][64]

// NJ = H / 128: a warp of the second product owns 8 * NJ output columns
template <int NJ>
__global__ void __launch_bounds__(HS_T) k_hs_dan_head(const __grid_constant__ HsDanArgs q) {
  extern __shared__ __align__(16) float hh_smem[];
  __shared__ __align__(8) unsigned long long bar_mem;
  const HsHeadArgs& p = q.h;
  constexpr int K = 128 * NJ;          // H
  constexpr int CS = K / HS_BN;        // cluster size
  constexpr int SA = K + 4;
  const int D = q.D;
  float* sA = hh_smem;                       // [32][K + 4] activation tile
  float* sB = sA + HS_BM * SA;               // [64][K + 4] rows [64 j, 64 j + 64) of W1
  float* sT = sB + HS_BN * SA;               // [4][32][72] partial tiles; tile 0 then holds Hd, tiles 0/1 the final (dY, dY zhat)
  float* sD = sT + HS_TILE_FLOATS;           // [32][68] dHd slice
  float* sP = sD + HS_BM * HD_DP;            // [32][K + 4] partial dA of this CTA's hidden slice (read by the cluster)
  float* sW2 = sP + HS_BM * SA;              // [D][64]
  float* sLg = sW2 + 32 * HS_BN;             // [32][32] partial logits (read by the cluster)
  float* sZs = sLg + 32 * 32;                // [32][33] logits (a lane per row in the loss: odd pitch)
  float* sDz = sZs + 32 * 33;                // [32][33] dLogits
  float* sYs = sDz + 32 * 33;                // [32][33] one-hot domain labels
  float* sMean = sYs + 32 * 33;              // (32 * 33 floats is a multiple of 16 bytes)
  float* sInv = sMean + K;
  float* sGam = sInv + K;
  float* sBet = sGam + K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x;                  // rank in the cluster == hidden / output column slice
  const int lt = blockIdx.y;                 // tile among the adversary rows
  const int tile = q.row_base / HS_BM + lt, row0 = tile * HS_BM;
  const int s = hs_seg_of(p.sg, row0);
  const int t0 = p.sg.off[s] / HS_BM, t1 = p.sg.off[s + 1] / HS_BM;
  const uint32_t bar = smem_u32(&bar_mem);
  hh_stamp(p.dbg, 1, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)((HS_BM + HS_BN) * K * sizeof(float)));
  hs_bulk_rows(sB, SA, q.W1 + (size_t)j * HS_BN * K, K, HS_BN, (uint32_t)(K * sizeof(float)), bar, HS_BM);
  hs_bulk_rows(sA, SA, p.Z + (size_t)row0 * K, K, HS_BM, (uint32_t)(K * sizeof(float)), bar, 0);
  // epilogue operands of this thread (row erow, columns 64 j + ec4 .. +3): pre-BatchNorm values for zhat
  const int erow = tid >> 4, ec4 = (tid & 15) << 2;
  const float4 zv = *reinterpret_cast<const float4*>(p.Z + (size_t)(row0 + erow) * K + j * HS_BN + ec4);
  // dropout keep bytes of the float4 groups this thread transforms, requested before anything waits
  constexpr int fpr = K >> 2, rpp = HS_T / fpr, NR = HS_BM / rpp;
  const int tr0 = tid / fpr, tc = (tid - tr0 * fpr) << 2;
  const bool drop = p.p_drop > 0.f;
  const int n_rows_val = q.n_rows_dev ? *q.n_rows_dev : q.Rd;       // device scalars of the loss, requested now
  const float ls = q.loss_scale_dev ? *q.loss_scale_dev : 1.f;
  uint32_t kbv[NR];
#pragma unroll
  for (int u = 0; u < NR; ++u) kbv[u] = drop ? hs_keep4_raw(p.keep, (size_t)(row0 + tr0 + u * rpp) * K + tc) : 0x01010101u;
  // threads [0, K) combine the statistics, the others stage this CTA's columns of W2
  for (int i = tid - K; i >= 0 && i < D * HS_BN; i += HS_T - K) {
    const int d = i >> 6, c = i & (HS_BN - 1);
    sW2[i] = q.W2[(size_t)d * K + j * HS_BN + c];
  }
  hh_segment_consts(p, s, t0, t1, j == 0 && tile == t0, sMean, sInv, sGam, sBet, HS_T);
  hh_stamp(p.dbg, 1, 1);
  mbar_wait(bar, 0);
  __syncthreads();
  hh_stamp(p.dbg, 1, 2);
  // BatchNorm -> Dropout -> ReLU in place on the A operand
  const float drop_scale = 1.0f / (1.0f - p.p_drop);
#pragma unroll
  for (int u = 0; u < NR; ++u) {
    const int r = tr0 + u * rpp;
    float* a = sA + r * SA + tc;
    const float4 y = hh_act4(*reinterpret_cast<const float4*>(a), sMean, sInv, sGam, sBet, tc, drop, hs_keep4_bits(kbv[u]), drop_scale);
    *reinterpret_cast<float4*>(a) = y;
    if (j == 0) *reinterpret_cast<float4*>(p.A + (size_t)(row0 + r) * K + tc) = y;
  }
  __syncthreads();
  hh_stamp(p.dbg, 1, 3);
  // ---- Hd[:, 64 j ..] = relu(A W1[64 j .., :]^T + b1)
  {
    float acc[2][2][4];
    hs_tile_mma<0>(sA, sB, K, acc);
    hs_tile_store_partials(sT, acc);     // (sT does not alias the weight panel here: the panel is used again below)
  }
  __syncthreads();
  hh_stamp(p.dbg, 1, 4);
  {
    float4 v = hs_tile_sum4(sT, erow, ec4);
    const float4 bb = *reinterpret_cast<const float4*>(q.b1 + j * HS_BN + ec4);
    v.x = fmaxf(v.x + bb.x, 0.f); v.y = fmaxf(v.y + bb.y, 0.f); v.z = fmaxf(v.z + bb.z, 0.f); v.w = fmaxf(v.w + bb.w, 0.f);
    *reinterpret_cast<float4*>(q.Hd + (size_t)(lt * HS_BM + erow) * K + j * HS_BN + ec4) = v;
    *reinterpret_cast<float4*>(sT + erow * HS_TP + ec4) = v;      // own elements only
    // partial domain logits over this CTA's 64 hidden columns: 16 lanes per row
    for (int d = 0; d < D; ++d) {
      const float4 w = *reinterpret_cast<const float4*>(sW2 + d * HS_BN + ec4);
      float pl = v.x * w.x;
      pl = fmaf(v.y, w.y, pl);
      pl = fmaf(v.z, w.z, pl);
      pl = fmaf(v.w, w.w, pl);
      pl += __shfl_xor_sync(0xffffffffu, pl, 8);
      pl += __shfl_xor_sync(0xffffffffu, pl, 4);
      pl += __shfl_xor_sync(0xffffffffu, pl, 2);
      pl += __shfl_xor_sync(0xffffffffu, pl, 1);
      if ((tid & 15) == 0) sLg[erow * 32 + d] = pl;
    }
  }
  hh_stamp(p.dbg, 1, 5);
  hh_cluster_sync();
  hh_stamp(p.dbg, 1, 6);
  // ---- full logits (every CTA of the cluster forms the same values: partials summed in rank order)
  {
    const uint32_t lg = smem_u32(sLg);
    for (int i = tid; i < HS_BM * D; i += HS_T) {
      const int r = i / D, d = i - r * D;
      float z = 0.f;
#pragma unroll
      for (int c = 0; c < CS; ++c) z += hh_ld_cluster(hh_map_rank(lg + (uint32_t)(r * 32 + d) * 4u, (uint32_t)c));
      sZs[r * 33 + d] = z + (q.b2 ? q.b2[d] : 0.f);
      const int gr = lt * HS_BM + r;
      const int id = q.dom_ids[q.src ? q.src[q.src_off + gr] : gr];
      sYs[r * 33 + d] = (d == id) ? 1.f : 0.f;
    }
  }
  __syncthreads();
  hh_stamp(p.dbg, 1, 7);
  // ---- CE against the domain label, one lane per row (every CTA of the cluster needs dLogits; rank 0 keeps the loss terms
  //      and writes the head's row outputs)
  if (warp == 0) {
    const int gr = lt * HS_BM + lane;
    const int n_act = min(n_rows_val, q.Rd);
    const float nn = (float)(q.n_rows_dev ? max(n_rows_val, 1) : 1);
    float lterm, hit;
    ce_row_thread(gr < n_act, sZs + lane * 33, D, sYs + lane * 33, sYs + lane * 33, nullptr, nn, 1.0f, ls, sDz + lane * 33, lterm, hit);
    lterm = warp_sum(lterm);
    hit = warp_sum(hit);
    if (j == 0 && lane == 0) {
      if (lterm != 0.f) atomicAdd(q.loss2, lterm);
      if (hit != 0.f) atomicAdd(q.loss2 + 1, hit);
    }
  }
  __syncthreads();
  hh_stamp(p.dbg, 1, 8);
  if (j == 0) {
    for (int i = tid; i < HS_BM * D; i += HS_T) {
      const int r = i / D, d = i - r * D;
      const size_t o = (size_t)(lt * HS_BM + r) * D + d;
      q.dlog[o] = sZs[r * 33 + d];
      q.dlabel[o] = sYs[r * 33 + d];
      q.ddlog[o] = sDz[r * 33 + d];
    }
  }
  // ---- dHd[:, 64 j ..] = (ddlog W2[:, 64 j ..]) [Hd > 0]
  {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int d = 0; d < D; ++d) {
      const float dz = sDz[erow * 33 + d];
      const float4 w = *reinterpret_cast<const float4*>(sW2 + d * HS_BN + ec4);
      g.x = fmaf(dz, w.x, g.x);
      g.y = fmaf(dz, w.y, g.y);
      g.z = fmaf(dz, w.z, g.z);
      g.w = fmaf(dz, w.w, g.w);
    }
    const float4 hd = *reinterpret_cast<const float4*>(sT + erow * HS_TP + ec4);
    if (!(hd.x > 0.f)) g.x = 0.f;
    if (!(hd.y > 0.f)) g.y = 0.f;
    if (!(hd.z > 0.f)) g.z = 0.f;
    if (!(hd.w > 0.f)) g.w = 0.f;
    *reinterpret_cast<float4*>(q.dHd + (size_t)(lt * HS_BM + erow) * K + j * HS_BN + ec4) = g;
    *reinterpret_cast<float4*>(sD + erow * HD_DP + ec4) = g;
  }
  __syncthreads();
  hh_stamp(p.dbg, 1, 9);
  // ---- partial dA[32, K] = dHd[:, 64 j ..] W1[64 j .., :]  (the same panel, now read as [k][n]); warp w -> columns [8 NJ w, ..)
  {
    const int g = lane >> 2, t = lane & 3;
    const int nc0 = warp * 8 * NJ;
    float acc[2][NJ][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][jj][c] = 0.f;
#pragma unroll 2
    for (int kk = 0; kk < HS_BN; kk += 8) {
      uint32_t ah[2][4], al[2][4], bh[NJ][2], bl[NJ][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        hs_split(sD[(16 * i + g) * HD_DP + kk + t], ah[i][0], al[i][0]);
        hs_split(sD[(16 * i + g + 8) * HD_DP + kk + t], ah[i][1], al[i][1]);
        hs_split(sD[(16 * i + g) * HD_DP + kk + t + 4], ah[i][2], al[i][2]);
        hs_split(sD[(16 * i + g + 8) * HD_DP + kk + t + 4], ah[i][3], al[i][3]);
      }
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        const int n = nc0 + 8 * jj + g;
        hs_split(sB[(kk + t) * SA + n], bh[jj][0], bl[jj][0]);
        hs_split(sB[(kk + t + 4) * SA + n], bh[jj][1], bl[jj][1]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          hs_mma(acc[i][jj], al[i], bh[jj]);
          hs_mma(acc[i][jj], ah[i], bl[jj]);
          hs_mma(acc[i][jj], ah[i], bh[jj]);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        const int c = nc0 + 8 * jj + 2 * t;
        *reinterpret_cast<float2*>(sP + (16 * i + g) * SA + c) = make_float2(acc[i][jj][0], acc[i][jj][1]);
        *reinterpret_cast<float2*>(sP + (16 * i + g + 8) * SA + c) = make_float2(acc[i][jj][2], acc[i][jj][3]);
      }
  }
  hh_stamp(p.dbg, 1, 10);
  hh_cluster_sync();
  hh_stamp(p.dbg, 1, 11);
  // ---- dA[:, 64 j ..] = grl_scale * sum over the cluster's partials (rank order); dY, tile sums
  {
    const uint32_t pa = smem_u32(sP + erow * SA + j * HS_BN + ec4);
    float4 da = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < CS; ++c) {
      const float4 u = hh_ld_cluster4(hh_map_rank(pa, (uint32_t)c));
      da.x += u.x; da.y += u.y; da.z += u.z; da.w += u.w;
    }
    da.x *= q.grl_scale; da.y *= q.grl_scale; da.z *= q.grl_scale; da.w *= q.grl_scale;
    const float4 av = *reinterpret_cast<const float4*>(sA + erow * SA + j * HS_BN + ec4);
    const float4 mu = *reinterpret_cast<const float4*>(sMean + j * HS_BN + ec4), is = *reinterpret_cast<const float4*>(sInv + j * HS_BN + ec4);
    float4 d, dz;
    d.x = av.x > 0.f ? da.x * drop_scale : 0.f;
    d.y = av.y > 0.f ? da.y * drop_scale : 0.f;
    d.z = av.z > 0.f ? da.z * drop_scale : 0.f;
    d.w = av.w > 0.f ? da.w * drop_scale : 0.f;
    dz.x = d.x * ((zv.x - mu.x) * is.x);
    dz.y = d.y * ((zv.y - mu.y) * is.y);
    dz.z = d.z * ((zv.z - mu.z) * is.z);
    dz.w = d.w * ((zv.w - mu.w) * is.w);
    *reinterpret_cast<float4*>(p.dY + (size_t)(row0 + erow) * K + j * HS_BN + ec4) = d;
    *reinterpret_cast<float4*>(sT + erow * HS_TP + ec4) = d;
    *reinterpret_cast<float4*>(sT + (size_t)HS_BM * HS_TP + erow * HS_TP + ec4) = dz;
  }
  __syncthreads();
  if (tid < 2 * HS_BN) {
    const int which = tid >> 6, c = tid & (HS_BN - 1);
    const float* src = sT + (size_t)which * HS_BM * HS_TP + c;
    float acc1 = 0.f;
#pragma unroll 8
    for (int i = 0; i < HS_BM; ++i) acc1 += src[i * HS_TP];
    p.part_out[((size_t)tile * 2 + which) * K + j * HS_BN + c] = acc1;
  }
  hh_stamp(p.dbg, 1, 12);
  hh_cluster_sync();
  hh_stamp(p.dbg, 1, 13);   // no CTA leaves while a peer may still read its partial tile
}

template <int NJ>
static int hh_launch_dan(const HsDanArgs& a, cudaStream_t st) {
  constexpr int K = 128 * NJ, CS = K / HS_BN;
  const size_t smem = ((size_t)(2 * HS_BM + HS_BN) * (K + 4) + HS_TILE_FLOATS + HS_BM * HD_DP + 32 * HS_BN + 32 * 32 + 3 * 32 * 33 + 4 * (size_t)K) * sizeof(float);
  int rc = hs_set_smem(k_hs_dan_head<NJ>, smem, "hs_dan_head");
  if (rc != SCNB_OK) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS, a.Rd / HS_BM);
  cfg.blockDim = dim3(HS_T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_hs_dan_head<NJ>, a);
  if (e != cudaSuccess) {
    scnb_set_error("hs_dan_head: launch failed: %s", cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  return SCNB_OK;
}

// Returns SCNB_ERR_UNSUPPORTED (-3) for shapes it is not built for (H not 128 or 256, D > 32).
extern "C" int scnb_hs_dan_head(const float* Z, float* A, const float* part_in, float* stats, const float* gamma, const float* beta,
                                const uint8_t* keep, float p_drop, int H, int R, int row_base, int n_seg, const int32_t* seg_off_host,
                                const float* W1, const float* b1, const float* W2, const float* b2, int D, const int32_t* n_rows_dev,
                                const int32_t* dom_ids, const int32_t* src, int src_off, const float* loss_scale_dev, float dan_weight,
                                float* Hd, float* dHd, float* dlog, float* dlabel, float* ddlog, float* loss2, float* dY,
                                float* part_out, void* stream) {
  SCNB_CHECK_ARG(W1 && b1 && W2 && dom_ids && Hd && dHd && dlog && dlabel && ddlog && loss2 && D > 0, "hs_dan_head: bad arguments");
  if ((H != 128 && H != 256) || D > 32) return SCNB_ERR_UNSUPPORTED;
  HsDanArgs a;
  int rc = hh_fill(a.h, Z, A, part_in, stats, gamma, beta, keep, p_drop, dY, part_out, H, R, n_seg, seg_off_host, "hs_dan_head");
  if (rc != SCNB_OK) return rc;
  SCNB_CHECK_ARG(row_base >= 0 && row_base < R && row_base % HS_BM == 0, "hs_dan_head: row_base must be a multiple of 32 inside the batch");
  bool on_boundary = false;
  for (int i = 0; i <= n_seg; ++i) on_boundary = on_boundary || a.h.sg.off[i] == row_base;
  SCNB_CHECK_ARG(on_boundary, "hs_dan_head: the adversary rows must start at a segment boundary");
  SCNB_CHECK_ARG((((uintptr_t)W1 | (uintptr_t)b1 | (uintptr_t)Hd | (uintptr_t)dHd) % 16) == 0, "hs_dan_head: operands must be 16-byte aligned");
  a.row_base = row_base; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2; a.D = D; a.n_rows_dev = n_rows_dev; a.dom_ids = dom_ids;
  a.src = src; a.src_off = src_off; a.loss_scale_dev = loss_scale_dev; a.grl_scale = -dan_weight; a.Hd = Hd; a.dHd = dHd;
  a.dlog = dlog; a.dlabel = dlabel; a.ddlog = ddlog; a.loss2 = loss2; a.Rd = R - row_base;
  rc = (H == 128) ? hh_launch_dan<1>(a, (cudaStream_t)stream) : hh_launch_dan<2>(a, (cudaStream_t)stream);
  return rc;
}

// =====================================================================================================================
// scnb_hs_bwd_in_unmix_planes: the input end of the backward in one launch -- BatchNorm backward of application 0
// (scnb_hs_bwd_in), the MixUp adjoint over the rows that used each base row (scnb_unmix_rows; dataprep.py:665-689 transposed)
// and the bf16 hi/lo operand planes of dZ1 for the tensor-core dW1 kernel (scnb_unmix_rows_planes).  dZ[0] is formed on the
// fly for the <= 3 rows every base row gathers (an elementwise function of dY[0], Z[0] and the segment constants) and never
// written.  grid = (ceil(n_base / 8), H / 128), 128 threads: a CTA owns 8 base rows (one 16-byte chunk of every plane row)
// and 128 columns; it also sums dZ[0] over R / gridDim.x rows of its own for the bias gradient of the first Linear.
// =====================================================================================================================
__device__ __forceinline__ float hh_sel3(int s, float a0, float a1, float a2) { return s == 0 ? a0 : (s == 1 ? a1 : a2); }

__global__ void __launch_bounds__(128) k_hs_bwd_in_unmix(const float* __restrict__ dY, const float* __restrict__ Z, int H, int R, HsSeg sg,
                                                         const float* __restrict__ stats, const float* __restrict__ part,
                                                         const float* __restrict__ bmean, const float* __restrict__ gamma,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ db,
                                                         const int32_t* __restrict__ inv, const float* __restrict__ coef, int n_base,
                                                         float* __restrict__ dZ1, uint8_t* __restrict__ pl_hi, uint8_t* __restrict__ pl_lo) {
  __shared__ int s_inv[8][3], s_seg[8][3];
  __shared__ float s_coef[8][3];
  const int r0 = blockIdx.x * 8;
  if (threadIdx.x < 24) {
    const int j = threadIdx.x / 3, k = threadIdx.x % 3, r = r0 + j;
    const int i = r < n_base ? inv[k * n_base + r] : -1;
    s_inv[j][k] = i;
    s_seg[j][k] = i >= 0 ? hs_seg_of(sg, i) : 0;
    s_coef[j][k] = r < n_base ? coef[k * n_base + r] : 0.f;
  }
  __syncthreads();
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= H) return;
  // segment constants of this column
  float mu[3], is[3], m1[3], m2[3], gs[3];
  const float gm = gamma[c];
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    mu[s] = is[s] = m1[s] = m2[s] = gs[s] = 0.f;
    if (s < sg.n_seg) {
      mu[s] = stats[((size_t)s * 3 + 0) * H + c];
      is[s] = stats[((size_t)s * 3 + 2) * H + c];
      if (part) {
        const int t0 = sg.off[s] / HS_BM, t1 = sg.off[s + 1] / HS_BM;
        float s1, s2;
        hs_combine_sums(part, H, c, t0, t1, s1, s2);
        const float inv_n = 1.0f / (float)((t1 - t0) * HS_BM);
        m1[s] = s1 * inv_n;
        m2[s] = s2 * inv_n;
        if (blockIdx.x == 0 && t1 > t0) {   // one CTA row: BatchNorm parameter gradients
          atomicAdd(dbeta + c, s1);
          atomicAdd(dgamma + c, s2);
        }
      } else {
        m1[s] = bmean[((size_t)s * 2 + 0) * H + c];
        m2[s] = bmean[((size_t)s * 2 + 1) * H + c];
      }
      gs[s] = gm * is[s];
    }
  }
  auto dz0 = [&](int i, int s) {
    const float dy = dY[(size_t)i * H + c];
    const float zh = (Z[(size_t)i * H + c] - hh_sel3(s, mu[0], mu[1], mu[2])) * hh_sel3(s, is[0], is[1], is[2]);
    return hh_sel3(s, gs[0], gs[1], gs[2]) * (dy - hh_sel3(s, m1[0], m1[1], m1[2]) - zh * hh_sel3(s, m2[0], m2[1], m2[2]));
  };
  if (db) {   // bias gradient of the Linear in front of application 0: column sums of dZ[0] over all rows
    const int rpc = (R + (int)gridDim.x - 1) / (int)gridDim.x;
    const int lo = blockIdx.x * rpc, hi = min(R, lo + rpc);
    float acc = 0.f;
    for (int i0 = lo; i0 < hi; i0 += 8) {      // eight rows' loads in flight (a rolled loop pays a round trip per row)
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = (i0 + u < hi) ? dz0(i0 + u, hs_seg_of(sg, i0 + u)) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += t[u];
    }
    if (hi > lo) atomicAdd(db + c, acc);
  }
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float x = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = s_inv[j][k];
      if (i >= 0) x = fmaf(s_coef[j][k], dz0(i, s_seg[j][k]), x);
    }
    v[j] = x;
    if (r0 + j < n_base) dZ1[(size_t)(r0 + j) * H + c] = x;
  }
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
  const size_t off = ((size_t)(r0 >> 5) * H * 32) * 2 + (size_t)c * 64 + swz_chunk<32>((uint32_t)(r0 & 31) >> 3, (uint32_t)c);
  *reinterpret_cast<uint4*>(pl_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  *reinterpret_cast<uint4*>(pl_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

extern "C" int scnb_hs_bwd_in_unmix_planes(const float* dY, const float* Z, int H, int R, int n_seg, const int32_t* seg_off_host,
                                           const float* stats, const float* part, const float* gamma, float* dgamma, float* dbeta,
                                           const float* bmean_in, float* db, const int32_t* inv3, const float* coef3, int n_base,
                                           float* dZ1, void* planes_hi, void* planes_lo, void* stream) {
  SCNB_CHECK_ARG(dY && Z && stats && (part || bmean_in) && gamma && (!part || (dgamma && dbeta)) && inv3 && coef3 && dZ1 && planes_hi &&
                     planes_lo, "hs_bwd_in_unmix_planes: null argument");
  SCNB_CHECK_ARG(H % 32 == 0 && R % HS_BM == 0 && n_base >= 0 && (((uintptr_t)planes_hi | (uintptr_t)planes_lo) % 16) == 0,
                 "hs_bwd_in_unmix_planes: H %% 32, rows %% 32 and 16-byte aligned planes");
  if (n_base == 0) return SCNB_OK;
  HsSeg sg;
  int rc = hs_fill_seg(sg, n_seg, seg_off_host, R, "hs_bwd_in_unmix_planes");
  if (rc != SCNB_OK) return rc;
  k_hs_bwd_in_unmix<<<dim3(scnb_cdiv(n_base, 8), scnb_cdiv(H, 128)), 128, 0, (cudaStream_t)stream>>>(
      dY, Z, H, R, sg, stats, part, bmean_in, gamma, dgamma, dbeta, db, inv3, coef3, n_base, dZ1, (uint8_t*)planes_hi, (uint8_t*)planes_lo);
  SCNB_CHECK_LAUNCH("hs_bwd_in_unmix_planes");
  return SCNB_OK;
}
