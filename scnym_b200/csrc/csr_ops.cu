// CSR-side kernels of the scNym hot path (sm_100a).
//   K0/K1  scnb_csr_row_offsets / scnb_csr_gather_rows : device row gather (+ log1p_drop augmentation on nnz only)
//   K2a    scnb_csr_linear_fwd{,_i64}                  : Z = X_csr * W1T + b1   (warp per row, W1T gene-major)
//   K8     scnb_csr_linear_bwd_dw                      : dW1T += X_csr^T * dZ   (vector atomics, v1)
// Reference semantics: scnym/dataprep.py:143-155 (row densify), :721-729 (log1p_drop),
// scnym/model.py:211 (first nn.Linear).  The dense [B,G] batch of the reference is never formed.
#include "optim.cuh"

// ---------------------------------------------------------------------------------------------
// Row offsets: b_indptr[i] = sum_{j<i} len(rows[j]);  single CTA, n up to a few 100k.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_row_offsets(const int64_t* __restrict__ indptr, const int64_t* __restrict__ rows,
                                                      int64_t row0, int n, int32_t* __restrict__ b_indptr,
                                                      const int32_t* __restrict__ carry_in) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = carry_in ? *carry_in : 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    int i = base + tid;
    int len = 0;
    if (i < n) {
      int64_t r = rows ? rows[i] : (row0 + i);
      len = (int)(indptr[r + 1] - indptr[r]);
    }
    int x = len;  // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;  // inclusive over warps
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + (wid ? warp_tot[wid - 1] : 0) + x - len;
    if (i < n) b_indptr[i] = excl;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (tid == 0) b_indptr[n] = carry_s;
}

extern "C" int scnb_csr_row_offsets(const int64_t* indptr, const int64_t* rows, int64_t row0, int n, int32_t* b_indptr,
                                    const int32_t* carry_in_dev, void* stream) {
  SCNB_CHECK_ARG(indptr && b_indptr && n >= 0, "csr_row_offsets: bad arguments");
  k_row_offsets<<<1, 1024, 0, (cudaStream_t)stream>>>(indptr, rows, row0, n, b_indptr, carry_in_dev);
  SCNB_CHECK_LAUNCH("csr_row_offsets");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Gather rows into a compact batch CSR; optional log1p_drop on the stored entries only
// (expm1(0)=0, dropout(0)=0, log1p(0)=0 => support is preserved).
// One 128-thread CTA per row (a 256+256-row training batch then covers the chip); rows [0,nA) come
// from source A, rows [nA, nA+nB) from source B, so the whole [U_raw ; L_aug] batch is one launch.
// keep_mask (if given) is aligned with the batch nnz order.  gene_cnt (if given) accumulates the
// per-gene number of non-zero entries written (histogram for scnb_csr_to_csc).
// ---------------------------------------------------------------------------------------------
struct GatherSrc {
  const int64_t* indptr;
  const int32_t* indices;
  const float* data;
  const int64_t* rows;  // NULL: row0 + i
  int64_t row0;
  int n;
  int augment;          // augmentation scheme of rows [aug_begin, aug_end) (row index local to this source), AUG_* below
  int aug_begin, aug_end;
  uint32_t stream_id;
};

#define GR_T 256  // threads per gathered row

// Augmentation schemes of scnym/dataprep.py:730-754 (AUGMENTATION_SCHEMES), all on the STORED entries only: every stage maps 0
// to 0 (expm1, Poisson sampling with its rate-0 guard, dropout, masking, log1p), so the support of a row is preserved.
//   1 log1p_drop          ExpMinusOne -> InputDropout(p) -> LibrarySizeNormalize(log1p)                 (dataprep.py:721-729)
//   2 log1p_mask          ExpMinusOne -> GeneMasking(p, p_apply) -> LibrarySizeNormalize(log1p)         (:405-465)
//   3 log1p_poisson       ExpMinusOne -> PoissonSample(depth 1) -> LibrarySizeNormalize(log1p)          (:277-341)
//   4 log1p_poisson_drop  ExpMinusOne -> PoissonSample(depth U(0.1, 2)) -> InputDropout(p) -> LibrarySizeNormalize(log1p)
//   5 count_poisson       PoissonSample(depth 1) on raw counts, no normalisation
#define AUG_NONE 0
#define AUG_DROP 1
#define AUG_MASK 2
#define AUG_POIS 3
#define AUG_POIS_DROP 4
#define AUG_COUNT_POIS 5

struct AugExtra {
  int n_genes;             // G (gene masking draws floor(G * p) genes per row, from ALL genes)
  float p_apply;           // GeneMasking: probability that the batch is masked at all
  const float* inj_pois;   // injected Poisson samples by batch nnz position (tests), or NULL
  const float* inj_row_r;  // injected per-row depth factors (tests), or NULL
  int ctr_off;             // added to the RNG step counter rng[1] (1 when the batch of the next step is assembled early)
};

__device__ __forceinline__ float u01(uint32_t w) { return ((float)(w >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)

// Poisson(lam) from a counter-based stream: entry `pos` of logical stream `strm` (re-evaluating it gives the same draw).
// lam < 10: inversion by sequential search; else PTRS, Hormann's transformed rejection with squeeze (the algorithm numpy
// and torch use for large rates); the rare exact acceptance test runs in double precision.
__device__ float poisson_draw(float lam, uint64_t pos, uint64_t seed, uint64_t strm) {
  if (!(lam > 0.f)) return 0.f;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  if (lam < 10.f) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)pos, (uint32_t)(pos >> 32), (uint32_t)strm, (uint32_t)(strm >> 32)), key);
    const float u = u01(r.x);
    float p = expf(-lam), F = p;
    int k = 0;
    while (u > F && k < 256) {
      ++k;
      p *= lam / (float)k;
      F += p;
    }
    return (float)k;
  }
  const float slam = sqrtf(lam), b = 0.931f + 2.53f * slam, a = -0.059f + 0.02483f * b;
  const float invalpha = 1.1239f + 1.1328f / (b - 3.4f), vr = 0.9277f - 3.6224f / (b - 2.f);
  uint4 r = make_uint4(0, 0, 0, 0);
  for (int trial = 0; trial < 64; ++trial) {
    if ((trial & 1) == 0) {
      const uint64_t sub = strm + ((uint64_t)(1 + (trial >> 1)) << 40);   // a fresh counter block every two trials
      r = philox4x32_10(make_uint4((uint32_t)pos, (uint32_t)(pos >> 32), (uint32_t)sub, (uint32_t)(sub >> 32)), key);
    }
    const float U = u01((trial & 1) ? r.z : r.x) - 0.5f, V = u01((trial & 1) ? r.w : r.y);
    const float us = 0.5f - fabsf(U);
    const float k = floorf((2.f * a / us + b) * U + lam + 0.43f);
    if (us >= 0.07f && V <= vr) return k;
    if (k < 0.f || (us < 0.013f && V > us)) continue;
    const double lhs = (double)logf(V) + (double)logf(invalpha) - (double)logf(a / (us * us) + b);
    const double rhs = -(double)lam + (double)k * log((double)lam) - lgamma((double)k + 1.0);
    if (lhs <= rhs) return k;
  }
  return floorf(lam + 0.5f);
}

__device__ __forceinline__ float gather_block_sum(float v, float* red8) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red8[threadIdx.x >> 5] = v;
  __syncthreads();
  const float t = ((red8[0] + red8[1]) + (red8[2] + red8[3])) + ((red8[4] + red8[5]) + (red8[6] + red8[7]));
  __syncthreads();
  return t;
}

#define GR_CACHE 2048  // entries of a row kept in shared memory between the two augmentation passes

// keep decision of the input dropout: injected mask by batch position, else one Philox call per
// four entries of a thread (entries tid + GR_T*k, k = 4q..4q+3 share counter `pos0 + tid + 4*GR_T*q`).
__device__ __forceinline__ bool gather_keep(const uint8_t* mask, int pos, const uint4& rnd, int k, float p) {
  if (mask) return mask[pos] != 0;
  const uint32_t w = (k & 3) == 0 ? rnd.x : (k & 3) == 1 ? rnd.y : (k & 3) == 2 ? rnd.z : rnd.w;
  return (float)(w >> 8) * (1.0f / 16777216.0f) >= p;
}

#define GR_MASKBITS 32768   // entries of a row whose gene-masking decisions are held in shared memory (exact sampling)

// EXT = false: schemes 0 / 1 only (the step's hot path: the Poisson sampler and the masking scan are compiled out)
// CACHE = false: no shared-memory row cache -- the second pass re-reads the row and recomputes its (counter-based) random
// stages.  32 bytes of shared memory per CTA: this form fits on an SM next to a kernel that holds nearly all of it (the
// dW1 + optimizer kernel), which is where the batch of the NEXT step is assembled when its rows arrive late (host-mapped
// matrices, scnb_csr_fetch_rows2).
template <bool EXT, bool CACHE = true>
__global__ void __launch_bounds__(GR_T) k_gather_rows(GatherSrc A, GatherSrc B, const int32_t* __restrict__ b_indptr,
                                                     int32_t* __restrict__ b_idx, float* __restrict__ b_val,
                                                     const uint8_t* __restrict__ keep_mask, const uint64_t* __restrict__ rng,
                                                     float p_drop, float target_sum, int32_t* __restrict__ gene_cnt, AugExtra X) {
  PDL_ENTRY();
  __shared__ float red8[8];
  constexpr int NCACHE = CACHE ? GR_CACHE : 0;
  __shared__ float e_s[CACHE ? GR_CACHE : 1];
  __shared__ int32_t c_s[CACHE ? GR_CACHE : 1];
  __shared__ uint32_t mbits[EXT ? GR_MASKBITS / 32 : 1];   // gene masking: bit j = entry j of the row is masked
  const int tid = threadIdx.x;
  const int i = blockIdx.x;
  const bool inA = i < A.n;
  const GatherSrc& S = inA ? A : B;
  const int li = inA ? i : i - A.n;
  const int64_t r = S.rows ? S.rows[li] : (S.row0 + li);
  const int64_t s = S.indptr[r];
  const int len = (int)(S.indptr[r + 1] - s);
  const int d = b_indptr[i];
  int mode = (S.augment && li >= S.aug_begin && li < S.aug_end) ? S.augment : AUG_NONE;
  if (!EXT && mode != AUG_NONE) mode = AUG_DROP;
  if (mode == AUG_NONE) {
    for (int j = tid; j < len; j += GR_T) {
      const int c = S.indices[s + j];
      const float v = S.data[s + j];
      b_idx[d + j] = c;
      b_val[d + j] = v;
      if (gene_cnt && v != 0.f) atomicAdd(gene_cnt + c, 1);
    }
    return;
  }
  const bool drop = !EXT || mode == AUG_DROP || mode == AUG_POIS_DROP;   // Bernoulli input dropout, scaled by 1/(1-p)
  const bool pois = EXT && (mode == AUG_POIS || mode == AUG_POIS_DROP || mode == AUG_COUNT_POIS);
  const float inv_keep = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  uint64_t seed = 0, strm = 0;
  if (rng) {
    seed = rng[0];
    strm = (rng[1] + (uint64_t)X.ctr_off) * 64ull + S.stream_id;   // ctr_off = 1: the batch of the NEXT step, assembled early
  }
  // per-row Poisson depth factor: U(0.1, 2.0) as `rand * (0.1 - 2.0) + 2.0` (dataprep.py:318-321); 1 otherwise
  float row_r = 1.0f;
  if (EXT && mode == AUG_POIS_DROP) {
    const float u = X.inj_row_r ? X.inj_row_r[i] : philox_uniform(seed, strm + 33, (uint64_t)i);
    row_r = u * (0.1f - 2.0f) + 2.0f;
  }
  // gene masking: with probability p_apply (one decision per batch, dataprep.py:439-442) every row loses floor(G*p) genes
  // drawn without replacement from ALL G genes; restricted to the row's stored entries that is Knuth's selection sampling
  // over the entries (entry t is taken with probability (m - taken) / (G - t)): exact, one pass of thread 0.
  bool mask_on = false;
  if (EXT && mode == AUG_MASK && !keep_mask) {
    mask_on = philox_uniform(seed, strm + 35, 0) <= X.p_apply;
    if (mask_on) {
      for (int w = tid; w < GR_MASKBITS / 32; w += GR_T) mbits[w] = 0u;
      __syncthreads();
      if (tid == 0) {
        const int G = X.n_genes;
        int left = (int)floorf((float)G * p_drop);
        const int n_scan = min(len, GR_MASKBITS);
        for (int t = 0; t < n_scan && left > 0; ++t) {
          const float u = philox_uniform(seed, strm + 34, (uint64_t)d + (uint64_t)t);
          if (u * (float)(G - t) < (float)left) {
            mbits[t >> 5] |= 1u << (t & 31);
            --left;
          }
        }
      }
      __syncthreads();
    }
  }
  // value of entry j after every random stage and before the library-size normalisation
  auto entry = [&](int j, int k, uint4& rnd) -> float {
    const float x = S.data[s + j];
    float e = (EXT && mode == AUG_COUNT_POIS) ? x : expm1f(x);
    if (pois) e = X.inj_pois ? (x == 0.f ? 0.f : X.inj_pois[d + j]) : (e > 0.f ? poisson_draw(e * row_r, (uint64_t)(d + j), seed, strm + 32) : 0.f);
    // (explicitly rounded product / sum below: both passes of the cache-free form, and the cached form, must produce the
    //  same bits whatever the compiler would otherwise contract into an FMA)
    if (drop) e = __fmul_rn(e, gather_keep(keep_mask, d + j, rnd, k, p_drop) ? inv_keep : 0.f);
    if (EXT && mode == AUG_MASK) {
      if (keep_mask) e = keep_mask[d + j] ? e : 0.f;
      else if (mask_on) {
        const bool m = j < GR_MASKBITS ? ((mbits[j >> 5] >> (j & 31)) & 1u) : (philox_uniform(seed, strm + 34, (uint64_t)d + (uint64_t)j) < p_drop);
        if (m) e = 0.f;
      }
    }
    return e;
  };
  // pass 1: one trip to DRAM for values and indices; the transformed value is cached in shared memory
  float sum = 0.f;
  uint4 rnd = make_uint4(0, 0, 0, 0);
  for (int j = tid, k = 0; j < len; j += GR_T, ++k) {
    if (drop && !keep_mask && (k & 3) == 0) {
      const uint64_t ctr = (uint64_t)(d + tid) + (uint64_t)(4 * GR_T) * (uint64_t)(k >> 2);
      rnd = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)strm, (uint32_t)(strm >> 32)),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    }
    const float e = entry(j, k, rnd);
    sum = __fadd_rn(sum, e);
    if (j < NCACHE) {
      e_s[j] = e;
      c_s[j] = S.indices[s + j];
    }
  }
  sum = gather_block_sum(sum, red8);   // (also orders the shared-memory cache writes before the reads below)
  for (int j = tid, k = 0; j < len; j += GR_T, ++k) {
    float e;
    int c;
    if (j < NCACHE) {
      e = e_s[j];
      c = c_s[j];
    } else {  // rows longer than the cache: recompute (every random stage is counter-based: the same draw again)
      if (drop && !keep_mask && ((k & 3) == 0 || j - GR_T < NCACHE)) {
        const uint64_t ctr = (uint64_t)(d + tid) + (uint64_t)(4 * GR_T) * (uint64_t)(k >> 2);
        rnd = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)strm, (uint32_t)(strm >> 32)),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      }
      e = entry(j, k, rnd);
      c = S.indices[s + j];
    }
    // count_poisson returns the sampled counts; every other scheme ends in LibrarySizeNormalize(log1p)
    // (NaN if every entry of the row was dropped, as in the reference)
    const float v = (EXT && mode == AUG_COUNT_POIS) ? e : log1pf((e / sum) * target_sum);
    b_idx[d + j] = c;
    b_val[d + j] = v;
    if (gene_cnt && v != 0.f) atomicAdd(gene_cnt + c, 1);
  }
}

extern "C" int scnb_csr_gather_rows(const int64_t* indptr, const int32_t* indices, const float* data, const int64_t* rows,
                                    int64_t row0, int n, const int32_t* b_indptr, int32_t* b_idx, float* b_val, int augment,
                                    int aug_row_begin, int aug_row_end, const uint8_t* keep_mask, const uint64_t* rng,
                                    uint32_t stream_id, float p_drop, int32_t* gene_cnt, void* stream) {
  SCNB_CHECK_ARG(indptr && indices && data && b_indptr && b_idx && b_val && n >= 0, "csr_gather_rows: bad arguments");
  SCNB_CHECK_ARG(!augment || keep_mask || rng, "csr_gather_rows: augmentation needs keep_mask or rng state");
  if (n == 0) return SCNB_OK;
  GatherSrc A = {indptr, indices, data, rows, row0, n, augment, aug_row_begin, aug_row_end, stream_id};
  GatherSrc B = {nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, 0};
  k_gather_rows<false><<<n, GR_T, 0, (cudaStream_t)stream>>>(A, B, b_indptr, b_idx, b_val, keep_mask, rng, p_drop, 1e6f, gene_cnt, AugExtra{0, 0.f, nullptr, nullptr, 0});
  SCNB_CHECK_LAUNCH("csr_gather_rows");
  return SCNB_OK;
}

// Two sources in one launch: rows [0,nA) raw or augmented from A, rows [nA,nA+nB) from B
// (the MixMatch batch [U_raw ; L_aug]: A = unlabeled, B = labeled with augment_B = 1).
extern "C" int scnb_csr_gather_rows2(const int64_t* indptr_A, const int32_t* indices_A, const float* data_A,
                                     const int64_t* rows_A, int nA, int augment_A, uint32_t stream_id_A,
                                     const int64_t* indptr_B, const int32_t* indices_B, const float* data_B,
                                     const int64_t* rows_B, int nB, int augment_B, uint32_t stream_id_B,
                                     const int32_t* b_indptr, int32_t* b_idx, float* b_val, const uint8_t* keep_mask,
                                     const uint64_t* rng, float p_drop, int32_t* gene_cnt, int ctr_off, void* stream) {
  SCNB_CHECK_ARG(b_indptr && b_idx && b_val && nA >= 0 && nB >= 0, "csr_gather_rows2: bad arguments");
  SCNB_CHECK_ARG(nA == 0 || (indptr_A && indices_A && data_A), "csr_gather_rows2: source A missing");
  SCNB_CHECK_ARG(nB == 0 || (indptr_B && indices_B && data_B), "csr_gather_rows2: source B missing");
  SCNB_CHECK_ARG(!(augment_A || augment_B) || keep_mask || rng, "csr_gather_rows2: augmentation needs keep_mask or rng state");
  if (nA + nB == 0) return SCNB_OK;
  GatherSrc A = {indptr_A, indices_A, data_A, rows_A, 0, nA, augment_A, 0, nA, stream_id_A};
  GatherSrc B = {indptr_B, indices_B, data_B, rows_B, 0, nB, augment_B, 0, nB, stream_id_B};
  scnb_launch_pdl(k_gather_rows<false>, dim3(nA + nB), dim3(GR_T), 0, (cudaStream_t)stream, A, B, b_indptr, b_idx, b_val, keep_mask, rng, p_drop, 1e6f,
                  gene_cnt, AugExtra{0, 0.f, nullptr, nullptr, ctr_off});
  SCNB_CHECK_LAUNCH("csr_gather_rows2");
  return SCNB_OK;
}

// Sampled rows of two HOST-resident matrices (page-locked + mapped, scnb_host_register_mapped) -> a device staging CSR laid
// out like the batch (rows [A ; B], offsets = b_indptr, also written as the int64 `st_indptr` the gather expects), verbatim.
// The reads travel over PCIe (~2 us latency, ~45 GB/s): a persistent grid of at most one light CTA per SM -- no shared
// memory, 8 loads in flight per thread, ~1 MB outstanding in total -- so that the step's own kernels keep their SM
// resources while the upload of the NEXT step's rows is in flight (the augmenting gather itself, with its shared-memory
// row cache, then runs on the staged rows in HBM).  Replaces the host-side minibatch assembly + `.to(device)` of
// scnym/dataprep.py:114-168 / trainer.py:589-599.
__global__ void __launch_bounds__(256) k_fetch_rows(GatherSrc A, GatherSrc B, const int32_t* __restrict__ b_indptr,
                                                    int64_t* __restrict__ st_indptr, int32_t* __restrict__ st_idx,
                                                    float* __restrict__ st_val) {
  const int tid = threadIdx.x, n = A.n + B.n;
  if (blockIdx.x == 0)
    for (int i = tid; i <= n; i += 256) st_indptr[i] = (int64_t)b_indptr[i];
  auto start_of = [&](int i) -> int64_t {
    const bool inA = i < A.n;
    const GatherSrc& S = inA ? A : B;
    const int li = inA ? i : i - A.n;
    return S.indptr[S.rows ? S.rows[li] : (S.row0 + li)];
  };
  int i = blockIdx.x;
  int64_t s = i < n ? start_of(i) : 0;
  for (; i < n; i += gridDim.x) {
    const int nx = i + gridDim.x;
    const int64_t s_next = nx < n ? start_of(nx) : 0;     // the next row's offset (a host read) is in flight during this row
    const bool inA = i < A.n;
    const int32_t* __restrict__ ix = (inA ? A.indices : B.indices) + s;
    const float* __restrict__ vv = (inA ? A.data : B.data) + s;
    const int d = b_indptr[i], len = b_indptr[i + 1] - d;
    for (int j0 = 0; j0 < len; j0 += 4 * 256) {
      int c[4];
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * 256 + tid;
        if (j < len) {
          c[u] = ix[j];
          v[u] = vv[j];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * 256 + tid;
        if (j < len) {
          st_idx[d + j] = c[u];
          st_val[d + j] = v[u];
        }
      }
    }
    s = s_next;
  }
}

extern "C" int scnb_csr_fetch_rows2(const int64_t* indptr_A, const int32_t* indices_A, const float* data_A, const int64_t* rows_A,
                                    int nA, const int64_t* indptr_B, const int32_t* indices_B, const float* data_B,
                                    const int64_t* rows_B, int nB, const int32_t* b_indptr, int64_t* st_indptr, int32_t* st_idx,
                                    float* st_val, int max_ctas, void* stream) {
  SCNB_CHECK_ARG(b_indptr && st_indptr && st_idx && st_val && nA >= 0 && nB >= 0, "csr_fetch_rows2: bad arguments");
  SCNB_CHECK_ARG(nA == 0 || (indptr_A && indices_A && data_A), "csr_fetch_rows2: source A missing");
  SCNB_CHECK_ARG(nB == 0 || (indptr_B && indices_B && data_B), "csr_fetch_rows2: source B missing");
  if (nA + nB == 0) return SCNB_OK;
  GatherSrc A = {indptr_A, indices_A, data_A, rows_A, 0, nA, 0, 0, 0, 0};
  GatherSrc B = {indptr_B, indices_B, data_B, rows_B, 0, nB, 0, 0, 0, 0};
  int grid = max_ctas > 0 ? max_ctas : 148;
  if (grid > nA + nB) grid = nA + nB;
  k_fetch_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, b_indptr, st_indptr, st_idx, st_val);
  SCNB_CHECK_LAUNCH("csr_fetch_rows2");
  return SCNB_OK;
}

// scnb_csr_gather_rows2 without the shared-memory row cache (see k_gather_rows<.., CACHE = false>): same arguments, same
// results bit for bit.
extern "C" int scnb_csr_gather_rows2_nocache(const int64_t* indptr_A, const int32_t* indices_A, const float* data_A,
                                             const int64_t* rows_A, int nA, int augment_A, uint32_t stream_id_A,
                                             const int64_t* indptr_B, const int32_t* indices_B, const float* data_B,
                                             const int64_t* rows_B, int nB, int augment_B, uint32_t stream_id_B,
                                             const int32_t* b_indptr, int32_t* b_idx, float* b_val, const uint8_t* keep_mask,
                                             const uint64_t* rng, float p_drop, int32_t* gene_cnt, int ctr_off, void* stream) {
  SCNB_CHECK_ARG(b_indptr && b_idx && b_val && nA >= 0 && nB >= 0, "csr_gather_rows2_nocache: bad arguments");
  SCNB_CHECK_ARG(nA == 0 || (indptr_A && indices_A && data_A), "csr_gather_rows2_nocache: source A missing");
  SCNB_CHECK_ARG(nB == 0 || (indptr_B && indices_B && data_B), "csr_gather_rows2_nocache: source B missing");
  SCNB_CHECK_ARG(!(augment_A || augment_B) || keep_mask || rng, "csr_gather_rows2_nocache: augmentation needs keep_mask or rng state");
  if (nA + nB == 0) return SCNB_OK;
  GatherSrc A = {indptr_A, indices_A, data_A, rows_A, 0, nA, augment_A, 0, nA, stream_id_A};
  GatherSrc B = {indptr_B, indices_B, data_B, rows_B, 0, nB, augment_B, 0, nB, stream_id_B};
  k_gather_rows<false, false><<<nA + nB, GR_T, 0, (cudaStream_t)stream>>>(A, B, b_indptr, b_idx, b_val, keep_mask, rng, p_drop, 1e6f,
                                                                          gene_cnt, AugExtra{0, 0.f, nullptr, nullptr, ctr_off});
  SCNB_CHECK_LAUNCH("csr_gather_rows2_nocache");
  return SCNB_OK;
}

// The other augmentation schemes of scnym/dataprep.py:730-754 (`scheme`: 2 log1p_mask, 3 log1p_poisson, 4 log1p_poisson_drop,
// 5 count_poisson; 1 = log1p_drop is also accepted) on rows [0, n) of one source.  Randomness: `keep_mask` (Bernoulli dropout
// keeps of scheme 1 / 4, or the per-entry gene-masking keeps of scheme 2 with the batch-level apply decision folded in),
// `inj_poisson` (sampled values by batch nnz position) and `inj_row_depth` (per-row uniforms of the depth factor) inject the
// reference's draws in tests; NULL: counter-based Philox streams of (rng[0], rng[1], stream_id).
extern "C" int scnb_csr_gather_rows_aug(const int64_t* indptr, const int32_t* indices, const float* data, const int64_t* rows,
                                        int64_t row0, int n, const int32_t* b_indptr, int32_t* b_idx, float* b_val, int scheme,
                                        int n_genes, float p_drop, float p_apply, const uint8_t* keep_mask, const float* inj_poisson,
                                        const float* inj_row_depth, const uint64_t* rng, uint32_t stream_id, int32_t* gene_cnt,
                                        int ctr_off, void* stream) {
  SCNB_CHECK_ARG(indptr && indices && data && b_indptr && b_idx && b_val && n >= 0, "csr_gather_rows_aug: bad arguments");
  SCNB_CHECK_ARG(scheme >= AUG_DROP && scheme <= AUG_COUNT_POIS, "csr_gather_rows_aug: unknown augmentation scheme %d", scheme);
  SCNB_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f && p_apply >= 0.f && p_apply <= 1.f && n_genes > 0, "csr_gather_rows_aug: bad parameters");
  const bool need_keep = scheme == AUG_DROP || scheme == AUG_POIS_DROP || scheme == AUG_MASK;
  const bool need_pois = scheme == AUG_POIS || scheme == AUG_POIS_DROP || scheme == AUG_COUNT_POIS;
  SCNB_CHECK_ARG(rng || ((!need_keep || keep_mask) && (!need_pois || inj_poisson) && (scheme != AUG_POIS_DROP || inj_row_depth)),
                 "csr_gather_rows_aug: needs rng state or every injected draw of the scheme");
  if (n == 0) return SCNB_OK;
  GatherSrc A = {indptr, indices, data, rows, row0, n, scheme, 0, n, stream_id};
  GatherSrc B = {nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, 0};
  k_gather_rows<true><<<n, GR_T, 0, (cudaStream_t)stream>>>(A, B, b_indptr, b_idx, b_val, keep_mask, rng, p_drop, 1e6f, gene_cnt,
                                                            AugExtra{n_genes, p_apply, inj_poisson, inj_row_depth, ctr_off});
  SCNB_CHECK_LAUNCH("csr_gather_rows_aug");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// K2a  Z[n,H0] = X_csr[n,G] * W1T[G,H0] + b1.   WPR warps per row (the row's nnz are split in WPR
// contiguous chunks; partial sums meet in shared memory and are added in a fixed order, so the
// result does not depend on timing).  A lane owns NV float4 column groups (h = (i*32+lane)*4); the
// (col,val) stream is loaded coalesced 32 at a time and broadcast by shuffle; W1T rows are read as
// 512-byte coalesced segments.  Small batches (train: 512 rows) use WPR=8 so that the launch covers
// every SM; large row ranges (predict) use WPR=1, where the 8 rows of a CTA walk the gene axis
// together and share W1T rows through L1.
// ---------------------------------------------------------------------------------------------
template <int NV, int WPR, typename IdxT>
__global__ void __launch_bounds__(256) k_csr_linear_fwd(const IdxT* __restrict__ indptr, const int32_t* __restrict__ idx,
                                                        const float* __restrict__ val, int n, const float4* __restrict__ W1T,
                                                        const float4* __restrict__ b1, int H0v, float4* __restrict__ Z) {
  PDL_ENTRY();
  constexpr int RPC = 8 / WPR;  // rows per CTA
  __shared__ float4 part[WPR > 1 ? 8 : 1][NV * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * RPC + warp / WPR;
  const int sub = warp % WPR;
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < n) {
    long long s = (long long)indptr[row], e = (long long)indptr[row + 1];
    if (WPR > 1) {
      const long long chunk = (((e - s) + WPR - 1) / WPR + 31) & ~31ll;  // whole 32-entry groups per warp
      s = s + sub * chunk;
      e = min(e, s + chunk);
    }
    for (long long base = s; base < e; base += 32) {
      const int cnt = (int)min((long long)32, e - base);
      int c = 0;
      float v = 0.f;
      if (lane < cnt) {
        c = idx[base + lane];
        v = val[base + lane];
      }
#pragma unroll 8
      for (int k = 0; k < cnt; ++k) {
        const int ck = __shfl_sync(0xffffffffu, c, k);
        const float vk = __shfl_sync(0xffffffffu, v, k);
        const float4* wrow = W1T + (size_t)ck * H0v + lane;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 w = __ldg(wrow + i * 32);
          acc[i].x = fmaf(vk, w.x, acc[i].x);
          acc[i].y = fmaf(vk, w.y, acc[i].y);
          acc[i].z = fmaf(vk, w.z, acc[i].z);
          acc[i].w = fmaf(vk, w.w, acc[i].w);
        }
      }
    }
  }
  if (WPR > 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) part[warp][i * 32 + lane] = acc[i];
    __syncthreads();
    if (sub != 0 || row >= n) return;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 t = acc[i];
#pragma unroll
      for (int w = 1; w < WPR; ++w) {
        const float4 o = part[warp + w][i * 32 + lane];
        t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
      }
      acc[i] = t;
    }
  } else if (row >= n) {
    return;
  }
  float4* zrow = Z + (size_t)row * H0v + lane;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 b = b1 ? __ldg(b1 + i * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    zrow[i * 32] = make_float4(acc[i].x + b.x, acc[i].y + b.y, acc[i].z + b.z, acc[i].w + b.w);
  }
}

// Generic fallback for any H0: lane owns columns h = lane + 32*i, processed in chunks of 8.
template <typename IdxT>
__global__ void __launch_bounds__(256) k_csr_linear_fwd_generic(const IdxT* __restrict__ indptr, const int32_t* __restrict__ idx,
                                                                const float* __restrict__ val, int n,
                                                                const float* __restrict__ W1T, const float* __restrict__ b1,
                                                                int H0, float* __restrict__ Z) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const long long s = (long long)indptr[row], e = (long long)indptr[row + 1];
  for (int h0 = 0; h0 < H0; h0 += 256) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (long long base = s; base < e; base += 32) {
      const int cnt = (int)min((long long)32, e - base);
      int c = 0;
      float v = 0.f;
      if (lane < cnt) {
        c = idx[base + lane];
        v = val[base + lane];
      }
      for (int k = 0; k < cnt; ++k) {
        const int ck = __shfl_sync(0xffffffffu, c, k);
        const float vk = __shfl_sync(0xffffffffu, v, k);
        const float* wrow = W1T + (size_t)ck * H0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int h = h0 + lane + 32 * i;
          if (h < H0) acc[i] = fmaf(vk, __ldg(wrow + h), acc[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int h = h0 + lane + 32 * i;
      if (h < H0) Z[(size_t)row * H0 + h] = acc[i] + (b1 ? b1[h] : 0.f);
    }
  }
}

template <int NV, typename IdxT>
static void launch_csr_fwd_nv(int wpr, const IdxT* indptr, const int32_t* idx, const float* val, int n, const float* W1T,
                              const float* b1, int H0, float* Z, cudaStream_t st) {
  const float4* W = (const float4*)W1T;
  const float4* B = (const float4*)b1;
  float4* Zv = (float4*)Z;
  if (wpr == 8)
    scnb_launch_pdl(k_csr_linear_fwd<NV, 8, IdxT>, dim3(n), dim3(256), 0, st, indptr, idx, val, n, W, B, H0 / 4, Zv);
  else if (wpr == 4)
    scnb_launch_pdl(k_csr_linear_fwd<NV, 4, IdxT>, dim3(scnb_cdiv(n, 2)), dim3(256), 0, st, indptr, idx, val, n, W, B, H0 / 4, Zv);
  else if (wpr == 2)
    scnb_launch_pdl(k_csr_linear_fwd<NV, 2, IdxT>, dim3(scnb_cdiv(n, 4)), dim3(256), 0, st, indptr, idx, val, n, W, B, H0 / 4, Zv);
  else
    scnb_launch_pdl(k_csr_linear_fwd<NV, 1, IdxT>, dim3(scnb_cdiv(n, 8)), dim3(256), 0, st, indptr, idx, val, n, W, B, H0 / 4, Zv);
}

template <typename IdxT>
static int launch_csr_linear_fwd(const IdxT* indptr, const int32_t* idx, const float* val, int n, const float* W1T,
                                 const float* b1, int H0, float* Z, cudaStream_t st) {
  if (n == 0) return SCNB_OK;
  const bool vec = (H0 % 128 == 0) && (((uintptr_t)W1T | (uintptr_t)Z | (uintptr_t)b1) % 16 == 0);
  const int nv = H0 / 128;
  // warps per row: enough CTAs (8 warps each) to put >= 2 on every one of the 148 SMs
  int wpr = 1;
  while (wpr < 8 && (long long)n * wpr < 8ll * 2 * 148) wpr *= 2;
  if (vec && nv == 1)
    launch_csr_fwd_nv<1, IdxT>(wpr, indptr, idx, val, n, W1T, b1, H0, Z, st);
  else if (vec && nv == 2)
    launch_csr_fwd_nv<2, IdxT>(wpr, indptr, idx, val, n, W1T, b1, H0, Z, st);
  else if (vec && nv == 4)
    launch_csr_fwd_nv<4, IdxT>(wpr, indptr, idx, val, n, W1T, b1, H0, Z, st);
  else if (vec && nv == 8)
    launch_csr_fwd_nv<8, IdxT>(wpr, indptr, idx, val, n, W1T, b1, H0, Z, st);
  else
    k_csr_linear_fwd_generic<IdxT><<<scnb_cdiv(n, 8), 256, 0, st>>>(indptr, idx, val, n, W1T, b1, H0, Z);
  SCNB_CHECK_LAUNCH("csr_linear_fwd");
  return SCNB_OK;
}

extern "C" int scnb_csr_linear_fwd(const int32_t* b_indptr, const int32_t* b_idx, const float* b_val, int n, const float* W1T,
                                   const float* b1, int H0, float* Z, void* stream) {
  SCNB_CHECK_ARG(b_indptr && b_idx && b_val && W1T && Z && n >= 0 && H0 > 0, "csr_linear_fwd: bad arguments");
  return launch_csr_linear_fwd<int32_t>(b_indptr, b_idx, b_val, n, W1T, b1, H0, Z, (cudaStream_t)stream);
}

// Dataset-CSR variant (int64 indptr, contiguous row range starting at indptr[0]); used by predict.
extern "C" int scnb_csr_linear_fwd_i64(const int64_t* indptr, const int32_t* indices, const float* data, int n,
                                       const float* W1T, const float* b1, int H0, float* Z, void* stream) {
  SCNB_CHECK_ARG(indptr && indices && data && W1T && Z && n >= 0 && H0 > 0, "csr_linear_fwd_i64: bad arguments");
  return launch_csr_linear_fwd<int64_t>(indptr, indices, data, n, W1T, b1, H0, Z, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// K8 (v1)  dW1T[g,:] += x[r,g] * dZ[r,:]  with 128-bit vector atomics (red.global.add.v4.f32).
// Caller zeroes dW1T.  Summation order is not deterministic (fp32 atomics).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_csr_linear_bwd_dw_v4(const int32_t* __restrict__ indptr, const int32_t* __restrict__ idx,
                                                              const float* __restrict__ val, int n,
                                                              const float4* __restrict__ dZ, int H0v, float4* __restrict__ dW1T) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int s = indptr[row], e = indptr[row + 1];
  // hv0 loop bounds are warp-uniform so that every lane takes part in the shuffles
  for (int hv0 = 0; hv0 < H0v; hv0 += 32) {
    const int hv = hv0 + lane;
    const bool ok = hv < H0v;
    const float4 g = ok ? dZ[(size_t)row * H0v + hv] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = s; base < e; base += 32) {
      const int cnt = min(32, e - base);
      int c = 0;
      float v = 0.f;
      if (lane < cnt) {
        c = idx[base + lane];
        v = val[base + lane];
      }
      for (int k = 0; k < cnt; ++k) {
        const int ck = __shfl_sync(0xffffffffu, c, k);
        const float vk = __shfl_sync(0xffffffffu, v, k);
        if (ok && vk != 0.f) atomicAdd(dW1T + (size_t)ck * H0v + hv, make_float4(vk * g.x, vk * g.y, vk * g.z, vk * g.w));
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_csr_linear_bwd_dw_scalar(const int32_t* __restrict__ indptr, const int32_t* __restrict__ idx,
                                                                  const float* __restrict__ val, int n,
                                                                  const float* __restrict__ dZ, int H0, float* __restrict__ dW1T) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int s = indptr[row], e = indptr[row + 1];
  for (int h = lane; h < H0; h += 32) {
    const float g = dZ[(size_t)row * H0 + h];
    for (int j = s; j < e; ++j) {
      const float v = val[j];
      if (v != 0.f) atomicAdd(dW1T + (size_t)idx[j] * H0 + h, v * g);
    }
  }
}

extern "C" int scnb_csr_linear_bwd_dw(const int32_t* b_indptr, const int32_t* b_idx, const float* b_val, int n,
                                      const float* dZ, int H0, float* dW1T, void* stream) {
  SCNB_CHECK_ARG(b_indptr && b_idx && b_val && dZ && dW1T && n >= 0 && H0 > 0, "csr_linear_bwd_dw: bad arguments");
  if (n == 0) return SCNB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (H0 % 4 == 0 && (((uintptr_t)dZ | (uintptr_t)dW1T) % 16 == 0))
    k_csr_linear_bwd_dw_v4<<<scnb_cdiv(n, 8), 256, 0, st>>>(b_indptr, b_idx, b_val, n, (const float4*)dZ, H0 / 4, (float4*)dW1T);
  else
    k_csr_linear_bwd_dw_scalar<<<scnb_cdiv(n, 8), 256, 0, st>>>(b_indptr, b_idx, b_val, n, dZ, H0, dW1T);
  SCNB_CHECK_LAUNCH("csr_linear_bwd_dw");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Batch CSR -> gene-major copy ("batch CSC").  gene_cnt[G] holds the per-gene number of non-zero
// stored entries (accumulated by scnb_csr_gather_rows).  scan: c_ptr = exclusive scan, cursor =
// c_ptr, gene_cnt <- 0 (ready for the next step).  fill: entry (row, value) of every stored
// non-zero goes to c_ent[cursor[gene]++].  The order of a gene's entries is the arrival order of the
// atomics (fp32 summation order of dW1 may differ between runs, as with the atomics of v1).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_csc_scan(int32_t* __restrict__ gene_cnt, int G, int32_t* __restrict__ c_ptr,
                                                   int32_t* __restrict__ cursor, int use_smem) {
  PDL_ENTRY();
  // one sweep: the counts are staged in shared memory with coalesced loads (when they fit), thread t
  // scans its contiguous genes [t*per, (t+1)*per) there, and the three outputs are written coalesced
  extern __shared__ int32_t cnt_s[];
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = (G + 1023) / 1024;
  const int g0 = min(G, tid * per), g1 = min(G, g0 + per);
  int32_t* cnt = use_smem ? cnt_s : gene_cnt;
  if (use_smem) {
    for (int g = tid; g < G; g += 1024) cnt_s[g] = gene_cnt[g];
    __syncthreads();
  }
  int tsum = 0;
  for (int g = g0; g < g1; ++g) tsum += cnt[g];
  int x = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  int excl = (wid ? warp_tot[wid - 1] : 0) + x - tsum;
  if (use_smem) {
    for (int g = g0; g < g1; ++g) {
      const int v = cnt_s[g];
      cnt_s[g] = excl;
      excl += v;
    }
    __syncthreads();
    for (int g = tid; g < G; g += 1024) {
      const int e = cnt_s[g];
      c_ptr[g] = e;
      cursor[g] = e;
      gene_cnt[g] = 0;
    }
  } else {
    for (int g = g0; g < g1; ++g) {
      const int v = gene_cnt[g];
      c_ptr[g] = excl;
      cursor[g] = excl;
      gene_cnt[g] = 0;
      excl += v;
    }
  }
  if (tid == 1023) c_ptr[G] = warp_tot[31];
}

// one CTA per batch row (coalesced reads of the row, one atomic per non-zero entry)
__global__ void __launch_bounds__(128) k_csc_fill(const int32_t* __restrict__ b_indptr, const int32_t* __restrict__ b_idx,
                                                  const float* __restrict__ b_val, int n, int32_t* __restrict__ cursor,
                                                  int2* __restrict__ c_ent) {
  PDL_ENTRY();
  const int row = blockIdx.x;
  const int s = b_indptr[row], e = b_indptr[row + 1];
  for (int j = s + threadIdx.x; j < e; j += 128) {
    const float v = b_val[j];
    if (v != 0.f) {
      const int pos = atomicAdd(cursor + b_idx[j], 1);
      c_ent[pos] = make_int2(row, __float_as_int(v));
    }
  }
}

extern "C" int scnb_csr_to_csc(const int32_t* b_indptr, const int32_t* b_idx, const float* b_val, int n, int G,
                               int32_t* gene_cnt, int32_t* c_ptr, int32_t* cursor, void* c_ent, void* stream) {
  SCNB_CHECK_ARG(b_indptr && b_idx && b_val && gene_cnt && c_ptr && cursor && c_ent && n >= 0 && G > 0,
                 "csr_to_csc: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t scan_smem = (size_t)G * sizeof(int32_t);
  const int use_smem = scan_smem <= 200 * 1024;
  if (use_smem && scan_smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_csc_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem);
    if (e != cudaSuccess) {
      scnb_set_error("csr_to_csc: cannot reserve %zu bytes of shared memory: %s", scan_smem, cudaGetErrorString(e));
      return SCNB_ERR_CUDA;
    }
  }
  scnb_launch_pdl(k_csc_scan, dim3(1), dim3(1024), use_smem ? scan_smem : 0, st, gene_cnt, G, c_ptr, cursor, use_smem);
  SCNB_CHECK_LAUNCH("csr_to_csc/scan");
  if (n > 0) {
    scnb_launch_pdl(k_csc_fill, dim3(n), dim3(128), 0, st, b_indptr, b_idx, b_val, n, cursor, (int2*)c_ent);
    SCNB_CHECK_LAUNCH("csr_to_csc/fill");
  }
  return SCNB_OK;
}

// per-gene histogram of the non-zero stored entries of a batch CSR (when the batch was not
// assembled by scnb_csr_gather_rows, which counts on the fly)
__global__ void k_csc_count(const int32_t* __restrict__ b_idx, const float* __restrict__ b_val, int nnz,
                            int32_t* __restrict__ gene_cnt) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += gridDim.x * blockDim.x)
    if (b_val[j] != 0.f) atomicAdd(gene_cnt + b_idx[j], 1);
}

extern "C" int scnb_csc_count(const int32_t* b_idx, const float* b_val, int nnz, int32_t* gene_cnt, void* stream) {
  SCNB_CHECK_ARG(b_idx && b_val && gene_cnt && nnz >= 0, "csc_count: bad arguments");
  if (nnz == 0) return SCNB_OK;
  const int blocks = scnb_cdiv(nnz, 256) < 148 * 8 ? scnb_cdiv(nnz, 256) : 148 * 8;
  k_csc_count<<<blocks, 256, 0, (cudaStream_t)stream>>>(b_idx, b_val, nnz, gene_cnt);
  SCNB_CHECK_LAUNCH("csc_count");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// K8+K9 fused:  for every gene g  grad[g,:] = sum_{(r,v) in column g} v * dZ[r,:]  and, in the same
// pass, the optimizer update of W1T[g,:] and its two state rows (scnym/model.py:211 backward +
// optimizer.step(), scnym/trainer.py:671).  The first layer holds ~96 % of all parameters: this
// kernel reads p, s1, s2 once and writes them once (24 B per parameter), and the gradient never
// goes to HBM unless `grad_out` is given (parity tests, data-parallel all-reduce).
// Grid = (H0/64 column slices, gene blocks).  A CTA stages its 64-column slice of dZ in shared
// memory ([nb][64] fp32, 128 KB at nb=512) once, then its 32 half-warps walk genes g = gb*32+hw,
// += 32*gridDim.y; 16 lanes x float4 cover the slice.  Entries are loaded 16 at a time (one 128-B
// line per half-warp) and broadcast by shuffle; the next gene's p/s1/s2 are prefetched before the
// current gene's entries are consumed.  When nb*256 B exceeds shared memory (large batches) dZ is
// read through L1/L2 instead (STAGED=false).
// ---------------------------------------------------------------------------------------------
#define W1_NT 1024          // threads per CTA of the fused dW1+optimizer kernel
#define W1_HW (W1_NT / 16)  // half-warps (genes in flight) per CTA
template <int OPT, bool STAGED>
__global__ void __launch_bounds__(W1_NT) k_w1_bwd_optim(const int32_t* __restrict__ c_ptr, const int2* __restrict__ c_ent,
                                                      const float4* __restrict__ dZ, int nb, int H0v, int G,
                                                      float4* __restrict__ W, float4* __restrict__ grad_out,
                                                      float4* __restrict__ S1, float4* __restrict__ S2,
                                                      const float* __restrict__ hp, const int64_t* __restrict__ st) {
  PDL_ENTRY();
  extern __shared__ float4 dzs[];  // [nb][16]
  __shared__ OptimConsts cs;
  const int tid = threadIdx.x;
  const int slice = blockIdx.x;
  if (OPT != OPT_NONE && tid == 0) cs = optim_consts(hp, st);
  if (STAGED) {
    for (int i = tid; i < nb * 16; i += W1_NT) dzs[i] = dZ[(size_t)(i >> 4) * H0v + slice * 16 + (i & 15)];
  }
  __syncthreads();
  OptimConsts c;
  if (OPT != OPT_NONE) c = cs;
  // A warp walks two genes at a time (one per half-warp); every loop below has a warp-uniform trip
  // count (the longer of the two genes; a finished half multiplies by v = 0), so the shuffles use the
  // full mask and the warp never diverges.
  const int lane = tid & 31, q = tid & 15, half16 = lane & 16;
  const int hw = tid >> 4;
  const int col = slice * 16 + q;
  const int gstride = gridDim.y * W1_HW;
  const int g_first = blockIdx.y * W1_HW + hw;
  const int n_iter = (G - (blockIdx.y * W1_HW + (hw & ~1)) + gstride - 1) / gstride;  // uniform across the warp
  // software pipeline: column ranges are fetched two genes ahead, entries and parameter/state rows one ahead
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f), a = p, b = p;
  int js = 0, je = 0, jsn = 0, jen = 0;
  int2 ent = make_int2(0, 0);
  if (g_first < G) {
    js = c_ptr[g_first];
    je = c_ptr[g_first + 1];
    if (OPT != OPT_NONE) {
      const size_t o = (size_t)g_first * H0v + col;
      p = W[o];
      a = S1[o];
      if (OPT != OPT_SGD) b = S2[o];
    }
    if (js + q < je) ent = c_ent[js + q];
  }
  if (g_first + gstride < G) {
    jsn = c_ptr[g_first + gstride];
    jen = c_ptr[g_first + gstride + 1];
  }
  for (int it = 0; it < n_iter; ++it) {
    const int g = g_first + it * gstride;
    const int gn = g + gstride, gnn = gn + gstride;
    float4 pn = make_float4(0.f, 0.f, 0.f, 0.f), an = pn, bn = pn;
    int2 entn = make_int2(0, 0);
    int jsnn = 0, jenn = 0;
    if (gn < G) {
      if (jsn + q < jen) entn = c_ent[jsn + q];
      if (OPT != OPT_NONE) {
        const size_t o = (size_t)gn * H0v + col;
        pn = W[o];
        an = S1[o];
        if (OPT != OPT_SGD) bn = S2[o];
      }
    }
    if (gnn < G) {
      jsnn = c_ptr[gnn];
      jenn = c_ptr[gnn + 1];
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int len = je - js;
    const int len_max = max(len, __shfl_xor_sync(0xffffffffu, len, 16));
    for (int base = 0; base < len_max; base += 16) {
      const int2 e_cur = ent;
      // next batch of 16 entries of this gene (if any), issued before the current batch is consumed
      ent = make_int2(0, 0);
      if (base + 16 + q < len) ent = c_ent[js + base + 16 + q];
      const int cnt = min(16, len_max - base);
#pragma unroll 4
      for (int k = 0; k < cnt; ++k) {
        const int r = __shfl_sync(0xffffffffu, e_cur.x, half16 + k);
        const float v = __int_as_float(__shfl_sync(0xffffffffu, e_cur.y, half16 + k));
        const float4 d = STAGED ? dzs[r * 16 + q] : __ldg(dZ + (size_t)r * H0v + col);
        acc.x = fmaf(v, d.x, acc.x);
        acc.y = fmaf(v, d.y, acc.y);
        acc.z = fmaf(v, d.z, acc.z);
        acc.w = fmaf(v, d.w, acc.w);
      }
    }
    if (g < G) {
      const size_t o = (size_t)g * H0v + col;
      if (grad_out) grad_out[o] = acc;
      if (OPT != OPT_NONE) {
        optim_update<OPT>(p.x, acc.x, a.x, b.x, c);
        optim_update<OPT>(p.y, acc.y, a.y, b.y, c);
        optim_update<OPT>(p.z, acc.z, a.z, b.z, c);
        optim_update<OPT>(p.w, acc.w, a.w, b.w, c);
        W[o] = p;
        S1[o] = a;
        if (OPT != OPT_SGD) S2[o] = b;
      }
    }
    p = pn; a = an; b = bn;
    js = jsn; je = jen; ent = entn;
    jsn = jsnn; jen = jenn;
  }
}

template <int OPT>
static int launch_w1_bwd_optim(const int32_t* c_ptr, const void* c_ent, const float* dZ, int nb, int H0, int G, float* W1T,
                               float* grad_out, float* s1, float* s2, const float* hp, const int64_t* st4, cudaStream_t st) {
  const int slices = H0 / 64;
  const size_t smem = (size_t)nb * 256;
  const bool staged = smem <= 200 * 1024;
  // one CTA per SM when the dZ slice is staged (its shared memory fills the SM); more, smaller shares otherwise
  int gblocks = staged ? (148 / slices > 0 ? 148 / slices : 1) : (148 * 2 / slices > 0 ? 148 * 2 / slices : 1);
  if (gblocks > scnb_cdiv(G, W1_HW)) gblocks = scnb_cdiv(G, W1_HW);
  dim3 grid(slices, gblocks);
  if (staged) {
    cudaError_t e = cudaFuncSetAttribute(k_w1_bwd_optim<OPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      scnb_set_error("csr_w1_bwd_optim: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return SCNB_ERR_CUDA;
    }
    scnb_launch_pdl(k_w1_bwd_optim<OPT, true>, grid, dim3(W1_NT), smem, st, c_ptr, (const int2*)c_ent, (const float4*)dZ, nb, H0 / 4, G, (float4*)W1T,
                                                       (float4*)grad_out, (float4*)s1, (float4*)s2, hp, st4);
  } else {
    k_w1_bwd_optim<OPT, false><<<grid, W1_NT, 0, st>>>(c_ptr, (const int2*)c_ent, (const float4*)dZ, nb, H0 / 4, G, (float4*)W1T,
                                                     (float4*)grad_out, (float4*)s1, (float4*)s2, hp, st4);
  }
  SCNB_CHECK_LAUNCH("csr_w1_bwd_optim");
  return SCNB_OK;
}

extern "C" int scnb_csr_w1_bwd_optim(int opt, const int32_t* c_ptr, const void* c_ent, const float* dZ, int nb, int H0, int G,
                                     float* W1T, float* grad_out, float* state1, float* state2, const float* hp8,
                                     const int64_t* state4, void* stream) {
  SCNB_CHECK_ARG(c_ptr && c_ent && dZ && nb >= 0 && H0 > 0 && G > 0, "csr_w1_bwd_optim: bad arguments");
  SCNB_CHECK_ARG(H0 % 64 == 0 && ((uintptr_t)dZ % 16 == 0), "csr_w1_bwd_optim: H0 must be a multiple of 64 and dZ 16-byte aligned");
  SCNB_CHECK_ARG(opt == OPT_NONE ? (grad_out != nullptr) : (W1T && state1 && hp8 && state4 && (opt == OPT_SGD || state2)),
                 "csr_w1_bwd_optim: missing buffers for the requested mode");
  SCNB_CHECK_ARG((((uintptr_t)W1T | (uintptr_t)grad_out | (uintptr_t)state1 | (uintptr_t)state2) % 16) == 0,
                 "csr_w1_bwd_optim: parameter/state rows must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  switch (opt) {
    case OPT_NONE: return launch_w1_bwd_optim<OPT_NONE>(c_ptr, c_ent, dZ, nb, H0, G, W1T, grad_out, state1, state2, hp8, state4, st);
    case OPT_ADADELTA: return launch_w1_bwd_optim<OPT_ADADELTA>(c_ptr, c_ent, dZ, nb, H0, G, W1T, grad_out, state1, state2, hp8, state4, st);
    case OPT_ADAM: return launch_w1_bwd_optim<OPT_ADAM>(c_ptr, c_ent, dZ, nb, H0, G, W1T, grad_out, state1, state2, hp8, state4, st);
    case OPT_ADAMW: return launch_w1_bwd_optim<OPT_ADAMW>(c_ptr, c_ent, dZ, nb, H0, G, W1T, grad_out, state1, state2, hp8, state4, st);
    case OPT_SGD: return launch_w1_bwd_optim<OPT_SGD>(c_ptr, c_ent, dZ, nb, H0, G, W1T, grad_out, state1, state2, hp8, state4, st);
    default: scnb_set_error("csr_w1_bwd_optim: unknown optimizer %d", opt); return SCNB_ERR_ARG;
  }
}

// ---------------------------------------------------------------------------------------------
// Step prologue, one single-CTA launch: (1) sampler hand-off -- row ids / MixUp randomness of step
// (t % n_tab) from the per-epoch tables (DataLoader iteration, scnym/trainer.py:574-582);
// (2) labels / domain ids of the sampled rows (SingleCellDS.__getitem__, scnym/dataprep.py:157-162);
// (3) row offsets of the batch CSR [U rows ; L rows]; (4) zero the loss scalars; (5) advance the
// device-side step counters (optimizer t, RNG counter).
// ---------------------------------------------------------------------------------------------
struct PrepArgs {
  const int64_t* l_tab; const int64_t* u_tab; const float* key_tab; const float* gam_tab; int n_tab;
  int64_t* state4; int L, U;
  int64_t* l_rows; int64_t* u_rows; float* keys; float* gam;
  const int32_t* lab_y; const int32_t* lab_dom; const int32_t* unl_dom; int want_dom;
  int32_t* lab_ids; int32_t* base_dom;
  const int64_t* lab_indptr; const int64_t* unl_indptr;
  int32_t* b_indptr; float* zero_buf; int n_zero;
  int phases;   // bit 0: sampler hand-off, labels, row offsets of the batch whose table row is state4[0] % n_tab;
                // bit 1: advance the step counters and zero the loss scalars.  3 = the whole prologue (one launch at the start
                // of a step); 1 / 2 split it so that the batch of step t+1 can be assembled while step t is still running.
};

__global__ void __launch_bounds__(1024) k_step_prep(PrepArgs a) {
  PDL_ENTRY();
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int L = a.L, U = a.U, nb = L + U;
  const int64_t srow = a.l_tab ? (a.state4[0] % (int64_t)a.n_tab) : 0;
  if (tid == 0) carry_s = 0;
  if (a.phases & 2)
    for (int i = tid; i < a.n_zero; i += 1024) a.zero_buf[i] = 0.f;
  __syncthreads();
  for (int base = 0; (a.phases & 1) && base < nb; base += 1024) {
    const int i = base + tid;
    int len = 0;
    if (i < nb) {
      if (a.key_tab) {
        a.keys[i] = a.key_tab[srow * nb + i];
        a.gam[i] = a.gam_tab[srow * nb + i];
      }
      if (i < U) {
        const int64_t r = a.u_tab ? a.u_tab[srow * U + i] : a.u_rows[i];
        if (a.u_tab) a.u_rows[i] = r;
        if (a.want_dom) a.base_dom[i] = a.unl_dom ? a.unl_dom[r] : 1;
        len = (int)(a.unl_indptr[r + 1] - a.unl_indptr[r]);
      } else {
        const int li = i - U;
        const int64_t r = a.l_tab ? a.l_tab[srow * L + li] : a.l_rows[li];
        if (a.l_tab) a.l_rows[li] = r;
        a.lab_ids[li] = a.lab_y ? a.lab_y[r] : 0;
        if (a.want_dom) a.base_dom[i] = a.lab_dom ? a.lab_dom[r] : 0;
        len = (int)(a.lab_indptr[r + 1] - a.lab_indptr[r]);
      }
    }
    int x = len;  // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int carry = carry_s;
    if (i < nb) a.b_indptr[i] = carry + (wid ? warp_tot[wid - 1] : 0) + x - len;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (tid == 0) {
    if (a.phases & 1) a.b_indptr[nb] = carry_s;
    if (a.phases & 2) {
      a.state4[0] += 1;  // optimizer step index t (1-based when read by the optimizer kernels)
      a.state4[2] += 1;  // RNG counter
    }
  }
}

extern "C" int scnb_step_prep(const int64_t* l_tab, const int64_t* u_tab, const float* key_tab, const float* gam_tab, int n_tab,
                              int64_t* state4, int L, int U, int64_t* l_rows, int64_t* u_rows, float* keys, float* gam,
                              const int32_t* lab_y, const int32_t* lab_dom, const int32_t* unl_dom, int want_dom,
                              int32_t* lab_ids, int32_t* base_dom, const int64_t* lab_indptr, const int64_t* unl_indptr,
                              int32_t* b_indptr, float* zero_buf, int n_zero, int phases, void* stream) {
  SCNB_CHECK_ARG(state4 && L > 0 && U >= 0 && l_rows && lab_ids && lab_indptr && b_indptr, "step_prep: bad arguments");
  SCNB_CHECK_ARG(phases >= 1 && phases <= 3, "step_prep: phases must be 1 (batch), 2 (counters) or 3 (both)");
  SCNB_CHECK_ARG(U == 0 || (u_rows && unl_indptr), "step_prep: unlabeled buffers missing");
  SCNB_CHECK_ARG(!l_tab || n_tab > 0, "step_prep: empty sampler table");
  SCNB_CHECK_ARG(!l_tab || U == 0 || u_tab, "step_prep: unlabeled sampler table missing");
  SCNB_CHECK_ARG(!key_tab || (gam_tab && keys && gam), "step_prep: key/gamma buffers missing");
  SCNB_CHECK_ARG(!want_dom || base_dom, "step_prep: base_dom missing");
  SCNB_CHECK_ARG(n_zero == 0 || zero_buf, "step_prep: zero_buf missing");
  PrepArgs a = {l_tab, u_tab, key_tab, gam_tab, n_tab, state4, L, U, l_rows, u_rows, keys, gam, lab_y, lab_dom, unl_dom,
                want_dom, lab_ids, base_dom, lab_indptr, unl_indptr, b_indptr, zero_buf, n_zero, phases};
  scnb_launch_pdl(k_step_prep, dim3(1), dim3(1024), 0, (cudaStream_t)stream, a);
  SCNB_CHECK_LAUNCH("step_prep");
  return SCNB_OK;
}

// out[i] = table ? table[rows ? rows[i] : i] : fill   (labels / domain ids of the sampled rows)
__global__ void k_gather_i32(const int32_t* __restrict__ table, const int64_t* __restrict__ rows, int n, int32_t fill,
                             int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = table ? table[rows ? rows[i] : i] : fill;
}

extern "C" int scnb_gather_i32(const int32_t* table, const int64_t* rows, int n, int32_t fill, int32_t* out, void* stream) {
  SCNB_CHECK_ARG(out && n >= 0, "gather_i32: bad arguments");
  if (n == 0) return SCNB_OK;
  k_gather_i32<<<scnb_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(table, rows, n, fill, out);
  SCNB_CHECK_LAUNCH("gather_i32");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Device-resident sampler hand-off: copy row ids / MixUp randomness of step (state4[0] % n_tab)
// from per-epoch tables into the engine's input buffers, so that an epoch is pure graph replay.
// Replaces the host DataLoader iteration of scnym/trainer.py:574-582 (+ main.py:48-68 `repeater`).
// ---------------------------------------------------------------------------------------------
__global__ void k_fetch_step(const int64_t* __restrict__ l_tab, const int64_t* __restrict__ u_tab,
                             const float* __restrict__ key_tab, const float* __restrict__ gam_tab, int L, int U, int n_tab,
                             const int64_t* __restrict__ state4, int64_t* __restrict__ l_rows, int64_t* __restrict__ u_rows,
                             float* __restrict__ keys, float* __restrict__ gam) {
  const int64_t s = state4[0] % (int64_t)n_tab;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L) l_rows[i] = l_tab[s * L + i];
  if (u_tab && i < U) u_rows[i] = u_tab[s * U + i];
  if (key_tab && i < L + U) {
    keys[i] = key_tab[s * (L + U) + i];
    gam[i] = gam_tab[s * (L + U) + i];
  }
}

extern "C" int scnb_fetch_step(const int64_t* l_tab, const int64_t* u_tab, const float* key_tab, const float* gam_tab, int L,
                               int U, int n_tab, const int64_t* state4, int64_t* l_rows, int64_t* u_rows, float* keys,
                               float* gam, void* stream) {
  SCNB_CHECK_ARG(l_tab && state4 && l_rows && L > 0 && U >= 0 && n_tab > 0, "fetch_step: bad arguments");
  SCNB_CHECK_ARG(!u_tab || u_rows, "fetch_step: u_rows missing");
  SCNB_CHECK_ARG(!key_tab || (gam_tab && keys && gam), "fetch_step: key/gamma buffers missing");
  k_fetch_step<<<scnb_cdiv(L + U, 256), 256, 0, (cudaStream_t)stream>>>(l_tab, u_tab, key_tab, gam_tab, L, U, n_tab, state4, l_rows,
                                                                        u_rows, keys, gam);
  SCNB_CHECK_LAUNCH("fetch_step");
  return SCNB_OK;
}
