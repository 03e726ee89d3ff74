// Pseudolabel, MixUp planning and loss kernels of the scNym hot path (sm_100a).
//   scnb_pseudolabel        K-view softmax mean, confidence, sharpening        (losses.py:660-737, 529-573)
//   scnb_mixmatch_plan      confident/unconfident compaction, injected-key permutation, Beta gamma,
//                           row maps of the student super-batch and their inverses
//                           (losses.py:811-875, dataprep.py:665-689, losses.py:1195-1213)
//   scnb_soft_ce_fwd_bwd    soft-target CE + dLogits                            (losses.py:67-151)
//   scnb_softmax_mse_fwd_bwd  sum_c (softmax(z)-y)^2 masked mean + dLogits      (losses.py:880-913)
//   scnb_predict_head       ensemble-mean logits -> argmax / softmax            (predict.py:178-192)
#include "common.cuh"
#include "loss_rows.cuh"

// ---------------------------------------------------------------------------------------------
// Pseudolabels.  One warp per unlabeled row.  logits: [K][U][C].
// ---------------------------------------------------------------------------------------------
// One warp, one row `u` of q: pseudolabel of unlabeled cell u (u < U) or one-hot label of labeled row u-U.
__device__ __forceinline__ void pseudolabel_row(int u, const float* __restrict__ logits, int K, int U, int C, float T,
                                                     int sharpen, float min_conf, float* __restrict__ q,
                                                     float* __restrict__ conf, uint8_t* __restrict__ confident,
                                                     float* __restrict__ scratch, const int32_t* __restrict__ lab_ids,
                                                     int L) {
  const int lane = threadIdx.x & 31;
  if (u >= U) {  // rows [U, U+L) of q: one-hot labels of the labeled rows (SingleCellDS, dataprep.py:157-160)
    if (lab_ids && u < U + L) {
      const int id = lab_ids[u - U];
      for (int c = lane; c < C; c += 32) q[(size_t)u * C + c] = (c == id) ? 1.f : 0.f;
    }
    return;
  }
  float* qb = scratch + (size_t)u * C;  // mean probabilities
  for (int c = lane; c < C; c += 32) qb[c] = 0.f;
  __syncwarp();
  const float invK = 1.0f / (float)K;
  for (int k = 0; k < K; ++k) {
    const float* z = logits + ((size_t)k * U + u) * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, z[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(z[c] - m);
    s = warp_sum(s);
    for (int c = lane; c < C; c += 32) qb[c] += expf(z[c] - m) / s;
  }
  __syncwarp();
  // mean over K, max/argmax (first maximal index, as torch.max)
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = qb[c] * invK;
    qb[c] = v;
    if (v > best) { best = v; besti = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  if (besti >= C) besti = 0;   // a row of NaN / -inf compares false everywhere: torch.max returns index 0 there
  if (lane == 0) {
    conf[u] = best;
    confident[u] = best >= min_conf ? 1 : 0;
  }
  __syncwarp();
  float* qo = q + (size_t)u * C;
  if (!sharpen || T == 1.0f) {
    for (int c = lane; c < C; c += 32) qo[c] = qb[c];
  } else if (T == 0.0f) {
    for (int c = lane; c < C; c += 32) qo[c] = (c == besti) ? 1.f : 0.f;
  } else {
    const float e = 1.0f / T;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += powf(qb[c], e);
    s = warp_sum(s);
    for (int c = lane; c < C; c += 32) qo[c] = powf(qb[c], e) / s;
  }
}

__global__ void __launch_bounds__(256) k_pseudolabel(const float* __restrict__ logits, int K, int U, int C, float T,
                                                     int sharpen, float min_conf, float* __restrict__ q,
                                                     float* __restrict__ conf, uint8_t* __restrict__ confident,
                                                     float* __restrict__ scratch, const int32_t* __restrict__ lab_ids,
                                                     int L) {
  const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  pseudolabel_row(u, logits, K, U, C, T, sharpen, min_conf, q, conf, confident, scratch, lab_ids, L);
}

extern "C" int scnb_pseudolabel(const float* teacher_logits, int K, int U, int C, float T, int sharpen, float min_confidence,
                                float* q, float* conf, uint8_t* confident, float* scratch_UC, const int32_t* lab_ids, int L,
                                void* stream) {
  SCNB_CHECK_ARG(teacher_logits && q && conf && confident && scratch_UC && K > 0 && U >= 0 && C > 0, "pseudolabel: bad arguments");
  const int rows = U + (lab_ids ? L : 0);
  if (rows == 0) return SCNB_OK;
  k_pseudolabel<<<scnb_cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(teacher_logits, K, U, C, T, sharpen, min_confidence, q, conf,
                                                                        confident, scratch_UC, lab_ids, L);
  SCNB_CHECK_LAUNCH("pseudolabel");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// MixMatch plan.  Base rows: [0,U) unlabeled (raw), [U,U+L) labeled (augmented).
// Super-batch rows: [0,U) U-pass = [mixed confident ; raw unconfident], [U,U+L) L-pass (mixed),
// [U+L, U+L+n_dan) DAN = [labeled ; unlabeled (all or confident only)].
// plan_i layout (int32): see offsets below.  All of it lives in device memory.
// ---------------------------------------------------------------------------------------------
struct PlanPtrs {
  int32_t* counts;    // [8]: n_conf, n_mm, n_dan, R_active, U, L, n_dan_unl, 0
  float* pconf;       // [1]: n_dan_unl / max(U,1)  (losses.py:1203)
  int32_t* seg_off;   // [4]
  int32_t* cpos;      // [U]   position among confident (or among unconfident) rows
  int32_t* cat2base;  // [U+L]
  int32_t* rank;      // [U+L]
  int32_t* ridx;      // [U+L]
  int32_t* srcA;      // [Rmax]
  int32_t* srcB;      // [Rmax]
  float* gam;         // [Rmax]
  int32_t* inv3;      // [3][U+L]
  float* coef3;       // [3][U+L]
};

__device__ void plan_compact_body(const uint8_t* __restrict__ confident, int U, int L, int use_dan, int dan_conf_only,
                                  const PlanPtrs& P, int* wtot, int* carry_s, int* n_conf_s) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) (*carry_s) = 0;
  __syncthreads();
  // exclusive scan of the confident flags
  for (int base = 0; base < U; base += 1024) {
    const int i = base + tid;
    const int f = (i < U && confident[i]) ? 1 : 0;
    int x = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wtot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int t = wtot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      wtot[lane] = t;
    }
    __syncthreads();
    const int carry = (*carry_s);
    const int excl = carry + (wid ? wtot[wid - 1] : 0) + x - f;
    if (i < U) P.cpos[i] = f ? excl : (i - excl);  // rank among confident, or among unconfident
    __syncthreads();
    if (tid == 1023) (*carry_s) = carry + wtot[31];
    __syncthreads();
  }
  if (tid == 0) {
    const int n_conf = (*carry_s);
    (*n_conf_s) = n_conf;
    const int n_dan_unl = use_dan ? (dan_conf_only ? n_conf : U) : 0;
    const int n_dan = use_dan ? L + n_dan_unl : 0;
    P.counts[0] = n_conf;
    P.counts[1] = n_conf + L;
    P.counts[2] = n_dan;
    P.counts[3] = U + L + n_dan;
    P.counts[4] = U;
    P.counts[5] = L;
    P.counts[6] = n_dan_unl;
    P.counts[7] = 0;
    if (P.pconf) P.pconf[0] = (float)n_dan_unl / (float)max(U, 1);
    P.seg_off[0] = 0;
    P.seg_off[1] = U;
    P.seg_off[2] = U + L;
    P.seg_off[3] = U + L + n_dan;
  }
  __syncthreads();
  const int n_conf = (*n_conf_s);
  for (int i = tid; i < U; i += 1024)
    if (confident[i]) P.cat2base[P.cpos[i]] = i;
  for (int i = tid; i < L; i += 1024) P.cat2base[n_conf + i] = U + i;
}

__global__ void __launch_bounds__(1024) k_plan_compact(const uint8_t* __restrict__ confident, int U, int L, int use_dan,
                                                       int dan_conf_only, PlanPtrs P) {
  __shared__ int wtot[32];
  __shared__ int carry_s;
  __shared__ int n_conf_s;
  plan_compact_body(confident, U, L, use_dan, dan_conf_only, P, wtot, &carry_s, &n_conf_s);
}

// rank[j] = #{k < n_mm : key[k] < key[j] or (key[k]==key[j] and k<j)};  ridx[rank[j]] = j
// (== torch.argsort(keys[:n_mm], stable=True), standing in for torch.randperm(n_mm)).
__global__ void __launch_bounds__(256) k_plan_rank(const float* __restrict__ keys, PlanPtrs P, int n_max) {
  __shared__ float tile[256];
  const int n = P.counts[1];
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.x * 256 >= n) return;
  const float kj = j < n ? keys[j] : 0.f;
  int r = 0;
  for (int base = 0; base < n; base += 256) {
    const int k = base + threadIdx.x;
    tile[threadIdx.x] = k < n ? keys[k] : INFINITY;
    __syncthreads();
    const int lim = min(256, n - base);
    for (int t = 0; t < lim; ++t) {
      const float kv = tile[t];
      r += (kv < kj || (kv == kj && base + t < j)) ? 1 : 0;
    }
    __syncthreads();
  }
  if (j < n) {
    P.rank[j] = r;
    P.ridx[r] = j;
  }
}

__device__ __forceinline__ void plan_maps_item(int i, const uint8_t* __restrict__ confident, const float* __restrict__ gamma_raw,
                                               int U, int L, int use_dan, int dan_conf_only, int mixup_on, int Rmax,
                                               const PlanPtrs& P) {
  const int n_conf = P.counts[0], n_dan_unl = P.counts[6];
  const int nb = U + L;
  // ---- forward maps over super-batch rows ----
  if (i < Rmax) {
    int a = 0, b = 0;
    float g = 1.f;
    if (i < U) {                       // U-pass
      if (i < n_conf) {
        a = P.cat2base[i];
        if (mixup_on) {
          b = P.cat2base[P.ridx[i]];
          const float gr = gamma_raw[i];
          g = fmaxf(gr, 1.f - gr);
        } else b = a;
      } else {                         // unconfident rows appended un-mixed, original order
        a = -1;                        // filled below from the inverse side
        b = -1;
      }
    } else if (i < U + L) {            // L-pass
      const int j = n_conf + (i - U);
      a = i;                           // base row U + (i-U)
      if (mixup_on) {
        b = P.cat2base[P.ridx[j]];
        const float gr = gamma_raw[j];
        g = fmaxf(gr, 1.f - gr);
      } else b = a;
    } else {                           // DAN
      const int k = i - U - L;
      if (k < L) a = U + k;
      else if (k - L < n_dan_unl) a = dan_conf_only ? P.cat2base[k - L] : (k - L);
      else a = 0;                      // inactive tail
      b = a;
    }
    if (a >= 0) {
      P.srcA[i] = a;
      P.srcB[i] = b;
      P.gam[i] = g;
    }
  }
  // ---- inverse maps over base rows (and the unconfident forward rows) ----
  if (i < nb) {
    int iA, iB = -1, iD = -1;
    float cA, cB = 0.f, cD = 0.f;
    if (i < U) {
      if (confident[i]) {
        const int j = P.cpos[i];
        iA = j;
        float g = 1.f;
        if (mixup_on) { const float gr = gamma_raw[j]; g = fmaxf(gr, 1.f - gr); }
        cA = g;
        if (mixup_on) {
          const int jp = P.rank[j];    // cat row whose partner is j
          iB = jp < n_conf ? jp : U + (jp - n_conf);
          const float gr = gamma_raw[jp];
          cB = 1.f - fmaxf(gr, 1.f - gr);
        }
        if (use_dan) { iD = U + L + L + (dan_conf_only ? j : i); cD = 1.f; }
      } else {
        const int row = n_conf + P.cpos[i];
        iA = row;
        cA = 1.f;
        P.srcA[row] = i;
        P.srcB[row] = i;
        P.gam[row] = 1.f;
        if (use_dan && !dan_conf_only) { iD = U + L + L + i; cD = 1.f; }
      }
    } else {
      const int l = i - U, j = n_conf + l;
      iA = U + l;
      float g = 1.f;
      if (mixup_on) { const float gr = gamma_raw[j]; g = fmaxf(gr, 1.f - gr); }
      cA = g;
      if (mixup_on) {
        const int jp = P.rank[j];
        iB = jp < n_conf ? jp : U + (jp - n_conf);
        const float gr = gamma_raw[jp];
        cB = 1.f - fmaxf(gr, 1.f - gr);
      }
      if (use_dan) { iD = U + L + l; cD = 1.f; }
    }
    P.inv3[i] = iA;            P.coef3[i] = cA;
    P.inv3[nb + i] = iB;       P.coef3[nb + i] = cB;
    P.inv3[2 * nb + i] = iD;   P.coef3[2 * nb + i] = cD;
  }
}

__global__ void __launch_bounds__(256) k_plan_maps(const uint8_t* __restrict__ confident, const float* __restrict__ gamma_raw,
                                                   int U, int L, int use_dan, int dan_conf_only, int mixup_on, int Rmax,
                                                   PlanPtrs P) {
  plan_maps_item(blockIdx.x * 256 + threadIdx.x, confident, gamma_raw, U, L, use_dan, dan_conf_only, mixup_on, Rmax, P);
}

// Whole plan in ONE single-CTA launch for training-sized batches (U+L <= PLAN_SMEM_KEYS): compaction,
// O(n^2) ranking of the permutation keys out of shared memory, forward and inverse row maps.
#define PLAN_SMEM_KEYS 4096
__global__ void __launch_bounds__(1024) k_plan_all(const uint8_t* __restrict__ confident, const float* __restrict__ keys,
                                                   const float* __restrict__ gamma_raw, int U, int L, int use_dan,
                                                   int dan_conf_only, int mixup_on, int Rmax, PlanPtrs P) {
  PDL_ENTRY();
  __shared__ int wtot[32];
  __shared__ int carry_s, n_conf_s;
  __shared__ float skeys[PLAN_SMEM_KEYS];
  __shared__ int srank[PLAN_SMEM_KEYS];
  plan_compact_body(confident, U, L, use_dan, dan_conf_only, P, wtot, &carry_s, &n_conf_s);
  __syncthreads();
  const int tid = threadIdx.x, nb = U + L;
  if (mixup_on) {
    // stable argsort of the n permutation keys: bitonic sort of (key, index) pairs in shared memory
    // (lexicographic order on the pair == torch.argsort(keys, stable=True)); padded to a power of two
    const int n = n_conf_s + L;
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int k = tid; k < np2; k += 1024) {
      skeys[k] = k < n ? keys[k] : INFINITY;
      srank[k] = k;
    }
    __syncthreads();
    for (int kk = 2; kk <= np2; kk <<= 1) {
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < np2; i += 1024) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const float ka = skeys[i], kb = skeys[ixj];
            const int ia = srank[i], ib = srank[ixj];
            const bool a_gt_b = (ka > kb) || (ka == kb && ia > ib);
            const bool up = (i & kk) == 0;
            if (a_gt_b == up) {
              skeys[i] = kb; skeys[ixj] = ka;
              srank[i] = ib; srank[ixj] = ia;
            }
          }
        }
        __syncthreads();
      }
    }
    for (int r = tid; r < n; r += 1024) {
      const int j = srank[r];   // padding (key = +inf, index >= n) sorts after every real key
      P.ridx[r] = j;
      P.rank[j] = r;
    }
  }
  __syncthreads();
  const int lim = Rmax > nb ? Rmax : nb;
  for (int i = tid; i < lim; i += 1024)
    plan_maps_item(i, confident, gamma_raw, U, L, use_dan, dan_conf_only, mixup_on, Rmax, P);
}

extern "C" int scnb_mixmatch_plan(const uint8_t* confident, const float* perm_keys, const float* gamma_raw, int U, int L,
                                  int use_dan, int dan_conf_only, int mixup_on, int32_t* counts8, int32_t* seg_off4,
                                  int32_t* work_i /* [U + 3*(U+L)] */, int32_t* srcA, int32_t* srcB, float* gam, int Rmax,
                                  int32_t* inv3, float* coef3, float* pconf1, void* stream) {
  SCNB_CHECK_ARG(confident && perm_keys && gamma_raw && counts8 && seg_off4 && work_i && srcA && srcB && gam && inv3 && coef3,
                 "mixmatch_plan: null pointer");
  SCNB_CHECK_ARG(U >= 0 && L >= 1 && Rmax >= U + L + (use_dan ? U + L : 0), "mixmatch_plan: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = U + L;
  PlanPtrs P;
  P.counts = counts8;
  P.pconf = pconf1;
  P.seg_off = seg_off4;
  P.cpos = work_i;
  P.cat2base = work_i + U;
  P.rank = work_i + U + nb;
  P.ridx = work_i + U + 2 * nb;
  P.srcA = srcA;
  P.srcB = srcB;
  P.gam = gam;
  P.inv3 = inv3;
  P.coef3 = coef3;
  if (nb <= PLAN_SMEM_KEYS) {
    scnb_launch_pdl(k_plan_all, dim3(1), dim3(1024), 0, st, confident, perm_keys, gamma_raw, U, L, use_dan, dan_conf_only, mixup_on, Rmax, P);
    SCNB_CHECK_LAUNCH("mixmatch_plan/all");
    return SCNB_OK;
  }
  k_plan_compact<<<1, 1024, 0, st>>>(confident, U, L, use_dan, dan_conf_only, P);
  SCNB_CHECK_LAUNCH("mixmatch_plan/compact");
  if (mixup_on) {
    k_plan_rank<<<scnb_cdiv(nb, 256), 256, 0, st>>>(perm_keys, P, nb);
    SCNB_CHECK_LAUNCH("mixmatch_plan/rank");
  }
  k_plan_maps<<<scnb_cdiv(Rmax > nb ? Rmax : nb, 256), 256, 0, st>>>(confident, gamma_raw, U, L, use_dan, dan_conf_only,
                                                                      mixup_on, Rmax, P);
  SCNB_CHECK_LAUNCH("mixmatch_plan/maps");
  return SCNB_OK;
}

__global__ void __launch_bounds__(256) k_soft_ce(const float* __restrict__ Zl, int n_max, int C,
                                                 const int32_t* __restrict__ n_rows_dev, const float* __restrict__ Y,
                                                 const int32_t* __restrict__ mapA, const int32_t* __restrict__ mapB,
                                                 const float* __restrict__ gam, int map_off,
                                                 const float* __restrict__ class_weight, const int32_t* __restrict__ n_norm_dev,
                                                 int n_norm, float grad_scale, const float* __restrict__ loss_scale_dev,
                                                 float* __restrict__ dZ, float* __restrict__ loss_out, int count_acc,
                                                 float* __restrict__ row_loss) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n_max) return;
  ce_row(i, Zl, n_max, C, n_rows_dev, Y, mapA, mapB, gam, map_off, class_weight, n_norm_dev, n_norm, grad_scale, loss_scale_dev, dZ,
         loss_out, count_acc, row_loss);
}

extern "C" int scnb_soft_ce_fwd_bwd(const float* logits, int n_max, int C, const int32_t* n_rows_dev, const float* Y,
                                    const int32_t* mapA, const int32_t* mapB, const float* gam, int map_off,
                                    const float* class_weight, const int32_t* n_norm_dev, int n_norm, float grad_scale,
                                    const float* loss_scale_dev, float* dlogits, float* loss_out2, int count_acc,
                                    float* row_loss_out, void* stream) {
  SCNB_CHECK_ARG(logits && Y && loss_out2 && n_max >= 0 && C > 0, "soft_ce_fwd_bwd: bad arguments");
  if (n_max == 0) return SCNB_OK;
  k_soft_ce<<<scnb_cdiv(n_max, 8), 256, 0, (cudaStream_t)stream>>>(logits, n_max, C, n_rows_dev, Y, mapA, mapB, gam, map_off,
                                                                     class_weight, n_norm_dev, n_norm, grad_scale,
                                                                     loss_scale_dev, dlogits, loss_out2, count_acc, row_loss_out);
  SCNB_CHECK_LAUNCH("soft_ce_fwd_bwd");
  return SCNB_OK;
}

__global__ void __launch_bounds__(256) k_softmax_mse(const float* __restrict__ Zl, int n_max, int C,
                                                     const int32_t* __restrict__ n_conf_dev, const float* __restrict__ Y,
                                                     const int32_t* __restrict__ mapA, const int32_t* __restrict__ mapB,
                                                     const float* __restrict__ gam, int map_off, int n_norm,
                                                     const float* __restrict__ weight_dev, float weight,
                                                     float* __restrict__ dZ, float* __restrict__ loss_out) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n_max) return;
  mse_row(i, Zl, n_max, C, n_conf_dev, Y, mapA, mapB, gam, map_off, n_norm, weight_dev, weight, dZ, loss_out);
}

extern "C" int scnb_softmax_mse_fwd_bwd(const float* logits, int n_max, int C, const int32_t* n_conf_dev, const float* Y,
                                        const int32_t* mapA, const int32_t* mapB, const float* gam, int map_off, int n_norm,
                                        const float* weight_dev, float weight, float* dlogits, float* loss_out, void* stream) {
  SCNB_CHECK_ARG(logits && Y && loss_out && n_max >= 0 && C > 0, "softmax_mse_fwd_bwd: bad arguments");
  if (n_max == 0) return SCNB_OK;
  k_softmax_mse<<<scnb_cdiv(n_max, 8), 256, 0, (cudaStream_t)stream>>>(logits, n_max, C, n_conf_dev, Y, mapA, mapB, gam, map_off,
                                                                         n_norm, weight_dev, weight, dlogits, loss_out);
  SCNB_CHECK_LAUNCH("softmax_mse_fwd_bwd");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// One-hot gather:  out[i, :] = onehot(ids[src[i]])  (domain targets of the DAN rows), zeros when id < 0.
// Also usable to build label matrices from integer labels.
// ---------------------------------------------------------------------------------------------
__global__ void k_onehot_rows(const int32_t* __restrict__ ids, const int32_t* __restrict__ src, int src_off, int n, int C,
                              float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * C) return;
  const int i = t / C, c = t - i * C;
  const int r = src ? src[src_off + i] : i;
  out[t] = (ids[r] == c) ? 1.f : 0.f;
}

extern "C" int scnb_onehot_rows(const int32_t* ids, const int32_t* src, int src_off, int n, int C, float* out, void* stream) {
  SCNB_CHECK_ARG(ids && out && n >= 0 && C > 0, "onehot_rows: bad arguments");
  if (n == 0) return SCNB_OK;
  k_onehot_rows<<<scnb_cdiv((long long)n * C, 256), 256, 0, (cudaStream_t)stream>>>(ids, src, src_off, n, C, out);
  SCNB_CHECK_LAUNCH("onehot_rows");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Predict head: score = logits_sum / n_models; pred = argmax; prob = softmax(score).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_predict_head(const float* __restrict__ logits_sum, float inv_models, int n, int C,
                                                      int64_t* __restrict__ pred, float* __restrict__ prob,
                                                      float* __restrict__ score) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const float* z = logits_sum + (size_t)i * C;
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = z[c] * inv_models;
    if (v > m) { m = v; am = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oi < am)) { m = om; am = oi; }
  }
  if (am >= C) am = 0;         // all-NaN / all -inf row: a valid class index, as torch.max gives
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(z[c] * inv_models - m);
  s = warp_sum(s);
  for (int c = lane; c < C; c += 32) {
    const float v = z[c] * inv_models;
    if (score) score[(size_t)i * C + c] = v;
    if (prob) prob[(size_t)i * C + c] = expf(v - m) / s;
  }
  if (lane == 0 && pred) pred[i] = am;
}

extern "C" int scnb_predict_head(const float* logits_sum, int n_models, int n, int C, int64_t* pred, float* prob, float* score,
                                 void* stream) {
  SCNB_CHECK_ARG(logits_sum && n_models > 0 && n >= 0 && C > 0, "predict_head: bad arguments");
  if (n == 0) return SCNB_OK;
  k_predict_head<<<scnb_cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(logits_sum, 1.0f / (float)n_models, n, C, pred, prob, score);
  SCNB_CHECK_LAUNCH("predict_head");
  return SCNB_OK;
}

// MixMatch student head: rows [0,U) interpolation-consistency MSE, rows [U,U+L) soft-target CE
// (scnym/model.py:229 classifier; scnym/losses.py:880-923), dA = dLogits Wc.
template <int NV>
__global__ void __launch_bounds__(256) k_classif_mixmatch(const float* __restrict__ A, const float* __restrict__ Wc,
                                                          const float* __restrict__ bc, int H, int C, int U, int L,
                                                          const int32_t* __restrict__ n_conf_dev, const float* __restrict__ Y,
                                                          const int32_t* __restrict__ mapA, const int32_t* __restrict__ mapB,
                                                          const float* __restrict__ gam, const float* __restrict__ class_weight,
                                                          float unsup_weight, float* __restrict__ logits,
                                                          float* __restrict__ dlogits, float* __restrict__ dA,
                                                          float* __restrict__ loss8) {
  PDL_ENTRY();
  extern __shared__ __align__(16) float Ws[];
  // The rows' loss terms meet in shared memory first: one global atomic per CTA and scalar instead of one per row
  // (hundreds of same-address atomics serialise in L2 and were most of this kernel's duration).
  __shared__ float sl[8];
  if (threadIdx.x < 8) sl[threadIdx.x] = 0.f;
  stage_weights(Ws, Wc, C * H);
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r < U + L) {
    float4 a[NV];
    if (C <= 32) {
      float* wsm = Ws + (size_t)C * H + (threadIdx.x >> 5) * HEAD_WS;
      float* zs = wsm + 32 * 33;
      float* dzs = zs + 32;
      head_logits_s<NV>(A + (size_t)r * H, Ws, bc, C, H, wsm, a, lane);
      if (lane < C) logits[(size_t)r * C + lane] = zs[lane];
      // the row functions address row i of a [rows, C] array: hand them the shared-memory row through a shifted base
      if (r < U)
        mse_row(r, zs - (size_t)r * C, U, C, n_conf_dev, Y, mapA, mapB, gam, 0, U, nullptr, unsup_weight, dzs - (size_t)r * C, sl + 2);
      else
        ce_row(r - U, zs - (size_t)(r - U) * C, L, C, nullptr, Y, mapA, mapB, gam, U, class_weight, nullptr, L, 1.0f, nullptr,
               dzs - (size_t)(r - U) * C, sl, 1, nullptr);
      __syncwarp();
      if (lane < C) dlogits[(size_t)r * C + lane] = dzs[lane];
      head_backprop<NV>(dzs, Ws, C, H, 1.0f, nullptr, dA + (size_t)r * H, lane);
    } else {
      head_logits<NV>(A + (size_t)r * H, Ws, bc, C, H, logits + (size_t)r * C, a, lane);
      if (r < U)
        mse_row(r, logits, U, C, n_conf_dev, Y, mapA, mapB, gam, 0, U, nullptr, unsup_weight, dlogits, sl + 2);
      else
        ce_row(r - U, logits + (size_t)U * C, L, C, nullptr, Y, mapA, mapB, gam, U, class_weight, nullptr, L, 1.0f, nullptr,
               dlogits + (size_t)U * C, sl, 1, nullptr);
      __syncwarp();
      head_backprop<NV>(dlogits + (size_t)r * C, Ws, C, H, 1.0f, nullptr, dA + (size_t)r * H, lane);
    }
  }
  __syncthreads();
  if (threadIdx.x < 4 && sl[threadIdx.x] != 0.f) atomicAdd(loss8 + threadIdx.x, sl[threadIdx.x]);
}

// Supervised head: CE against one-hot integer labels (Trainer step, scnym/trainer.py:189-256).
template <int NV>
__global__ void __launch_bounds__(256) k_classif_supervised(const float* __restrict__ A, const float* __restrict__ Wc,
                                                            const float* __restrict__ bc, int H, int C, int n,
                                                            const float* __restrict__ Y, const float* __restrict__ class_weight,
                                                            float* __restrict__ logits, float* __restrict__ dlogits,
                                                            float* __restrict__ dA, float* __restrict__ loss8) {
  PDL_ENTRY();
  extern __shared__ __align__(16) float Ws[];
  __shared__ float sl[8];
  if (threadIdx.x < 8) sl[threadIdx.x] = 0.f;
  stage_weights(Ws, Wc, C * H);
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r < n) {
    float4 a[NV];
    if (C <= 32) {
      float* wsm = Ws + (size_t)C * H + (threadIdx.x >> 5) * HEAD_WS;
      float* zs = wsm + 32 * 33;
      float* dzs = zs + 32;
      head_logits_s<NV>(A + (size_t)r * H, Ws, bc, C, H, wsm, a, lane);
      if (lane < C) logits[(size_t)r * C + lane] = zs[lane];
      ce_row(r, zs - (size_t)r * C, n, C, nullptr, Y, nullptr, nullptr, nullptr, 0, class_weight, nullptr, n, 1.0f, nullptr,
             dzs - (size_t)r * C, sl, 1, nullptr);
      __syncwarp();
      if (lane < C) dlogits[(size_t)r * C + lane] = dzs[lane];
      head_backprop<NV>(dzs, Ws, C, H, 1.0f, nullptr, dA + (size_t)r * H, lane);
    } else {
      head_logits<NV>(A + (size_t)r * H, Ws, bc, C, H, logits + (size_t)r * C, a, lane);
      ce_row(r, logits, n, C, nullptr, Y, nullptr, nullptr, nullptr, 0, class_weight, nullptr, n, 1.0f, nullptr, dlogits, sl, 1,
             nullptr);
      __syncwarp();
      head_backprop<NV>(dlogits + (size_t)r * C, Ws, C, H, 1.0f, nullptr, dA + (size_t)r * H, lane);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 && sl[threadIdx.x] != 0.f) atomicAdd(loss8 + threadIdx.x, sl[threadIdx.x]);
}

// Domain adversary tail: Hd = relu(Linear(H,H)(embed)) -> Linear(H,D) -> CE against the domain label
// (scnym/model.py:377-383, scnym/losses.py:1209-1243) -> dHd = (ddlog W2) * [Hd > 0].
template <int NV>
__global__ void __launch_bounds__(256) k_dan_tail(const float* __restrict__ Hd, const float* __restrict__ W2,
                                                  const float* __restrict__ b2, int H, int D, int Rd,
                                                  const int32_t* __restrict__ n_rows_dev, const int32_t* __restrict__ dom_ids,
                                                  const int32_t* __restrict__ src, int src_off,
                                                  const float* __restrict__ loss_scale_dev, float* __restrict__ dlog,
                                                  float* __restrict__ dlabel, float* __restrict__ ddlog,
                                                  float* __restrict__ dHd, float* __restrict__ loss2) {
  PDL_ENTRY();
  extern __shared__ __align__(16) float Ws[];
  __shared__ float sl[8];
  if (threadIdx.x < 8) sl[threadIdx.x] = 0.f;
  stage_weights(Ws, W2, D * H);
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r < Rd) {
    float4 a[NV];
    const int id = dom_ids[src ? src[src_off + r] : r];
    if (D <= 32) {
      float* wsm = Ws + (size_t)D * H + (threadIdx.x >> 5) * HEAD_WS;
      float* zs = wsm + 32 * 33;
      float* dzs = zs + 32;
      float* ys = wsm;               // the [C][33] tile is free once the logits are formed: one-hot label row
      head_logits_s<NV>(Hd + (size_t)r * H, Ws, b2, D, H, wsm, a, lane);
      if (lane < D) {
        dlog[(size_t)r * D + lane] = zs[lane];
        const float y = (lane == id) ? 1.f : 0.f;
        dlabel[(size_t)r * D + lane] = y;
        ys[lane] = y;
      }
      __syncwarp();
      ce_row(r, zs - (size_t)r * D, Rd, D, n_rows_dev, ys - (size_t)r * D, nullptr, nullptr, nullptr, 0, nullptr, n_rows_dev, 0, 1.0f,
             loss_scale_dev, dzs - (size_t)r * D, sl, 1, nullptr);
      __syncwarp();
      if (lane < D) ddlog[(size_t)r * D + lane] = dzs[lane];
      head_backprop<NV>(dzs, Ws, D, H, 1.0f, a, dHd + (size_t)r * H, lane);
    } else {
      head_logits<NV>(Hd + (size_t)r * H, Ws, b2, D, H, dlog + (size_t)r * D, a, lane);
      for (int c = lane; c < D; c += 32) dlabel[(size_t)r * D + c] = (c == id) ? 1.f : 0.f;
      __syncwarp();
      ce_row(r, dlog, Rd, D, n_rows_dev, dlabel, nullptr, nullptr, nullptr, 0, nullptr, n_rows_dev, 0, 1.0f, loss_scale_dev, ddlog,
             sl, 1, nullptr);
      __syncwarp();
      head_backprop<NV>(ddlog + (size_t)r * D, Ws, D, H, 1.0f, a, dHd + (size_t)r * H, lane);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 && sl[threadIdx.x] != 0.f) atomicAdd(loss2 + threadIdx.x, sl[threadIdx.x]);
}

// Teacher head: classifier on the K eval-mode views of cell u, then pseudolabel (+ label one-hots).
template <int NV>
__global__ void __launch_bounds__(256) k_teacher_head(const float* __restrict__ TA, const float* __restrict__ Wc,
                                                      const float* __restrict__ bc, int H, int C, int K, int U, float T,
                                                      int sharpen, float min_conf, float* __restrict__ t_logits,
                                                      float* __restrict__ q, float* __restrict__ conf,
                                                      uint8_t* __restrict__ confident, float* __restrict__ scratch,
                                                      const int32_t* __restrict__ lab_ids, int L, const float* __restrict__ bn_gamma,
                                                      const float* __restrict__ bn_beta, const float* __restrict__ bn_rmean,
                                                      const float* __restrict__ bn_rvar) {
  PDL_ENTRY();
  extern __shared__ __align__(16) float Ws[];
  const int lane = threadIdx.x & 31;
  const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bn_gamma && K == 1 && C <= 32) {
    // TA holds the PRE-BatchNorm output of the last hidden Linear: eval-mode BatchNorm -> ReLU (running statistics,
    // scnym/model.py:219-227 under model.eval()) is applied as the row is loaded -- one launch less on the teacher's chain.
    // The row and the column constants are requested before the weights are staged.
    float4 a[NV];
    if (u < U) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c4 = i * 32 + lane;
        const float4 z = *reinterpret_cast<const float4*>(TA + (size_t)u * H + c4 * 4);
        const float4 m = reinterpret_cast<const float4*>(bn_rmean)[c4], v = reinterpret_cast<const float4*>(bn_rvar)[c4];
        const float4 g = reinterpret_cast<const float4*>(bn_gamma)[c4], b = reinterpret_cast<const float4*>(bn_beta)[c4];
        a[i].x = fmaxf((z.x - m.x) * (1.0f / sqrtf(v.x + 1e-5f)) * g.x + b.x, 0.f);
        a[i].y = fmaxf((z.y - m.y) * (1.0f / sqrtf(v.y + 1e-5f)) * g.y + b.y, 0.f);
        a[i].z = fmaxf((z.z - m.z) * (1.0f / sqrtf(v.z + 1e-5f)) * g.z + b.z, 0.f);
        a[i].w = fmaxf((z.w - m.w) * (1.0f / sqrtf(v.w + 1e-5f)) * g.w + b.w, 0.f);
      }
    }
    stage_weights(Ws, Wc, C * H);
    if (u < U) {
      float* wsm = Ws + (size_t)C * H + (threadIdx.x >> 5) * HEAD_WS;
      float* zs = wsm + 32 * 33;
      float* qs = zs + 32;
      head_logits_s_regs<NV>(a, Ws, bc, C, H, wsm, lane);
      if (lane < C) t_logits[(size_t)u * C + lane] = zs[lane];
      pseudolabel_row(u, zs - (size_t)u * C, 1, U, C, T, sharpen, min_conf, q, conf, confident, qs - (size_t)u * C, lab_ids, L);
      return;
    }
    pseudolabel_row(u, t_logits, K, U, C, T, sharpen, min_conf, q, conf, confident, scratch, lab_ids, L);
    return;
  }
  stage_weights(Ws, Wc, C * H);
  if (u < U && K == 1 && C <= 32) {   // one view: logits and the mean-probability scratch stay in shared memory
    float4 a[NV];
    float* wsm = Ws + (size_t)C * H + (threadIdx.x >> 5) * HEAD_WS;
    float* zs = wsm + 32 * 33;
    float* qs = zs + 32;
    head_logits_s<NV>(TA + (size_t)u * H, Ws, bc, C, H, wsm, a, lane);
    if (lane < C) t_logits[(size_t)u * C + lane] = zs[lane];
    pseudolabel_row(u, zs - (size_t)u * C, 1, U, C, T, sharpen, min_conf, q, conf, confident, qs - (size_t)u * C, lab_ids, L);
    return;
  }
  if (u < U) {
    float4 a[NV];
    for (int k = 0; k < K; ++k) {
      const size_t row = (size_t)k * U + u;
      head_logits<NV>(TA + row * H, Ws, bc, C, H, t_logits + row * C, a, lane);
    }
  }
  pseudolabel_row(u, t_logits, K, U, C, T, sharpen, min_conf, q, conf, confident, scratch, lab_ids, L);
}

static bool head_fusable(int H, int C, const void* p0, const void* p1, const void* p2) {
  const uintptr_t bits = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2;
  return (H % 128 == 0) && H <= 1024 && (size_t)C * H * sizeof(float) <= 96 * 1024 && (bits % 16 == 0);
}
#define HEAD_DISPATCH(KERNEL, rows, smem, st, ...)                                                   \
  do {                                                                                               \
    const int nv__ = H / 128;                                                                        \
    dim3 grid__(scnb_cdiv(rows, 8));                                                                 \
    if ((smem) > 48 * 1024) {                                                                        \
      cudaFuncSetAttribute(KERNEL<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem));     \
      cudaFuncSetAttribute(KERNEL<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem));     \
      cudaFuncSetAttribute(KERNEL<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem));     \
      cudaFuncSetAttribute(KERNEL<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem));     \
    }                                                                                                \
    if (nv__ == 1) scnb_launch_pdl(KERNEL<1>, grid__, dim3(256), smem, st, __VA_ARGS__);                                \
    else if (nv__ == 2) scnb_launch_pdl(KERNEL<2>, grid__, dim3(256), smem, st, __VA_ARGS__);                           \
    else if (nv__ <= 4) { if (nv__ == 4) scnb_launch_pdl(KERNEL<4>, grid__, dim3(256), smem, st, __VA_ARGS__); else return SCNB_ERR_UNSUPPORTED; } \
    else if (nv__ == 8) scnb_launch_pdl(KERNEL<8>, grid__, dim3(256), smem, st, __VA_ARGS__);                           \
    else return SCNB_ERR_UNSUPPORTED;                                                                \
  } while (0)

// Returns SCNB_ERR_UNSUPPORTED (-3) when the shape cannot be fused (caller falls back to the unfused sequence).
extern "C" int scnb_classif_mixmatch_fused(const float* A, const float* Wc, const float* bc, int H, int C, int U, int L,
                                           const int32_t* n_conf_dev, const float* Y, const int32_t* mapA, const int32_t* mapB,
                                           const float* gam, const float* class_weight, float unsup_weight, float* logits,
                                           float* dlogits, float* dA, float* loss8, void* stream) {
  SCNB_CHECK_ARG(A && Wc && Y && mapA && logits && dlogits && dA && loss8 && H > 0 && C > 0 && U >= 0 && L >= 0,
                 "classif_mixmatch_fused: bad arguments");
  if (!head_fusable(H, C, A, Wc, dA) || (H / 128 != 1 && H / 128 != 2 && H / 128 != 4 && H / 128 != 8)) return SCNB_ERR_UNSUPPORTED;
  if (U + L == 0) return SCNB_OK;
  const size_t smem = ((size_t)C * H + (C <= 32 ? 8 * HEAD_WS : 0)) * sizeof(float);
  HEAD_DISPATCH(k_classif_mixmatch, U + L, smem, (cudaStream_t)stream, A, Wc, bc, H, C, U, L, n_conf_dev, Y, mapA, mapB, gam,
                class_weight, unsup_weight, logits, dlogits, dA, loss8);
  SCNB_CHECK_LAUNCH("classif_mixmatch_fused");
  return SCNB_OK;
}

extern "C" int scnb_classif_supervised_fused(const float* A, const float* Wc, const float* bc, int H, int C, int n, const float* Y,
                                             const float* class_weight, float* logits, float* dlogits, float* dA, float* loss8,
                                             void* stream) {
  SCNB_CHECK_ARG(A && Wc && Y && logits && dlogits && dA && loss8 && H > 0 && C > 0 && n >= 0, "classif_supervised_fused: bad arguments");
  if (!head_fusable(H, C, A, Wc, dA) || (H / 128 != 1 && H / 128 != 2 && H / 128 != 4 && H / 128 != 8)) return SCNB_ERR_UNSUPPORTED;
  if (n == 0) return SCNB_OK;
  const size_t smem = ((size_t)C * H + (C <= 32 ? 8 * HEAD_WS : 0)) * sizeof(float);
  HEAD_DISPATCH(k_classif_supervised, n, smem, (cudaStream_t)stream, A, Wc, bc, H, C, n, Y, class_weight, logits, dlogits, dA, loss8);
  SCNB_CHECK_LAUNCH("classif_supervised_fused");
  return SCNB_OK;
}

extern "C" int scnb_dan_tail_fused(const float* Hd, const float* W2, const float* b2, int H, int D, int Rd,
                                   const int32_t* n_rows_dev, const int32_t* dom_ids, const int32_t* src, int src_off,
                                   const float* loss_scale_dev, float* dlog, float* dlabel, float* ddlog, float* dHd, float* loss2,
                                   void* stream) {
  SCNB_CHECK_ARG(Hd && W2 && dom_ids && dlog && dlabel && ddlog && dHd && loss2 && H > 0 && D > 0 && Rd >= 0,
                 "dan_tail_fused: bad arguments");
  if (!head_fusable(H, D, Hd, W2, dHd) || (H / 128 != 1 && H / 128 != 2 && H / 128 != 4 && H / 128 != 8)) return SCNB_ERR_UNSUPPORTED;
  if (Rd == 0) return SCNB_OK;
  const size_t smem = ((size_t)D * H + (D <= 32 ? 8 * HEAD_WS : 0)) * sizeof(float);
  HEAD_DISPATCH(k_dan_tail, Rd, smem, (cudaStream_t)stream, Hd, W2, b2, H, D, Rd, n_rows_dev, dom_ids, src, src_off, loss_scale_dev,
                dlog, dlabel, ddlog, dHd, loss2);
  SCNB_CHECK_LAUNCH("dan_tail_fused");
  return SCNB_OK;
}

static int teacher_head_launch(const float* TA, const float* Wc, const float* bc, int H, int C, int K, int U, float T, int sharpen,
                               float min_confidence, float* t_logits, float* q, float* conf, uint8_t* confident, float* scratch_UC,
                               const int32_t* lab_ids, int L, const float* bn_gamma, const float* bn_beta, const float* bn_rmean,
                               const float* bn_rvar, void* stream) {
  SCNB_CHECK_ARG(TA && Wc && t_logits && q && conf && confident && scratch_UC && H > 0 && C > 0 && K > 0 && U >= 0,
                 "teacher_head_fused: bad arguments");
  if (!head_fusable(H, C, TA, Wc, Wc) || (H / 128 != 1 && H / 128 != 2 && H / 128 != 4 && H / 128 != 8)) return SCNB_ERR_UNSUPPORTED;
  const int rows = U + (lab_ids ? L : 0);
  if (rows == 0) return SCNB_OK;
  const size_t smem = ((size_t)C * H + (C <= 32 ? 8 * HEAD_WS : 0)) * sizeof(float);
  HEAD_DISPATCH(k_teacher_head, rows, smem, (cudaStream_t)stream, TA, Wc, bc, H, C, K, U, T, sharpen, min_confidence, t_logits, q,
                conf, confident, scratch_UC, lab_ids, L, bn_gamma, bn_beta, bn_rmean, bn_rvar);
  SCNB_CHECK_LAUNCH("teacher_head_fused");
  return SCNB_OK;
}

extern "C" int scnb_teacher_head_fused(const float* TA, const float* Wc, const float* bc, int H, int C, int K, int U, float T,
                                       int sharpen, float min_confidence, float* t_logits, float* q, float* conf,
                                       uint8_t* confident, float* scratch_UC, const int32_t* lab_ids, int L, void* stream) {
  return teacher_head_launch(TA, Wc, bc, H, C, K, U, T, sharpen, min_confidence, t_logits, q, conf, confident, scratch_UC, lab_ids, L,
                             nullptr, nullptr, nullptr, nullptr, stream);
}

// The same head on the PRE-BatchNorm output TZ of the last hidden Linear: eval-mode BatchNorm (running statistics) -> ReLU is
// applied while the row is loaded.  One view (K == 1) and at most 32 classes; -3 (SCNB_ERR_UNSUPPORTED) otherwise.
extern "C" int scnb_teacher_head_bn_fused(const float* TZ, const float* gamma, const float* beta, const float* running_mean,
                                          const float* running_var, const float* Wc, const float* bc, int H, int C, int K, int U,
                                          float T, int sharpen, float min_confidence, float* t_logits, float* q, float* conf,
                                          uint8_t* confident, float* scratch_UC, const int32_t* lab_ids, int L, void* stream) {
  SCNB_CHECK_ARG(gamma && beta && running_mean && running_var, "teacher_head_bn_fused: null BatchNorm argument");
  if (K != 1 || C > 32) return SCNB_ERR_UNSUPPORTED;
  SCNB_CHECK_ARG((((uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)running_mean | (uintptr_t)running_var) % 16) == 0,
                 "teacher_head_bn_fused: BatchNorm vectors must be 16-byte aligned");
  return teacher_head_launch(TZ, Wc, bc, H, C, K, U, T, sharpen, min_confidence, t_logits, q, conf, confident, scratch_UC, lab_ids, L,
                             gamma, beta, running_mean, running_var, stream);
}

// Small-output Linear for inference: out[r, :] (+)= A[r, :] W^T + b with C <= a few dozen outputs (the classifier,
// scnym/model.py:229, in Predicter.predict scnym/predict.py:178-192).  One warp per row, weights in shared memory.
template <int NV>
__global__ void __launch_bounds__(256) k_head_logits(const float* __restrict__ A, const float* __restrict__ W,
                                                     const float* __restrict__ b, int H, int C, int n, float* __restrict__ out,
                                                     int accumulate) {
  PDL_ENTRY();
  extern __shared__ __align__(16) float Ws[];
  stage_weights(Ws, W, C * H);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n; r += gridDim.x * wpb) {
    float4 a[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) a[i] = *reinterpret_cast<const float4*>(A + (size_t)r * H + (i * 32 + lane) * 4);
    for (int c = 0; c < C; ++c) {
      float p = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + (size_t)c * H + (i * 32 + lane) * 4);
        p = fmaf(a[i].x, w.x, p);
        p = fmaf(a[i].y, w.y, p);
        p = fmaf(a[i].z, w.z, p);
        p = fmaf(a[i].w, w.w, p);
      }
      p = warp_sum(p);
      if (lane == (c & 31)) {
        float v = p + (b ? b[c] : 0.f);
        if (accumulate) v += out[(size_t)r * C + c];
        out[(size_t)r * C + c] = v;
      }
    }
  }
}

extern "C" int scnb_head_logits(const float* A, const float* W, const float* b, int H, int C, int n, float* out, int accumulate,
                                void* stream) {
  SCNB_CHECK_ARG(A && W && out && H > 0 && C > 0 && n >= 0, "head_logits: bad arguments");
  if (!head_fusable(H, C, A, W, W) || (H / 128 != 1 && H / 128 != 2 && H / 128 != 4 && H / 128 != 8)) return SCNB_ERR_UNSUPPORTED;
  if (n == 0) return SCNB_OK;
  const size_t smem = (size_t)C * H * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  const int nv = H / 128;
  const int blocks = scnb_cdiv(n, 8) < 148 * 8 ? scnb_cdiv(n, 8) : 148 * 8;
#define HL_GO(NV_)                                                                                                  \
  do {                                                                                                              \
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_head_logits<NV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    k_head_logits<NV_><<<blocks, 256, smem, st>>>(A, W, b, H, C, n, out, accumulate);                               \
  } while (0)
  if (nv == 1) HL_GO(1);
  else if (nv == 2) HL_GO(2);
  else if (nv == 4) HL_GO(4);
  else HL_GO(8);
#undef HL_GO
  SCNB_CHECK_LAUNCH("head_logits");
  return SCNB_OK;
}
