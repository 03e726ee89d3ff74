// CSR-side kernels of the scNym hot path (sm_100a).
//   K0/K1  scnb_csr_row_offsets / scnb_csr_gather_rows : device row gather (+ log1p_drop augmentation on nnz only)
//   K2a    scnb_csr_linear_fwd{,_i64}                  : Z = X_csr * W1T + b1   (warp per row, W1T gene-major)
//   K8     scnb_csr_linear_bwd_dw                      : dW1T += X_csr^T * dZ   (vector atomics, v1)
// Reference semantics: scnym/dataprep.py:143-155 (row densify), :721-729 (log1p_drop),
// scnym/model.py:211 (first nn.Linear).  The dense [B,G] batch of the reference is never formed.
#include "common.cuh"

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
// One warp per row.  keep_mask (if given) is aligned with the batch nnz order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_rows(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                     const float* __restrict__ data, const int64_t* __restrict__ rows,
                                                     int64_t row0, int n, const int32_t* __restrict__ b_indptr,
                                                     int32_t* __restrict__ b_idx, float* __restrict__ b_val, int augment,
                                                     int aug_row_begin, int aug_row_end, const uint8_t* __restrict__ keep_mask,
                                                     const uint64_t* __restrict__ rng, uint32_t stream_id, float p_drop,
                                                     float target_sum) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int64_t r = rows ? rows[i] : (row0 + i);
  const int64_t s = indptr[r];
  const int len = (int)(indptr[r + 1] - s);
  const int d = b_indptr[i];
  const bool aug = augment && i >= aug_row_begin && i < aug_row_end;
  if (!aug) {
    for (int j = lane; j < len; j += 32) {
      b_idx[d + j] = indices[s + j];
      b_val[d + j] = data[s + j];
    }
    return;
  }
  const float inv_keep = 1.0f / (1.0f - p_drop);
  uint64_t seed = 0, strm = 0;
  if (!keep_mask) {
    seed = rng[0];
    strm = rng[1] * 64ull + stream_id;
  }
  float sum = 0.f;
  for (int j = lane; j < len; j += 32) {
    float e = expm1f(data[s + j]);
    bool keep = dropout_keep(keep_mask, (uint64_t)(d + j), seed, strm, p_drop);
    e = e * (keep ? inv_keep : 0.f);
    sum += e;
  }
  sum = warp_sum(sum);
  for (int j = lane; j < len; j += 32) {
    float e = expm1f(data[s + j]);
    bool keep = dropout_keep(keep_mask, (uint64_t)(d + j), seed, strm, p_drop);
    e = e * (keep ? inv_keep : 0.f);
    float v = log1pf((e / sum) * target_sum);  // NaN if every entry of the row was dropped, as in the reference
    b_idx[d + j] = indices[s + j];
    b_val[d + j] = v;
  }
}

extern "C" int scnb_csr_gather_rows(const int64_t* indptr, const int32_t* indices, const float* data, const int64_t* rows,
                                    int64_t row0, int n, const int32_t* b_indptr, int32_t* b_idx, float* b_val, int augment,
                                    int aug_row_begin, int aug_row_end, const uint8_t* keep_mask, const uint64_t* rng,
                                    uint32_t stream_id, float p_drop, void* stream) {
  SCNB_CHECK_ARG(indptr && indices && data && b_indptr && b_idx && b_val && n >= 0, "csr_gather_rows: bad arguments");
  SCNB_CHECK_ARG(!augment || keep_mask || rng, "csr_gather_rows: augmentation needs keep_mask or rng state");
  if (n == 0) return SCNB_OK;
  k_gather_rows<<<scnb_cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(indptr, indices, data, rows, row0, n, b_indptr, b_idx,
                                                                     b_val, augment, aug_row_begin, aug_row_end, keep_mask,
                                                                     rng, stream_id, p_drop, 1e6f);
  SCNB_CHECK_LAUNCH("csr_gather_rows");
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// K2a  Z[n,H0] = X_csr[n,G] * W1T[G,H0] + b1.   One warp per row; lane owns NV float4 column
// groups (h = (i*32+lane)*4); the (col,val) stream is loaded coalesced 32 at a time and
// broadcast by shuffle; W1T rows are read as 512-byte coalesced segments.
// ---------------------------------------------------------------------------------------------
template <int NV, typename IdxT>
__global__ void __launch_bounds__(256) k_csr_linear_fwd(const IdxT* __restrict__ indptr, const int32_t* __restrict__ idx,
                                                        const float* __restrict__ val, int n, const float4* __restrict__ W1T,
                                                        const float4* __restrict__ b1, int H0v, float4* __restrict__ Z) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const long long s = (long long)indptr[row], e = (long long)indptr[row + 1];
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
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

template <typename IdxT>
static int launch_csr_linear_fwd(const IdxT* indptr, const int32_t* idx, const float* val, int n, const float* W1T,
                                 const float* b1, int H0, float* Z, cudaStream_t st) {
  if (n == 0) return SCNB_OK;
  dim3 grid(scnb_cdiv(n, 8)), block(256);
  const bool vec = (H0 % 128 == 0) && (((uintptr_t)W1T | (uintptr_t)Z | (uintptr_t)b1) % 16 == 0);
  const int nv = H0 / 128;
  if (vec && nv == 1)
    k_csr_linear_fwd<1, IdxT><<<grid, block, 0, st>>>(indptr, idx, val, n, (const float4*)W1T, (const float4*)b1, H0 / 4, (float4*)Z);
  else if (vec && nv == 2)
    k_csr_linear_fwd<2, IdxT><<<grid, block, 0, st>>>(indptr, idx, val, n, (const float4*)W1T, (const float4*)b1, H0 / 4, (float4*)Z);
  else if (vec && nv == 4)
    k_csr_linear_fwd<4, IdxT><<<grid, block, 0, st>>>(indptr, idx, val, n, (const float4*)W1T, (const float4*)b1, H0 / 4, (float4*)Z);
  else if (vec && nv == 8)
    k_csr_linear_fwd<8, IdxT><<<grid, block, 0, st>>>(indptr, idx, val, n, (const float4*)W1T, (const float4*)b1, H0 / 4, (float4*)Z);
  else
    k_csr_linear_fwd_generic<IdxT><<<grid, block, 0, st>>>(indptr, idx, val, n, W1T, b1, H0, Z);
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
