// Gene re-indexing of a CSR atlas to the model's gene order (sm_100a), the step in front of batched predict.
// Reference: scnym/utils.py:155-233 (`build_classification_matrix`; caller scnym/api.py:670-675) maps every
// sample gene to its model column with an O(genes^2) Python loop and copies columns through a LIL matrix.
// Here the host builds `col_map[sample column] = model column or -1` once (a hash join) and the device
//   1. counts the surviving entries of every row (one warp per row) and scans the counts into the new indptr,
//   2. rewrites every row in the model's column order (one CTA per row).
// Step 2 needs the survivors of a row SORTED by their new column.  The new columns of a row are distinct
// (col_map is injective on its non-negative entries), so the CTA marks them in a shared-memory bitmap over the
// model's genes; the position of an entry in the output row is then the number of marked bits below its own
// (per-word prefix popcount + popcount of the low bits of its word): O(nnz + G/32) per row, no sort.
// HBM traffic: 4 B per stored entry in pass 1, 8 B read + 8 B written per entry in pass 2 (col_map stays in L1/L2).
#include <stdlib.h>

#include "common.cuh"

#define REMAP_ROWS_PER_CTA 64   // rows whose counts one CTA of the counting pass scans locally

// Pass 1: len[r] = #{p in row r : col_map[indices[p]] >= 0}; inclusive scan inside the CTA's 64 rows is written to
// out_indptr[1 + r], the CTA total to cta_sums[blockIdx.x].
__global__ void __launch_bounds__(256) k_remap_count(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                     const int32_t* __restrict__ col_map, int n_src_cols, long long n_rows,
                                                     int64_t* __restrict__ out_indptr, int64_t* __restrict__ cta_sums) {
  __shared__ int lens[REMAP_ROWS_PER_CTA];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r0 = (long long)blockIdx.x * REMAP_ROWS_PER_CTA;
  for (int j = warp; j < REMAP_ROWS_PER_CTA; j += 8) {
    const long long r = r0 + j;
    int cnt = 0;
    if (r < n_rows) {
      const int64_t s = indptr[r], e = indptr[r + 1];
      for (int64_t p = s + lane; p < e; p += 32) {
        const int32_t c = indices[p];
        cnt += ((unsigned)c < (unsigned)n_src_cols && col_map[c] >= 0) ? 1 : 0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) lens[j] = cnt;
  }
  __syncthreads();
  if (warp == 0) {   // inclusive scan of the 64 counts by one warp (two per lane)
    const int a = lens[2 * lane], b = lens[2 * lane + 1];
    int x = a + b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    const long long ra = r0 + 2 * lane;
    if (ra < n_rows) out_indptr[1 + ra] = (int64_t)(x - b);
    if (ra + 1 < n_rows) out_indptr[2 + ra] = (int64_t)x;
    if (lane == 31) cta_sums[blockIdx.x] = (int64_t)x;
    if (blockIdx.x == 0 && lane == 0) out_indptr[0] = 0;
  }
}

// Exclusive scan of the CTA totals in place (single CTA; 4 M rows = 65 536 totals = 64 rounds of 1024).
__global__ void __launch_bounds__(1024) k_remap_scan_sums(int64_t* __restrict__ cta_sums, int n) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n; b0 += 1024) {
    const int i = b0 + tid;
    const long long v = i < n ? cta_sums[i] : 0;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      long long t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const long long carry = carry_s;
    if (i < n) cta_sums[i] = carry + (wid ? warp_tot[wid - 1] : 0) + x - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_remap_add_offsets(int64_t* __restrict__ out_indptr, const int64_t* __restrict__ cta_sums,
                                                           long long n_rows) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) out_indptr[1 + r] += cta_sums[r / REMAP_ROWS_PER_CTA];
}

// Pass 2: one CTA per row.  Shared memory: bitmap[nw] over the destination genes, then prefix[nw].
// Measured (profiles/r01_ncu_full_v12_remap.txt): 65 536 rows of ~1 000 entries in 556 us, L2 throughput 70 %, DRAM 21 %:
// the col_map gathers (one 32-byte sector per 4-byte entry, the shared-memory carve-out leaves little L1) and the placing
// stores keep L2 busy.  Staging the placed row in shared memory to write full lines was tried and was slower (+12 %).
// A persistent form (one CTA per SM, col_map in shared memory, a warp per row with its own bitmap / prefix / 1024-entry
// staging area so that rows leave as full lines) was also tried: correct, but 1.54 ms against 0.73 ms per 75 776-row
// chunk -- eight warps per SM walking rows one after the other hide far less latency than sixteen CTAs per SM.
// A thread owns REMAP_E entries of a 128 x REMAP_E tile of the row; their loads (index, then col_map and value) are issued
// together so that a row costs three memory latencies, not three per entry.  Rows that fit one tile (<= 1024 stored entries:
// every row at 5 % of 20 000 genes) keep new column and value in registers between the marking and the placing phase;
// longer rows walk their tiles twice.
#define REMAP_E 8
__global__ void __launch_bounds__(128) k_remap_fill(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                    const float* __restrict__ data, const int32_t* __restrict__ col_map,
                                                    int n_src_cols, int nw, const int64_t* __restrict__ out_indptr,
                                                    int32_t* __restrict__ out_indices, float* __restrict__ out_data) {
  extern __shared__ uint32_t remap_sm[];
  uint32_t* bm = remap_sm;
  uint32_t* pre = remap_sm + nw;
  __shared__ uint32_t warp_tot[4];
  const long long r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t s = indptr[r], e = indptr[r + 1];
  const int64_t o0 = out_indptr[r];
  if (out_indptr[r + 1] == o0) return;     // nothing of this row survives (uniform over the CTA)
  const bool one_tile = (e - s) <= 128 * REMAP_E;
  int32_t d[REMAP_E];
  float v[REMAP_E];
  // new columns of a tile's entries: index loads first, then the dependent col_map loads, all independent of each other
  auto load_tile = [&](int64_t t0, bool want_val) {
    int32_t c[REMAP_E];
#pragma unroll
    for (int j = 0; j < REMAP_E; ++j) {
      const int64_t p = t0 + tid + 128 * j;
      c[j] = p < e ? indices[p] : -1;
    }
#pragma unroll
    for (int j = 0; j < REMAP_E; ++j) {
      const int64_t p = t0 + tid + 128 * j;
      d[j] = (unsigned)c[j] < (unsigned)n_src_cols ? col_map[c[j]] : -1;
      if (want_val) v[j] = p < e ? data[p] : 0.f;
    }
  };
  for (int w = tid; w < nw; w += 128) bm[w] = 0u;
  if (one_tile) load_tile(s, true);        // in flight during the clear
  __syncthreads();
  if (one_tile) {
#pragma unroll
    for (int j = 0; j < REMAP_E; ++j)
      if (d[j] >= 0) atomicOr(&bm[d[j] >> 5], 1u << (d[j] & 31));
  } else {
    for (int64_t t0 = s; t0 < e; t0 += 128 * REMAP_E) {
      load_tile(t0, false);
#pragma unroll
      for (int j = 0; j < REMAP_E; ++j)
        if (d[j] >= 0) atomicOr(&bm[d[j] >> 5], 1u << (d[j] & 31));
    }
  }
  __syncthreads();
  // exclusive prefix popcount over the bitmap words: thread t owns words [t * per, t * per + per)
  const int per = (nw + 127) / 128;
  const int w0 = tid * per, w1 = min(nw, w0 + per);
  uint32_t mine = 0;
  for (int w = w0; w < w1; ++w) mine += __popc(bm[w]);
  uint32_t x = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[wid] = x;
  __syncthreads();
  uint32_t run = x - mine;
  for (int k = 0; k < wid; ++k) run += warp_tot[k];
  for (int w = w0; w < w1; ++w) {
    pre[w] = run;
    run += __popc(bm[w]);
  }
  __syncthreads();
  auto place_tile = [&]() {
#pragma unroll
    for (int j = 0; j < REMAP_E; ++j) {
      if (d[j] >= 0) {
        const uint32_t word = bm[d[j] >> 5];
        const int64_t q = o0 + pre[d[j] >> 5] + __popc(word & ((1u << (d[j] & 31)) - 1u));
        out_indices[q] = d[j];
        out_data[q] = v[j];
      }
    }
  };
  if (one_tile) {
    place_tile();
  } else {
    for (int64_t t0 = s; t0 < e; t0 += 128 * REMAP_E) {
      load_tile(t0, true);
      place_tile();
    }
  }
}

// number of int64 words of caller-provided workspace the counting pass needs for n_rows rows
extern "C" int scnb_csr_remap_workspace_len(long long n_rows) {
  return (int)((n_rows + REMAP_ROWS_PER_CTA - 1) / REMAP_ROWS_PER_CTA + 1);
}

extern "C" int scnb_csr_remap_count(const int64_t* indptr, const int32_t* indices, const int32_t* col_map, int n_src_cols,
                                    long long n_rows, int64_t* out_indptr, void* workspace, void* stream) {
  SCNB_CHECK_ARG(indptr && indices && col_map && out_indptr && workspace && n_rows >= 0 && n_src_cols > 0,
                 "csr_remap_count: bad arguments");
  SCNB_CHECK_ARG(n_rows <= 2000000000LL, "csr_remap_count: at most 2e9 rows per call (got %lld)", n_rows);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t* sums = (int64_t*)workspace;
  if (n_rows == 0) {
    cudaMemsetAsync(out_indptr, 0, sizeof(int64_t), st);
    return SCNB_OK;
  }
  const int n_cta = scnb_cdiv(n_rows, REMAP_ROWS_PER_CTA);
  k_remap_count<<<n_cta, 256, 0, st>>>(indptr, indices, col_map, n_src_cols, n_rows, out_indptr, sums);
  SCNB_CHECK_LAUNCH("csr_remap_count");
  k_remap_scan_sums<<<1, 1024, 0, st>>>(sums, n_cta);
  SCNB_CHECK_LAUNCH("csr_remap_count(scan)");
  k_remap_add_offsets<<<scnb_cdiv(n_rows, 256), 256, 0, st>>>(out_indptr, sums, n_rows);
  SCNB_CHECK_LAUNCH("csr_remap_count(offsets)");
  return SCNB_OK;
}

extern "C" int scnb_csr_remap_fill(const int64_t* indptr, const int32_t* indices, const float* data, const int32_t* col_map,
                                   int n_src_cols, int n_dst_cols, long long n_rows, const int64_t* out_indptr,
                                   int32_t* out_indices, float* out_data, void* stream) {
  SCNB_CHECK_ARG(indptr && indices && data && col_map && out_indptr && n_rows >= 0 && n_src_cols > 0 && n_dst_cols > 0,
                 "csr_remap_fill: bad arguments");
  SCNB_CHECK_ARG(n_rows <= 2000000000LL, "csr_remap_fill: at most 2e9 rows per call (got %lld)", n_rows);
  if (n_rows == 0) return SCNB_OK;
  SCNB_CHECK_ARG(out_indices && out_data, "csr_remap_fill: missing output buffers");
  const int nw = (n_dst_cols + 31) / 32;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = 2 * sizeof(uint32_t) * (size_t)nw;
  SCNB_CHECK_ARG(smem <= 200 * 1024, "csr_remap_fill: %d destination genes need %zu bytes of shared memory (limit 200 KB)",
                 n_dst_cols, smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_remap_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      scnb_set_error("csr_remap_fill: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return SCNB_ERR_CUDA;
    }
  }
  k_remap_fill<<<(unsigned)n_rows, 128, smem, st>>>(indptr, indices, data, col_map, n_src_cols, nw, out_indptr,
                                                  out_indices, out_data);
  SCNB_CHECK_LAUNCH("csr_remap_fill");
  return SCNB_OK;
}
