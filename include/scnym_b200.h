/*
 * libscnym_b200 -- C ABI of the B200-native scNym hot path (MixMatch+DAN train step, batched predict).
 *
 * The reference (calico/scnym v0.4.0) has no FFI/plugin layer: its hot path is Python calling
 * PyTorch ops.  This header is the boundary introduced directly UNDER the reference's Python
 * surfaces (scnym.model / losses / trainer / predict); each entry point names the reference code
 * (file:line under /root/reference) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes
 * binding a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a BORROWED device pointer (e.g. torch.Tensor.data_ptr()); the library
 *    allocates nothing and keeps no global state besides a thread-local error string;
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it, performs no
 *    host synchronisation and is CUDA-graph capturable;
 *  - return 0 on success, negative on error (scnb_last_error() has the message);
 *  - all floating point is fp32, column indices int32, dataset indptr int64, batch indptr int32;
 *  - row-major matrices; W1T is the first-layer weight stored gene-major [G, H0]
 *    (== embed.0.weight.t().contiguous()).
 *  - `*_dev` arguments are device-resident scalars/arrays so that data-dependent sizes
 *    (pseudolabel confidence thresholding) never reach the host.
 *  - dropout: `keep_mask` (uint8, 1 = keep) injects the Bernoulli draw for parity tests; when NULL
 *    the kernel draws Philox4x32-10(seed = rng[0], stream = rng[1]*64 + stream_id, index).
 */
#ifndef SCNYM_B200_H
#define SCNYM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* scnb_last_error(void);
int scnb_version(void);
int scnb_device_arch(void);

/* ---- K0/K1: batch assembly.  Replaces SingleCellDS.__getitem__ + default collate
 *      (scnym/dataprep.py:114-168: per-cell CSR row -> dense fp32 vector, stacked) and the
 *      `log1p_drop` augmentation ExpMinusOne -> InputDropout(0.1) -> LibrarySizeNormalize
 *      (scnym/dataprep.py:257-274, 468-496, 213-254, 720-729), applied on stored nnz only. ---- */
int scnb_csr_row_offsets(const int64_t* indptr, const int64_t* rows /*NULL: row0+i*/, int64_t row0, int n,
                         int32_t* b_indptr /*[n+1]*/, const int32_t* carry_in_dev /*NULL: 0*/, void* stream);
int scnb_gather_i32(const int32_t* table, const int64_t* rows, int n, int32_t fill, int32_t* out, void* stream);
int scnb_csr_gather_rows(const int64_t* indptr, const int32_t* indices, const float* data, const int64_t* rows,
                         int64_t row0, int n, const int32_t* b_indptr, int32_t* b_idx, float* b_val, int augment,
                         int aug_row_begin, int aug_row_end, const uint8_t* keep_mask /*per batch nnz*/,
                         const uint64_t* rng, uint32_t stream_id, float p_drop,
                         int32_t* gene_cnt /*NULL or [G]: += per-gene count of non-zero entries (for scnb_csr_to_csc)*/,
                         void* stream);

/* Two sources in one launch (the MixMatch batch [U_raw ; L_aug]): rows [0,nA) from A, [nA,nA+nB) from B. */
int scnb_csr_gather_rows2(const int64_t* indptr_A, const int32_t* indices_A, const float* data_A, const int64_t* rows_A, int nA,
                          int augment_A, uint32_t stream_id_A, const int64_t* indptr_B, const int32_t* indices_B,
                          const float* data_B, const int64_t* rows_B, int nB, int augment_B, uint32_t stream_id_B,
                          const int32_t* b_indptr, int32_t* b_idx, float* b_val, const uint8_t* keep_mask,
                          const uint64_t* rng, float p_drop, int32_t* gene_cnt, int ctr_off /* as scnb_csr_gather_rows_aug */,
                          void* stream);
/* scnb_csr_gather_rows2 without its shared-memory row cache (the second pass re-reads the row): 32 bytes of shared memory
 * per CTA, so it runs next to a kernel that holds nearly all of an SM's shared memory.  Same arguments and results. */
int scnb_csr_gather_rows2_nocache(const int64_t* indptr_A, const int32_t* indices_A, const float* data_A, const int64_t* rows_A,
                                  int nA, int augment_A, uint32_t stream_id_A, const int64_t* indptr_B, const int32_t* indices_B,
                                  const float* data_B, const int64_t* rows_B, int nB, int augment_B, uint32_t stream_id_B,
                                  const int32_t* b_indptr, int32_t* b_idx, float* b_val, const uint8_t* keep_mask,
                                  const uint64_t* rng, float p_drop, int32_t* gene_cnt, int ctr_off, void* stream);
/* Sampled rows of two HOST-resident matrices (page-locked + mapped: scnb_host_register_mapped) -> a device staging CSR with
 * the batch's row order [A ; B] and offsets (b_indptr from scnb_step_prep; st_indptr: the same offsets as int64, nA+nB+1),
 * copied verbatim over PCIe by a persistent grid of <= max_ctas (0: 148) light CTAs.  The staged rows are then the source
 * of scnb_csr_gather_rows2.  Replaces host-side minibatch assembly + the per-step `.to(device)` of
 * scnym/dataprep.py:114-168, scnym/trainer.py:589-599. */
int scnb_csr_fetch_rows2(const int64_t* indptr_A, const int32_t* indices_A, const float* data_A, const int64_t* rows_A, int nA,
                         const int64_t* indptr_B, const int32_t* indices_B, const float* data_B, const int64_t* rows_B, int nB,
                         const int32_t* b_indptr, int64_t* st_indptr, int32_t* st_idx, float* st_val, int max_ctas, void* stream);
/* The other augmentation schemes of scnym/dataprep.py:730-754 (AUGMENTATION_SCHEMES) on rows [0, n) of one source, on the
 * STORED entries only (every stage maps 0 to 0).  scheme: 1 log1p_drop (:721-729), 2 log1p_mask (GeneMasking :405-465),
 * 3 log1p_poisson (PoissonSample :277-341), 4 log1p_poisson_drop (depth ~ U(0.1, 2), then InputDropout), 5 count_poisson
 * (Poisson on raw counts, no normalisation).  Randomness: counter-based Philox streams of (rng[0], rng[1], stream_id), or
 * injected draws (tests): keep_mask (dropout keeps / gene-masking keeps by batch nnz position), inj_poisson (sampled values by
 * batch nnz position), inj_row_depth (per-row uniforms of the depth factor). */
int scnb_csr_gather_rows_aug(const int64_t* indptr, const int32_t* indices, const float* data, const int64_t* rows,
                             int64_t row0, int n, const int32_t* b_indptr, int32_t* b_idx, float* b_val, int scheme,
                             int n_genes, float p_drop, float p_apply, const uint8_t* keep_mask, const float* inj_poisson,
                             const float* inj_row_depth, const uint64_t* rng, uint32_t stream_id, int32_t* gene_cnt,
                             int ctr_off /* added to the RNG step counter rng[1]: 1 when the next step's batch is assembled early */,
                             void* stream);

/* Step prologue in one launch: sampler hand-off (as scnb_fetch_step; tables may be NULL, then
 * l_rows/u_rows are read as staged), labels/domain ids of the sampled rows
 * (scnym/dataprep.py:157-162), batch row offsets [U rows ; L rows], zeroing of `zero_buf`
 * (loss scalars) and the step counters of scnb_step_begin. */
int scnb_step_prep(const int64_t* l_tab, const int64_t* u_tab, const float* key_tab, const float* gam_tab, int n_tab,
                   int64_t* state4, int L, int U, int64_t* l_rows, int64_t* u_rows, float* keys, float* gam,
                   const int32_t* lab_y, const int32_t* lab_dom, const int32_t* unl_dom, int want_dom, int32_t* lab_ids,
                   int32_t* base_dom, const int64_t* lab_indptr, const int64_t* unl_indptr, int32_t* b_indptr,
                   float* zero_buf, int n_zero,
                   int phases /* 1: assemble the batch of table row state4[0] % n_tab; 2: advance counters + zero; 3: both */,
                   void* stream);

/* Device-resident sampler hand-off (replaces DataLoader iteration, scnym/trainer.py:574-582):
 * copies the row ids / MixUp randomness of step (state4[0] % n_tab) from per-epoch tables. */
int scnb_fetch_step(const int64_t* l_tab, const int64_t* u_tab, const float* key_tab, const float* gam_tab, int L, int U,
                    int n_tab, const int64_t* state4, int64_t* l_rows, int64_t* u_rows, float* keys, float* gam,
                    void* stream);
/* HOST function (no stream): gather CSR rows of a host-resident matrix into (pinned) staging
 * buffers for the host-buffer entry; replaces per-cell densify + collate (dataprep.py:114-168). */
int scnb_host_gather_rows(const int64_t* indptr, const int32_t* indices, const float* data, const int64_t* rows, int n,
                          int64_t* out_indptr, int32_t* out_indices, float* out_data, long long out_capacity,
                          int n_threads);
/* Labeled and unlabeled rows of one MixMatch batch in one call (worker threads started once, next rows prefetched);
 * n_b == 0 gathers the first matrix only. */
int scnb_host_gather_rows2(const int64_t* indptr_a, const int32_t* indices_a, const float* data_a, const int64_t* rows_a, int n_a,
                           int64_t* out_indptr_a, int32_t* out_indices_a, float* out_data_a, long long cap_a,
                           const int64_t* indptr_b, const int32_t* indices_b, const float* data_b, const int64_t* rows_b, int n_b,
                           int64_t* out_indptr_b, int32_t* out_indices_b, float* out_data_b, long long cap_b, int n_threads);

/* ---- K2a/K8: first layer nn.Linear(n_genes, n_hidden_init) (scnym/model.py:211) on CSR input.
 *      fwd: Z = X W1T + b1.   bwd: dW1T += X^T dZ (caller zeroes dW1T).  dX is never formed: the
 *      reference computes it only because inputs carry requires_grad (scnym/trainer.py:596-599)
 *      and never reads it. ---- */
int scnb_csr_linear_fwd(const int32_t* b_indptr, const int32_t* b_idx, const float* b_val, int n, const float* W1T,
                        const float* b1, int H0, float* Z, void* stream);
int scnb_csr_linear_fwd_i64(const int64_t* indptr /*first row of the range*/, const int32_t* indices, const float* data,
                            int n, const float* W1T, const float* b1, int H0, float* Z, void* stream);
int scnb_csr_linear_bwd_dw(const int32_t* b_indptr, const int32_t* b_idx, const float* b_val, int n, const float* dZ,
                           int H0, float* dW1T, void* stream);

/* ---- K2b: dense tensor-core form of the first layer (densify-then-tcgen05 GEMM, bf16x3 split with fp32
 *      accumulation in TMEM).  scnb_w1_planes splits W1T into two bf16 planes laid out as 128-byte-swizzled
 *      shared-memory tile images (ceil(G/64) tiles of H0*64 elements per plane); scnb_csr_linear_fwd_tc{,_i64}
 *      computes Z = X W1T + b1 from a batch CSR (int32 indptr) or a row range of a dataset CSR (int64 indptr).
 *      ksplits == 0: choose the number of gene-axis splits automatically; ksplits < 0: the same, and the caller has
 *      already zeroed Z (the partial products of the splits are added atomically).  H0 in {32,...,256}, multiple of 32. ---- */
int scnb_w1_planes(const float* W1T, int G, int H0, void* planes_hi, void* planes_lo, void* stream);
int scnb_csr_linear_fwd_tc(const int32_t* b_indptr, const int32_t* b_idx, const float* b_val, int n, int G, int H0,
                           const void* planes_hi, const void* planes_lo, const float* b1, float* Z, int ksplits, void* stream);
int scnb_csr_linear_fwd_tc_i64(const int64_t* indptr, const int32_t* indices, const float* data, int n, int G, int H0,
                               const void* planes_hi, const void* planes_lo, const float* b1, float* Z, int ksplits,
                               void* stream);
/* The split forward with its per-row starting points precomputed (scnym/model.py:211 on the batch CSR, same result as
 * scnb_csr_linear_fwd_tc with automatic splits): scnb_csr_linear_fwd_tc_geometry reports the number of gene-axis splits
 * and the genes per split for a batch of n rows; scnb_row_fwd_splits fills fwd_splits[(ksplits + 1) * n] (position of
 * the first entry of row r with gene >= s * genes_per_split); scnb_csr_linear_fwd_tc_splits (ksplits <= 0 as above) then
 * starts every CTA at its entry without a binary search at the head of the step. */
int scnb_csr_linear_fwd_tc_geometry(int n, int G, int H0, int* ksplits, int* genes_per_split);
int scnb_row_fwd_splits(const int32_t* b_indptr, const int32_t* b_idx, int n, int G, int H0, int32_t* fwd_splits, void* stream);
int scnb_csr_linear_fwd_tc_splits(const int32_t* b_indptr, const int32_t* b_idx, const float* b_val, int n, int G, int H0,
                                  const void* planes_hi, const void* planes_lo, const float* b1, float* Z, int ksplits,
                                  const int32_t* fwd_splits, void* stream);

/* Tensor-core form of the fused dW1 + optimizer kernel (replaces the backward of nn.Linear in scnym/model.py:211 plus
 * torch.optim's step on that weight, scnym/main.py:374-396): D^T[H0, genes] = dZ1^T X as a tcgen05 GEMM (bf16x3, fp32 in
 * TMEM, one CTA per SM owning scnb_w1tc_gene_block(G) consecutive genes) with the optimizer in the epilogue.
 * scnb_row_gene_splits finds, for every batch row, where each gene block starts (splits: [(ceil(G/block)+1) * nb] int32);
 * the dZ1 operand is given as bf16 planes built by scnb_w1_planes(dZ1, nb, H0, ...).  opt / grad_out as in
 * scnb_csr_w1_bwd_optim.  planes_hi_out / planes_lo_out (both or neither; opt >= 0 only): the updated W1T is also
 * written as the bf16 operand planes scnb_csr_linear_fwd_tc reads, so the next step needs no scnb_w1_planes pass. */
int scnb_w1tc_gene_block(int G);
int scnb_row_gene_splits(const int32_t* b_indptr, const int32_t* b_idx, int nb, int G, int32_t* splits, void* stream);
/* scnb_row_gene_splits and scnb_row_fwd_splits in one launch (both tables depend only on the batch). */
int scnb_row_gene_splits2(const int32_t* b_indptr, const int32_t* b_idx, int nb, int G, int H0, int32_t* splits,
                          int32_t* fwd_splits, void* stream);
int scnb_csr_w1_bwd_optim_tc(int opt, const int32_t* splits, const int32_t* b_idx, const float* b_val, int nb, int H0, int G,
                             const void* dz_planes_hi, const void* dz_planes_lo, float* W1T, float* grad_out, float* state1,
                             float* state2, const float* hp8, const int64_t* state4, void* planes_hi_out, void* planes_lo_out,
                             void* stream);

/* Tensor-core hidden layers for batched predict: nn.Linear weights as B planes (scnb_linear_planes), eval-mode
 * BatchNorm->ReLU that also emits its output as tiled bf16 A planes (scnb_bn_eval_planes; A_out / pre_out / planes
 * may each be NULL), and Y = A W^T + b from planes only (scnb_linear_fwd_tc_planes).  Plane sizes in bf16 elements:
 * A: ceil(n/256)*256 * K;  W: N * ceil(K/32)*32. */
int scnb_linear_planes(const float* W /*[N,K]*/, int N, int K, void* planes_hi, void* planes_lo, void* stream);
int scnb_bn_eval_planes(const float* Z, int n_rows, int H, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, int relu, float* A_out, float* pre_out, void* planes_hi, void* planes_lo,
                        void* stream);
int scnb_linear_fwd_tc_planes(const void* a_hi, const void* a_lo, int n, int K, int N, const void* w_hi, const void* w_lo,
                              const float* bias, float* Y, void* stream);
/* scnb_csr_linear_fwd_tc_i64 (or scnb_linear_fwd_tc_planes) + scnb_bn_eval_planes in ONE launch: a layer of batched predict
 * (scnym/model.py:211-224, eval mode) ends in the operand planes of the next Linear -- fp32 Z is neither written nor read
 * back.  Bit-identical planes. */
int scnb_linear_fwd_tc_planes_bn_planes(const void* a_hi, const void* a_lo, int n, int K, int N, const void* w_hi, const void* w_lo,
                                        const float* bias, const float* gamma, const float* beta, const float* running_mean,
                                        const float* running_var, int relu, void* out_hi, void* out_lo, void* stream);
/* ... and the LAST hidden Linear + BatchNorm + ReLU + classifier (scnym/model.py:218-229) in one launch: planes in, logits [n, C]
 * (+= when accumulate: ensembles) and, optionally, the pre-ReLU BatchNorm output (the embedding, scnym/api.py:700) out.  C <= 16. */
int scnb_linear_fwd_tc_planes_bn_head(const void* a_hi, const void* a_lo, int n, int K, int N, const void* w_hi, const void* w_lo,
                                      const float* bias, const float* gamma, const float* beta, const float* running_mean,
                                      const float* running_var, int relu, const float* Wc, const float* bc, int C, float* logits,
                                      int accumulate, float* pre_out, void* stream);
int scnb_csr_linear_fwd_tc_bn_planes_i64(const int64_t* indptr, const int32_t* indices, const float* data, int n, int G, int H0,
                                         const void* planes_hi, const void* planes_lo, const float* b1, const float* gamma,
                                         const float* beta, const float* running_mean, const float* running_var, int relu,
                                         void* a_hi_out, void* a_lo_out, void* stream);
/* out[n,C] (+)= A[n,H] W[C,H]^T + b for a small C (classifier logits of predict; one warp per row). */
int scnb_head_logits(const float* A, const float* W, const float* b, int H, int C, int n, float* out, int accumulate,
                     void* stream);

/* ---- K8+K9 fused: first-layer weight gradient AND optimizer.step() for W1T in one pass
 *      (scnym/model.py:211 backward + scnym/trainer.py:671 / torch.optim rules, scnym/main.py:36-41,374-396).
 *      The batch is first transposed to gene-major order ("batch CSC"): gene_cnt[G] is the histogram
 *      accumulated by scnb_csr_gather_rows (or scnb_csc_count); scnb_csr_to_csc scans it into
 *      c_ptr[G+1], resets gene_cnt to 0 and fills c_ent[nnz] = {int32 row, fp32 value} pairs.
 *      scnb_csr_w1_bwd_optim then computes, per gene g, grad = sum v*dZ[row,:] and applies optimizer
 *      `opt` to W1T[g,:], state1[g,:], state2[g,:] (opt = -1: gradient only, written to grad_out).
 *      grad_out may be NULL when opt >= 0 (the gradient then never reaches HBM).  Requires H0 % 64 == 0. ---- */
int scnb_csc_count(const int32_t* b_idx, const float* b_val, int nnz, int32_t* gene_cnt, void* stream);
int scnb_csr_to_csc(const int32_t* b_indptr, const int32_t* b_idx, const float* b_val, int n, int G, int32_t* gene_cnt,
                    int32_t* c_ptr /*[G+1]*/, int32_t* cursor /*[G] scratch*/, void* c_ent /*[nnz] 8-byte pairs*/,
                    void* stream);
int scnb_csr_w1_bwd_optim(int opt, const int32_t* c_ptr, const void* c_ent, const float* dZ /*[nb,H0]*/, int nb, int H0,
                          int G, float* W1T, float* grad_out, float* state1, float* state2, const float* hp8,
                          const int64_t* state4, void* stream);

/* ---- K5: hidden nn.Linear layers, classifier and DAN head (scnym/model.py:218,170,229,377-383)
 *      and their backward GEMMs.  C = alpha*op(A)op(B) + beta*C (+bias[N]) (+ReLU). ---- */
int scnb_sgemm(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
               int ldc, float alpha, float beta, const float* bias, int relu, const float* gate /*ReLU-bwd mask, C layout*/,
               void* stream);
/* scnb_sgemm sends products of at least `flops` = 2 M N K (default 2e9, 0 = never) to the tcgen05 kernel (bf16x3 split,
 * fp32 accumulation in TMEM, operands converted from fp32 inside the kernel, or beforehand into planes -- below); smaller ones keep the 3xTF32 mma.sync kernel. */
int scnb_gemm_tc_min_flops(long long flops);
int scnb_gemm_tc_min_flops_get(long long* flops_out);
/* Planes form of the tcgen05 product (tc_gemm.cu): when the caller has registered a workspace for the stream a large product
 * is issued on, both operands are first written once as bf16 hi / lo planes tiled [256 rows][32 k] in the shared-memory image
 * the MMA wants (scnb_gemm_tc_workspace_bytes(M, N, K) bytes), and the product itself only bulk-copies tiles and issues MMAs
 * (256 x 256 tile of C per CTA, two accumulators in TMEM).  The library allocates nothing: the engine owns the buffers
 * (1024-byte aligned) and registers them before any graph capture; buf = NULL unregisters the stream. */
int scnb_gemm_tc_workspace(void* stream, void* buf, long long bytes);
int scnb_gemm_tc_workspace_bytes(int M, int N, int K, long long* bytes_out);
int scnb_gemm_tc_planes_count(long long* count_out);   /* products issued in the planes form so far */
/* Planes written by the PRODUCER of a tensor: after scnb_gemm_tc_bind_planes(X, rows, cols, hi, lo) (cols a multiple of 32,
 * planes of ceil(rows / 256) * 256 x cols bf16 each, 1024-byte aligned) scnb_bn_act_fwd (output A == X) and scnb_bn_act_bwd
 * (output dZ == X) also write X as operand planes, and a product that reads X as its row-major A operand (lda == cols) skips
 * the conversion pass; a weight-gradient product (transA) whose two operands are both bound reads those same planes as
 * MN-major operands.  hi = lo = NULL unbinds.  scnb_gemm_tc_bound_count: operands taken from bound planes so far. */
int scnb_gemm_tc_bind_planes(const float* X, int rows, int cols, void* hi, void* lo);
int scnb_gemm_tc_bound_count(long long* count_out);
int scnb_colsum(const float* A, int M, int N, int lda, float* out, int accumulate, void* stream);
int scnb_relu_bwd(const float* g, const float* y, float* out, long long n, void* stream);

/* ---- K4: nn.BatchNorm1d -> nn.Dropout -> nn.ReLU (scnym/model.py:212-217, 219-224, 173-182),
 *      train mode with per-segment batch statistics or eval mode with running statistics.
 *      stats: [n_seg][3][H] = mean, biased var, invstd.  pre_out (optional) receives the
 *      pre-ReLU value (api.py:700 embedding = last BN output). ---- */
int scnb_bn_act_fwd(const float* Z, float* A, float* pre_out, int H, int n_seg, const int32_t* seg_off_dev, int max_rows,
                    int train, const float* gamma, const float* beta, const float* running_mean,
                    const float* running_var, float* stats, const uint8_t* keep_mask, const uint64_t* rng,
                    uint32_t stream_id, float p_drop, int relu,
                    float* dp_partial_out /*[n_seg*2*H + n_seg]: local sum z, sum z^2, row counts; no apply*/,
                    const float* dp_sums /*same layout after the cross-rank all-reduce*/, void* stream);
int scnb_bn_act_bwd(const float* dA, const float* Z, const float* A, int H, int n_seg, const int32_t* seg_off_dev,
                    int max_rows, const float* gamma, const float* stats, const uint8_t* keep_mask, const uint64_t* rng,
                    uint32_t stream_id, float p_drop, int relu, float* dZ, float* dgamma /*+=*/, float* dbeta /*+=*/,
                    float* dp_partial_out /*local sum dy, sum dy*zhat, counts*/, const float* dp_sums, void* stream);

/* ---- data-parallel BatchNorm in ONE launch: the per-segment column sums are all-reduced inside the kernel through
 *      peer-addressable exchange buffers (CUDA IPC over NVLink / NVSwitch), replacing the partial -> NCCL all-reduce ->
 *      apply protocol of scnb_bn_act_fwd/bwd.  Equivalence target: nn.BatchNorm1d of scnym/model.py:217 on the
 *      concatenated global batch.  peer_bufs: device array [world] of the ranks' exchange buffers as
 *      addressable from this process (own buffer at [rank]); each buffer holds scnb_bn_peer_bytes(H, n_seg, n_slots,
 *      world) bytes, zeroed before the first step.  slot: index of this BatchNorm call within the step (< n_slots);
 *      epoch: device scalar > 0 that changes every step (the optimizer step count).  global_sums_out (may be NULL):
 *      [n_seg][2][H] global sums + [n_seg] global row counts, as scnb_bn_running_update's dp_counts expects.
 *      scnb_peer_alloc/open/close/free: exchange buffers as cudaMalloc memory + 64-byte CUDA IPC handles. ---- */
int scnb_bn_peer_bytes(int H, int n_seg, int n_slots, int world);   /* returns the byte count (> 0), not a status */
int scnb_bn_act_fwd_peer(const float* Z, float* A, float* pre_out, int H, int n_seg, const int32_t* seg_off_dev, int max_rows,
                         const float* gamma, const float* beta, float* stats, const uint8_t* keep_mask, const uint64_t* rng,
                         uint32_t stream_id, float p_drop, int relu, const uint64_t* peer_bufs, int rank, int world, int slot,
                         int n_slots, const int64_t* epoch, long long buf_bytes, float* global_sums_out, void* stream);
int scnb_bn_act_bwd_peer(const float* dA, const float* Z, const float* A, int H, int n_seg, const int32_t* seg_off_dev,
                         int max_rows, const float* gamma, const float* stats, const uint8_t* keep_mask, const uint64_t* rng,
                         uint32_t stream_id, float p_drop, int relu, float* dZ, float* dgamma, float* dbeta,
                         const uint64_t* peer_bufs, int rank, int world, int slot, int n_slots, const int64_t* epoch,
                         long long buf_bytes, void* stream);
int scnb_peer_alloc(long long bytes, void** ptr_out, void* handle64_out);
int scnb_peer_open(const void* handle64, void** ptr_out);
int scnb_peer_close(void* ptr);
int scnb_peer_free(void* ptr);
/* Host memory that kernels can store into (cudaHostAllocMapped) and a <= 256-float publish kernel: the per-step loss
 * read-back of the host-buffer pipeline (replaces the `loss.item()` of scnym/trainer.py:666-681) without a copy-engine
 * transfer. */
int scnb_host_alloc_mapped(long long bytes, void** host_ptr_out, void** dev_ptr_out);
int scnb_host_free_mapped(void* host_ptr);
int scnb_publish_f32(const float* src, int n, float* dst_mapped, void* stream);
/* Page-lock + map an existing host array (cudaHostRegisterMapped): dev_ptr_out is what kernels dereference.  Feeds the row
 * gather (scnb_step_prep / scnb_csr_gather_rows2) straight from host-resident CSR matrices -- the host-side minibatch
 * assembly of scnym/dataprep.py:114-168 + the per-step copy of scnym/trainer.py:589-599 without host threads. */
int scnb_host_register_mapped(void* host_ptr, long long bytes, void** dev_ptr_out);
int scnb_host_unregister(void* host_ptr);

int scnb_bn_eval_bwd(const float* dA, const float* A, int relu, int rows, int H, const float* gamma,
                     const float* running_var, float* dZ, void* stream);
int scnb_bn_running_update(float* running_mean, float* running_var, int64_t* num_batches_tracked, const float* stats0,
                           const float* stats1, const float* stats2, const float* stats3 /*shared block: one per application*/,
                           const int32_t* seg_off_dev, int n_seg, int H, float momentum,
                           const float* dp_cnt /*NULL or global rows per segment*/, void* stream);

/* ---- K3: SampleMixUp (scnym/dataprep.py:638-705) folded behind the first layer (linearity):
 *      out[i] = gam[i] Z[srcA[i]] + (1-gam[i]) Z[srcB[i]];  adjoint gathers through inverse maps. ---- */
int scnb_mix_rows(const float* Z, int H, const int32_t* srcA, const int32_t* srcB, const float* gam,
                  const int32_t* n_active_dev, int max_rows, float* out, void* stream);
int scnb_unmix_rows(const float* dOut, int H, const int32_t* inv3 /*[3][n_base]*/, const float* coef3, int n_base,
                    float* dZ, void* stream);
/* scnb_unmix_rows that also writes dZ as the bf16 hi/lo operand planes scnb_csr_w1_bwd_optim_tc reads (the layout
 * scnb_w1_planes(dZ, n_base, H, ...) produces; planes hold ceil(n_base/32)*32*H bf16 each, zeroed once by the caller). */
int scnb_unmix_rows_planes(const float* dOut, int H, const int32_t* inv3, const float* coef3, int n_base, float* dZ,
                           void* planes_hi, void* planes_lo, void* stream);

/* ---- Fused hidden stack (hidden_ops.cu): `[Linear -> BatchNorm1d(train) -> Dropout -> ReLU] x n` of CellTypeCLF.embed
 *      (scnym/model.py:210-233) on the student super-batch and its adjoint, with BatchNorm -> Dropout -> ReLU applied while
 *      the GEMM's A operand is staged in shared memory.  Unit of work: a 32-row x 64-column tile; BatchNorm statistics
 *      travel between launches as per-tile partials `part[rows/32][2][H]` ((mean, M2) forward, (sum dY, sum dY*zhat)
 *      backward) that every consumer combines per segment in a fixed order.  seg_off_host: HOST int32 [n_seg + 1] row
 *      offsets (multiples of 32, n_seg <= 3); stats: [n_seg][3][H] = mean, biased var, invstd (as scnb_bn_act_fwd);
 *      dropout stream ids / injected masks as scnb_bn_act_fwd.  All buffers 16-byte aligned, H % 64 == 0. ---- */
/* out = MixUp(Z) rows (dataprep.py:665-689; srcA == NULL: out = Z) + tile statistics of out.  keep_out (NULL or uint8 [R, H]):
 * also draws the dropout keep bytes of application 0 (Philox stream `stream_id` of (rng[0], rng[1]), keep probability 1-p_drop):
 * the kernel that PRODUCES a pre-BatchNorm tensor emits its dropout mask once, every consumer of those rows reads bytes. */
int scnb_hs_mix(const float* Z, int H, const int32_t* srcA, const int32_t* srcB, const float* gam, int R, float* out,
                float* part, uint8_t* keep_out, const uint64_t* rng, uint32_t stream_id, float p_drop, void* stream);
/* A = relu(drop(bn(Z))) of the last application (input of the heads); writes stats */
int scnb_hs_act(const float* Z, float* A, int H, int R, int n_seg, const int32_t* seg_off_host, const float* part,
                const float* gamma, const float* beta, float* stats, const uint8_t* keep_mask, const uint64_t* rng,
                uint32_t stream_id, float p_drop, void* stream);
/* Aout = relu(drop(bn(Zin))) (materialised; keep_mask = keep bytes [R, K] of that dropout, generated by the producer of Zin or
 * injected), Zout = Aout W^T + bias (W [N, K]; 3xTF32), tile statistics of Zout; writes stats_in; keep_out as scnb_hs_mix (the
 * mask of the application whose input Zout is).  running_mean / running_var non-NULL: eval mode (the MixMatch teacher,
 * scnym/losses.py:660-737): running statistics, no dropout; part_in / stats_in / Aout / part_out may then be NULL.
 * part_in == NULL in train mode: the (global, data-parallel) statistics are already in stats_in. */
int scnb_hs_fwd(const float* Zin, float* Aout, const float* part_in, float* stats_in, const float* gamma, const float* beta,
                const uint8_t* keep_mask, float p_drop, const float* W, const float* bias, float* Zout, float* part_out,
                uint8_t* keep_out, const uint64_t* rng, uint32_t out_stream_id, float out_p_drop, int K, int N, int R, int n_seg,
                const int32_t* seg_off_host, const float* running_mean, const float* running_var, void* stream);
/* Heads fused with the ends of the hidden stack (head_ops.cu), one launch per branch from the pre-BatchNorm Z of the LAST
 * application to dY of that application and its tile sums (the input of scnb_hs_bwd) -- replaces scnb_hs_act, the head
 * kernels and scnb_hs_bwd_act.  Common arguments: Z / A [R, H], part_in (tile statistics of Z; NULL: `stats` already holds the
 * global segment statistics), stats (written unless given), gamma / beta of the last BatchNorm, keep = dropout keep bytes [R, H]
 * of that application (scnb_hs_fwd's keep_out), dY [R, H], part_out [R/32][2][H].  Both return -3 (SCNB_ERR_UNSUPPORTED) for
 * shapes they are not built for (H not 128 / 256, more than 32 classes / domains).
 *   scnb_hs_classif_head: rows [0, U + L) (whole segments): activation (scnym/model.py:219-227), classifier (model.py:229),
 *     MSE on rows [0, U) / soft-target CE on rows [U, U + L) with dLogits (scnym/losses.py:880-923; arguments as
 *     scnb_classif_mixmatch_fused), dA = dLogits Wc, dY = dA [A > 0] / (1-p).
 *   scnb_hs_dan_head: rows [row_base, R) (whole segments): activation, Hd = relu(A W1^T + b1), domain logits Hd W2^T + b2, CE
 *     against the domain label (scnym/model.py:377-383, scnym/losses.py:1209-1243; arguments as scnb_dan_tail_fused),
 *     dHd = (ddlog W2) [Hd > 0], gradient reversal dA = -dan_weight dHd W1 (model.py:338), dY.  Hd / dHd [R - row_base, H] and
 *     dlog / dlabel / ddlog [R - row_base, D] are kept for the head's weight gradients.  A thread-block cluster of H/64 CTAs
 *     shares a 32-row tile through distributed shared memory. */
int scnb_hs_classif_head(const float* Z, float* A, const float* part_in, float* stats, const float* gamma, const float* beta,
                         const uint8_t* keep, float p_drop, int H, int R, int n_seg, const int32_t* seg_off_host, const float* Wc,
                         const float* bc, int C, int U, int L, const int32_t* n_conf_dev, const float* Y, const int32_t* mapA,
                         const int32_t* mapB, const float* gam, const float* class_weight, float unsup_weight, float* logits,
                         float* dlogits, float* loss8, float* dY, float* part_out, void* stream);
int scnb_hs_dan_head(const float* Z, float* A, const float* part_in, float* stats, const float* gamma, const float* beta,
                     const uint8_t* keep, float p_drop, int H, int R, int row_base, int n_seg, const int32_t* seg_off_host,
                     const float* W1, const float* b1, const float* W2, const float* b2, int D, const int32_t* n_rows_dev,
                     const int32_t* dom_ids, const int32_t* src, int src_off, const float* loss_scale_dev, float dan_weight,
                     float* Hd, float* dHd, float* dlog, float* dlabel, float* ddlog, float* loss2, float* dY, float* part_out,
                     void* stream);
/* dY = dA * [A > 0] / (1-p) of the last application + tile sums */
int scnb_hs_bwd_act(const float* dA, const float* Z, const float* A, int H, int R, int n_seg, const int32_t* seg_off_host,
                    const float* stats, float p_drop, float* dY, float* part, void* stream);
/* dZout = BatchNorm backward of dYin (materialised; dgamma / dbeta += segment totals), dA = dZout W (W [K, N]),
 * dYout = dA * [Aprev > 0] / (1-p_prev) + tile sums against (Zprev, stats_prev) */
int scnb_hs_bwd(const float* dYin, const float* Zin, const float* stats_in, const float* part_in, const float* gamma,
                float* dgamma, float* dbeta, float* dZout, const float* W, const float* Zprev, const float* Aprev,
                const float* stats_prev, float p_drop_prev, float* dYout, float* part_out, int K, int N, int R, int n_seg,
                const int32_t* seg_off_host, const float* bmean_in /* data parallel: global means instead of part_in */, void* stream);
/* dW[M, N] += dZ[R, M]^T A[R, N], db[M] += column sums of dZ (db may be NULL): weight / bias gradient of a hidden Linear
 *  in one launch; accumulates with atomics into the zeroed gradient arena; M, N % 64 == 0 */
int scnb_hs_dw(const float* dZ, int M, const float* A, int N, int R, float* dW, float* db, void* stream);
/* dZ of application 0 from dY (feeds the MixUp adjoint and the first-layer weight gradient) */
int scnb_hs_bwd_in(const float* dY, const float* Z, int H, int R, int n_seg, const int32_t* seg_off_host, const float* stats,
                   const float* part, const float* gamma, float* dgamma, float* dbeta, float* dZ, const float* bmean_in,
                   float* db /* NULL or [H]: += column sums of dZ (bias gradient of the Linear in front) */, void* stream);
/* scnb_hs_bwd_in, scnb_unmix_rows and the operand planes of scnb_unmix_rows_planes in one launch: dZ1[r] = sum_k coef3[k][r] *
 * dZ0[inv3[k][r]] with dZ0 (the BatchNorm backward of application 0) formed on the fly and never written; dgamma / dbeta / db as
 * scnb_hs_bwd_in; planes as scnb_unmix_rows_planes (MixMatch steps on the tensor-core dW1 path). */
int scnb_hs_bwd_in_unmix_planes(const float* dY, const float* Z, int H, int R, int n_seg, const int32_t* seg_off_host,
                                const float* stats, const float* part, const float* gamma, float* dgamma, float* dbeta,
                                const float* bmean_in, float* db, const int32_t* inv3, const float* coef3, int n_base, float* dZ1,
                                void* planes_hi, void* planes_lo, void* stream);
/* Data parallel (one process per GPU; equivalence target: the reference step on the CONCATENATED global batch): exchange of
 * the per-segment statistics between a producer of tile partials and their consumer, through peer memory (CUDA IPC buffers
 * over NVLink / NVSwitch, `peer_bufs[world]` as scnb_bn_act_fwd_peer).  kind 0 (forward): out = stats[n_seg][3][H] (global
 * mean, biased var, invstd), cnt_out[n_seg] = global rows; kind 1 (backward): out = bmean[n_seg][2][H] (global means of dY and
 * dY*zhat), and this rank's LOCAL sums are added to dgamma / dbeta.  Consumers then take part_in = NULL. */
int scnb_hs_debug_read(unsigned long long* out32_host);   /* SCNB_HS_DBG=1: %globaltimer phase stamps of CTA (0,0) (tools/hs_phases.py) */
int scnb_hh_debug_read(unsigned long long* out32_host);   /* the same for the fused head kernels (tools/head_phases.py) */
int scnb_hs_peer_bytes(int H_max, int n_seg, int n_slots, int world);   /* bytes of the exchange region */
int scnb_hs_xchg(int kind, const float* part, int H, int R, int n_seg, const int32_t* seg_off_host, const uint64_t* peer_bufs,
                 int rank, int world, int slot, int n_slots, int H_max, const int64_t* epoch, long long buf_bytes, float* out,
                 float* cnt_out, float* dgamma, float* dbeta, void* stream);

/* ---- K6: MixMatchLoss._generate_labels + sharpen_labels (scnym/losses.py:660-737, 529-573) and
 *      the confident/unconfident split, MixUp permutation and DAN row selection
 *      (scnym/losses.py:811-875, 1195-1213).  perm_keys stand in for torch.randperm
 *      (ridx = stable argsort(keys[:n_conf+L])), gamma_raw for Beta(alpha,alpha).sample(). ---- */
int scnb_pseudolabel(const float* teacher_logits /*[K][U][C]*/, int K, int U, int C, float T, int sharpen,
                     float min_confidence, float* q, float* conf, uint8_t* confident, float* scratch_UC,
                     const int32_t* lab_ids /*NULL or [L]: also write one-hot label rows q[U..U+L)*/, int L, void* stream);
int scnb_mixmatch_plan(const uint8_t* confident, const float* perm_keys, const float* gamma_raw, int U, int L, int use_dan,
                       int dan_conf_only, int mixup_on, int32_t* counts8, int32_t* seg_off4,
                       int32_t* work_i /*[U + 3(U+L)]*/, int32_t* srcA, int32_t* srcB, float* gam, int Rmax, int32_t* inv3,
                       float* coef3, float* pconf1, void* stream);

/* ---- K7: losses with fused dLogits.  cross_entropy (scnym/losses.py:67-151; supervised call
 *      :920-923, domain call :1224-1243), MSE on probabilities with confidence masking
 *      (scnym/main.py:504-505, scnym/losses.py:880-913).  loss_out[0] += loss, loss_out[1] += hits. ---- */
int scnb_soft_ce_fwd_bwd(const float* logits, int n_max, int C, const int32_t* n_rows_dev, const float* Y,
                         const int32_t* mapA, const int32_t* mapB, const float* gam, int map_off,
                         const float* class_weight, const int32_t* n_norm_dev, int n_norm, float grad_scale,
                         const float* loss_scale_dev, float* dlogits, float* loss_out2, int count_acc,
                         float* row_loss_out /*optional [n_max]: unreduced losses*/, void* stream);
int scnb_softmax_mse_fwd_bwd(const float* logits, int n_max, int C, const int32_t* n_conf_dev, const float* Y,
                             const int32_t* mapA, const int32_t* mapB, const float* gam, int map_off, int n_norm,
                             const float* weight_dev, float weight, float* dlogits, float* loss_out, void* stream);
int scnb_onehot_rows(const int32_t* ids, const int32_t* src, int src_off, int n, int C, float* out, void* stream);

/* ---- Fused heads (one warp per row; weights staged in shared memory).  Each replaces a
 *      Linear -> loss -> Linear-backward chain of four launches on the step's critical path:
 *      classifier + MixMatch losses + dA (scnym/model.py:229, scnym/losses.py:880-923);
 *      classifier + supervised CE + dA (scnym/trainer.py:215-233);
 *      DAN output layer + domain CE + ReLU-gated dHd (scnym/model.py:377-383, scnym/losses.py:1209-1243);
 *      teacher classifier + pseudolabels (scnym/losses.py:676-735).
 *      Need H % 128 == 0, H <= 1024, C*H*4 <= 96 KB; return -3 (unsupported) otherwise. ---- */
int scnb_classif_mixmatch_fused(const float* A /*[U+L,H]*/, const float* Wc, const float* bc, int H, int C, int U, int L,
                                const int32_t* n_conf_dev, const float* Y, const int32_t* mapA, const int32_t* mapB,
                                const float* gam, const float* class_weight, float unsup_weight, float* logits, float* dlogits,
                                float* dA /*[U+L,H]*/, float* loss8, void* stream);
int scnb_classif_supervised_fused(const float* A, const float* Wc, const float* bc, int H, int C, int n, const float* Y,
                                  const float* class_weight, float* logits, float* dlogits, float* dA, float* loss8,
                                  void* stream);
int scnb_dan_tail_fused(const float* Hd /*[Rd,H] post-ReLU*/, const float* W2, const float* b2, int H, int D, int Rd,
                        const int32_t* n_rows_dev, const int32_t* dom_ids, const int32_t* src, int src_off,
                        const float* loss_scale_dev, float* dlog, float* dlabel, float* ddlog, float* dHd, float* loss2,
                        void* stream);
int scnb_teacher_head_fused(const float* TA /*[K*U,H]*/, const float* Wc, const float* bc, int H, int C, int K, int U, float T,
                            int sharpen, float min_confidence, float* t_logits, float* q, float* conf, uint8_t* confident,
                            float* scratch_UC, const int32_t* lab_ids, int L, void* stream);
/* scnb_teacher_head_fused on the PRE-BatchNorm output TZ [U, H] of the last hidden Linear: eval-mode BatchNorm (running
 * statistics, scnym/model.py:219-227 under model.eval(); scnym/losses.py:660-737) -> ReLU applied while a row is loaded.
 * One view (K == 1), C <= 32; returns -3 otherwise. */
int scnb_teacher_head_bn_fused(const float* TZ, const float* gamma, const float* beta, const float* running_mean,
                               const float* running_var, const float* Wc, const float* bc, int H, int C, int K, int U, float T,
                               int sharpen, float min_confidence, float* t_logits, float* q, float* conf, uint8_t* confident,
                               float* scratch_UC, const int32_t* lab_ids, int L, void* stream);

/* ---- K9/K10: optimizer.step() (scnym/trainer.py:234,671,1089; torch.optim.{Adadelta,Adam,AdamW,SGD}
 *      selected at scnym/main.py:36-41,374-396) over the flat parameter arena, and the mean-teacher
 *      EMA (scnym/losses.py:382-406).  opt: 0 adadelta, 1 adam, 2 adamw, 3 sgd(momentum).
 *      hp8 = [lr, weight_decay, beta1, beta2, eps, rho, momentum, 0] (device);
 *      state4 = [t, rng_seed, rng_counter, mixmatch_step] (device int64). ---- */
int scnb_step_begin(int64_t* state4, void* stream);
int scnb_mm_step_inc(int64_t* state4, void* stream);
/* optimizer.zero_grad() (scnym/trainer.py:608) for the accumulation targets of the step: zero up to two 16-byte aligned fp32
 * buffers of na / nb floats (multiples of 4) in one launch. */
int scnb_zero2(float* a, long long na, float* b, long long nb, void* stream);
int scnb_optim_step(int opt, float* params, const float* grads, float* state1, float* state2, long long n,
                    const float* hp8, const int64_t* state4, void* stream);
int scnb_record_step(const float* loss8, float* hist, int hist_cap, const float* conf, int U, float* conf_ring, int ring_cap,
                     const int64_t* state4, void* stream);
/* Data-parallel gradient exchange fused with the optimizer over peer memory (replaces ncclAllReduce(AVG) of the gradient
 * arena followed by scnb_optim_step; equivalence target: torch.optim.step() of scnym/main.py:374-396 on the gradient of
 * the global batch).  Every rank's exchange buffer (scnb_peer_alloc) holds a 1 KB flag block at off_flags, its gradient
 * arena at off_grad and its parameter arena at off_flat (`total` floats each, total % 4 == 0).  Rank r reduces slice r of
 * all ranks' gradients with peer loads, updates that slice (state1/state2: local arrays of `total` floats, only the
 * owned slice is touched) and stores the new parameters into every rank's parameter arena; the call returns (in stream
 * order) when all ranks have done so.  epoch: device scalar > 0 that changes every step. */
int scnb_dp_reduce_optim_bcast(int opt, const uint64_t* peer_bufs, int rank, int world, long long off_flags, long long off_grad,
                               long long off_flat, long long total, float* state1, float* state2, const float* hp8,
                               const int64_t* state4, const int64_t* epoch, void* stream);
/* Push form of the same exchange.  Ownership: rank r owns floats [r*S, (r+1)*S) of the arena, S = scnb_dp_slice_floats
 * (the even share rounded up to 256 floats).  scnb_csr_w1_bwd_push_tc is the data-parallel form of
 * scnb_csr_w1_bwd_optim_tc(opt = -1): every gradient row of W1T (arena floats [0, G*H0)) is stored, while the kernel runs,
 * into region `rank` of the OWNER's inbox (off_inbox of its exchange buffer, `world` regions of S floats) instead of a local
 * gradient arena.  scnb_dp_reduce_optim_bcast2 then takes the gradients of arena floats [0, pushed_floats) of its slice
 * from its own inbox (local memory) and pulls only the rest from the peers; off_inbox < 0: identical to
 * scnb_dp_reduce_optim_bcast.  The link then carries gradient stores during the dW1 kernel and parameter stores during
 * the exchange, instead of both (plus load round trips) inside the exchange. */
int scnb_dp_slice_floats(long long total, int world, long long* slice_floats_out);
int scnb_csr_w1_bwd_push_tc(const int32_t* splits, const int32_t* b_idx, const float* b_val, int nb, int H0, int G,
                            const void* dz_planes_hi, const void* dz_planes_lo, const uint64_t* peer_bufs, int rank, int world,
                            long long off_inbox, long long total, float* grad_out /* NULL, or also keep the local gradient */,
                            void* stream);
int scnb_dp_reduce_optim_bcast2(int opt, const uint64_t* peer_bufs, int rank, int world, long long off_flags, long long off_grad,
                                long long off_flat, long long off_inbox, long long pushed_floats, long long total, float* state1,
                                float* state2, const float* hp8, const int64_t* state4, const int64_t* epoch, void* stream);

int scnb_ema_update(float* teacher, const float* params, long long n, float decay, const int64_t* state4, void* stream);

/* ---- Predict head: Predicter.predict (scnym/predict.py:178-192): ensemble-mean logits,
 *      argmax, softmax. ---- */
int scnb_predict_head(const float* logits_sum, int n_models, int n, int C, int64_t* pred, float* prob, float* score,
                      void* stream);

/* ---- Gene re-indexing in front of predict: utils.build_classification_matrix (scnym/utils.py:155-233, called at
 *      scnym/api.py:670-675) on a CSR matrix resident in HBM.  col_map[sample column] = model column or -1 (built on the
 *      host by one hash join; must be injective on its non-negative entries).  Two calls, both asynchronous:
 *      scnb_csr_remap_count writes the indptr of the re-indexed matrix (out_indptr[n_rows+1], starting at 0) using
 *      `workspace` (scnb_csr_remap_workspace_len(n_rows) int64 words); scnb_csr_remap_fill writes every row's surviving
 *      entries sorted by model column (bitmap + prefix popcount per row, no sort).  indptr addresses `indices`/`data`
 *      absolutely, so a row range of a larger matrix is passed as indptr + first_row. ---- */
int scnb_csr_remap_workspace_len(long long n_rows);
int scnb_csr_remap_count(const int64_t* indptr, const int32_t* indices, const int32_t* col_map, int n_src_cols,
                         long long n_rows, int64_t* out_indptr, void* workspace, void* stream);
int scnb_csr_remap_fill(const int64_t* indptr, const int32_t* indices, const float* data, const int32_t* col_map,
                        int n_src_cols, int n_dst_cols, long long n_rows, const int64_t* out_indptr, int32_t* out_indices,
                        float* out_data, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCNYM_B200_H */
