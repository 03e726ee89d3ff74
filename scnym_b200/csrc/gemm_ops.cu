// Hidden-stack GEMMs of the scNym hot path (sm_100a): fp32 SIMT "panel" kernel with a cp.async pipeline.
//   nn.Linear forward   C[M,N] = A[M,K] W[N,K]^T + b        (NT; scnym/model.py:218,229,170,377-383)
//   activation gradient dA[M,K'] = dZ[M,N'] W[N',K']        (NN)
//   weight gradient     dW[N',K'] += dZ[rows,N']^T A[rows,K'] (TN, split over rows = K of the product)
// The products of this path are tiny (<= 1024 x 256 x 256 at batch 256), so a tile's K loop is a
// chain of L2 round trips unless several K chunks are in flight: operand panels of KC k-values are
// fetched with cp.async into a 3-stage shared-memory ring (zero-filled at the edges), and the tile
// shape is picked so that the grid covers the 148 SMs (down to 16x32).  fp32 FMA throughout: the
// parity budget (1e-4) rules out single-pass TF32/BF16.
#include "common.cuh"
#include <stdlib.h>

#define PK_THREADS 256
#define PK_STAGES 3
#define PK_PAD 4

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// panel whose shared-memory row is a matrix row (k contiguous in memory): [ROWS][KC + PAD]
template <int ROWS, int KC>
__device__ __forceinline__ void load_panel_rows(float* __restrict__ sm, const float* __restrict__ X, int ld, int row0, int nrows,
                                                int k0, int kend, int tid) {
  constexpr int CPR = KC / 4, S = KC + PK_PAD;
  for (int ch = tid; ch < ROWS * CPR; ch += PK_THREADS) {
    const int r = ch / CPR, kq = ch % CPR;
    const int gr = row0 + r, gk = k0 + kq * 4;
    const bool ok = gr < nrows && gk < kend;
    cp_async16(sm + r * S + kq * 4, ok ? X + (size_t)gr * ld + gk : X, ok);
  }
}
// panel whose shared-memory row is a k index (the m / n index is contiguous in memory): [KC][COLS + PAD]
template <int COLS, int KC>
__device__ __forceinline__ void load_panel_cols(float* __restrict__ sm, const float* __restrict__ X, int ld, int col0, int ncols,
                                                int k0, int kend, int tid) {
  constexpr int CPR = COLS / 4, S = COLS + PK_PAD;
  for (int ch = tid; ch < KC * CPR; ch += PK_THREADS) {
    const int k = ch / CPR, cq = ch % CPR;
    const int gk = k0 + k, gc = col0 + cq * 4;
    const bool ok = gk < kend && gc < ncols;
    cp_async16(sm + k * S + cq * 4, ok ? X + (size_t)gk * ld + gc : X, ok);
  }
}

template <int N>
struct VecN { float v[N]; };
template <int N>
__device__ __forceinline__ VecN<N> lds_vec(const float* p) {
  VecN<N> r;
  if constexpr (N == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (N == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    static_assert(N == 1, "micro-tile width must be 1, 2 or 4");
    r.v[0] = p[0];
  }
  return r;
}

// LAY: 0 = NT (A[M,K], B[N,K]),  1 = NN (A[M,K], B[K,N]),  2 = TN (A[K,M], B[K,N])
template <int BM, int BN, int KC, int LAY>
__global__ void __launch_bounds__(PK_THREADS) k_sgemm_panel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                            const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                                                            float alpha, float beta, const float* __restrict__ bias, int relu,
                                                            const float* __restrict__ gate, int k_chunk) {
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int SA = (LAY == 2) ? (BM + PK_PAD) : (KC + PK_PAD);   // row stride of the A panel
  constexpr int SB = (LAY == 0) ? (KC + PK_PAD) : (BN + PK_PAD);   // row stride of the B panel
  constexpr int A_FLOATS = (LAY == 2) ? KC * SA : BM * SA;
  constexpr int B_FLOATS = (LAY == 0) ? BN * SB : KC * SB;
  extern __shared__ __align__(16) float pk_smem[];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * k_chunk, ke = min(K, kb + k_chunk);
  const int nchunks = (ke - kb + KC - 1) / KC;
  auto stageA = [&](int st) { return pk_smem + st * (A_FLOATS + B_FLOATS); };
  auto stageB = [&](int st) { return pk_smem + st * (A_FLOATS + B_FLOATS) + A_FLOATS; };
  auto issue = [&](int chunk) {
    if (chunk < nchunks) {
      const int st = chunk % PK_STAGES, k0 = kb + chunk * KC;
      if (LAY == 2) load_panel_cols<BM, KC>(stageA(st), A, lda, m0, M, k0, ke, tid);
      else load_panel_rows<BM, KC>(stageA(st), A, lda, m0, M, k0, ke, tid);
      if (LAY == 0) load_panel_rows<BN, KC>(stageB(st), B, ldb, n0, N, k0, ke, tid);
      else load_panel_cols<BN, KC>(stageB(st), B, ldb, n0, N, k0, ke, tid);
    }
    cp_async_commit();  // always commit (possibly empty) so that the group count stays uniform
  };
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  issue(0);
  issue(1);
  for (int ch = 0; ch < nchunks; ++ch) {
    cp_async_wait<1>();   // chunk `ch` has landed (at most the newest group is still in flight)
    __syncthreads();      // ... for every thread; also: everyone is done with the buffer refilled next
    issue(ch + 2);
    const float* As = stageA(ch % PK_STAGES);
    const float* Bs = stageB(ch % PK_STAGES);
    if (LAY == 0) {
#pragma unroll 4
      for (int k4 = 0; k4 < KC; k4 += 4) {
        float4 a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * SA + k4);
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const float4*>(Bs + (tx + 16 * j) * SB + k4);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            float t = acc[i][j];
            t = fmaf(a[i].x, b[j].x, t);
            t = fmaf(a[i].y, b[j].y, t);
            t = fmaf(a[i].z, b[j].z, t);
            t = fmaf(a[i].w, b[j].w, t);
            acc[i][j] = t;
          }
      }
    } else if (LAY == 1) {
#pragma unroll 2
      for (int k4 = 0; k4 < KC; k4 += 4) {
        float4 a[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * SA + k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const VecN<TN> b = lds_vec<TN>(Bs + (k4 + kk) * SB + tx * TN);
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av, b.v[j], acc[i][j]);
          }
        }
      }
    } else {
#pragma unroll 8
      for (int k = 0; k < KC; ++k) {
        const VecN<TM> a = lds_vec<TM>(As + k * SA + ty * TM);
        const VecN<TN> b = lds_vec<TN>(Bs + k * SB + tx * TN);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a.v[i], b.v[j], acc[i][j]);
      }
    }
  }
  cp_async_wait<0>();
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ((LAY == 2) ? ty * TM + i : ty + 16 * i);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + ((LAY == 0) ? tx + 16 * j : tx * TN + j);
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (split) {  // C already holds beta*C; partial products are accumulated atomically
        atomicAdd(C + (size_t)m * ldc + n, v);
        continue;
      }
      if (beta != 0.f) v += beta * C[(size_t)m * ldc + n];
      if (bias) v += bias[n];
      if (relu) v = fmaxf(v, 0.f);
      if (gate && !(gate[(size_t)m * ldc + n] > 0.f)) v = 0.f;  // ReLU backward gate (same layout as C)
      C[(size_t)m * ldc + n] = v;
    }
  }
}

template <int BM, int BN, int KC, int LAY>
static int launch_panel(dim3 grid, cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                        float* C, int ldc, float alpha, float beta, const float* bias, int relu, const float* gate, int k_chunk) {
  constexpr int SA = (LAY == 2) ? (BM + PK_PAD) : (KC + PK_PAD);
  constexpr int SB = (LAY == 0) ? (KC + PK_PAD) : (BN + PK_PAD);
  constexpr int A_FLOATS = (LAY == 2) ? KC * SA : BM * SA;
  constexpr int B_FLOATS = (LAY == 0) ? BN * SB : KC * SB;
  constexpr int smem = PK_STAGES * (A_FLOATS + B_FLOATS) * (int)sizeof(float);
  static bool attr_set = false;  // per instantiation; the attribute is idempotent, a race is harmless
  if (smem > 48 * 1024 && !attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_sgemm_panel<BM, BN, KC, LAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      scnb_set_error("sgemm: cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return SCNB_ERR_CUDA;
    }
    attr_set = true;
  }
  k_sgemm_panel<BM, BN, KC, LAY><<<grid, PK_THREADS, smem, st>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate,
                                                               k_chunk);
  return SCNB_OK;
}

// Returns 1 if the product was launched here, 0 if the shape/alignment needs the generic kernel, <0 on error.
int scnb_sgemm_panel_try(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                         int ldc, float alpha, float beta, const float* bias, int relu, const float* gate, cudaStream_t st) {
  if (transA && transB) return 0;
  const int lay = transA ? 2 : (transB ? 0 : 1);
  // 16-byte cp.async chunks: contiguous extents must be multiples of 4 floats and 16-byte aligned
  if ((lda % 4) || (ldb % 4) || (((uintptr_t)A | (uintptr_t)B) % 16)) return 0;
  if (lay == 0 && (K % 4)) return 0;
  if (lay == 1 && ((K % 4) || (N % 4))) return 0;
  if (lay == 2 && ((M % 4) || (N % 4))) return 0;
  auto ctas = [&](int a, int b) { return (long long)scnb_cdiv(M, a) * scnb_cdiv(N, b); };
  int bm = 64, bn = 64, splits = 1, k_chunk = K;
  const bool splittable = !bias && !relu && !gate && (beta == 1.f) && K >= 128;
  if (lay == 2) {
    // weight-gradient shape: small output, long K -> 64x64 tiles split over K (KC = 32 per stage)
    if (splittable) {
      const long long c = ctas(64, 64);
      while (splits < 32 && c * splits < 2 * 148 && K / (splits * 2) >= 32) splits *= 2;
    }
    k_chunk = scnb_cdiv(scnb_cdiv(K, splits), 32) * 32;
    splits = scnb_cdiv(K, k_chunk);
    dim3 grid(scnb_cdiv(N, 64), scnb_cdiv(M, 64), splits);
    int rc = launch_panel<64, 64, 32, 2>(grid, st, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, k_chunk);
    return rc < 0 ? rc : 1;
  }
  // forward / activation-gradient shapes: largest tile that still gives >= 148 CTAs, else the smallest
  const int cand[4][2] = {{64, 64}, {32, 64}, {32, 32}, {16, 32}};
  int pick = 3;
  for (int i = 0; i < 4; ++i)
    if (ctas(cand[i][0], cand[i][1]) >= 148) { pick = i; break; }
  bm = cand[pick][0];
  bn = cand[pick][1];
  dim3 grid(scnb_cdiv(N, bn), scnb_cdiv(M, bm), 1);
  int rc;
#define PK_GO(BM_, BN_)                                                                                                      \
  rc = (lay == 0) ? launch_panel<BM_, BN_, 64, 0>(grid, st, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, K) \
                  : launch_panel<BM_, BN_, 64, 1>(grid, st, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, K)
  if (pick == 0) PK_GO(64, 64);
  else if (pick == 1) PK_GO(32, 64);
  else if (pick == 2) PK_GO(32, 32);
  else PK_GO(16, 32);
#undef PK_GO
  return rc < 0 ? rc : 1;
}

// ---------------------------------------------------------------------------------------------
// Tensor-core variant for the same three product shapes: 3xTF32 (error-compensated) mma.sync.
//   x = hi + lo with hi = trunc_tf32(x), lo = x - hi;  A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi
// keeps ~20 mantissa bits per product, which the 1e-4 parity budget needs (a single TF32 pass does
// not: SURVEY.md §7).  One CTA = one 32x32 output tile; its 4 warps split the K chunk held in shared
// memory four ways (each runs m16n8k8 over a 32x32 warp tile) and meet in shared memory, so the
// serial MMA chain per warp is K/32 steps and a 1024x256 output still spreads over 256 CTAs.
// Operand panels are fetched whole (K chunk <= 256) with cp.async: one L2 round trip per chunk.
// Fragment reads are bank-conflict free: row strides are = 4 (k-contiguous panels) or 8
// (k-row panels) modulo 32 words.
// ---------------------------------------------------------------------------------------------
#define TG_KC 256
#define TG_SR (TG_KC + 4)   // stride of a [32][KC] panel (k contiguous)
#define TG_SC 40            // stride of a [KC][32] panel (k rows)

// The tensor core ignores the low 13 mantissa bits of a tf32 operand register: hi = x as it is (read as trunc(x)),
// lo = x - trunc(x) (exact in fp32).  Two full-rate instructions per element; `cvt.rna.tf32.f32` is emulated on sm_100a
// with ~4 instructions per conversion, which made these loops issue-bound (profiles/r02_ncu_hidden_stack.txt).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x);
  lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// [32][kc] panel from a matrix whose rows are k-contiguous (fixed 64-chunk rows: shifts, no divisions)
__device__ __forceinline__ void tg_load_rows(float* __restrict__ sm, const float* __restrict__ X, int ld, int row0, int nrows,
                                             int k0, int kend, int kc, int tid, int nthr) {
  const int kq = tid & 63;               // 16-byte chunk within the row, fixed per thread (nthr is a multiple of 64)
  const int gk = k0 + kq * 4;
  if (kq * 4 >= kc) return;
  const bool kok = gk < kend;
  const float* src = X + (size_t)row0 * ld + gk;
  float* dst = sm + kq * 4;
  for (int r = tid >> 6; r < 32; r += nthr >> 6) {
    const bool ok = kok && (row0 + r) < nrows;
    cp_async16(dst + r * TG_SR, ok ? src + (size_t)r * ld : X, ok);
  }
}
// [kc][32] panel from a matrix stored [k][cols]
__device__ __forceinline__ void tg_load_cols(float* __restrict__ sm, const float* __restrict__ X, int ld, int col0, int ncols,
                                             int k0, int kend, int kc, int tid, int nthr) {
  const int cq = tid & 7;
  const int gc = col0 + cq * 4;
  const bool cok = gc < ncols;
  const float* src = X + (size_t)k0 * ld + gc;
  float* dst = sm + cq * 4;
  for (int k = tid >> 3; k < kc; k += nthr >> 3) {
    const bool ok = cok && (k0 + k) < kend;
    cp_async16(dst + k * TG_SC, ok ? src + (size_t)k * ld : X, ok);
  }
}

template <int LAY, int NW>
__global__ void __launch_bounds__(32 * NW) k_tgemm(int M, int N, int K, const float* __restrict__ A, int lda,
                                               const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc, float alpha,
                                               float beta, const float* __restrict__ bias, int relu,
                                               const float* __restrict__ gate, int k_chunk) {
  PDL_ENTRY();
  constexpr int A_FLOATS = (LAY == 2) ? TG_KC * TG_SC : 32 * TG_SR;
  constexpr int B_FLOATS = (LAY == 0) ? 32 * TG_SR : TG_KC * TG_SC;
  extern __shared__ __align__(16) float tg_smem[];
  float* As = tg_smem;
  float* Bs = tg_smem + A_FLOATS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int kb = blockIdx.z * k_chunk, ke = min(K, kb + k_chunk);
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
  for (int k0 = kb; k0 < ke; k0 += TG_KC) {
    const int kc = min(TG_KC, ((ke - k0) + 8 * NW - 1) / (8 * NW) * (8 * NW));  // chunk width, padded (zero-filled beyond ke)
    if (k0 > kb) __syncthreads();                        // previous chunk fully consumed
    if (LAY == 2) tg_load_cols(As, A, lda, m0, M, k0, ke, kc, tid, 32 * NW);
    else tg_load_rows(As, A, lda, m0, M, k0, ke, kc, tid, 32 * NW);
    if (LAY == 0) tg_load_rows(Bs, B, ldb, n0, N, k0, ke, kc, tid, 32 * NW);
    else tg_load_cols(Bs, B, ldb, n0, N, k0, ke, kc, tid, 32 * NW);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // this warp's share of the chunk
    const int kw = kc / NW;
    for (int kk = warp * kw; kk < (warp + 1) * kw; kk += 8) {
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float x0, x1, x2, x3;
        if (LAY == 2) {  // A stored [k][m]
          x0 = As[(kk + t) * TG_SC + 16 * i + g];
          x1 = As[(kk + t) * TG_SC + 16 * i + g + 8];
          x2 = As[(kk + t + 4) * TG_SC + 16 * i + g];
          x3 = As[(kk + t + 4) * TG_SC + 16 * i + g + 8];
        } else {         // A stored [m][k]
          x0 = As[(16 * i + g) * TG_SR + kk + t];
          x1 = As[(16 * i + g + 8) * TG_SR + kk + t];
          x2 = As[(16 * i + g) * TG_SR + kk + t + 4];
          x3 = As[(16 * i + g + 8) * TG_SR + kk + t + 4];
        }
        split_tf32(x0, ah[i][0], al[i][0]);
        split_tf32(x1, ah[i][1], al[i][1]);
        split_tf32(x2, ah[i][2], al[i][2]);
        split_tf32(x3, ah[i][3], al[i][3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float y0, y1;
        if (LAY == 0) {  // B stored [n][k]
          y0 = Bs[(8 * j + g) * TG_SR + kk + t];
          y1 = Bs[(8 * j + g) * TG_SR + kk + t + 4];
        } else {         // B stored [k][n]
          y0 = Bs[(kk + t) * TG_SC + 8 * j + g];
          y1 = Bs[(kk + t + 4) * TG_SC + 8 * j + g];
        }
        split_tf32(y0, bh[j][0], bl[j][0]);
        split_tf32(y1, bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mma_tf32(acc[i][j], al[i], bh[j]);
          mma_tf32(acc[i][j], ah[i], bl[j]);
          mma_tf32(acc[i][j], ah[i], bh[j]);
        }
    }
  }
  // cross-warp reduction of the four K quarters (the operand panels are dead: reuse their memory)
  __syncthreads();
  float* red = tg_smem;  // [NW][32][33]
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 16 * i + g, c = 8 * j + 2 * t;
      red[(warp * 32 + r) * 33 + c] = acc[i][j][0];
      red[(warp * 32 + r) * 33 + c + 1] = acc[i][j][1];
      red[(warp * 32 + r + 8) * 33 + c] = acc[i][j][2];
      red[(warp * 32 + r + 8) * 33 + c + 1] = acc[i][j][3];
    }
  __syncthreads();
  const bool split = gridDim.z > 1;
  for (int e = tid; e < 1024; e += 32 * NW) {
    const int r = e >> 5, c = e & 31;
    const int m = m0 + r, n = n0 + c;
    if (m >= M || n >= N) continue;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) v += red[(w * 32 + r) * 33 + c];
    v *= alpha;
    if (split) {
      atomicAdd(C + (size_t)m * ldc + n, v);
      continue;
    }
    if (beta != 0.f) v += beta * C[(size_t)m * ldc + n];
    if (bias) v += bias[n];
    if (relu) v = fmaxf(v, 0.f);
    if (gate && !(gate[(size_t)m * ldc + n] > 0.f)) v = 0.f;
    C[(size_t)m * ldc + n] = v;
  }
}

template <int LAY, int NW>
static int launch_tgemm(dim3 grid, cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                        int ldc, float alpha, float beta, const float* bias, int relu, const float* gate, int k_chunk) {
  constexpr int A_FLOATS = (LAY == 2) ? TG_KC * TG_SC : 32 * TG_SR;
  constexpr int B_FLOATS = (LAY == 0) ? 32 * TG_SR : TG_KC * TG_SC;
  constexpr int panels = (A_FLOATS + B_FLOATS) * (int)sizeof(float), redbuf = NW * 32 * 33 * (int)sizeof(float);
  constexpr int smem = panels > redbuf ? panels : redbuf;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_tgemm<LAY, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      scnb_set_error("sgemm: cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return SCNB_ERR_CUDA;
    }
    attr_set = true;
  }
  scnb_launch_pdl(k_tgemm<LAY, NW>, grid, dim3(32 * NW), smem, st, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, k_chunk);
  return SCNB_OK;
}

// Returns 1 if launched, 0 if not applicable, <0 on error.
int scnb_tgemm_try(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                   float alpha, float beta, const float* bias, int relu, const float* gate, cudaStream_t st) {
  if (transA && transB) return 0;
  const int lay = transA ? 2 : (transB ? 0 : 1);
  if ((lda % 4) || (ldb % 4) || (((uintptr_t)A | (uintptr_t)B) % 16)) return 0;
  if (lay == 0 && (K % 4)) return 0;
  if (lay == 1 && ((K % 4) || (N % 4))) return 0;
  if (lay == 2 && ((M % 4) || (N % 4))) return 0;
  const long long tiles = (long long)scnb_cdiv(M, 32) * scnb_cdiv(N, 32);
  int splits = 1;
  const bool splittable = !bias && !relu && !gate && (beta == 1.f) && K >= 128;
  if (splittable)
    while (splits < 32 && tiles * splits < 2 * 148 && K / (splits * 2) >= 64) splits *= 2;
  int k_chunk = scnb_cdiv(scnb_cdiv(K, splits), 32) * 32;
  splits = scnb_cdiv(K, k_chunk);
  dim3 grid(scnb_cdiv(N, 32), scnb_cdiv(M, 32), splits);
  static int nw = 0;  // warps per CTA = ways the K chunk is split inside the CTA (SCNB_TG_WARPS = 4 | 8 | 16)
  if (nw == 0) {
    const char* e = getenv("SCNB_TG_WARPS");
    nw = e ? atoi(e) : 8;
    if (nw != 4 && nw != 8 && nw != 16) nw = 8;
  }
  int rc;
#define TG_GO(NW_)                                                                                                              \
  rc = (lay == 0)   ? launch_tgemm<0, NW_>(grid, st, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, k_chunk)   \
       : (lay == 1) ? launch_tgemm<1, NW_>(grid, st, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, k_chunk)   \
                    : launch_tgemm<2, NW_>(grid, st, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, bias, relu, gate, k_chunk)
  if (nw == 4) TG_GO(4);
  else if (nw == 8) TG_GO(8);
  else TG_GO(16);
#undef TG_GO
  return rc < 0 ? rc : 1;
}
