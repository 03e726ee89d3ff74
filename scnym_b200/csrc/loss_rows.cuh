// Row functions of the loss / head kernels (one warp per row), shared by loss_ops.cu and head_ops.cu.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// Soft-target cross entropy, mean over `n_norm` rows, with dLogits.  One warp per row.
//   label row i = gam[i]*Y[mapA[i]] + (1-gam[i])*Y[mapB[i]]   (maps optional -> Y[i])
//   loss_i = -w_i * sum_c y_c * log_softmax(z)_c ;  w_i = max_c(cw_c*y_c) if class weights
//   dz = grad_scale * (w_i/n_norm) * (softmax(z)*sum_c y_c - y)
// Rows >= *n_rows_dev are inactive: dz = 0.  loss_out[0] += loss (atomic); loss_out[1] += #argmax hits.
// ---------------------------------------------------------------------------------------------
// One warp, one row `i` of the soft-target cross entropy (see k_soft_ce).
__device__ __forceinline__ void ce_row(int i, const float* __restrict__ Zl, int n_max, int C,
                                                 const int32_t* __restrict__ n_rows_dev, const float* __restrict__ Y,
                                                 const int32_t* __restrict__ mapA, const int32_t* __restrict__ mapB,
                                                 const float* __restrict__ gam, int map_off,
                                                 const float* __restrict__ class_weight, const int32_t* __restrict__ n_norm_dev,
                                                 int n_norm, float grad_scale, const float* __restrict__ loss_scale_dev,
                                                 float* __restrict__ dZ, float* __restrict__ loss_out, int count_acc,
                                                 float* __restrict__ row_loss) {
  const int lane = threadIdx.x & 31;
  const int n_rows = n_rows_dev ? min(*n_rows_dev, n_max) : n_max;
  float* dz = dZ ? dZ + (size_t)i * C : nullptr;
  if (i >= n_rows) {
    if (dz)
      for (int c = lane; c < C; c += 32) dz[c] = 0.f;
    if (row_loss && lane == 0) row_loss[i] = 0.f;
    return;
  }
  const float nn = (float)(n_norm_dev ? max(*n_norm_dev, 1) : max(n_norm, 1));
  const float ls = loss_scale_dev ? *loss_scale_dev : 1.f;
  const float* z = Zl + (size_t)i * C;
  int a = i, b = i;
  float g = 1.f;
  if (mapA) {
    a = mapA[map_off + i];
    b = mapB ? mapB[map_off + i] : a;
    g = gam ? gam[map_off + i] : 1.f;
  }
  const float* ya = Y + (size_t)a * C;
  const float* yb = Y + (size_t)b * C;
  const float g1 = 1.f - g;
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = z[c];
    if (v > m) { m = v; am = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oi < am)) { m = om; am = oi; }
  }
  if (am >= C) am = 0;         // all-NaN / all -inf row (see pseudolabel_row)
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(z[c] - m);
  s = warp_sum(s);
  const float lse = logf(s);
  float dot = 0.f, ysum = 0.f, w = -INFINITY;
  float ym = -INFINITY;  // argmax of the dominant label row (the original, un-mixed label)
  int yam = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float y = (a == b) ? ya[c] : g * ya[c] + g1 * yb[c];
    dot += y * ((z[c] - m) - lse);
    ysum += y;
    if (class_weight) w = fmaxf(w, class_weight[c] * y);
    if (ya[c] > ym) { ym = ya[c]; yam = c; }
  }
  if (count_acc) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, ym, o);
      const int oi = __shfl_xor_sync(0xffffffffu, yam, o);
      if (om > ym || (om == ym && oi < yam)) { ym = om; yam = oi; }
    }
    if (yam >= C) yam = 0;
  }
  dot = warp_sum(dot);
  ysum = warp_sum(ysum);
  w = class_weight ? warp_max(w) : 1.f;
  const float li = -dot * w;
  if (dz) {
    const float k = grad_scale * ls * w / nn;
    for (int c = lane; c < C; c += 32) {
      const float y = (a == b) ? ya[c] : g * ya[c] + g1 * yb[c];
      const float p = expf(z[c] - m) / s;
      dz[c] = k * (p * ysum - y);
    }
  }
  if (lane == 0) {
    if (row_loss) row_loss[i] = li * ls;
    atomicAdd(loss_out, li * ls / nn);
    if (count_acc && yam == am) atomicAdd(loss_out + 1, 1.0f);
  }
}


// ---------------------------------------------------------------------------------------------
// Interpolation-consistency MSE on probabilities (losses.py:880-913):
//   L_u = (1/n_norm) sum_{i < n_conf} sum_c (softmax(z_i)_c - y'_ic)^2 ;  rows >= n_conf get zero loss/grad.
//   dz = p * (g - sum_c g_c p_c),  g = 2 (p - y') * weight / n_norm
// ---------------------------------------------------------------------------------------------
// One warp, one row `i` of the masked MSE on probabilities (see k_softmax_mse).
__device__ __forceinline__ void mse_row(int i, const float* __restrict__ Zl, int n_max, int C,
                                                     const int32_t* __restrict__ n_conf_dev, const float* __restrict__ Y,
                                                     const int32_t* __restrict__ mapA, const int32_t* __restrict__ mapB,
                                                     const float* __restrict__ gam, int map_off, int n_norm,
                                                     const float* __restrict__ weight_dev, float weight,
                                                     float* __restrict__ dZ, float* __restrict__ loss_out) {
  const int lane = threadIdx.x & 31;
  const int n_conf = n_conf_dev ? min(*n_conf_dev, n_max) : n_max;
  float* dz = dZ ? dZ + (size_t)i * C : nullptr;
  if (i >= n_conf) {
    if (dz)
      for (int c = lane; c < C; c += 32) dz[c] = 0.f;
    return;
  }
  const float wgt = weight_dev ? *weight_dev : weight;
  const float* z = Zl + (size_t)i * C;
  int a = i, b = i;
  float g = 1.f;
  if (mapA) {
    a = mapA[map_off + i];
    b = mapB ? mapB[map_off + i] : a;
    g = gam ? gam[map_off + i] : 1.f;
  }
  const float* ya = Y + (size_t)a * C;
  const float* yb = Y + (size_t)b * C;
  const float g1 = 1.f - g;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, z[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(z[c] - m);
  s = warp_sum(s);
  float sq = 0.f, gp = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float y = (a == b) ? ya[c] : g * ya[c] + g1 * yb[c];
    const float p = expf(z[c] - m) / s;
    const float d = p - y;
    sq = fmaf(d, d, sq);
    gp = fmaf(d, p, gp);
  }
  sq = warp_sum(sq);
  gp = warp_sum(gp);
  const float inv_n = 1.0f / (float)max(n_norm, 1);
  if (dz) {
    const float k = 2.0f * wgt * inv_n;
    for (int c = lane; c < C; c += 32) {
      const float y = (a == b) ? ya[c] : g * ya[c] + g1 * yb[c];
      const float p = expf(z[c] - m) / s;
      dz[c] = k * p * ((p - y) - gp);
    }
  }
  if (lane == 0) atomicAdd(loss_out, sq * inv_n);
}


// ---------------------------------------------------------------------------------------------
// Fused heads.  The classifier (H -> C), the last layer of the domain adversary (H -> D) and their
// losses are a few kFLOP per row: as separate GEMM / loss / GEMM launches they cost four dependent
// kernels each on the step's critical path.  Here one warp owns one row: it forms the C (or D)
// logits against the weight matrix staged in shared memory, runs the loss row function on them and
// immediately back-propagates dLogits through the same weights.  H must be a multiple of 128
// (lane owns float4 groups h = (i*32 + lane)*4) and C*H floats must fit in shared memory;
// otherwise the engine keeps the unfused sequence.
// ---------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void head_logits(const float* __restrict__ a_row, const float* __restrict__ Ws,
                                            const float* __restrict__ bias, int C, int H, float* __restrict__ z_row,
                                            float4 (&a)[NV], int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = *reinterpret_cast<const float4*>(a_row + (i * 32 + lane) * 4);
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
    if (lane == (c & 31)) z_row[c] = p + (bias ? bias[c] : 0.f);
  }
  __syncwarp();
}
// out_row[h] = scale * sum_c dz[c] * W[c][h]   (optionally gated by act[h] > 0: ReLU backward)
template <int NV>
__device__ __forceinline__ void head_backprop(const float* __restrict__ dz_row, const float* __restrict__ Ws, int C, int H,
                                              float scale, const float4* gate_act, float* __restrict__ out_row, int lane) {
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < C; ++c) {
    const float d = dz_row[c];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 w = *reinterpret_cast<const float4*>(Ws + (size_t)c * H + (i * 32 + lane) * 4);
      acc[i].x = fmaf(d, w.x, acc[i].x);
      acc[i].y = fmaf(d, w.y, acc[i].y);
      acc[i].z = fmaf(d, w.z, acc[i].z);
      acc[i].w = fmaf(d, w.w, acc[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 v = make_float4(scale * acc[i].x, scale * acc[i].y, scale * acc[i].z, scale * acc[i].w);
    if (gate_act) {
      const float4 gt = gate_act[i];
      if (!(gt.x > 0.f)) v.x = 0.f;
      if (!(gt.y > 0.f)) v.y = 0.f;
      if (!(gt.z > 0.f)) v.z = 0.f;
      if (!(gt.w > 0.f)) v.w = 0.f;
    }
    *reinterpret_cast<float4*>(out_row + (i * 32 + lane) * 4) = v;
  }
}
__device__ __forceinline__ void stage_weights(float* Ws, const float* __restrict__ W, int n) {
  for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) *reinterpret_cast<float4*>(Ws + i) = *reinterpret_cast<const float4*>(W + i);
  __syncthreads();
}

// Short-chain variant for C <= 32 outputs (the usual number of cell types / domains): the per-class partial dot products of
// all lanes meet in a per-warp shared-memory tile [C][33] and lane c adds its class's 32 partials -- one transposed pass
// instead of C dependent butterfly reductions -- and the logits stay in shared memory (`zs[C]`) for the loss row function
// and the back-propagation, instead of a store -> load round trip through global memory.  Per-warp scratch: HEAD_WS floats.
#define HEAD_WS (32 * 33 + 64)
// (the row's activation already in registers: lane holds columns (i*32 + lane)*4 .. +3 of a[i])
template <int NV>
__device__ __forceinline__ void head_logits_s_regs(const float4 (&a)[NV], const float* __restrict__ Ws, const float* __restrict__ bias,
                                                   int C, int H, float* __restrict__ wsm, int lane) {
  float* sp = wsm;               // [C][33]
  float* zs = wsm + 32 * 33;     // [32] logits
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
    sp[c * 33 + lane] = p;
  }
  __syncwarp();
  if (lane < C) {
    float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      z0 += sp[lane * 33 + i];
      z1 += sp[lane * 33 + i + 1];
      z2 += sp[lane * 33 + i + 2];
      z3 += sp[lane * 33 + i + 3];
    }
    zs[lane] = ((z0 + z1) + (z2 + z3)) + (bias ? bias[lane] : 0.f);
  }
  __syncwarp();
}
template <int NV>
__device__ __forceinline__ void head_logits_s(const float* __restrict__ a_row, const float* __restrict__ Ws,
                                              const float* __restrict__ bias, int C, int H, float* __restrict__ wsm,
                                              float4 (&a)[NV], int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = *reinterpret_cast<const float4*>(a_row + (i * 32 + lane) * 4);
  head_logits_s_regs<NV>(a, Ws, bias, C, H, wsm, lane);
}

// ---------------------------------------------------------------------------------------------
// Thread-per-row forms of the two loss rows for kernels that hold a tile of logits AND of label rows in shared memory
// (head_ops.cu): with C <= 32 classes a warp-per-row call spends its time in five butterfly reductions over mostly idle
// lanes; here a lane owns a row and loops over the classes.  Same expressions as ce_row / mse_row; the sums over classes run
// in class order instead of the butterfly order.  y: the (mixed) soft label row, ya: the dominant un-mixed label row (its
// argmax is the class the accuracy counts against).  `dz` doubles as scratch for the exponentials.  The loss terms are
// returned (the caller reduces them over its rows), nothing is added atomically.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ce_row_thread(bool active, const float* __restrict__ z, int C, const float* __restrict__ y,
                                              const float* __restrict__ ya, const float* __restrict__ class_weight, float nn,
                                              float grad_scale, float ls, float* __restrict__ dz, float& loss_term, float& hit) {
  loss_term = 0.f;
  hit = 0.f;
  if (!active) {
    for (int c = 0; c < C; ++c) dz[c] = 0.f;
    return;
  }
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int c = 0; c < C; ++c) {
    const float v = z[c];
    if (v > m) { m = v; am = c; }
  }
  if (am >= C) am = 0;
  float s = 0.f;
  for (int c = 0; c < C; ++c) {
    const float e = expf(z[c] - m);
    dz[c] = e;
    s += e;
  }
  const float lse = logf(s);
  float dot = 0.f, ysum = 0.f, w = -INFINITY, ym = -INFINITY;
  int yam = 0x7fffffff;
  for (int c = 0; c < C; ++c) {
    const float yc = y[c];
    dot += yc * ((z[c] - m) - lse);
    ysum += yc;
    if (class_weight) w = fmaxf(w, class_weight[c] * yc);
    if (ya[c] > ym) { ym = ya[c]; yam = c; }
  }
  if (yam >= C) yam = 0;
  if (!class_weight) w = 1.f;
  const float li = -dot * w;
  const float k = grad_scale * ls * w / nn;
  for (int c = 0; c < C; ++c) {
    const float p = dz[c] / s;
    dz[c] = k * (p * ysum - y[c]);
  }
  loss_term = li * ls / nn;
  hit = (yam == am) ? 1.f : 0.f;
}

__device__ __forceinline__ void mse_row_thread(bool active, const float* __restrict__ z, int C, const float* __restrict__ y, int n_norm,
                                               float wgt, float* __restrict__ dz, float& loss_term) {
  loss_term = 0.f;
  if (!active) {
    for (int c = 0; c < C; ++c) dz[c] = 0.f;
    return;
  }
  float m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, z[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) {
    const float e = expf(z[c] - m);
    dz[c] = e;
    s += e;
  }
  float sq = 0.f, gp = 0.f;
  for (int c = 0; c < C; ++c) {
    const float p = dz[c] / s;
    const float d = p - y[c];
    sq = fmaf(d, d, sq);
    gp = fmaf(d, p, gp);
  }
  const float inv_n = 1.0f / (float)max(n_norm, 1);
  const float k = 2.0f * wgt * inv_n;
  for (int c = 0; c < C; ++c) {
    const float p = dz[c] / s;
    dz[c] = k * p * ((p - y[c]) - gp);
  }
  loss_term = sq * inv_n;
}

// The same two rows with FOUR lanes per row (lane = 4 * row + part; part p owns classes p, p + 4, ...): the chains of
// exponentials and divisions are a quarter as long and a warp still finishes 8 rows at once.  All 32 lanes must call (the
// group reductions are warp shuffles); `row_ok` = false marks lanes whose row does not exist.  The loss terms are returned by
// part 0 only (the others return 0), so that a plain warp sum counts every row once.
__device__ __forceinline__ float grp4_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float grp4_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ void grp4_argmax(float& m, int& am) {   // first index of the largest value
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oi < am)) { m = om; am = oi; }
  }
}
__device__ __forceinline__ void ce_row_group4(bool row_ok, bool active, int part, const float* __restrict__ z, int C,
                                              const float* __restrict__ y, const float* __restrict__ ya,
                                              const float* __restrict__ class_weight, float nn, float grad_scale, float ls,
                                              float* __restrict__ dz, float& loss_term, float& hit) {
  loss_term = 0.f;
  hit = 0.f;
  const bool live = row_ok && active;
  float m = -INFINITY;
  int am = 0x7fffffff;
  if (live)
    for (int c = part; c < C; c += 4) {
      const float v = z[c];
      if (v > m) { m = v; am = c; }
    }
  grp4_argmax(m, am);
  if (am >= C) am = 0;
  float s = 0.f;
  if (live)
    for (int c = part; c < C; c += 4) {
      const float e = expf(z[c] - m);
      dz[c] = e;
      s += e;
    }
  s = grp4_sum(s);
  const float lse = live ? logf(s) : 0.f;
  float dot = 0.f, ysum = 0.f, w = -INFINITY, ym = -INFINITY;
  int yam = 0x7fffffff;
  if (live)
    for (int c = part; c < C; c += 4) {
      const float yc = y[c];
      dot += yc * ((z[c] - m) - lse);
      ysum += yc;
      if (class_weight) w = fmaxf(w, class_weight[c] * yc);
      if (ya[c] > ym) { ym = ya[c]; yam = c; }
    }
  dot = grp4_sum(dot);
  ysum = grp4_sum(ysum);
  w = class_weight ? grp4_max(w) : 1.f;
  grp4_argmax(ym, yam);
  if (yam >= C) yam = 0;
  if (!row_ok) return;
  if (!active) {
    for (int c = part; c < C; c += 4) dz[c] = 0.f;
    return;
  }
  const float li = -dot * w;
  const float k = grad_scale * ls * w / nn;
  for (int c = part; c < C; c += 4) {
    const float p = dz[c] / s;
    dz[c] = k * (p * ysum - y[c]);
  }
  if (part == 0) {
    loss_term = li * ls / nn;
    hit = (yam == am) ? 1.f : 0.f;
  }
}

__device__ __forceinline__ void mse_row_group4(bool row_ok, bool active, int part, const float* __restrict__ z, int C,
                                               const float* __restrict__ y, int n_norm, float wgt, float* __restrict__ dz,
                                               float& loss_term) {
  loss_term = 0.f;
  const bool live = row_ok && active;
  float m = -INFINITY;
  if (live)
    for (int c = part; c < C; c += 4) m = fmaxf(m, z[c]);
  m = grp4_max(m);
  float s = 0.f;
  if (live)
    for (int c = part; c < C; c += 4) {
      const float e = expf(z[c] - m);
      dz[c] = e;
      s += e;
    }
  s = grp4_sum(s);
  float sq = 0.f, gp = 0.f;
  if (live)
    for (int c = part; c < C; c += 4) {
      const float p = dz[c] / s;
      const float d = p - y[c];
      sq = fmaf(d, d, sq);
      gp = fmaf(d, p, gp);
    }
  sq = grp4_sum(sq);
  gp = grp4_sum(gp);
  if (!row_ok) return;
  if (!active) {
    for (int c = part; c < C; c += 4) dz[c] = 0.f;
    return;
  }
  const float inv_n = 1.0f / (float)max(n_norm, 1);
  const float k = 2.0f * wgt * inv_n;
  for (int c = part; c < C; c += 4) {
    const float p = dz[c] / s;
    dz[c] = k * p * ((p - y[c]) - gp);
  }
  if (part == 0) loss_term = sq * inv_n;
}
