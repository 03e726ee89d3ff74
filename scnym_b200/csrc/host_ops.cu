// Host-side batch assembly for the host-buffer (end-to-end) entry: gathers CSR rows of a
// host-resident matrix into pinned staging buffers with a few threads.  Stands in for the
// reference's per-cell `SingleCellDS.__getitem__` + collate (scnym/dataprep.py:114-168) when the
// expression matrix lives in host memory; the dense [B,G] batch is still never formed.
#include <string.h>

#include <thread>
#include <vector>

#include "common.cuh"

extern "C" int scnb_host_gather_rows(const int64_t* indptr, const int32_t* indices, const float* data, const int64_t* rows,
                                     int n, int64_t* out_indptr, int32_t* out_indices, float* out_data, long long out_capacity,
                                     int n_threads) {
  SCNB_CHECK_ARG(indptr && indices && data && rows && out_indptr && out_indices && out_data && n >= 0,
                 "host_gather_rows: bad arguments");
  out_indptr[0] = 0;
  for (int i = 0; i < n; ++i) out_indptr[i + 1] = out_indptr[i] + (indptr[rows[i] + 1] - indptr[rows[i]]);
  SCNB_CHECK_ARG(out_indptr[n] <= out_capacity, "host_gather_rows: staging capacity %lld < batch nnz %lld", out_capacity,
                 (long long)out_indptr[n]);
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n) n_threads = n > 0 ? n : 1;
  auto work = [&](int t) {
    for (int i = t; i < n; i += n_threads) {
      const int64_t s = indptr[rows[i]], len = indptr[rows[i] + 1] - s, d = out_indptr[i];
      memcpy(out_indices + d, indices + s, (size_t)len * sizeof(int32_t));
      memcpy(out_data + d, data + s, (size_t)len * sizeof(float));
    }
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Peer-addressable exchange buffers for the data-parallel BatchNorm kernels (one process per GPU): plain cudaMalloc
// memory exported / opened through CUDA IPC, so that a kernel on one GPU can store into every other GPU's buffer over
// NVLink / NVSwitch.  handle: 64 opaque bytes (cudaIpcMemHandle_t) that the host side ships to the other ranks.
// ---------------------------------------------------------------------------------------------
extern "C" int scnb_peer_alloc(long long bytes, void** ptr_out, void* handle64_out) {
  SCNB_CHECK_ARG(bytes > 0 && ptr_out && handle64_out, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    scnb_set_error("peer_alloc: %s", cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  memcpy(handle64_out, &h, sizeof(h));
  *ptr_out = p;
  return SCNB_OK;
}

extern "C" int scnb_peer_open(const void* handle64, void** ptr_out) {
  SCNB_CHECK_ARG(handle64 && ptr_out, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    scnb_set_error("peer_open: %s", cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  *ptr_out = p;
  return SCNB_OK;
}

extern "C" int scnb_peer_close(void* ptr) {
  if (!ptr) return SCNB_OK;
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) {
    scnb_set_error("peer_close: %s", cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  return SCNB_OK;
}

extern "C" int scnb_peer_free(void* ptr) {
  if (!ptr) return SCNB_OK;
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    scnb_set_error("peer_free: %s", cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  return SCNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Host memory a kernel can store into (cudaHostAllocMapped): the per-step loss scalars of the host-buffer pipeline
// (hostfeed.py; the reference reads `loss.item()` every step, scnym/trainer.py:666-681) are published by a one-CTA kernel
// instead of a device-to-host memcpy, which the copy engine queues behind the 4.6 MB upload of a later batch.
// ---------------------------------------------------------------------------------------------
extern "C" int scnb_host_alloc_mapped(long long bytes, void** host_ptr_out, void** dev_ptr_out) {
  SCNB_CHECK_ARG(bytes > 0 && host_ptr_out && dev_ptr_out, "host_alloc_mapped: bad arguments");
  void *h = nullptr, *d = nullptr;
  cudaError_t e = cudaHostAlloc(&h, (size_t)bytes, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&d, h, 0);
  if (e != cudaSuccess) {
    if (h) cudaFreeHost(h);
    scnb_set_error("host_alloc_mapped: %s", cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  memset(h, 0, (size_t)bytes);
  *host_ptr_out = h;
  *dev_ptr_out = d;
  return SCNB_OK;
}

extern "C" int scnb_host_free_mapped(void* host_ptr) {
  if (!host_ptr) return SCNB_OK;
  cudaError_t e = cudaFreeHost(host_ptr);
  if (e != cudaSuccess) {
    scnb_set_error("host_free_mapped: %s", cudaGetErrorString(e));
    return SCNB_ERR_CUDA;
  }
  return SCNB_OK;
}

// Page-lock an existing host array in place and map it into the device's address space (cudaHostRegisterMapped): the row
// gather of the training step then reads the sampled rows straight out of host memory over PCIe -- no host threads, no
// staging blob, no copy-engine transfer (the reference assembles every minibatch on the host, scnym/dataprep.py:114-168).
extern "C" int scnb_host_register_mapped(void* host_ptr, long long bytes, void** dev_ptr_out) {
  SCNB_CHECK_ARG(host_ptr && bytes > 0 && dev_ptr_out, "host_register_mapped: bad arguments");
  void* d = nullptr;
  cudaError_t e = cudaHostRegister(host_ptr, (size_t)bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
  if (e == cudaSuccess) {
    e = cudaHostGetDevicePointer(&d, host_ptr, 0);
    if (e != cudaSuccess) cudaHostUnregister(host_ptr);
  }
  if (e != cudaSuccess) {
    scnb_set_error("host_register_mapped: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return SCNB_ERR_CUDA;
  }
  *dev_ptr_out = d;
  return SCNB_OK;
}

extern "C" int scnb_host_unregister(void* host_ptr) {
  if (!host_ptr) return SCNB_OK;
  cudaError_t e = cudaHostUnregister(host_ptr);
  if (e != cudaSuccess) {
    scnb_set_error("host_unregister: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return SCNB_ERR_CUDA;
  }
  return SCNB_OK;
}

// (no system-scope fence: the host reads after an event behind this launch, and a fence would wait for the SM's other
// traffic to host memory -- 20 us when an upload is in flight)
__global__ void k_publish_f32(const float* __restrict__ src, int n, float* __restrict__ dst) {
  for (int i = threadIdx.x; i < n; i += 32) dst[i] = src[i];
}

// dst_mapped[i] = src[i], i < n <= 256 (dst_mapped: device pointer of scnb_host_alloc_mapped memory); visible to the host
// once the stream has passed this launch
extern "C" int scnb_publish_f32(const float* src, int n, float* dst_mapped, void* stream) {
  SCNB_CHECK_ARG(src && dst_mapped && n > 0 && n <= 256, "publish_f32: bad arguments");
  k_publish_f32<<<1, 32, 0, (cudaStream_t)stream>>>(src, n, dst_mapped);
  SCNB_CHECK_LAUNCH("publish_f32");
  return SCNB_OK;
}

// Both matrices of a MixMatch batch (labeled and unlabeled rows) in ONE call: the worker threads are started once and
// share the n_a + n_b rows, and the source rows of the next few copies are prefetched (the gather is a latency-bound
// random read of host memory).  Same outputs as two scnb_host_gather_rows calls.
extern "C" int scnb_host_gather_rows2(const int64_t* indptr_a, const int32_t* indices_a, const float* data_a, const int64_t* rows_a,
                                      int n_a, int64_t* out_indptr_a, int32_t* out_indices_a, float* out_data_a,
                                      long long cap_a, const int64_t* indptr_b, const int32_t* indices_b, const float* data_b,
                                      const int64_t* rows_b, int n_b, int64_t* out_indptr_b, int32_t* out_indices_b,
                                      float* out_data_b, long long cap_b, int n_threads) {
  SCNB_CHECK_ARG(indptr_a && indices_a && data_a && rows_a && out_indptr_a && out_indices_a && out_data_a && n_a >= 0,
                 "host_gather_rows2: bad arguments (first matrix)");
  SCNB_CHECK_ARG(n_b == 0 || (indptr_b && indices_b && data_b && rows_b && out_indptr_b && out_indices_b && out_data_b && n_b > 0),
                 "host_gather_rows2: bad arguments (second matrix)");
  out_indptr_a[0] = 0;
  for (int i = 0; i < n_a; ++i) out_indptr_a[i + 1] = out_indptr_a[i] + (indptr_a[rows_a[i] + 1] - indptr_a[rows_a[i]]);
  SCNB_CHECK_ARG(out_indptr_a[n_a] <= cap_a, "host_gather_rows2: staging capacity %lld < batch nnz %lld", cap_a,
                 (long long)out_indptr_a[n_a]);
  if (n_b > 0) {
    out_indptr_b[0] = 0;
    for (int i = 0; i < n_b; ++i) out_indptr_b[i + 1] = out_indptr_b[i] + (indptr_b[rows_b[i] + 1] - indptr_b[rows_b[i]]);
    SCNB_CHECK_ARG(out_indptr_b[n_b] <= cap_b, "host_gather_rows2: staging capacity %lld < batch nnz %lld", cap_b,
                   (long long)out_indptr_b[n_b]);
  }
  const int n = n_a + n_b;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n) n_threads = n > 0 ? n : 1;
  auto src_of = [&](int i, const int32_t*& ix, const float*& vv, int64_t& len) {
    if (i < n_a) {
      const int64_t s = indptr_a[rows_a[i]];
      len = indptr_a[rows_a[i] + 1] - s;
      ix = indices_a + s;
      vv = data_a + s;
    } else {
      const int64_t s = indptr_b[rows_b[i - n_a]];
      len = indptr_b[rows_b[i - n_a] + 1] - s;
      ix = indices_b + s;
      vv = data_b + s;
    }
  };
  auto work = [&](int t) {
    for (int i = t; i < n; i += n_threads) {
      const int nx = i + n_threads;
      if (nx < n) {   // touch the head of the next row this thread will copy
        const int32_t* pix;
        const float* pvv;
        int64_t plen;
        src_of(nx, pix, pvv, plen);
        for (int64_t o = 0; o < plen; o += 16) {
          __builtin_prefetch(pix + o);
          __builtin_prefetch(pvv + o);
        }
      }
      const int32_t* ix;
      const float* vv;
      int64_t len;
      src_of(i, ix, vv, len);
      if (i < n_a) {
        memcpy(out_indices_a + out_indptr_a[i], ix, (size_t)len * sizeof(int32_t));
        memcpy(out_data_a + out_indptr_a[i], vv, (size_t)len * sizeof(float));
      } else {
        memcpy(out_indices_b + out_indptr_b[i - n_a], ix, (size_t)len * sizeof(int32_t));
        memcpy(out_data_b + out_indptr_b[i - n_a], vv, (size_t)len * sizeof(float));
      }
    }
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  }
  return SCNB_OK;
}
