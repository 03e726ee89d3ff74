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
