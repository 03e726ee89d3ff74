"""Data structures and batch operations (reference surface: scnym/dataprep.py).

`DeviceCSR` keeps a log1p(CPM) expression matrix resident in HBM as CSR (int64 indptr, int32
column indices, fp32 values); `CSRBatch` is the compact CSR of one minibatch produced by the
device row-gather kernel.  The reference instead densifies every cell on the host
(`SingleCellDS.__getitem__`, dataprep.py:114-168) and ships [B, G] dense batches to the GPU.
"""
from typing import Optional, Union

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


class DeviceCSR(object):
    """Expression matrix resident on one GPU."""

    def __init__(self, indptr: torch.Tensor, indices: torch.Tensor, data: torch.Tensor, n_cols: int,
                 max_row_nnz: Optional[int] = None) -> None:
        if indptr.dtype != torch.int64 or indices.dtype != torch.int32 or data.dtype != torch.float32:
            raise TypeError("DeviceCSR expects int64 indptr, int32 indices, float32 data")
        self.indptr, self.indices, self.data = indptr, indices, data
        self.n_rows = int(indptr.numel() - 1)
        self.n_cols = int(n_cols)
        if max_row_nnz is None:
            max_row_nnz = int((indptr[1:] - indptr[:-1]).max().item()) if self.n_rows > 0 else 0
        self.max_row_nnz = int(max_row_nnz)

    @property
    def shape(self):
        return (self.n_rows, self.n_cols)

    @property
    def device(self):
        return self.data.device

    @property
    def nnz(self):
        return int(self.data.numel())

    @classmethod
    def from_arrays(cls, indptr, indices, data, n_cols, device="cuda"):
        ip = torch.as_tensor(np.asarray(indptr, dtype=np.int64))
        ix = torch.as_tensor(np.asarray(indices, dtype=np.int32))
        dv = torch.as_tensor(np.asarray(data, dtype=np.float32))
        mx = int(np.diff(np.asarray(indptr)).max()) if len(indptr) > 1 else 0
        return cls(ip.to(device), ix.to(device), dv.to(device), n_cols, mx)

    @classmethod
    def from_scipy(cls, X, device="cuda"):
        from scipy import sparse
        if not sparse.isspmatrix_csr(X):
            X = sparse.csr_matrix(X)
        if not X.has_sorted_indices:
            X = X.sorted_indices()
        return cls.from_arrays(X.indptr, X.indices, X.data, X.shape[1], device)

    @classmethod
    def from_any(cls, X, device="cuda"):
        """np.ndarray or scipy CSR (the two types `SingleCellDS` accepts, dataprep.py:67-72)."""
        if isinstance(X, DeviceCSR):
            return X
        if isinstance(X, np.ndarray):
            from scipy import sparse
            return cls.from_scipy(sparse.csr_matrix(X.astype(np.float32)), device)
        return cls.from_scipy(X, device)

    def row_slice(self, start: int, stop: int) -> "DeviceCSR":
        """Contiguous row range as a view (shares indices/data; indptr is re-based on use)."""
        ip = self.indptr[start:stop + 1]
        return DeviceCSR(ip, self.indices, self.data, self.n_cols, self.max_row_nnz)


class CSRBatch(object):
    """Compact CSR minibatch on device (int32 indptr)."""

    def __init__(self, indptr: torch.Tensor, indices: torch.Tensor, data: torch.Tensor, n_rows: int, n_cols: int) -> None:
        self.indptr, self.indices, self.data = indptr, indices, data
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)

    @property
    def shape(self):
        return (self.n_rows, self.n_cols)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def device(self):
        return self.data.device

    def to_dense(self) -> torch.Tensor:
        """Materialise [n_rows, n_cols] (debug / foreign criteria only; not on the hot path)."""
        ip = self.indptr.to(torch.int64)
        nnz = int(ip[-1].item())
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=self.device), ip[1:] - ip[:-1])
        out = torch.zeros((self.n_rows, self.n_cols), device=self.device, dtype=torch.float32)
        out[rows, self.indices[:nnz].to(torch.int64)] = self.data[:nnz]
        return out


def gather_rows(ds: DeviceCSR, rows: Optional[torch.Tensor] = None, row0: int = 0, n: Optional[int] = None,
                augment: bool = False, keep_mask: Optional[torch.Tensor] = None, rng: Optional[torch.Tensor] = None,
                stream_id: int = 0, p_drop: float = 0.1, out: Optional[CSRBatch] = None) -> CSRBatch:
    """Device row gather (+ optional `log1p_drop`, dataprep.py:721-729) into a compact CSRBatch."""
    if rows is not None:
        if rows.dtype != torch.int64:
            rows = rows.to(torch.int64)
        n = int(rows.numel())
    elif n is None:
        n = ds.n_rows - row0
    dev = ds.device
    cap = max(n * max(ds.max_row_nnz, 1), 1)
    if out is None:
        out = CSRBatch(torch.empty(n + 1, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.int32, device=dev),
                       torch.empty(cap, dtype=torch.float32, device=dev), n, ds.n_cols)
    if augment and keep_mask is None and rng is None:
        # module-level path: draw the Bernoulli(1-p) mask with torch's generator
        keep_mask = (torch.rand(cap, device=dev) >= p_drop).to(torch.uint8)
    st = _stream()
    call("scnb_csr_row_offsets", ptr(ds.indptr), ptr(rows), row0, n, ptr(out.indptr), None, st)
    call("scnb_csr_gather_rows", ptr(ds.indptr), ptr(ds.indices), ptr(ds.data), ptr(rows), row0, n, ptr(out.indptr),
         ptr(out.indices), ptr(out.data), int(augment), 0, n, ptr(keep_mask), ptr(rng), stream_id, float(p_drop), st)
    return out


class HostBatchFeeder(object):
    """Host-buffer entry: assembles one minibatch from a HOST-resident CSR matrix into pinned
    staging buffers (C threads, `scnb_host_gather_rows`) and copies it to fixed device staging
    tensors that a `TrainEngine` reads as a tiny DeviceCSR (rows 0..B-1).

    This is the path a caller takes when the expression matrix does not live in HBM; it is
    what `bench.py` times as `e2e` (host->device copies inside the timed region).
    """

    def __init__(self, indptr, indices, data, n_cols: int, batch: int, device="cuda", labels=None, domains=None,
                 n_threads: int = 4) -> None:
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float32)
        self.labels = None if labels is None else np.ascontiguousarray(labels, dtype=np.int32)
        self.domains = None if domains is None else np.ascontiguousarray(domains, dtype=np.int32)
        self.n_cols, self.B, self.n_threads = int(n_cols), int(batch), int(n_threads)
        self.max_row_nnz = int(np.diff(self.indptr).max()) if len(self.indptr) > 1 else 1
        self.cap = self.B * max(self.max_row_nnz, 1)
        pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory()
        self.p_indptr, self.p_idx, self.p_val = pin(self.B + 1, torch.int64), pin(self.cap, torch.int32), pin(self.cap, torch.float32)
        self.p_y, self.p_dom = pin(self.B, torch.int32), pin(self.B, torch.int32)
        dev = torch.device(device)
        self.d_indptr = torch.zeros(self.B + 1, dtype=torch.int64, device=dev)
        self.d_idx = torch.zeros(self.cap, dtype=torch.int32, device=dev)
        self.d_val = torch.zeros(self.cap, dtype=torch.float32, device=dev)
        self.d_y = torch.zeros(self.B, dtype=torch.int32, device=dev)
        self.d_dom = torch.zeros(self.B, dtype=torch.int32, device=dev)
        self.csr = DeviceCSR(self.d_indptr, self.d_idx, self.d_val, self.n_cols, self.max_row_nnz)
        self.csr.n_rows = self.B
        self.bytes_last = 0

    def feed(self, rows) -> int:
        """Gather `rows` (int64 numpy, len B) and enqueue the H2D copies on the current stream.
        Returns the number of bytes copied host->device."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        if rows.shape[0] != self.B:
            raise ValueError(f"expected {self.B} rows, got {rows.shape[0]}")
        h = _lib.lib()
        rc = h.scnb_host_gather_rows(self.indptr.ctypes.data, self.indices.ctypes.data, self.data.ctypes.data,
                                     rows.ctypes.data, self.B, self.p_indptr.data_ptr(), self.p_idx.data_ptr(),
                                     self.p_val.data_ptr(), self.cap, self.n_threads)
        if rc != 0:
            raise _lib.ScnbError(h.scnb_last_error().decode())
        nnz = int(self.p_indptr[self.B])
        self.d_indptr.copy_(self.p_indptr, non_blocking=True)
        self.d_idx[:nnz].copy_(self.p_idx[:nnz], non_blocking=True)
        self.d_val[:nnz].copy_(self.p_val[:nnz], non_blocking=True)
        nbytes = (self.B + 1) * 8 + nnz * 8
        if self.labels is not None:
            self.p_y.copy_(torch.from_numpy(self.labels[rows]))
            self.d_y.copy_(self.p_y, non_blocking=True)
            nbytes += self.B * 4
        if self.domains is not None:
            self.p_dom.copy_(torch.from_numpy(self.domains[rows]))
            self.d_dom.copy_(self.p_dom, non_blocking=True)
            nbytes += self.B * 4
        self.bytes_last = nbytes
        return nbytes
