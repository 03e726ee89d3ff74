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
