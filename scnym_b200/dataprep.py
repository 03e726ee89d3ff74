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


def remap_columns(ds: DeviceCSR, col_map: torch.Tensor, n_cols_out: int, row0: int = 0, n_rows: Optional[int] = None,
                  out: Optional[tuple] = None) -> DeviceCSR:
    """Rows [row0, row0+n_rows) of `ds` with column c moved to col_map[c] (dropped where col_map[c] < 0), every row
    sorted by its new column: the device form of `utils.build_classification_matrix` (scnym/utils.py:155-233).
    `col_map` (int32 [ds.n_cols], on the device) must be injective on its non-negative entries.
    `out = (indptr int64 [>= n_rows+1], indices int32 [cap], data fp32 [cap], workspace int64)` reuses caller buffers
    whose capacity covers the range's stored entries (no host synchronisation at all); otherwise exact-size outputs are
    allocated, which reads the new entry count back once."""
    n = ds.n_rows - row0 if n_rows is None else int(n_rows)
    if col_map.dtype != torch.int32 or col_map.numel() != ds.n_cols:
        raise TypeError("remap_columns: col_map must be int32 with one entry per input column")
    _lib.require_cuda(ds.data, col_map)
    st = _stream()
    dev = ds.device
    ip_in = ds.indptr[row0:]
    if out is None:
        ws = torch.empty(_lib.lib().scnb_csr_remap_workspace_len(n), dtype=torch.int64, device=dev)
        ip = torch.empty(n + 1, dtype=torch.int64, device=dev)
    else:
        ip, idx, val, ws = out
    call("scnb_csr_remap_count", ptr(ip_in), ptr(ds.indices), ptr(col_map), ds.n_cols, n, ptr(ip), ptr(ws), st)
    if out is None:
        total = int(ip[n].item()) if n > 0 else 0
        idx = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        val = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
    call("scnb_csr_remap_fill", ptr(ip_in), ptr(ds.indices), ptr(ds.data), ptr(col_map), ds.n_cols, int(n_cols_out), n, ptr(ip),
         ptr(idx), ptr(val), st)
    if out is None:
        idx, val = idx[:total], val[:total]
    return DeviceCSR(ip[:n + 1], idx, val, int(n_cols_out), max_row_nnz=min(ds.max_row_nnz, int(n_cols_out)))


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
         ptr(out.indices), ptr(out.data), int(augment), 0, n, ptr(keep_mask), ptr(rng), stream_id, float(p_drop), None, st)
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


# =================================================================================================
# Reference-compatible surface (scnym/dataprep.py): SingleCellDS, SampleMixUp, augmentations
# =================================================================================================
class SingleCellDS(torch.utils.data.Dataset):
    """Dataset of single cell profiles (scnym/dataprep.py:12-168).

    Same constructor, attributes (`X`, `y_labels`, `y`, `dom_labels`, `dom`, `transform`) and
    `__getitem__` contract as the reference, so foreign code can still index it on the host;
    the scnym_b200 trainers never do -- they call `device_csr()` once and sample rows on the GPU.
    """

    def __init__(self, X, y, domain=None, transform=None, num_classes: int = -1, num_domains: int = -1) -> None:
        from scipy import sparse
        super().__init__()
        if type(X) not in (np.ndarray, sparse.csr_matrix):
            raise TypeError(f"X is type {type(X)}, must `np.ndarray` or `sparse.csr_matrix`")
        if type(y) not in (np.ndarray, sparse.csr_matrix):
            raise TypeError(f"X is type {type(y)}, must `np.ndarray` or `sparse.csr_matrix`")
        if type(y) != np.ndarray:
            y = y.toarray()
        self.X = X
        self.y_labels = torch.from_numpy(y).long()
        self.y = torch.nn.functional.one_hot(self.y_labels, num_classes=num_classes).float()
        self.num_classes = int(self.y.shape[1]) if self.y.dim() > 1 else num_classes
        self.dom_labels = domain
        self.num_domains = num_domains
        if self.dom_labels is not None:
            self.dom = torch.nn.functional.one_hot(torch.from_numpy(np.asarray(self.dom_labels)).long(),
                                                   num_classes=num_domains).float()
        else:
            self.dom = np.zeros_like(self.y) - 1
        self.transform = transform
        if not self.X.shape[0] == self.y.shape[0]:
            raise ValueError("X rows %d not equal to y rows %d." % (self.X.shape[0], self.y.shape[0]))
        self._dev = None

    def __len__(self) -> int:
        return self.X.shape[0]

    def __getitem__(self, idx: int) -> dict:
        if type(idx) != int:
            raise TypeError(f"indices must be int, you passed {type(idx)}, {idx}")
        if idx < 0 or idx > len(self):
            raise ValueError("idx %d is invalid for dataset with %d examples." % (idx, len(self)))
        input_ = self.X[idx, ...]
        if type(input_) != np.ndarray:
            input_ = input_.toarray()
        input_ = torch.from_numpy(np.asarray(input_)).float()
        if input_.dim() > 1 and input_.size(0) == 1:
            input_ = input_.squeeze()
        sample = {"input": input_, "output": self.y[idx], "domain": self.dom[idx]}
        if self.transform is not None:
            sample = self.transform(sample)
        return sample

    def device_csr(self, device="cuda") -> DeviceCSR:
        """The expression matrix resident in HBM as CSR (uploaded once, cached)."""
        if self._dev is None or self._dev.device != torch.device(device):
            self._dev = DeviceCSR.from_any(self.X, device)
        return self._dev


class DeviceLoader(object):
    """Stand-in for `torch.utils.data.DataLoader(SingleCellDS, ...)` whose batches are drawn on
    the GPU.  Exposes `.dataset`, `.batch_size`, `.sampler`, `__len__` (read by
    `Trainer.__init__`, scnym/trainer.py:157-171) and iterates sample dicts
    {"input": CSRBatch, "output": one-hot [B, C], "domain": one-hot [B, D] or -1} for criteria
    that consume batches one by one; the fused trainers bypass iteration entirely."""

    def __init__(self, dataset: SingleCellDS, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                 sampler=None, device="cuda") -> None:
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.sampler = sampler if sampler is not None else ("RandomSampler" if shuffle else "SequentialSampler")
        self.weights = None
        if sampler is not None and hasattr(sampler, "weights"):
            self.weights = torch.as_tensor(sampler.weights, dtype=torch.float64)
            self.num_samples = int(sampler.num_samples)
        self.device = torch.device(device)

    def __len__(self) -> int:
        n = len(self.dataset) if self.weights is None else self.num_samples
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def epoch_order(self) -> torch.Tensor:
        """Row order of one epoch, on the device (shuffle / weighted sampling / sequential)."""
        n = len(self.dataset)
        if self.weights is not None:  # WeightedRandomSampler (main.py:229-232): with replacement
            return torch.multinomial(self.weights.to(self.device).float(), self.num_samples, replacement=True)
        if self.shuffle:
            return torch.randperm(n, device=self.device)
        return torch.arange(n, device=self.device)

    def __iter__(self):
        ds = self.dataset.device_csr(self.device)
        order = self.epoch_order()
        y = self.dataset.y.to(self.device)
        dom = torch.as_tensor(self.dataset.dom).to(self.device)
        for b in range(len(self)):
            rows = order[b * self.batch_size:(b + 1) * self.batch_size]
            yield {"input": gather_rows(ds, rows), "output": y[rows], "domain": dom[rows], "rows": rows}


def balance_classes(y: np.ndarray, class_min: int = 256) -> np.ndarray:
    """Class balancing by under/over-sampling (scnym/dataprep.py:171-210)."""
    classes, counts = np.unique(y, return_counts=True)
    min_count = max(int(np.min(counts)), class_min)
    all_idx = []
    for i, c in enumerate(classes):
        class_idx = np.where(y == c)[0].astype("int")
        rep = counts[i] < min_count
        if rep:
            print("Count for class %s is %d. Oversampling." % (c, counts[i]))
        all_idx += [np.random.choice(class_idx, size=min_count, replace=rep)]
    return np.concatenate(all_idx).astype("int")


def _dense_as_csr(x: torch.Tensor) -> DeviceCSR:
    """View a dense CUDA [B, G] tensor as a CSR matrix that stores every entry."""
    B, G = x.shape
    key = (x.device, B, G)
    cache = _dense_as_csr.cache
    if key not in cache:
        cache[key] = (torch.arange(B + 1, device=x.device, dtype=torch.int64) * G,
                      torch.arange(G, device=x.device, dtype=torch.int32).repeat(B))
    ip, ix = cache[key]
    return DeviceCSR(ip, ix, x.contiguous().view(-1).float(), G, G)


_dense_as_csr.cache = {}


class Log1pDrop(object):
    """`log1p_drop` = ExpMinusOne -> InputDropout(0.1) -> LibrarySizeNormalize(log1p)
    (scnym/dataprep.py:721-729) as ONE device kernel over the stored entries.  Accepts a sample
    dict whose "input" is a CSRBatch or a dense CUDA tensor and replaces it, as the reference
    Compose does.  `KEEP_INJECT` (test hook) supplies the Bernoulli keep mask."""
    KEEP_INJECT = []

    def __init__(self, p_drop: float = 0.1) -> None:
        self.p_drop = p_drop
        # mimic torchvision.transforms.Compose.transforms for code that introspects the scheme
        self.transforms = [ExpMinusOne(), InputDropout(p_drop=p_drop), LibrarySizeNormalize(log1p=True)]

    def __call__(self, sample: dict) -> dict:
        x = sample["input"]
        dense = isinstance(x, torch.Tensor)
        if dense:
            _lib.require_cuda(x)
            ds = _dense_as_csr(x.detach())
        else:
            ds = DeviceCSR(x.indptr.to(torch.int64), x.indices, x.data, x.n_cols, max(int(x.indices.numel() // max(x.n_rows, 1)), 1))
        keep = None
        if Log1pDrop.KEEP_INJECT:
            keep = Log1pDrop.KEEP_INJECT.pop(0)
            if dense:
                keep = keep.to(device=x.device, dtype=torch.uint8).contiguous().view(-1)
        n = ds.n_rows
        nnz_cap = int(ds.data.numel())
        out = CSRBatch(torch.empty(n + 1, dtype=torch.int32, device=ds.device), torch.empty(max(nnz_cap, 1), dtype=torch.int32, device=ds.device),
                       torch.empty(max(nnz_cap, 1), dtype=torch.float32, device=ds.device), n, ds.n_cols)
        if keep is None:
            keep = (torch.rand(max(nnz_cap, 1), device=ds.device) >= self.p_drop).to(torch.uint8)
        st = _stream()
        call("scnb_csr_row_offsets", ptr(ds.indptr), None, 0, n, ptr(out.indptr), None, st)
        call("scnb_csr_gather_rows", ptr(ds.indptr), ptr(ds.indices), ptr(ds.data), None, 0, n, ptr(out.indptr), ptr(out.indices),
             ptr(out.data), 1, 0, n, ptr(keep), None, 0, float(self.p_drop), None, st)
        sample["input"] = out.data[: n * ds.n_cols].view(n, ds.n_cols) if dense else out
        return sample


class LibrarySizeNormalize(object):
    """Host-tensor restatement kept for API compatibility (scnym/dataprep.py:213-254); the
    device path uses `Log1pDrop`."""

    def __init__(self, counts_per_cell_after: int = int(1e6), log1p: bool = True) -> None:
        self.counts_per_cell_after, self.log1p = counts_per_cell_after, log1p

    def __call__(self, sample: dict) -> dict:
        x = sample["input"]
        x = x / torch.sum(x, dim=1).reshape(-1, 1) * self.counts_per_cell_after
        sample["input"] = torch.log1p(x) if self.log1p else x
        return sample


class ExpMinusOne(object):
    def __call__(self, sample: dict) -> dict:
        sample["input"] = torch.expm1(sample["input"])
        return sample


class InputDropout(object):
    def __init__(self, p_drop: float = 0.1) -> None:
        self.p_drop = p_drop

    def __call__(self, sample: dict) -> dict:
        sample["input"] = torch.nn.functional.dropout(sample["input"], p=self.p_drop, inplace=False)
        return sample


def mixup(a, b, gamma):
    """gamma * a + (1 - gamma) * b (scnym/dataprep.py:571-594)."""
    return gamma * a + (1 - gamma) * b


class SampleMixUp(object):
    """MixUp on a sample batch (scnym/dataprep.py:597-705): `ridx = randperm(n)`,
    `gamma ~ Beta(alpha, alpha)` (max(gamma, 1-gamma) if keep_dominant_obs), every tensor in
    the dict mixed, `random_idx` recorded.  Used by the module-level path on dense tensors; the
    fused engine performs the same mix behind the first layer (`scnb_mix_rows`)."""
    INJECT = []  # test hook: (perm_keys, gamma_raw) tuples

    def __init__(self, alpha: float = 0.2, keep_dominant_obs: bool = False) -> None:
        self.alpha = alpha
        if alpha > 0.0:
            self.beta = torch.distributions.beta.Beta(self.alpha, self.alpha)
        self.keep_dominant_obs = keep_dominant_obs

    def __call__(self, sample: dict) -> dict:
        if self.alpha == 0.0:
            return sample
        input_, output = sample["input"], sample["output"]
        n = input_.size(0)
        if SampleMixUp.INJECT:
            keys, graw = SampleMixUp.INJECT.pop(0)
            ridx = torch.argsort(keys[:n], stable=True)
            gamma = graw[:n].clone()
        else:
            ridx = torch.randperm(n)
            gamma = self.beta.sample((n,))
        if self.keep_dominant_obs:
            gamma = torch.maximum(gamma, 1 - gamma)
        gamma = gamma.reshape(-1, 1).to(device=input_.device)
        ridx_d = ridx.to(input_.device)
        sample["input"] = mixup(input_, input_[ridx_d], gamma)
        sample["output"] = mixup(output, output[ridx_d], gamma)
        for k in [k for k in sample.keys() if k not in ("input", "output")]:
            if type(sample[k]) == torch.Tensor:
                sample[k] = mixup(sample[k], sample[k][ridx_d], gamma=gamma)
        sample["random_idx"] = ridx
        return sample


AUG_CODES = {"log1p_drop": 1, "log1p_mask": 2, "log1p_poisson": 3, "log1p_poisson_drop": 4, "count_poisson": 5}


class SparseAugment(object):
    """The other `AUGMENTATION_SCHEMES` of the reference (scnym/dataprep.py:730-754) as ONE device kernel over the stored
    entries (`scnb_csr_gather_rows_aug`): every stage maps 0 to 0, so the support of a row is preserved.
      log1p_mask          ExpMinusOne -> GeneMasking(p_drop=0.1, p_apply=0.5) -> LibrarySizeNormalize(log1p)   (:405-465)
      log1p_poisson       ExpMinusOne -> PoissonSample() -> LibrarySizeNormalize(log1p)                        (:277-341)
      log1p_poisson_drop  ExpMinusOne -> PoissonSample(depth=(0.1, 2.0)) -> InputDropout(0.1) -> LibrarySizeNormalize(log1p)
      count_poisson       PoissonSample() on raw counts
    Accepts a sample dict whose "input" is a CSRBatch or a dense CUDA tensor and replaces it, like the reference Compose.
    Test hooks (class-level queues, popped per call): KEEP_INJECT (0/1 keeps, dense [B, G] or per nnz), POIS_INJECT
    (sampled values, same layout), ROWU_INJECT (per-row uniforms of the depth factor)."""
    KEEP_INJECT, POIS_INJECT, ROWU_INJECT = [], [], []
    _seed = [0x5c9b200, 0]

    def __init__(self, scheme: str, p_drop: float = 0.1, p_apply: float = 0.5) -> None:
        if scheme not in AUG_CODES:
            raise ValueError(f"unknown augmentation scheme {scheme!r}")
        self.scheme, self.code, self.p_drop, self.p_apply = scheme, AUG_CODES[scheme], float(p_drop), float(p_apply)

    def __call__(self, sample: dict) -> dict:
        x = sample["input"]
        dense = isinstance(x, torch.Tensor)
        if dense:
            _lib.require_cuda(x)
            ds = _dense_as_csr(x.detach())
        else:
            ds = DeviceCSR(x.indptr.to(torch.int64), x.indices, x.data, x.n_cols, max(int(x.indices.numel() // max(x.n_rows, 1)), 1))
        dev = ds.device
        pop = lambda q, dt: (q.pop(0).to(device=dev, dtype=dt).contiguous().view(-1) if q else None)
        keep = pop(SparseAugment.KEEP_INJECT, torch.uint8) if self.code in (1, 2, 4) else None
        pois = pop(SparseAugment.POIS_INJECT, torch.float32) if self.code in (3, 4, 5) else None
        rowu = pop(SparseAugment.ROWU_INJECT, torch.float32) if self.code == 4 else None
        n, cap = ds.n_rows, max(int(ds.data.numel()), 1)
        out = CSRBatch(torch.empty(n + 1, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.int32, device=dev),
                       torch.empty(cap, dtype=torch.float32, device=dev), n, ds.n_cols)
        SparseAugment._seed[1] += 1
        rng = torch.tensor(SparseAugment._seed, dtype=torch.int64, device=dev)
        st = _stream()
        call("scnb_csr_row_offsets", ptr(ds.indptr), None, 0, n, ptr(out.indptr), None, st)
        call("scnb_csr_gather_rows_aug", ptr(ds.indptr), ptr(ds.indices), ptr(ds.data), None, 0, n, ptr(out.indptr), ptr(out.indices),
             ptr(out.data), self.code, ds.n_cols, self.p_drop, self.p_apply, ptr(keep), ptr(pois), ptr(rowu), ptr(rng), 0, None, 0, st)
        sample["input"] = out.data[: n * ds.n_cols].view(n, ds.n_cols) if dense else out
        return sample


def identity(x):
    return x


AUGMENTATION_SCHEMES = {
    "log1p_drop": Log1pDrop(p_drop=0.1),
    "log1p_mask": SparseAugment("log1p_mask", p_drop=0.1, p_apply=0.5),
    "log1p_poisson": SparseAugment("log1p_poisson"),
    "log1p_poisson_drop": SparseAugment("log1p_poisson_drop", p_drop=0.1),
    "count_poisson": SparseAugment("count_poisson"),
    "None": identity,
    "none": identity,
    None: identity,
}
