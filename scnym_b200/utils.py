"""Host-side helpers on the boundary of the hot path (reference: scnym/utils.py:127-233) and a
minimal AnnData stand-in (anndata is not installed in this image)."""
from typing import Union

import numpy as np
import pandas as pd
from scipy import sparse


def get_adata_asarray(adata) -> Union[np.ndarray, sparse.csr_matrix]:
    """`.X` as an in-memory array of the same kind (scnym/utils.py:127-152)."""
    if sparse.issparse(adata.X):
        return sparse.csr_matrix(adata.X)
    return np.array(adata.X)


def gene_column_map(model_genes: np.ndarray, sample_genes: np.ndarray) -> np.ndarray:
    """int32 [len(sample_genes)]: the model column of every sample column, -1 where the model lacks the gene
    (the mapping loop of scnym/utils.py:209-217 as one hash join).  A gene repeated in the sample keeps its LAST
    column, as the reference's column-batch assignment does; a gene repeated in the model is ambiguous and raises
    ValueError like the reference's `.item()`."""
    model_genes, sample_genes = np.asarray(model_genes), np.asarray(sample_genes)
    midx = pd.Index(model_genes)
    if midx.has_duplicates:
        dup = set(midx[midx.duplicated()])
        if dup & set(sample_genes.tolist()):
            raise ValueError("can only convert an array of size 1 to a Python scalar")
    pos = pd.Series(np.arange(len(model_genes)), index=midx)
    pos = pos[~pos.index.duplicated(keep="first")]
    dest = pos.reindex(pd.Index(sample_genes)).to_numpy()
    dest = np.where(np.isnan(dest), -1, dest).astype(np.int32)
    sidx = pd.Index(sample_genes)
    if sidx.has_duplicates:
        dest[sidx.duplicated(keep="last")] = -1
    return dest


def build_classification_matrix(X, model_genes: np.ndarray, sample_genes: np.ndarray, gene_batch_size: int = 512):
    """[Cells, len(model_genes)] matrix in the model's gene order; genes the sample lacks stay zero
    (scnym/utils.py:155-233).  The reference maps genes with an O(genes^2) Python loop and copies
    through a LIL matrix; here the map is one hash join (`gene_column_map`).  A `DeviceCSR` is re-indexed in HBM by
    `dataprep.remap_columns` (scnb_csr_remap_count / scnb_csr_remap_fill) and comes back as a `DeviceCSR`; the two host
    types the reference accepts come back as the same host type (`type(N)` matches `type(X)`, utils.py:181-182)."""
    from .dataprep import DeviceCSR
    if type(X) not in (np.ndarray, sparse.csr_matrix, DeviceCSR):
        raise TypeError(f"X is type {type(X)}, must `np.ndarray` or `sparse.csr_matrix`")
    model_genes, sample_genes = np.asarray(model_genes), np.asarray(sample_genes)
    if len(model_genes) == len(sample_genes) and np.all(model_genes == sample_genes):
        print("Gene names match exactly, returning input.")
        return X
    dest_i = gene_column_map(model_genes, sample_genes)
    present = dest_i >= 0
    n_cells = X.shape[0]
    if type(X) == DeviceCSR:
        import torch
        from .dataprep import remap_columns
        N = remap_columns(X, torch.as_tensor(dest_i).to(X.device), len(model_genes))
    elif type(X) == np.ndarray:
        N = np.zeros((n_cells, len(model_genes)))
        N[:, dest_i[present]] = X[:, np.where(present)[0]]
    else:
        X = sparse.csr_matrix(X)
        keep = present[X.indices]
        row_of = np.repeat(np.arange(n_cells), np.diff(X.indptr))
        lens = np.bincount(row_of[keep], minlength=n_cells)
        indptr = np.concatenate([[0], np.cumsum(lens)])
        N = sparse.csr_matrix((X.data[keep], dest_i[X.indices[keep]], indptr), shape=(n_cells, len(model_genes)))
        N.sort_indices()
    print("Found %d common genes." % int(present.sum()))
    return N


class _Names(pd.Index):
    pass


class SimpleAnnData(object):
    """The slice of the AnnData interface that `scnym_api` touches: `X`, `obs` (DataFrame),
    `var_names`, `obs_names`, `uns`, `obsm`, `shape`, boolean/index row slicing with `[rows, :]`,
    `obs_keys()`, `copy()`."""

    def __init__(self, X, obs: pd.DataFrame = None, var_names=None, obs_names=None) -> None:
        self.X = X
        n, g = X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=[str(i) for i in range(n)])
        if obs_names is not None:
            self.obs.index = pd.Index(obs_names)
        self.var_names = pd.Index(var_names if var_names is not None else [f"gene{i}" for i in range(g)])
        self.uns, self.obsm = {}, {}

    @property
    def obs_names(self):
        return self.obs.index

    @property
    def shape(self):
        return self.X.shape

    def obs_keys(self):
        return list(self.obs.columns)

    def var_names_make_unique(self):
        seen, out = {}, []
        for v in self.var_names:
            if v in seen:
                seen[v] += 1
                out.append(f"{v}-{seen[v]}")
            else:
                seen[v] = 0
                out.append(v)
        self.var_names = pd.Index(out)

    def __getitem__(self, key):
        rows, cols = key if isinstance(key, tuple) else (key, slice(None))
        rows = np.asarray(rows)
        X = self.X[rows] if cols == slice(None) else self.X[rows][:, cols]
        out = SimpleAnnData(X, self.obs.loc[rows].copy() if rows.dtype == bool else self.obs.iloc[rows].copy(),
                            self.var_names if cols == slice(None) else self.var_names[cols])
        return out

    def copy(self):
        out = SimpleAnnData(self.X.copy(), self.obs.copy(), self.var_names.copy())
        out.uns, out.obsm = dict(self.uns), dict(self.obsm)
        return out
