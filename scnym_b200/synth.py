"""Synthetic log1p(CPM) expression matrices generated directly on the GPU (bench / smoke inputs).

Follows the recipe of SURVEY.md §8d: per-row nnz ~ clip(Poisson(d*G), 1, G); counts
1 + Poisson(3*lambda[class, gene]) with lambda ~ Gamma(0.5, 2) + 0.05 per class; values
log1p(1e6 * count / rowsum) in (0, 13.82]; labels uniform.  Column indices are drawn by
stratified sampling (one uniform draw inside each of `nnz` equal-width integer strata of
[0, G)), which yields sorted, duplicate-free, gene-uniform indices without a per-row sort.
"""
from typing import Optional, Tuple

import torch

from .dataprep import DeviceCSR


def synth_device_csr(n_cells: int, n_genes: int, density: float, n_classes: int, seed: int, device="cuda",
                     chunk_rows: int = 32768) -> Tuple[DeviceCSR, torch.Tensor]:
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    lam = torch._standard_gamma(torch.full((n_classes, n_genes), 0.5, device=dev), generator=g) * 2.0 + 0.05
    labels = torch.randint(0, n_classes, (n_cells,), device=dev, generator=g)
    nnz_row = torch.poisson(torch.full((n_cells,), float(density * n_genes), device=dev), generator=g).clamp_(1, n_genes).long()
    indptr = torch.zeros(n_cells + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(nnz_row, 0)
    total = int(indptr[-1].item())
    indices = torch.empty(total, dtype=torch.int32, device=dev)
    data = torch.empty(total, dtype=torch.float32, device=dev)
    for r0 in range(0, n_cells, chunk_rows):
        r1 = min(n_cells, r0 + chunk_rows)
        k = nnz_row[r0:r1]
        s, e = int(indptr[r0].item()), int(indptr[r1].item())
        rows = torch.repeat_interleave(torch.arange(r1 - r0, device=dev), k)
        j = torch.arange(e - s, device=dev) - (indptr[r0:r1] - s)[rows]
        kk = k[rows]
        lo = (j * n_genes) // kk
        hi = ((j + 1) * n_genes) // kk
        u = torch.rand(e - s, device=dev, generator=g)
        col = lo + (u * (hi - lo).float()).long().clamp_(max=n_genes - 1)
        col = torch.minimum(col, hi - 1)
        rate = 3.0 * lam[labels[r0:r1][rows], col]
        cnt = 1.0 + torch.poisson(rate, generator=g)
        rowsum = torch.zeros(r1 - r0, device=dev).index_add_(0, rows, cnt)
        val = torch.log1p(1e6 * cnt / rowsum[rows])
        indices[s:e] = col.to(torch.int32)
        data[s:e] = val
        del rows, j, kk, lo, hi, u, col, rate, cnt, rowsum, val
    max_nnz = int(nnz_row.max().item())
    return DeviceCSR(indptr, indices, data, n_genes, max_nnz), labels
