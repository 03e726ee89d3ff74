"""Gene re-indexing kernels on one predict chunk (75 776 rows x 20 000 genes, 5 %): persistent form against the CTA-per-row
form (SCNB_REMAP_FORM=1).  python tools/bench_remap.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from scnym_b200 import _lib
from scnym_b200.synth import synth_device_csr
from scnym_b200.dataprep import remap_columns
c = bench.CFG
n = 75776
tile, _ = synth_device_csr(n, c["G"], c["density"], c["C"], seed=2, device="cuda")
g = torch.Generator().manual_seed(7)
cm = torch.randperm(c["G"], generator=g).to(torch.int32); cm[torch.rand(c["G"], generator=g) < 0.1] = -1
cm = cm.cuda()
nnz_in = tile.nnz
out = (torch.empty(n + 1, dtype=torch.int64, device="cuda"), torch.empty(nnz_in, dtype=torch.int32, device="cuda"),
       torch.empty(nnz_in, dtype=torch.float32, device="cuda"), torch.empty(_lib.lib().scnb_csr_remap_workspace_len(n), dtype=torch.int64, device="cuda"))
part = remap_columns(tile, cm, c["G"], 0, n, out=out)
torch.cuda.synchronize()
nnz_out = int(part.indptr[n].item())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    remap_columns(tile, cm, c["G"], 0, n, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
alg = 12.0 * nnz_in + 8.0 * nnz_out
print(f"form {os.environ.get('SCNB_REMAP_FORM', '2')}: {ms:.3f} ms per chunk, {alg / ms / 1e6:.0f} GB/s algorithmic, {n / ms / 1e3:.1f} M cells/s, checksum {int(part.indices[:nnz_out].to(torch.int64).sum().item())} {float(part.data[:nnz_out].double().sum().item()):.6f}")
