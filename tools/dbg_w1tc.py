"""Micro-driver for the tensor-core dW1 + optimizer kernel (phase stamps with SCNB_W1TC_DBG=1; target for ncu)."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scnym_b200._lib import call, lib, ptr
from scnym_b200.synth import synth_device_csr
from scnym_b200.dataprep import gather_rows
G, H0, n = 20000, 256, 512
ds, _ = synth_device_csr(4096, G, 0.05, 16, seed=0, device="cuda")
b = gather_rows(ds, torch.arange(n).cuda())
GT = lib().scnb_w1tc_gene_block(G); n_blk = (G + GT - 1) // GT
splits = torch.zeros((n_blk + 1) * n, dtype=torch.int32).cuda()
dz_hi = torch.zeros(n * H0, dtype=torch.int16).cuda(); dz_lo = torch.zeros(n * H0, dtype=torch.int16).cuda()
W = torch.randn(G, H0).cuda() * 0.01; s1 = torch.zeros(G, H0).cuda(); s2 = torch.zeros(G, H0).cuda()
hp = torch.tensor([1e-3, 1e-4, 0.9, 0.999, 1e-8, 0.9, 0.9, 0.0]).cuda(); state = torch.ones(4, dtype=torch.int64).cuda()
dZ = torch.randn(n, H0).cuda()
st = torch.cuda.current_stream().cuda_stream
call("scnb_row_gene_splits", ptr(b.indptr), ptr(b.indices), n, G, ptr(splits), st)
call("scnb_w1_planes", ptr(dZ), n, H0, ptr(dz_hi), ptr(dz_lo), st)
pl_hi = torch.zeros(((G + 63) // 64) * H0 * 64, dtype=torch.int16).cuda(); pl_lo = torch.zeros_like(pl_hi)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(4):
    if os.environ.get("W1TC_FLUSH", "0") == "1":
        flush.fill_(it)
    ev[0].record()
    call("scnb_csr_w1_bwd_optim_tc", 1, ptr(splits), ptr(b.indices), ptr(b.data), n, H0, G, ptr(dz_hi), ptr(dz_lo), ptr(W), None, ptr(s1), ptr(s2), ptr(hp), ptr(state), ptr(pl_hi), ptr(pl_lo), st)
    ev[1].record()
    torch.cuda.synchronize()
    print("launch", it, "ms", ev[0].elapsed_time(ev[1]), file=sys.stderr)
