"""Phases of the tensor-core first layer (split forward, training form) at the config-2 / config-3 shapes: burst timing with and
without the precomputed split table, and per-CTA %globaltimer stamps (SCNB_FWDTC_DBG=1).   python tools/exp_fwd.py [G]"""
import ctypes
import os
import sys

os.environ["SCNB_FWDTC_DBG"] = "1" if "--stamps" in sys.argv else "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scnym_b200._lib import call, ptr
from scnym_b200.synth import synth_device_csr
from scnym_b200.dataprep import gather_rows

args = [a for a in sys.argv[1:] if not a.startswith("--")]
G, H0, n = (int(args[0]) if args else 20000), 256, 512
ds, _ = synth_device_csr(4096, G, 0.05, 16, seed=0, device="cuda")
b = gather_rows(ds, torch.arange(n).cuda())
W = torch.randn(G, H0).cuda() * 0.01
n_kt = (G + 63) // 64
hi = torch.zeros(n_kt * H0 * 64, dtype=torch.int16).cuda(); lo = torch.zeros_like(hi)
st = torch.cuda.current_stream().cuda_stream
call("scnb_w1_planes", ptr(W), G, H0, ptr(hi), ptr(lo), st)
ks, gp = ctypes.c_int(0), ctypes.c_int(0)
call("scnb_csr_linear_fwd_tc_geometry", n, G, H0, ctypes.addressof(ks), ctypes.addressof(gp))
sp = torch.zeros((ks.value + 1) * n, dtype=torch.int32).cuda()
call("scnb_row_fwd_splits", ptr(b.indptr), ptr(b.indices), n, G, H0, ptr(sp), st)
Z = torch.zeros(n, H0).cuda()
b1 = torch.zeros(H0).cuda()
print(f"G={G}: {ks.value} gene splits of {gp.value} genes, nnz {int(b.indptr[-1])}")


def launch(table):
    if table:
        call("scnb_csr_linear_fwd_tc_splits", ptr(b.indptr), ptr(b.indices), ptr(b.data), n, G, H0, ptr(hi), ptr(lo), ptr(b1), ptr(Z), -1, ptr(sp), st)
    else:
        call("scnb_csr_linear_fwd_tc", ptr(b.indptr), ptr(b.indices), ptr(b.data), n, G, H0, ptr(hi), ptr(lo), ptr(b1), ptr(Z), -1, st)


for table in (0, 1):
    for _ in range(3):
        launch(table)
    torch.cuda.synchronize()
    if os.environ["SCNB_FWDTC_DBG"] == "1":
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        launch(table)
    e1.record()
    torch.cuda.synchronize()
    print(f"split table {table}: burst {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
