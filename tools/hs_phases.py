"""Phase timing inside the fused hidden-stack tile GEMM (k_hs_fwd): one launch at the config-2 shape with SCNB_HS_DBG=1,
%globaltimer stamps of CTA (0, 0) at the phase boundaries.   python tools/hs_phases.py"""
import ctypes
import os
import sys

os.environ["SCNB_HS_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scnym_b200 import _lib
from scnym_b200._lib import call, ptr

dev = torch.device("cuda", 0)
R, K, N = 1024, 256, 256
g = torch.Generator(device=dev).manual_seed(0)
Z = torch.randn(R, K, device=dev, generator=g)
W = torch.randn(N, K, device=dev, generator=g) * 0.05
b = torch.zeros(N, device=dev)
gam, bet = torch.ones(K, device=dev), torch.zeros(K, device=dev)
part = torch.zeros(R // 32, 2, K, device=dev)
part2 = torch.zeros(R // 32, 2, N, device=dev)
stats = torch.zeros(3, 3, K, device=dev)
A = torch.zeros(R, K, device=dev)
Zo = torch.zeros(R, N, device=dev)
rng = torch.tensor([123, 1], dtype=torch.int64, device=dev)
keep2 = torch.ones(R, N, dtype=torch.uint8, device=dev)
seg = np.asarray([0, 256, 512, 1024], dtype=np.int32)
st = torch.cuda.current_stream().cuda_stream
keep = torch.ones(R, K, dtype=torch.uint8, device=dev)
call("scnb_hs_mix", ptr(Z), K, None, None, None, R, ptr(Z), ptr(part), ptr(keep), ptr(rng), 16, 0.5, st)
names = ["start", "cp.async issued", "stats combined", "operands landed", "A transformed", "mma + reduce", "epilogue"]
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call("scnb_hs_fwd", ptr(Z), ptr(A), ptr(part), ptr(stats), ptr(gam), ptr(bet), ptr(keep), 0.5, ptr(W), ptr(b), ptr(Zo), ptr(part2),
         ptr(keep2), ptr(rng), 17, 0.5, K, N, R, 3, seg.ctypes.data, None, None, st)
    e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    call("scnb_hs_debug_read", ctypes.cast(buf, ctypes.c_void_p))
    t = np.array(buf[:7], dtype=np.int64)
    print(f"launch {it}: events {e0.elapsed_time(e1) * 1e3:.1f} us; CTA(0,0) phases (us since its start): " +
          ", ".join(f"{n} {(x - t[0]) / 1e3:.2f}" for n, x in zip(names, t)))
