"""Small hidden-layer products of config 2 (1024 x 256 x 256): the latency-oriented mma.sync kernel against the tcgen05
kernel with narrow tiles (SCNB_GEMM_TC_BN=16..256, default: auto).  Prints time per launch and the error against fp64.
    python tools/bench_gemm_small.py [M N K]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scnym_b200._lib import call, ptr
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (1024, 256, 256)
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
for lay, name, (m, n, k) in ((0, "fwd  Z = A W^T", (M, N, K)), (1, "dA = dZ W", (M, K, N)), (2, "dW = dZ^T A", (N, K, M))):
    tA, tB = lay == 2, lay == 0
    A = torch.randn((k, m) if tA else (m, k), device="cuda")
    B = torch.randn((n, k) if tB else (k, n), device="cuda")
    bias = torch.randn(n, device="cuda") if lay == 0 else None
    beta = 1.0 if lay == 2 else 0.0
    want = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    if bias is not None:
        want = want + bias.double()
    for mode, thr in (("mma.sync 3xTF32", 0), ("tcgen05 bf16x3", 1)):
        call("scnb_gemm_tc_min_flops", thr)
        C = torch.zeros(m, n, device="cuda")
        f = lambda: call("scnb_sgemm", int(tA), int(tB), m, n, k, ptr(A), A.shape[1], ptr(B), B.shape[1], ptr(C), n, 1.0, beta, ptr(bias), 0, None, st)
        f(); torch.cuda.synchronize()
        err = float((C.double() - want).abs().max() / want.abs().max())
        for _ in range(5): f()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(20): f()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20)
        print(f"lay {lay} {name:16s} M={m} N={n} K={k} {mode:16s}: {best * 1e3:6.2f} us/launch (back to back)  rel err {err:.2e}", flush=True)
