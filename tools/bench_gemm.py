"""Time scnb_sgemm on the hidden-layer shapes of config 5 (tcgen05 kernel) for the three nn.Linear layouts.
    python tools/bench_gemm.py [M N K]
Both forms are timed: operands split inside the kernel, and the planes form (workspace registered for the stream; the two
operand-plane passes are part of the time).
"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scnym_b200._lib import call, ptr
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (16384, 1024, 1024)
st = torch.cuda.current_stream().cuda_stream
for planes in (False, True):
  if planes:
    need = ctypes.c_longlong(0)
    nb = 0
    for (m, n, k) in ((M, N, K), (M, K, N), (N, K, M)):
        call("scnb_gemm_tc_workspace_bytes", m, n, k, ctypes.byref(need)); nb = max(nb, need.value)
    ws = torch.empty(nb + 1024, dtype=torch.uint8, device="cuda")
    call("scnb_gemm_tc_workspace", st, ws.data_ptr() + (-ws.data_ptr()) % 1024, nb)
  print("planes form" if planes else "in-kernel staging")
  for lay, name, (m, n, k) in ((0, "fwd  Z = A W^T", (M, N, K)), (1, "dA = dZ W", (M, K, N)), (2, "dW = dZ^T A", (N, K, M))):
      tA, tB = lay == 2, lay == 0
      A = torch.randn((k, m) if tA else (m, k), device="cuda")
      B = torch.randn((n, k) if tB else (k, n), device="cuda")
      C = torch.zeros(m, n, device="cuda")
      beta = 1.0 if lay == 2 else 0.0
      f = lambda: call("scnb_sgemm", int(tA), int(tB), m, n, k, ptr(A), A.shape[1], ptr(B), B.shape[1], ptr(C), n, 1.0, beta, None, 0, None, st)
      for _ in range(3): f()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      torch.cuda.synchronize(); e0.record()
      for _ in range(10): f()
      e1.record(); torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / 10
      fl = 2.0 * m * n * k
      print(f"lay {lay} {name:16s} M={m} N={n} K={k}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s fp32-equivalent, {3 * fl / ms / 1e9:7.1f} TFLOP/s issued bf16")
