"""Where does the epilogue of the tensor-core dW1 + optimizer kernel spend its time?  Times the kernel's experiment modes
(SCNB_W1TC_EXP: 0 production, 1 no parameter/state loads, 2 no stores, 3 no L2 prefetch, 4 L2 evict_last policy):
  burst   : CUDA events around 20 back-to-back launches (what bench.py's roofline uses)
  spaced  : events around single launches separated by a 48 MB streaming fill (the rest of a step polluting L2)
  phases  : per-CTA globaltimer stamps (SCNB_W1TC_DBG=1)
Usage: python tools/exp_w1tc.py [modes...]"""
import os, sys, torch
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
dZ = torch.randn(n, H0).cuda() * 1e-3
st = torch.cuda.current_stream().cuda_stream
call("scnb_row_gene_splits", ptr(b.indptr), ptr(b.indices), n, G, ptr(splits), st)
call("scnb_w1_planes", ptr(dZ), n, H0, ptr(dz_hi), ptr(dz_lo), st)
pl_hi = torch.zeros(((G + 63) // 64) * H0 * 64, dtype=torch.int16).cuda(); pl_lo = torch.zeros_like(pl_hi)
pollute = torch.empty(48 << 20, dtype=torch.uint8, device="cuda")
ALG = 8 * int(b.indptr[-1].item()) + 4 * n * H0 + 24 * G * H0


def launch():
    call("scnb_csr_w1_bwd_optim_tc", 1, ptr(splits), ptr(b.indices), ptr(b.data), n, H0, G, ptr(dz_hi), ptr(dz_lo), ptr(W), None,
         ptr(s1), ptr(s2), ptr(hp), ptr(state), ptr(pl_hi), ptr(pl_lo), st)


def ev():
    return torch.cuda.Event(enable_timing=True)


modes = [int(x) for x in sys.argv[1:]] or [0, 6, 8, 9]

# every form must produce the same bits as the register form (same MMA order, same optimizer arithmetic)
W0, s10, s20 = W.clone(), s1.clone(), s2.clone()
ref = None
for m in [0] + [x for x in modes if x != 0]:
    os.environ["SCNB_W1TC_EXP"] = str(m)
    W.copy_(W0); s1.copy_(s10); s2.copy_(s20); pl_hi.zero_(); pl_lo.zero_()
    for _ in range(2):
        launch()
    torch.cuda.synchronize()
    got = [x.clone() for x in (W, s1, s2, pl_hi, pl_lo)]
    if ref is None:
        ref = got
        assert not torch.equal(W, W0)
    else:
        ok = [bool(torch.equal(x, y)) for x, y in zip(got, ref)]
        print(f"mode {m}: identical to the register form (W, m, v, planes hi, lo): {ok}", flush=True)
for m in modes:
    os.environ["SCNB_W1TC_EXP"] = str(m)
    os.environ.pop("SCNB_W1TC_DBG", None)
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(20):
            launch()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    sp = []
    for i in range(12):
        pollute.fill_(i)
        e0, e1 = ev(), ev()
        e0.record(); launch(); e1.record()
        torch.cuda.synchronize()
        sp.append(e0.elapsed_time(e1))
    sp = sorted(sp[2:])
    print(f"mode {m}: burst {best * 1e3:.1f} us ({ALG / best / 1e6:.0f} GB/s algorithmic), spaced median {sp[len(sp) // 2] * 1e3:.1f} us min {sp[0] * 1e3:.1f} us",
          flush=True)
