"""Phase timing inside the fused head kernels (head_ops.cu) at the config-2 shape: SCNB_HS_DBG=1, %globaltimer stamps of
CTA (0, 0) at the phase boundaries.   python tools/head_phases.py"""
import ctypes
import os
import sys

os.environ["SCNB_HS_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scnym_b200._lib import call, ptr

dev = torch.device("cuda", 0)
R, H, C, D, U, L = 1024, 256, 16, 2, 256, 256
nb = U + L
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
Z = rn(R, H)
gam, bet = torch.ones(H, device=dev), torch.zeros(H, device=dev)
part, part_out = torch.zeros(R // 32, 2, H, device=dev), torch.zeros(R // 32, 2, H, device=dev)
stats = torch.zeros(3, 3, H, device=dev)
A, dY = torch.zeros(R, H, device=dev), torch.zeros(R, H, device=dev)
keep = (torch.rand(R, H, device=dev, generator=g) > 0.5).to(torch.uint8)
seg = np.asarray([0, U, nb, R], dtype=np.int32)
st = torch.cuda.current_stream().cuda_stream
rng = torch.tensor([123, 1], dtype=torch.int64, device=dev)
call("scnb_hs_mix", ptr(Z), H, None, None, None, R, ptr(Z), ptr(part), None, ptr(rng), 16, 0.5, st)
Wc, bc = rn(C, H) * 0.05, torch.zeros(C, device=dev)
Y = torch.softmax(rn(nb, C), 1)
ident = torch.arange(nb, dtype=torch.int32, device=dev)
gm = torch.full((nb,), 0.7, device=dev)
counts = torch.tensor([U, L, R - nb, R], dtype=torch.int32, device=dev)
logits, dlogits = torch.zeros(nb, C, device=dev), torch.zeros(nb, C, device=dev)
loss8 = torch.zeros(8, device=dev)
W1, b1 = rn(H, H) * 0.05, torch.zeros(H, device=dev)
W2, b2 = rn(D, H) * 0.05, torch.zeros(D, device=dev)
Rd = R - nb
dom = torch.randint(0, D, (nb,), device=dev, generator=g).to(torch.int32)
src = torch.arange(R, dtype=torch.int32, device=dev) % nb
Hd, dHd = torch.zeros(Rd, H, device=dev), torch.zeros(Rd, H, device=dev)
dlog, dlabel, ddlog = (torch.zeros(Rd, D, device=dev) for _ in range(3))


def classif():
    call("scnb_hs_classif_head", ptr(Z), ptr(A), ptr(part), ptr(stats), ptr(gam), ptr(bet), ptr(keep), 0.5, H, R, 3, seg.ctypes.data,
         ptr(Wc), ptr(bc), C, U, L, ptr(counts), ptr(Y), ptr(ident), ptr(ident), ptr(gm), None, 1.0, ptr(logits), ptr(dlogits),
         ptr(loss8), ptr(dY), ptr(part_out), st)


def dan():
    call("scnb_hs_dan_head", ptr(Z), ptr(A), ptr(part), ptr(stats), ptr(gam), ptr(bet), ptr(keep), 0.5, H, R, nb, 3, seg.ctypes.data,
         ptr(W1), ptr(b1), ptr(W2), ptr(b2), D, ptr(counts[2:]), ptr(dom), ptr(src), nb, None, 0.1, ptr(Hd), ptr(dHd), ptr(dlog),
         ptr(dlabel), ptr(ddlog), ptr(loss8[4:]), ptr(dY), ptr(part_out), st)


names = {0: ["start", "rows + labels + constants + weights", "activation + logits", "losses (lane per row)", "backprop", "dY + CTA sums", "tile sums (cluster)"],
         1: ["start", "constants", "operands landed", "A transformed", "mma 1", "Hd + partial logits", "cluster sync 1", "logits", "CE",
             "dHd", "mma 2", "cluster sync 2", "dY + sums", "cluster sync 3"]}
for kern, fn in ((0, classif), (1, dan)):
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 32)()
        call("scnb_hh_debug_read", ctypes.cast(buf, ctypes.c_void_p))
        t = np.array(buf[16 * kern: 16 * kern + len(names[kern])], dtype=np.int64)
        print(f"{'classif' if kern == 0 else 'dan'} head, launch {it}: events {e0.elapsed_time(e1) * 1e3:.1f} us; CTA(0,0) phases (us since its start): " +
              ", ".join(f"{n} {(x - t[0]) / 1e3:.2f}" for n, x in zip(names[kern], t)), flush=True)
