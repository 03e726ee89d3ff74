"""Local (one GPU, world = 1) cost of the exchange kernel against the plain elementwise optimizer pass over the same
elements: separates the kernel's own processing rate from NVLink.   python tools/exp_dp_local.py [--total 3950000]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scnym_b200._lib import call, ptr

ap = argparse.ArgumentParser()
ap.add_argument("--total", type=int, default=3950000)
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
dev = torch.device("cuda", 0)
total = a.total // 4 * 4
rnd = lambda b: (b + 255) // 256 * 256
off_flags, off_grad = 0, 1024
off_flat = off_grad + rnd(4 * total)
buf = torch.zeros(off_flat + rnd(4 * total), dtype=torch.uint8, device=dev)
ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=dev)
grad = buf[off_grad: off_grad + 4 * total].view(torch.float32)
flat = buf[off_flat: off_flat + 4 * total].view(torch.float32)
grad.normal_(); flat.normal_()
s1, s2 = torch.zeros(total, device=dev), torch.zeros(total, device=dev)
hp = torch.tensor([1e-3, 1e-4, 0.9, 0.999, 1e-8, 0.9, 0.9, 0.0], device=dev)
state = torch.ones(4, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, tag, cold=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if cold:
        for _ in range(a.iters):
            flush.zero_()
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
    else:
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        tot = e0.elapsed_time(e1)
    us = tot / a.iters * 1e3
    print(f"{tag:44s} {'cold' if cold else 'warm'} {us:7.1f} us  -> {28.0 * total / us / 1e3:6.0f} GB/s of 28 B/element", flush=True)

for cold in (False, True):
    timed(lambda: call("scnb_optim_step", 1, ptr(flat), ptr(grad), ptr(s1), ptr(s2), total, ptr(hp), ptr(state), st), "scnb_optim_step (Adam)", cold)
    timed(lambda: call("scnb_dp_reduce_optim_bcast", 1, ptr(ptrs), 0, 1, off_flags, off_grad, off_flat, total, ptr(s1), ptr(s2), ptr(hp),
                       ptr(state), ptr(state), st), f"dp_reduce_optim_bcast world=1 stages={os.environ.get('SCNB_DP_STAGES', 'auto')}", cold)
