"""Micro-benchmark of the fused gradient reduce-scatter + optimizer + parameter all-gather kernel under torchrun.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_dp_exchange.py [--total 5330000]
"""
import argparse, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scnym_b200._lib import call, ptr
from scnym_b200.parallel import PeerExchange

ap = argparse.ArgumentParser()
ap.add_argument("--total", type=int, default=5330000)
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
total = a.total // 4 * 4
rnd = lambda b: (b + 255) // 256 * 256
off_flags, off_grad = 0, 1024
off_flat = off_grad + rnd(4 * total)
px = PeerExchange.from_dist(off_flat + rnd(4 * total), dev)
grad, flat = px.tensor(off_grad, total), px.tensor(off_flat, total)
grad.normal_(); flat.normal_()
s1, s2 = torch.zeros(total, device=dev), torch.zeros(total, device=dev)
hp = torch.tensor([1e-3, 1e-4, 0.9, 0.999, 1e-8, 0.9, 0.9, 0.0], device=dev)
state = torch.zeros(4, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
def step():
    call("scnb_step_begin", ptr(state), st)
    call("scnb_dp_reduce_optim_bcast", 1, ptr(px.ptrs), rank, world, off_flags, off_grad, off_flat, total, ptr(s1), ptr(s2), ptr(hp), ptr(state), ptr(state), st)
for _ in range(5):
    step()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    step()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / a.iters * 1e3
link = 4.0 * total * (world - 1) / world
print(f"rank {rank}/{world}: {us:.1f} us per exchange of {4 * total / 1e6:.1f} MB; NVLink bytes per direction {link / 1e6:.1f} MB -> {link / us / 1e3:.0f} GB/s", flush=True)
dist.barrier()
os._exit(0)
