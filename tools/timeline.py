"""Kernel timeline of graph-replayed training steps (CUPTI through torch.profiler): prints, for one steady-state
step, every kernel with its stream, start offset and duration, so the critical path of the step DAG can be read off.
    python tools/timeline.py [--steps 6] [--config 2] > gpurun_out/timeline.txt
"""
import argparse
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--cells", type=int, default=20000)
    ap.add_argument("--config", type=int, default=2)
    a = ap.parse_args()
    bench.CFG, _ = bench.CONFIGS[a.config]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    comm = None
    if world > 1:   # under torchrun: data-parallel step, rank 0 prints its timeline
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        from scnym_b200.parallel import TorchDistComm
        comm = TorchDistComm()
    from scnym_b200.engine import TrainEngine
    model, dann, cfg, lab, y, unl = bench.build_train(dev, rank, n_lab=a.cells, n_unl=a.cells)
    eng = TrainEngine(model.to(dev), cfg, lab, y.to(torch.int32), bench.CFG["L"], unlabeled=unl, unlabeled_batch_size=bench.CFG["U"],
                      dann=dann, mode="mixmatch", comm=comm)
    eng.set_weights(1.0, 0.1)
    eng.sample_epoch_tables(a.steps + 12)
    for _ in range(10):
        eng.step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.steps):
            eng.step()
        torch.cuda.synchronize()
    if rank != 0:
        os._exit(0)
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    # split into steps at the first kernel of the step
    first = "k_zero2" if any(e["name"].startswith("k_zero2") for e in ev) else "k_step_prep"
    if os.environ.get("SCNB_ZERO_AT_END") == "1":
        first = "void k_csr_dense_fwd_tc"
    starts = [i for i, e in enumerate(ev) if e["name"].startswith(first)]
    k = len(starts) // 2
    seg = ev[starts[k]: starts[k + 1]] if k + 1 < len(starts) else ev[starts[k]:]
    t0 = seg[0]["ts"]
    nxt = ev[starts[k + 1]]["ts"] if k + 1 < len(starts) else None
    print(f"# step {k}: {len(seg)} kernels; next step starts at +{(nxt - t0) if nxt else float('nan'):.1f} us")
    streams = sorted({e["args"].get("stream") for e in seg})
    for e in seg:
        s = streams.index(e["args"].get("stream"))
        print(f"{e['ts'] - t0:8.1f} +{e['dur']:6.1f} = {e['ts'] - t0 + e['dur']:8.1f}  s{s}  {'    ' * s}{e['name'][:60]}")
    sys.stdout.flush()
    if world > 1:
        os._exit(0)


if __name__ == "__main__":
    main()
