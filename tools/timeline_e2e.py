"""GPU activity (kernels + copies) of steady-state steps of the host-buffer pipeline (hostfeed.HostTrainPipeline), from CUPTI
through torch.profiler: where the device waits inside an end-to-end step.   python tools/timeline_e2e.py > gpurun_out/tl.txt"""
import json, os, sys, tempfile
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from scnym_b200.hostfeed import HostTrainPipeline

dev = torch.device("cuda", 0)
c = bench.CFG
n_host = 20000
model, dann, cfg, lab, y, unl = bench.build_train(dev, 0, seed_off=7, n_lab=n_host, n_unl=n_host)
host = lambda t: t.cpu().numpy()
if os.environ.get("FEED", "staged") == "mapped":
    from scnym_b200.hostfeed import HostMappedTrainPipeline
    pipe = HostMappedTrainPipeline(model, cfg, lab=(host(lab.indptr), host(lab.indices), host(lab.data), host(y)),
                                   unl=(host(unl.indptr), host(unl.indices), host(unl.data)), n_genes=c["G"], batch_size=c["L"],
                                   unlabeled_batch_size=c["U"], dann=dann, device=dev, seed=5)
else:
    pipe = HostTrainPipeline(model, cfg, lab=(host(lab.indptr), host(lab.indices), host(lab.data), host(y)), unl=(host(unl.indptr), host(unl.indices), host(unl.data)),
                             n_genes=c["G"], batch_size=c["L"], unlabeled_batch_size=c["U"], dann=dann, device=dev, n_threads=4, seed=5, n_producers=3)
pipe.engine.set_weights(1.0, 0.1)
for _ in pipe.run(30):
    pass
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in pipe.run(12):
        pass
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
first = "k_zero2"
starts = [i for i, e in enumerate(ev) if e["name"].startswith(first)]
k = len(starts) // 2
t0 = ev[starts[k]]["ts"]
print(f"# steps at: " + ", ".join(f"{ev[i]['ts'] - t0:.0f}" for i in starts))
seg = ev[starts[k] - 3: starts[k + 2]]
streams = sorted({e["args"].get("stream") for e in seg})
for e in seg:
    s = streams.index(e["args"].get("stream"))
    nm = e["name"][:50] + (f" [{e['args'].get('bytes', '')} B]" if e.get("cat") != "kernel" else "")
    print(f"{e['ts'] - t0:8.1f} +{e['dur']:6.1f} = {e['ts'] - t0 + e['dur']:8.1f}  s{s}  {'  ' * s}{nm}")
