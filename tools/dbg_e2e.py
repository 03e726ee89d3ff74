"""Where the host-buffer (e2e) step goes: wall time spent waiting for the producers, enqueuing, and waiting for the device."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from scnym_b200.hostfeed import HostTrainPipeline
dev = torch.device("cuda", 0)
c = bench.CFG
n_host = 20000
model, dann, cfg, lab, y, unl = bench.build_train(dev, 0, seed_off=7, n_lab=n_host, n_unl=n_host)
host = lambda t: t.cpu().numpy()
n_prod = int(os.environ.get("NPROD", "3")); n_thr = int(os.environ.get("NTHR", "4"))
pipe = HostTrainPipeline(model, cfg, lab=(host(lab.indptr), host(lab.indices), host(lab.data), host(y)), unl=(host(unl.indptr), host(unl.indices), host(unl.data)),
                         n_genes=c["G"], batch_size=c["L"], unlabeled_batch_size=c["U"], dann=dann, device=dev, n_threads=n_thr, seed=5, n_producers=n_prod)
pipe.engine.set_weights(1.0, 0.1)
T = {"produce": 0.0, "h2d": 0.0, "step": 0.0, "read": 0.0}
def wrap(name, key):
    f = getattr(pipe, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); T[key] += time.perf_counter() - t0; return r
    setattr(pipe, name, g)
wrap("_enqueue_h2d", "h2d"); wrap("_enqueue_step", "step"); wrap("_read_loss", "read")
orig_prod = pipe._produce
tp = [0.0, 0]
seen = {}
def prod(k, j):
    if os.environ.get("NOPROD") and k in seen:      # experiment: re-upload the blob as it is (no host gather, no CPU writes)
        return seen[k]
    t0 = time.perf_counter(); r = orig_prod(k, j); tp[0] += time.perf_counter() - t0; tp[1] += 1
    seen[k] = r
    return r
pipe._produce = prod
for _ in pipe.run(20): pass
for k in T: T[k] = 0.0
tp[0], tp[1] = 0.0, 0
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 300
for _ in pipe.run(n): pass
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"producers {n_prod} x {n_thr} threads: {dt / n * 1e6:.0f} us/step -> {512 / (dt / n) / 1e6:.2f} M cells/s; per step: enqueue h2d {T['h2d'] / n * 1e6:.0f} us, enqueue step {T['step'] / n * 1e6:.0f} us, "
      f"wait for device (read loss) {T['read'] / n * 1e6:.0f} us, one _produce call {tp[0] / max(tp[1], 1) * 1e6:.0f} us")
