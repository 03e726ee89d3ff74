"""Why is the 4.6 MB upload of the host-buffer pipeline slow?  H2D of one pinned blob: alone, while training steps replay
on another stream, while producer threads gather into other blobs, and split over two streams."""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from scnym_b200.hostfeed import HostTrainPipeline
from scnym_b200.engine import TrainEngine

dev = torch.device("cuda", 0)
c = bench.CFG
model, dann, cfg, lab, y, unl = bench.build_train(dev, 0, seed_off=7, n_lab=20000, n_unl=20000)
host = lambda t: t.cpu().numpy()
pipe = HostTrainPipeline(model, cfg, lab=(host(lab.indptr), host(lab.indices), host(lab.data), host(y)), unl=(host(unl.indptr), host(unl.indices), host(unl.data)),
                         n_genes=c["G"], batch_size=c["L"], unlabeled_batch_size=c["U"], dann=dann, device=dev, n_threads=4, seed=5, n_producers=3)
pipe.engine.set_weights(1.0, 0.1)
for _ in pipe.run(10):
    pass
torch.cuda.synchronize()
nbytes = pipe.layout.total_bytes
src = pipe.blobs[0]
dst = pipe.land[0]
cs, cs2, comp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
own = torch.empty(nbytes, dtype=torch.uint8).pin_memory()

def h2d(tag, src=src, split=False, n=60):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record(cs)
        for _ in range(n):
            if split:
                h = nbytes // 2 // 256 * 256
                dst[:h].copy_(src[:h], non_blocking=True)
                cs2.wait_stream(cs)
                with torch.cuda.stream(cs2):
                    dst[h:].copy_(src[h:], non_blocking=True)
                cs.wait_stream(cs2)
            else:
                dst.copy_(src, non_blocking=True)
        e1.record(cs)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{tag:60s} {us:7.1f} us per {nbytes / 1e6:.2f} MB -> {nbytes / us / 1e3:5.1f} GB/s", flush=True)

h2d("alone (blob written by a producer)")
h2d("alone (fresh pinned tensor)", src=own)
h2d("alone, two halves on two streams", split=True)
eng = pipe.engine
stop = [False]
def steps():
    with torch.cuda.stream(comp):
        while not stop[0]:
            eng.step()
            comp.synchronize()
th = threading.Thread(target=steps); th.start(); time.sleep(0.2)
h2d("while training steps replay on another stream")
stop[0] = True; th.join()
stop[0] = False
def prod(k):
    i = 0
    while not stop[0]:
        pipe._produce(k, i); i += 1
ths = [threading.Thread(target=prod, args=(k,)) for k in (1, 2, 3)]
for t in ths: t.start()
time.sleep(0.2)
h2d("while three producers gather into other blobs")
stop[0] = True
for t in ths: t.join()
