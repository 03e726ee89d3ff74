"""Small driver for ncu: builds the config-2 engine and runs a few MixMatch+DAN steps (eager
launches, Philox dropout) so that the launch list shows one line per kernel of the step.
    python tools/profile_step.py [--steps 3] [--predict]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cells", type=int, default=20000)
    ap.add_argument("--predict", action="store_true")
    ap.add_argument("--single-stream", action="store_true")
    ap.add_argument("--config", type=int, default=2)
    a = ap.parse_args()
    bench.CFG, _ = bench.CONFIGS[a.config]
    dev = torch.device("cuda", 0)
    if a.predict:
        from scnym_b200.model import CellTypeCLF
        from scnym_b200.predict import EvalRunner
        from scnym_b200.synth import synth_device_csr
        c = bench.CFG
        tile, _ = synth_device_csr(1 << 16, c["G"], c["density"], c["C"], seed=2, device=dev)
        model = CellTypeCLF(n_genes=c["G"], n_cell_types=c["C"], n_hidden=c["H"], n_hidden_init=c["H0"], n_layers=c["n_layers"]).to(dev).eval()
        run = EvalRunner([model])
        run.run(tile, want_prob=True, want_embed=True)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        for _ in range(a.steps):
            run.run(tile, want_prob=True, want_embed=True)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    from scnym_b200.engine import TrainEngine
    model, dann, cfg, lab, y, unl = bench.build_train(dev, 0, n_lab=a.cells, n_unl=a.cells)
    cfg.use_cuda_graph = False
    cfg.multi_stream = not a.single_stream
    eng = TrainEngine(model.to(dev), cfg, lab, y.to(torch.int32), bench.CFG["L"], unlabeled=unl, unlabeled_batch_size=bench.CFG["U"],
                      dann=dann, mode="mixmatch")
    eng.set_weights(1.0, 0.1)
    eng.sample_epoch_tables(a.steps + 2)
    for i in range(a.steps + 2):
        if i == 2:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()   # ncu --profile-from-start off: steady-state steps only
        eng.step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("ok", eng.losses())


if __name__ == "__main__":
    main()
