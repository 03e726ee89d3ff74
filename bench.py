#!/usr/bin/env python
"""bench.py -- cells/s of the MixMatch+DAN training step and of batched predict on N B200s.

    python bench.py --gpus 1 --steps 200 --warmup 20               # BASELINE.json configs[1] ("config 2")
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # configs[2] ("config 3")
    python bench.py --impl reference [--gpus N]                     # the UNMODIFIED reference on the host cores

One "step" = one minibatch of the reference's SemiSupervisedTrainer loop (scnym/trainer.py:575-671): L labeled + U
unlabeled cells through MixMatch + DAN + optimizer.
`value`   : whole-job cells/s with the CSR datasets resident in HBM (CUDA events, max over ranks).
`e2e`     : the same step fed from HOST CSR buffers (pinned staging, H2D inside the timed region, loss read back
            every step); the host matrices are the full workload (not a sample).
`roofline`: the kernel with the largest share of the step (chosen among ALL kernels), timed with CUDA events.
`cpu_baseline` / `gpu_eager_baseline`: the reference's own classes (oracle/_ref, `pip --target` of the unmodified
            reference; the oracle port when that is absent) on the host cores / in eager PyTorch on the same GPU,
            train step and predict.
`predict` : BASELINE.json configs[3] ("config 4") shape, device-resident and end to end from host memory.
N = 1 times config 2 (the configuration the metric is quoted on); N > 1 times config 3 (the data-parallel workload
BASELINE.json names); `--config` overrides.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cells/sec MixMatch+DAN train step"
DTYPE = "f32 (first layer and dW1 as bf16x3 split products on tcgen05, fp32 accumulate; hidden GEMMs 3xTF32)"
WORKLOAD = ("config2: MixMatch+DAN semi-supervised train, G=20000 genes, 100k labeled + 100k unlabeled cells, "
            "~5% CSR density, L=U=256, n_hidden_init=n_hidden=256, n_layers=2, C=16, D=2 (inferred domains), K=1, "
            "T=0.5, alpha=0.3, Adam(lr=1e-3, wd=1e-4)")
CFG = dict(G=20000, n_lab=100_000, n_unl=100_000, density=0.05, L=256, U=256, H0=256, H=256, n_layers=2, C=16, D=2)
# The other BASELINE.json configs (`--config N`)
CONFIGS = {
    2: (CFG, WORKLOAD),
    3: (dict(G=30000, n_lab=125_000, n_unl=125_000, density=0.05, L=256, U=256, H0=256, H=256, n_layers=2, C=16, D=8, given_domains=True),
        "config3: MixMatch+DAN with 8 given domains, G=30000 genes, 5% CSR density, per-rank L=U=256 (cells sharded over ranks; "
        "125k labeled + 125k unlabeled per rank = 1M cells at 4 ranks), n_hidden_init=n_hidden=256, n_layers=2, C=16, K=1, "
        "Adam(lr=1e-3, wd=1e-4)"),
    5: (dict(G=20000, n_lab=40_000, n_unl=40_000, density=0.20, L=4096, U=4096, H0=256, H=1024, n_layers=3, C=16, D=8, given_domains=True),
        "config5: large-batch MixMatch+DAN, G=20000 genes, ~20% CSR density, L=U=4096, n_hidden_init=256, n_hidden=1024, n_layers=3, "
        "C=16, D=8 given domains, K=1, Adam"),
    6: (dict(G=20000, n_lab=40_000, n_unl=40_000, density=0.20, L=4096, U=4096, H0=1024, H=1024, n_layers=3, C=16, D=8, given_domains=True),
        "config5 with n_hidden_init=1024 (SURVEY 8d: dense-contraction stress): G=20000 genes, ~20% CSR density, L=U=4096, "
        "n_hidden=1024, n_layers=3, C=16, D=8 given domains, K=1, Adam"),
}


def config_dict(world):
    """The `config` object of the JSON line (identical for our arm and the reference arm)."""
    return {"workload": WORKLOAD,
            "l2": "inputs larger than L2: >= 1.6 GB of CSR sampled at random + the ~107 MB parameter / optimizer arena touched "
                  "every step; no explicit flush",
            "parallelism": f"dp{world}" if world > 1 else "single",
            "sampler": "row / MixUp tables of the K+W steps are drawn on the device before the timed region (one draw per "
                       "epoch in training; its cost is reported in `sampler_draw`)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def peaks():
    d = load_peaks()
    if "hbm_gbs" in d:
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tensor_peak():
    d = load_peaks()
    if "bf16_tflops_sustained" in d:
        return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1590.0, "fallback (B200_PROFILING.md dense bf16)"


def ncu_traffic(entry):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed `ncu --set full` captures
    (profiles/r0N_traffic.json: measured under ncu, never during a timed run)."""
    if CFG is not CONFIGS[2][0]:      # the captures are of the config-2 step
        return None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            tab = json.load(open(p))
            for key in (entry, entry.split(" ")[0] + "]" if "[" in entry else entry, entry.split("[")[0]):
                v = tab.get(key, {}).get("dram_bytes_per_launch") if isinstance(tab.get(key), dict) else None
                if v is not None:
                    return v
    return None


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_train(device, rank, seed_off=0, n_lab=None, n_unl=None, host_copy=False):
    from scnym_b200.engine import EngineConfig, TrainEngine
    from scnym_b200.model import CellTypeCLF, DANN
    from scnym_b200.synth import synth_device_csr
    c = CFG
    n_lab, n_unl = n_lab or c["n_lab"], n_unl or c["n_unl"]
    lab, y = synth_device_csr(n_lab, c["G"], c["density"], c["C"], seed=0 + 100 * rank + seed_off, device=device)
    unl, _ = synth_device_csr(n_unl, c["G"], c["density"], c["C"], seed=1 + 100 * rank + seed_off, device=device)
    torch.manual_seed(0)
    model = CellTypeCLF(n_genes=c["G"], n_cell_types=c["C"], n_hidden=c["H"], n_hidden_init=c["H0"], n_layers=c["n_layers"])
    dann = DANN(model, n_domains=c["D"])
    cfg = EngineConfig(n_augmentations=1, T=0.5, augment_pseudolabels=False, pseudolabel_min_confidence=0.0, mixup_alpha=0.3,
                       use_dan=True, optimizer="adam", lr=1e-3, weight_decay=1e-4, seed=1234 + rank)
    return model, dann, cfg, lab, y, unl


def given_domains(device, n_lab, n_unl, rank):
    """Domain ids when the config supplies them (labeled rows use domains [0, D/2), unlabeled [D/2, D))."""
    if not CFG.get("given_domains"):
        return None, None
    g = torch.Generator().manual_seed(77 + rank)
    D = CFG["D"]
    ld = torch.randint(0, max(D // 2, 1), (n_lab,), generator=g, dtype=torch.int32).to(device)
    ud = torch.randint(max(D // 2, 1), D, (n_unl,), generator=g, dtype=torch.int32).to(device)
    return ld, ud


def time_device_resident(args, device, rank, world):
    """`value`: graph-replayed steps, datasets + sampler tables resident in HBM."""
    from scnym_b200.engine import TrainEngine
    model, dann, cfg, lab, y, unl = build_train(device, rank)
    ld, ud = given_domains(device, lab.n_rows, unl.n_rows, rank)
    eng = TrainEngine(model.to(device), cfg, lab, y.to(torch.int32), CFG["L"], unlabeled=unl, unlabeled_batch_size=CFG["U"],
                      dann=dann, labeled_domain=ld, unlabeled_domain=ud, mode="mixmatch", comm=make_comm(world))
    eng.set_weights(unsup_weight=1.0, dan_weight=0.1)
    n_tab = args.steps + args.warmup
    gen = torch.Generator(device=device)
    gen.manual_seed(1234 + rank)
    # cost of one epoch's sampler draw (torch library kernels: randperm / rand / _standard_gamma), reported beside the line
    n_epoch = max(lab.n_rows // CFG["L"], 1)
    eng.sample_epoch_tables(n_epoch, generator=gen)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.sample_epoch_tables(n_epoch, generator=gen)
    torch.cuda.synchronize()
    draw_s = time.perf_counter() - t0
    eng.sample_epoch_tables(n_tab, generator=gen)
    for _ in range(args.warmup):
        eng.step()
    barrier(world)
    torch.cuda.synchronize()
    cs = ClockSampler(index=torch.cuda.current_device())
    if rank == 0:
        cs.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.cudart().cudaProfilerStart()     # `ncu --profile-from-start off` lists exactly the timed steps' launches
    e0.record()
    for _ in range(args.steps):
        eng.step()
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    barrier(world)
    clocks = cs.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    losses = eng.losses()
    assert np.isfinite(losses["loss"]), losses
    if world > 1:
        # data-parallel invariant, checked outside the timed region on every run: after the same number of steps every
        # rank holds bit-identical parameters, optimizer step count and BatchNorm running statistics
        bufs = [eng.arena.flat] + [b["bn"].running_mean for b in eng._blk] + [b["bn"].running_var for b in eng._blk]
        sig = torch.stack([torch.cat([x.double().sum().view(1), x.double().abs().max().view(1)]) for x in bufs]).view(-1)
        sig = torch.cat([sig, eng.state[:1].double()])
        hi, lo = sig.clone(), sig.clone()
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        assert torch.equal(hi, lo) and bool(torch.isfinite(hi).all()), "data-parallel replicas diverged"
        losses["replicas_identical"] = True
    sampler = {"steps_per_epoch": n_epoch, "ms_per_epoch_draw": draw_s * 1e3, "us_per_step_amortised": draw_s * 1e6 / n_epoch}
    return eng, ms, clocks, losses, sampler


def launch_key(name, a):
    """Aggregation key of a C-ABI call in the per-kernel timing: GEMM launches are told apart by layout and shape."""
    if name == "scnb_sgemm":
        return f"scnb_sgemm[{'T' if a[0] else 'N'}{'T' if a[1] else 'N'} {a[2]}x{a[3]}x{a[4]}]"
    if name == "scnb_hs_fwd":     # (Zin .. K, N, R at 16..18, running_mean at 21): the student's train-mode launches and the
        mode = "eval" if a[21] else "train"   # teacher's eval-mode ones are different kernels' worth of work
        return f"scnb_hs_fwd[{mode} {a[18]}x{a[17]}x{a[16]}]"
    if name == "scnb_hs_bwd":
        return f"scnb_hs_bwd[{a[17]}x{a[16]}x{a[15]}]"
    if name == "scnb_hs_dw":
        return f"scnb_hs_dw[{a[1]}x{a[3]}x{a[4]}]"
    return name


def time_kernels(eng, n_steps=12):
    """Eager pass with CUDA events around every C-ABI call -> per-launch-key mean duration (ms)."""
    from scnym_b200 import _lib
    stream = torch.cuda.current_stream()
    records = []
    orig = _lib.call

    def timed(name, *a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        r = orig(name, *a)
        e.record(stream)
        k = launch_key(name, a)
        records.append((k, s, e))
        LAST_ARGS[k] = (name, a)
        return r

    import scnym_b200.engine as E
    ws = getattr(eng, "_gemm_ws", None)
    if ws:   # the eager single-stream pass issues the large products on THIS stream: lend it the main chain's operand-plane workspace
        _lib.call("scnb_gemm_tc_workspace", stream.cuda_stream, ws[0].data_ptr() + (-ws[0].data_ptr()) % 1024, ws[0].numel() - 1024)
    graph_flag, ms_flag = eng.cfg.use_cuda_graph, eng.cfg.multi_stream
    eng.cfg.use_cuda_graph = False
    eng.cfg.multi_stream = False     # one stream, so that the events bracket each kernel
    E.call = timed
    try:
        for i in range(n_steps + 2):
            if i == 2:
                records.clear()
            eng.step()
        torch.cuda.synchronize()
    finally:
        E.call = orig
        eng.cfg.use_cuda_graph, eng.cfg.multi_stream = graph_flag, ms_flag
    agg = {}
    for name, s, e in records:
        agg.setdefault(name, []).append(s.elapsed_time(e))
    # median per launch (a single host hiccup inside one event pair must not make a 10 us kernel look dominant)
    per_launch = {k: float(np.median(v)) for k, v in agg.items()}
    counts = {k: len(v) // n_steps for k, v in agg.items()}
    per_step = {k: per_launch[k] * len(agg[k]) / n_steps for k in agg}
    return per_step, per_launch, counts


LAST_ARGS = {}   # launch key -> (entry point, arguments of its last launch in time_kernels; borrowed device pointers of the live engine)


def time_burst(key, reps=20):
    """Average duration (ms) of `reps` back-to-back launches of an entry point with the arguments of its last launch in
    the step: CUDA events on the launching stream around the whole burst, so that the launch latency of one call hides
    behind the execution of the previous one (a single event-bracketed launch of a 40 us kernel reads ~25 % high)."""
    from scnym_b200 import _lib
    name, a = LAST_ARGS[key]
    stream = torch.cuda.current_stream()
    a = a[:-1] + (stream.cuda_stream,)
    for _ in range(3):
        _lib.call(name, *a)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record(stream)
    for _ in range(reps):
        _lib.call(name, *a)
    e.record(stream)
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def kernel_cost(eng, key):
    """ALGORITHMIC cost per launch of a launch key: {"bound": "hbm", "bytes": B} for the byte-carrying
    kernels, {"bound": "tensor", "flops": F, "issued_flops": Fi} for the GEMMs (F = 2MNK of the fp32 product the
    reference computes; Fi = what the split-precision kernel issues to the tensor pipe).  None: no table entry."""
    nnz = int(eng.b_indptr[eng.nb].item())
    G, H0, nb, P, R = eng.G, eng.widths[0], eng.nb, eng.arena.total, eng.Rmax
    Hs = sum(eng.widths)
    hbm = lambda b: {"bound": "hbm", "bytes": float(b)}
    name = key.split("[")[0]
    if name == "scnb_sgemm":
        _, a = LAST_ARGS[key]
        M, N, K = a[2], a[3], a[4]
        return {"bound": "tensor", "flops": 2.0 * M * N * K, "issued_flops": 6.0 * M * N * K, "bytes": 4.0 * (M * K + K * N + M * N)}
    if name in ("scnb_hs_fwd", "scnb_hs_bwd", "scnb_hs_dw"):
        # tile GEMMs of the fused hidden stack (3xTF32 mma.sync): 2MNK of the fp32 product, three products issued; bytes:
        # operand / result tensors once (forward: Z in, A + Z out, W; backward: dY, Z in, dZ + dY out, W; dW: dZ, A in, dW out)
        _, a = LAST_ARGS[key]
        if name == "scnb_hs_fwd":
            K_, N_, R_ = a[16], a[17], a[18]
            by = 4.0 * (R_ * K_ * 2 + R_ * N_ + K_ * N_)
        elif name == "scnb_hs_bwd":
            K_, N_, R_ = a[15], a[16], a[17]
            by = 4.0 * (R_ * K_ * 3 + R_ * N_ * 2 + K_ * N_)
        else:
            K_, N_, R_ = a[1], a[3], a[4]
            by = 4.0 * (R_ * K_ + R_ * N_ + K_ * N_)
        return {"bound": "tensor", "flops": 2.0 * R_ * K_ * N_, "issued_flops": 6.0 * R_ * K_ * N_, "bytes": by}
    if name == "scnb_hs_dan_head":
        # adversary branch in one launch (head_ops.cu): Hd = A W1^T and dA = dHd W1 as 3xTF32 tile products; bytes: Z in, A, Hd,
        # dHd, dY out, W1 once
        Rd_, H_ = eng.Rd, eng.H
        return {"bound": "tensor", "flops": 4.0 * Rd_ * H_ * H_, "issued_flops": 12.0 * Rd_ * H_ * H_, "bytes": 4.0 * (5 * Rd_ * H_ + H_ * H_)}
    if name in ("scnb_csr_linear_fwd_tc", "scnb_csr_linear_fwd_tc_splits"):
        # bytes: CSR stream + the two bf16 planes of W1T once + Z; tensor work: densified bf16x3 product
        return {"bound": "tensor", "flops": 2.0 * nnz * H0, "issued_flops": 6.0 * nb * G * H0, "bytes": 8.0 * nnz + 4.0 * G * H0 + 4.0 * nb * H0}
    dp = eng.comm is not None
    table = {
        "scnb_csr_linear_fwd": hbm(8 * nnz + 4 * G * H0 + 4 * nb * H0),          # CSR stream + W1T once + Z
        "scnb_csr_linear_bwd_dw": hbm(8 * nnz + 4 * nb * H0 + 4 * G * H0),       # CSR stream + dZ + dW1T written once
        # batch entries + dZ + read p,m,v / write p,m,v of W1T (data parallel: the local gradient only, no update)
        "scnb_csr_w1_bwd_optim": hbm(8 * nnz + 4 * nb * H0 + (4 if dp else 24) * G * H0),
        "scnb_csr_w1_bwd_optim_tc": hbm(8 * nnz + 4 * nb * H0 + (4 if dp else 24) * G * H0),
        "scnb_optim_step": hbm(28 * (P - G * H0 if getattr(eng, "fused_w1", False) else P)),   # read p,g,m,v / write p,m,v
        "scnb_csr_gather_rows2": hbm(16 * nnz),                                  # dataset rows in, batch CSR out
        "scnb_csr_gather_rows": hbm(16 * nnz),
        "scnb_row_gene_splits": hbm(4 * nnz),
        "scnb_row_gene_splits2": hbm(4 * nnz),
        "scnb_w1_planes": hbm(8 * G * H0),
        "scnb_mix_rows": hbm(4 * H0 * (nb + R)),
        "scnb_unmix_rows": hbm(4 * H0 * (R + nb)),
        "scnb_unmix_rows_planes": hbm(4 * H0 * (R + nb) + 4 * nb * H0),
        "scnb_hs_bwd_in_unmix_planes": hbm(4 * H0 * (2 * R + nb) + 4 * nb * H0),   # dY, Z in; dZ1 + planes out
        "scnb_bn_act_fwd": hbm(8 * R * (Hs / len(eng.widths))),                 # Z in, A out
        "scnb_bn_act_bwd": hbm(16 * R * (Hs / len(eng.widths))),                # dA, Z, A in, dZ out
        "scnb_bn_act_fwd_peer": hbm(8 * R * (Hs / len(eng.widths))),
        "scnb_bn_act_bwd_peer": hbm(16 * R * (Hs / len(eng.widths))),
        "scnb_classif_mixmatch_fused": hbm(4 * nb * (2 * eng.H + 3 * eng.C)),
        "scnb_hs_classif_head": hbm(4 * nb * (3 * eng.H + 3 * eng.C)),           # Z in, A + dY out, logits / dlogits
        "scnb_dan_tail_fused": hbm(4 * eng.Rd * (2 * eng.H + 3 * max(eng.D, 1))) if eng.use_dan else None,
        "scnb_teacher_head_fused": hbm(4 * eng.U * (eng.H + 2 * eng.C)),
        "scnb_colsum": hbm(4 * R * eng.H),
        # fused hidden stack (one launch): every activation / gradient tensor of the stack written once and read once,
        # the hidden weights read twice (forward, backward), their gradients written once
        "scnb_hidden_stack": hbm(4 * (6 * R * Hs + 3 * (P - G * H0))),
        "scnb_teacher_stack": hbm(4 * (eng.U * (H0 + eng.C) + (P - G * H0))),
    }
    if dp:
        W = eng.comm.world
        # the exchange kernel reads its slice of every rank's gradients and of p, m, v, writes m, v and the new p into
        # every rank's arena
        table["scnb_dp_reduce_optim_bcast"] = hbm(4 * (P // W) * (W + 3 + 2 + W))
    return table.get(name)


def roofline_block(eng, per_step, per_launch, world, ms_per_step):
    """The kernel with the largest share of the step, over ALL kernels, against the roofline that bounds it."""
    dom = max(per_step, key=lambda k: per_step[k])
    cost = kernel_cost(eng, dom)
    hbm_peak, hbm_how = peaks()
    tf_peak, tf_how = tensor_peak()
    # timed alone: a burst of back-to-back launches when it needs no other rank, else the event-bracketed launches
    burstable = world == 1 and dom in LAST_ARGS and not dom.startswith("scnb_hidden_stack")
    t_dom = time_burst(dom) if burstable else per_launch[dom]
    roof = {"kernel": dom, "ms_per_launch": t_dom, "ms_per_launch_single_event_pair": per_launch[dom],
            "timing": "CUDA events around 20 back-to-back launches of the kernel with the step's arguments" if burstable else
                      "CUDA events around each launch of an eager step" + (" (includes waiting for the other ranks)" if world > 1 else ""),
            "step_share": per_step[dom] / sum(per_step.values()),
            "kernel_ms_per_step": {k: round(v, 5) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])}}
    if cost is None:
        roof.update({"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
                     "note": "no algorithmic-cost table entry for this kernel"})
    elif cost["bound"] == "tensor":
        ach = cost["flops"] / (t_dom * 1e-3) / 1e12
        roof.update({"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": ncu_traffic(dom),
                     "peak_source": tf_how, "algorithmic_flops_per_launch": cost["flops"], "issued_flops_per_launch": cost["issued_flops"],
                     "issued_tflops": cost["issued_flops"] / (t_dom * 1e-3) / 1e12,
                     "issued_frac": cost["issued_flops"] / (t_dom * 1e-3) / 1e12 / tf_peak,
                     "hbm_frac_of_same_kernel": cost["bytes"] / (t_dom * 1e-3) / 1e9 / hbm_peak})
    else:
        ach = cost["bytes"] / (t_dom * 1e-3) / 1e9
        if ach > hbm_peak and burstable:
            # back-to-back re-runs of a kernel whose tensors (partly) fit the 126 MB L2 read them from L2: the burst is not an
            # HBM measurement.  Use the event-bracketed launches inside the eager step (other kernels' data in between).
            t_dom = per_launch[dom]
            ach = cost["bytes"] / (t_dom * 1e-3) / 1e9
            roof["ms_per_launch"] = t_dom
            roof["timing"] = ("CUDA events around each launch of an eager step (a burst of back-to-back launches re-reads this "
                              "kernel's tensors from L2 and is not an HBM measurement)")
        roof.update({"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": ncu_traffic(dom),
                     "peak_source": hbm_how, "algorithmic_bytes_per_launch": cost["bytes"]})
    # the byte-carrying kernel (dW1 + optimizer) and the whole step against the HBM bound, always reported
    for k in per_step:
        if k.startswith("scnb_csr_w1_bwd_optim"):
            c = kernel_cost(eng, k)
            t = time_burst(k) if world == 1 else per_launch[k]
            roof["byte_carrier"] = {"kernel": k, "ms_per_launch": t, "algorithmic_bytes_per_launch": c["bytes"],
                                    "achieved_gbs": c["bytes"] / (t * 1e-3) / 1e9, "frac_of_hbm": c["bytes"] / (t * 1e-3) / 1e9 / hbm_peak,
                                    "traffic": ncu_traffic(k)}
    step_bytes = 8 * int(eng.b_indptr[eng.nb].item()) + 36 * eng.arena.total
    roof["step_algorithmic_bytes"] = step_bytes
    roof["step_frac_of_hbm"] = step_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak
    return roof


def time_e2e(args, device, rank, world, feed="mapped"):
    """`e2e`: the same step through the public host-buffer entries of `scnym_b200.hostfeed`: the FULL expression matrices
    of the workload live in HOST memory, every step's batch is brought to the device inside the timed region and the
    step's loss scalars are read on the host after every step.
      feed = "mapped" (`HostMappedTrainPipeline`): the matrices are page-locked in place and mapped; the step's own row
             gather reads the sampled rows over PCIe while the previous step runs; the loss lands in mapped host memory.
      feed = "staged" (`HostTrainPipeline`): producer threads gather each batch into pinned blobs, one H2D copy per step,
             device-side stash, same read-back; the three stages are pipelined across steps."""
    from scnym_b200.hostfeed import HostMappedTrainPipeline, HostTrainPipeline
    c = CFG
    model, dann, cfg, lab, y, unl = build_train(device, rank, seed_off=7)
    ld, ud = given_domains(device, lab.n_rows, unl.n_rows, rank)
    host = lambda t: t.cpu().numpy()
    lab_t = (host(lab.indptr), host(lab.indices), host(lab.data), host(y)) + ((host(ld),) if ld is not None else ())
    unl_t = (host(unl.indptr), host(unl.indices), host(unl.data)) + ((host(ud),) if ud is not None else ())
    n_host = lab.n_rows + unl.n_rows
    del lab, unl
    torch.cuda.empty_cache()
    if feed == "mapped":
        pipe = HostMappedTrainPipeline(model, cfg, lab=lab_t, unl=unl_t, n_genes=c["G"], batch_size=c["L"], unlabeled_batch_size=c["U"],
                                       dann=dann, device=device, seed=5 + rank, comm=make_comm(world))
    else:
        # host threads for batch assembly: the box's cores are shared by the ranks
        cores = max((os.cpu_count() or 8) // world, 1)
        n_thr = max(min(4, cores), 1)
        n_prod = max(min(int(os.environ.get("SCNB_E2E_PRODUCERS", "3")), cores // n_thr), 1)
        pipe = HostTrainPipeline(model, cfg, lab=lab_t, unl=unl_t, n_genes=c["G"], batch_size=c["L"], unlabeled_batch_size=c["U"],
                                 dann=dann, device=device, n_threads=n_thr, seed=5 + rank, comm=make_comm(world), n_producers=n_prod)
    pipe.engine.set_weights(1.0, 0.1)
    steps, warm = max(args.steps, 50), max(args.warmup, 5)
    for l in pipe.run(warm):
        pass
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for l in pipe.run(steps):
        pass
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert np.isfinite(l["loss"]), l
    out = (dt / steps * 1e3, pipe.h2d_bytes_last, pipe.d2h_bytes_last, n_host)
    pipe.close()
    return out


def time_predict(device, rank, passes=16):
    """Batched predict (BASELINE config 4 shape): 2^18-row tile x 16 passes = 4.19M cells, device-resident; then the same
    tile END TO END from host memory (pinned CSR -> device -> predictions / probabilities / embedding in host arrays)."""
    from scnym_b200.model import CellTypeCLF
    from scnym_b200.predict import EvalRunner, HostAtlasStream
    from scnym_b200.synth import synth_device_csr
    c = CONFIGS[2][0]
    tile, _ = synth_device_csr(1 << 18, c["G"], c["density"], c["C"], seed=2 + 100 * rank, device=device)
    torch.manual_seed(0)
    model = CellTypeCLF(n_genes=c["G"], n_cell_types=c["C"], n_hidden=c["H"], n_hidden_init=c["H0"], n_layers=c["n_layers"]).to(device).eval()
    run = EvalRunner([model])
    N = tile.n_rows
    pred = torch.empty(N, dtype=torch.int64, device=device)
    prob = torch.empty((N, c["C"]), device=device)
    emb = torch.empty((N, c["H"]), device=device)

    run.refresh_planes()

    def one_pass():
        for s in range(0, N, run.chunk):
            n = min(run.chunk, N - s)
            run.run_chunk(tile, s, n, pred[s:], prob[s:], None, emb[s:])

    one_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(passes):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cells = passes * N
    nnz = tile.nnz
    bytes_per_cell = 8.0 * nnz / N + 4 * (c["C"] + c["H"] + 2)
    # per-kernel time of one full chunk (CUDA events around every C-ABI call, eager)
    import scnym_b200.predict as Pm
    from scnym_b200 import _lib
    stream, recs, orig = torch.cuda.current_stream(), [], _lib.call

    def timed(name, *a):
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record(stream)
        r = orig(name, *a)
        e_.record(stream)
        recs.append((name, s_, e_))
        return r

    Pm.call = timed
    try:
        n = min(run.chunk, N)
        for i in range(3):
            if i == 1:
                recs.clear()
            run.run_chunk(tile, 0, n, pred, prob, None, emb)
        torch.cuda.synchronize()
    finally:
        Pm.call = orig
    per = {}
    for name, s_, e_ in recs:
        per[name] = per.get(name, 0.0) + s_.elapsed_time(e_) / 2
    extra = {"chunk_rows": n, "kernel_ms_per_chunk": {k: round(v, 4) for k, v in sorted(per.items(), key=lambda kv: -kv[1])}}
    pk = load_peaks()
    t_first = per.get("scnb_csr_linear_fwd_tc_i64") or per.get("scnb_csr_linear_fwd_gt_i64") or per.get("scnb_csr_linear_fwd_i64")
    if t_first:
        nnz_chunk = float((tile.indptr[n] - tile.indptr[0]).item())
        alg = 2.0 * nnz_chunk * c["H0"]
        peak_tf = float(pk.get("bf16_tflops", 1590.0))
        fl = {"kernel": max(per, key=lambda k: per[k]), "ms": t_first, "algorithmic_tflops": alg / (t_first * 1e-3) / 1e12,
              "algorithmic_flops_per_cell": alg / n}
        if "scnb_csr_linear_fwd_tc_i64" in per:
            issued = 2.0 * 3.0 * n * c["G"] * c["H0"]
            fl.update({"issued_tflops": issued / (t_first * 1e-3) / 1e12, "peak_tflops": peak_tf,
                       "issued_frac_of_bf16_peak": issued / (t_first * 1e-3) / 1e12 / peak_tf,
                       "issued_over_algorithmic": issued / alg,
                       "note": "densify-then-tcgen05 bf16x3: ISSUED flops (zeros multiplied, three products per MAC); the algorithm needs "
                               "2*nnz*H0 per cell"})
        extra["first_layer"] = fl
    # gene re-indexing in front of predict (scnym/utils.py:155-233): one chunk of the atlas whose genes
    # are a random permutation of the model's with 10 % absent, re-indexed in HBM by scnb_csr_remap_count + _fill
    try:
        from scnym_b200.dataprep import remap_columns
        g = torch.Generator().manual_seed(7)
        cm = torch.randperm(c["G"], generator=g).to(torch.int32)
        cm[torch.rand(c["G"], generator=g) < 0.1] = -1
        cm = cm.to(device)
        nnz_in = int((tile.indptr[n] - tile.indptr[0]).item())
        ws = torch.empty(_lib.lib().scnb_csr_remap_workspace_len(n), dtype=torch.int64, device=device)
        out = (torch.empty(n + 1, dtype=torch.int64, device=device), torch.empty(nnz_in, dtype=torch.int32, device=device),
               torch.empty(nnz_in, dtype=torch.float32, device=device), ws)
        part = remap_columns(tile, cm, c["G"], 0, n, out=out)
        torch.cuda.synchronize()
        nnz_out = int(part.indptr[n].item())
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(5):
            remap_columns(tile, cm, c["G"], 0, n, out=out)
        r1.record()
        torch.cuda.synchronize()
        t_re = r0.elapsed_time(r1) / 5
        alg = 12.0 * nnz_in + 8.0 * nnz_out
        hbm = float(pk.get("hbm_gbs", 6500.0))
        extra["gene_remap"] = {"kernels": "scnb_csr_remap_count + scnb_csr_remap_fill", "rows": n, "nnz_in": nnz_in, "nnz_out": nnz_out,
                               "ms_per_chunk": round(t_re, 4), "algorithmic_bytes": alg, "achieved_gbs": alg / (t_re * 1e-3) / 1e9,
                               "frac_of_hbm": alg / (t_re * 1e-3) / 1e9 / hbm, "cells_per_s": n / (t_re * 1e-3)}
        del out, ws, part
    except Exception as ex:  # a secondary figure must never take the bench line down
        extra["gene_remap"] = {"error": repr(ex)[:200]}
    # end to end: the tile in pinned HOST memory -> pred / prob / embedding in pinned host arrays
    try:
        ip_h = tile.indptr.cpu().numpy()
        ix_h = tile.indices.cpu().pin_memory()
        dv_h = tile.data.cpu().pin_memory()
        del tile
        torch.cuda.empty_cache()
        cap = int(max(ip_h[min(s + run.chunk, N)] - ip_h[s] for s in range(0, N, run.chunk)))
        hs = HostAtlasStream(run, cap, want_prob=True, want_score=False, want_embed=True)
        res = hs.run(ip_h, ix_h.numpy(), dv_h.numpy())           # warm-up (allocates the pinned outputs)
        t0 = time.perf_counter()
        res = hs.run(ip_h, ix_h.numpy(), dv_h.numpy(), out=res)
        dt = time.perf_counter() - t0
        agree = bool(torch.equal(res["pred"][:4096], pred[:4096].cpu()))
        extra["e2e"] = {"value": N / dt, "unit": "cells/s", "cells": N, "ms": dt * 1e3, "h2d_bytes": hs.h2d_bytes, "d2h_bytes": hs.d2h_bytes,
                        "h2d_gbs": hs.h2d_bytes / dt / 1e9, "d2h_gbs": hs.d2h_bytes / dt / 1e9, "matches_device_resident": agree,
                        "path": "scnym_b200.predict.HostAtlasStream (what Predicter.predict runs for a scipy CSR): pinned host CSR -> "
                                "chunked H2D -> kernels -> D2H of pred/prob/embed, three streams; bound by the H2D link (8 B per stored entry)"}
    except Exception as ex:
        extra["e2e"] = {"error": repr(ex)[:300]}
    return cells / (ms * 1e-3), ms, cells, bytes_per_cell, extra


# ----------------------------------------------------------------------------------------------------------------------
# Reference baselines (CPU and eager CUDA): the reference's own classes from oracle/_ref when installed, else the oracle port
# ----------------------------------------------------------------------------------------------------------------------
_POOL = {}


def _synth_pool(c=None, n_pool=4096):
    """A pool of synthetic cells (oracle generator, SURVEY §8d recipe) tiled to any length: the reference arm needs
    (W + K) x 256 rows per matrix; throughput per step does not depend on the rows being distinct."""
    from oracle import scnym_oracle as O
    c = c or CFG

    def synth(n, seed):
        key = (c["G"], c["density"], c["C"], seed, n_pool)
        if key not in _POOL or (len(_POOL[key][0]) - 1) < min(n_pool, n):
            _POOL[key] = O.synth_csr(min(n_pool, max(n, 1)), c["G"], c["density"], c["C"], seed=seed)
        ip, ii, dd, y = _POOL[key]
        m = len(ip) - 1
        if n <= m:
            e = ip[n]
            return ip[:n + 1].copy(), ii[:e].copy(), dd[:e].copy(), y[:n].copy()
        reps = (n + m - 1) // m
        lens = np.tile(np.diff(ip), reps)[:n]
        ip2 = np.zeros(n + 1, dtype=np.int64)
        ip2[1:] = np.cumsum(lens)
        return ip2, np.tile(ii, reps)[:ip2[-1]], np.tile(dd, reps)[:ip2[-1]], np.tile(y, reps)[:n]
    return synth


def reference_train(n_steps, warmup, device="cpu", dense=False):
    """(s_per_step, steps, kind) of the reference step loop on `device` ("cpu": all host threads)."""
    from oracle import ref_runner as RR
    c = dict(CFG)
    if RR.available():
        with contextlib.redirect_stdout(sys.stderr):
            t, steps = RR.time_train(c, _synth_pool(), n_steps, warmup, device=device, dense_batches=dense,
                                     threads=(os.cpu_count() or 1) if device == "cpu" else None,
                                     given_domains=bool(c.get("given_domains")))
        return t, steps, "reference"
    if device != "cpu":
        return None, 0, "unavailable"
    t_all, t_comp, done = cpu_port_step_time(n_steps, warmup)
    return (t_comp if dense else t_all), done, "port"


def cpu_port_step_time(n_steps, warmup, seed=0):
    """The oracle PORT of the reference step on host cores (used only when oracle/_ref is absent): per-step time incl.
    per-row densify (what SingleCellDS/DataLoader do) and compute-only."""
    from oracle import scnym_oracle as O
    c = CFG
    torch.set_num_threads(os.cpu_count() or 1)
    n_cells = 2048
    lp, li, ld, ly = O.synth_csr(n_cells, c["G"], c["density"], c["C"], seed=seed)
    up, ui, ud, _ = O.synth_csr(n_cells, c["G"], c["density"], c["C"], seed=seed + 1)
    from scnym_b200.model import CellTypeCLF, DANN
    torch.manual_seed(0)
    m = CellTypeCLF(n_genes=c["G"], n_cell_types=c["C"], n_hidden=c["H"], n_hidden_init=c["H0"], n_layers=c["n_layers"])
    dn = DANN(m, n_domains=c["D"])
    st = {k: v.detach().contiguous().clone() for k, v in m.state_dict().items()}
    dst = {"domain_clf." + k: v.detach().clone() for k, v in dn.domain_clf.state_dict().items()}
    om = O.OracleModel(st, n_layers=c["n_layers"], dan_state=dst)
    opt = O.OracleOptimizer(om.parameters(), O.OptimConfig(name="adam", lr=1e-3, weight_decay=1e-4))
    cfg = O.StepConfig(n_augmentations=1, T=0.5, augment_pseudolabels=False, unsup_weight=1.0, use_dan=True, dan_weight=0.1)
    rng = np.random.default_rng(seed)
    L, U, G = c["L"], c["U"], c["G"]
    widths = [c["H0"]] + [c["H"]] * (1 + max(c["n_layers"] - 1, 0))
    t_all, t_comp, done = 0.0, 0.0, 0
    for s in range(warmup + n_steps):
        a = time.perf_counter()
        lrow, urow = rng.choice(n_cells, L, replace=False), rng.choice(n_cells, U, replace=False)
        Lx = O.csr_to_dense(lp, li, ld, G, lrow)
        Ux = O.csr_to_dense(up, ui, ud, G, urow)
        Ly = torch.nn.functional.one_hot(torch.from_numpy(ly[lrow]), num_classes=c["C"]).float()
        rnd = O.StepRandomness(in_keep_L=torch.rand(L, G) < 0.9, perm_keys=torch.rand(L + U),
                               gamma_raw=torch.distributions.Beta(0.3, 0.3).sample((L + U,)),
                               drop_keep_U=[torch.rand(U, w) < 0.5 for w in widths],
                               drop_keep_L=[torch.rand(L, w) < 0.5 for w in widths],
                               drop_keep_D=[torch.rand(L + U, w) < 0.5 for w in widths])
        b = time.perf_counter()
        O.train_step(om, opt, cfg, Lx, Ly, Ux, rnd)
        e = time.perf_counter()
        if s >= warmup:
            t_all += e - a
            t_comp += e - b
            done += 1
    return t_all / done, t_comp / done, done


def reference_predict(n_cells, device="cpu"):
    """Seconds for the reference's `Predicter.predict` over n_cells synthetic cells (config-4 shape), or None."""
    from oracle import ref_runner as RR
    if not RR.available():
        return None
    c = dict(CONFIGS[2][0])
    with contextlib.redirect_stdout(sys.stderr), contextlib.redirect_stderr(open(os.devnull, "w")):
        return RR.time_predict(c, _synth_pool(c, 8192), n_cells, device=device, threads=(os.cpu_count() or 1) if device == "cpu" else None)


def make_comm(world):
    """Data-parallel communicator (NCCL) when launched under torchrun; None for a single process."""
    if world <= 1:
        return None
    from scnym_b200.parallel import TorchDistComm
    return TorchDistComm()


def barrier(world):
    if world > 1:
        torch.distributed.barrier()


def dp_parity(device, rank, world):
    """Outside any timed region: one INJECTED step of the production data-parallel path on all ranks, rank 0 compares the
    job against the CPU oracle on the concatenated global batch (tests/cuda_harness.py::dp_compare)."""
    from tests import cuda_harness as CH
    try:
        rep = CH.dp_compare(device, rank, world, make_comm(world))
        return rep
    except AssertionError as e:
        return {"parity": "FAILED", "error": str(e)[:300]}


def _claim_stdout():
    """Keep the real stdout for the ONE JSON line: everything else that writes to file descriptor 1 (NCCL's version banner,
    the reference's prints, tqdm) goes to stderr.  Returns a file object on the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w")
    return real


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-predict", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--config", type=int, default=None, choices=sorted(CONFIGS),
                    help="BASELINE.json config to time (default: 2 on one GPU, 3 -- the named data-parallel workload -- on several)")
    args = ap.parse_args()
    global CFG, WORKLOAD
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config is None:
        args.config = 2 if max(args.gpus, world) == 1 else 3
    CFG, WORKLOAD = CONFIGS[args.config]
    secondary = args.config in (5, 6)
    args.warmup = max(args.warmup, 3)
    cells_per_step = CFG["L"] + CFG["U"]

    if args.impl == "reference":
        if rank != 0:
            return
        t, done, kind = reference_train(args.steps, args.warmup, device="cpu", dense=False)
        v = cells_per_step / t
        cores = os.cpu_count() or 1
        what = ("the unmodified reference (oracle/_ref): SingleCellDS + DataLoader + MixMatchLoss + DANLoss + torch.optim.Adam through "
                "SemiSupervisedTrainer.train_epoch()") if kind == "reference" else "the CPU oracle port (dense batches incl. per-row densify)"
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(max(args.gpus, 1)),
                "cpu_baseline": {"value": v, "unit": "cells/s", "cores": cores, "kind": kind,
                                 "sample": f"{done} steps of {what} on {cores} host threads, one process "
                                           f"(the reference has no data-parallel mode: the same figure at every N)"},
                "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=out, flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (scnym_b200 has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    import __graft_entry__
    if rank == 0:
        with contextlib.redirect_stdout(sys.stderr):
            __graft_entry__.build()
    barrier(world)

    single_same = None
    if world > 1:
        # the SAME workload on one GPU without any exchange (every rank runs its own copy; rank 0 reports), so that the scaling
        # of this configuration can be read off one run: the driver's N = 1 line is config 2, a different workload
        _e, ms1, _c, _l, _s = time_device_resident(args, device, rank, 1)
        single_same = {"value": cells_per_step / (ms1 / args.steps * 1e-3), "unit": "cells/s", "ms_per_step": ms1 / args.steps,
                       "note": "this config on ONE GPU (no data parallelism), measured on rank 0 in the same run"}
        del _e
        torch.cuda.empty_cache()
        barrier(world)
    eng, ms, clocks, losses, sampler = time_device_resident(args, device, rank, world)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = cells_per_step * world / (ms_per_step * 1e-3)

    # per-kernel timing + roofline of the dominant kernel (every rank steps; rank 0 reports)
    per_step, per_launch, counts = time_kernels(eng)
    launches = int(eng.kernels_per_step or sum(counts.values()))
    roof = roofline_block(eng, per_step, per_launch, world, ms_per_step) if rank == 0 else None
    del eng
    torch.cuda.empty_cache()

    e2e = None
    if not args.no_e2e:
        feeds = {}
        for feed in ("mapped", "staged"):
            e2e_ms, h2d, d2h, n_host = time_e2e(args, device, rank, world, feed=feed)
            t = torch.tensor([e2e_ms], device=device, dtype=torch.float64)
            if world > 1:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            feeds[feed] = {"value": cells_per_step * world / (float(t.item()) * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": h2d,
                           "d2h_bytes_per_step": d2h, "host_cells_per_rank": n_host}
            torch.cuda.empty_cache()
        # both public host-buffer entries are timed; the headline is the one a user would pick (the faster)
        best = max(feeds, key=lambda k: feeds[k]["value"])
        e2e = dict(feeds[best], feed=("hostfeed.HostMappedTrainPipeline" if best == "mapped" else "hostfeed.HostTrainPipeline"),
                   feeds={k: v["value"] for k, v in feeds.items()},
                   note="mapped: the step's row gather reads the sampled rows of the page-locked host matrices over PCIe and the "
                        "loss is stored into mapped host memory; staged: producer threads -> pinned blob -> one H2D copy per step")

    dp = None
    if world > 1:
        dp = dp_parity(device, rank, world)

    predict = None
    if not args.no_predict and not secondary:
        # cells are sharded over the ranks (contiguous row ranges, no communication): every rank predicts its own
        # 4.19 M / world cells; the job's throughput is all cells over the slowest rank's time
        barrier(world)
        pv, pms, pcells, bpc, pextra = time_predict(device, rank, passes=max(16 // world, 1))
        t = torch.tensor([pms], device=device, dtype=torch.float64)
        e2e_p = torch.tensor([pextra.get("e2e", {}).get("ms", 0.0)], device=device, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            torch.distributed.all_reduce(e2e_p, op=torch.distributed.ReduceOp.MAX)
        pms = float(t.item())
        pcells *= world
        pv = pcells / (pms * 1e-3)
        if rank == 0:
            peak, _ = peaks()
            per_gpu = pv / world
            predict = {"value": pv, "unit": "cells/s", "cells": pcells, "ms": pms, "n_gpus": world,
                       "workload": "config4 shape: 2^18-row tile x16 passes (cells sharded over the ranks), G=20000, 5% CSR, "
                                   "outputs argmax+prob+embed", "bytes_per_cell": bpc,
                       "ceilings_per_gpu": {"hbm_cells_per_s": peak * 1e9 / bpc, "frac_of_hbm": per_gpu * bpc / 1e9 / peak,
                                            "fp32_fma_cells_per_s": 74e12 / (2.0 * 0.05 * 20000 * 256),
                                            "frac_of_fp32_fma": per_gpu / (74e12 / (2.0 * 0.05 * 20000 * 256)),
                                            "note": "SURVEY 8d: the sparse gather-FMA ceiling (74 TFLOP/s fp32 SIMT over 2*nnz*H0 flops per "
                                                    "cell) is the binding one for a SIMT first layer; the HBM ceiling for a streaming one"}}
            predict.update(pextra)
            if world > 1 and "e2e" in predict and "ms" in predict["e2e"]:
                predict["e2e"]["value"] = (1 << 18) * world / (float(e2e_p.item()) * 1e-3)
                predict["e2e"]["cells"] = (1 << 18) * world

    cpu = gpu_eager = None
    if rank == 0 and world == 1 and not secondary:
        cores = os.cpu_count() or 1
        if not args.no_cpu_baseline:
            t_e2e, done, kind = reference_train(12, 2, device="cpu", dense=False)
            t_comp, _, _ = reference_train(8, 2, device="cpu", dense=True)
            cpu = {"value": cells_per_step / t_e2e, "unit": "cells/s", "cores": cores, "kind": kind,
                   "sample": f"{done} steps of the same workload through "
                             + ("the unmodified reference's SemiSupervisedTrainer.train_epoch() (oracle/_ref) with its own SingleCellDS/DataLoader"
                                if kind == "reference" else "the CPU oracle port (dense batches incl. per-row densify)")
                             + f" on {cores} host threads", "compute_only": cells_per_step / t_comp}
            tp = reference_predict(65536, device="cpu")
            if tp is not None:
                cpu["predict"] = {"value": 65536 / tp, "unit": "cells/s", "cells": 65536, "kind": "reference",
                                  "sample": "Predicter.predict (scnym/predict.py:107-216) on 65 536 cells, batch 1024, CSR input, "
                                            f"{cores} host threads; the embedding pass of scnym_api (a second full forward) is not included"}
        if not args.no_gpu_eager:
            try:
                t_g, done_g, kind_g = reference_train(40, 5, device="cuda", dense=False)
                t_gc, _, _ = reference_train(40, 5, device="cuda", dense=True)
                if t_g is not None:
                    gpu_eager = {"value": cells_per_step / t_g, "unit": "cells/s", "kind": kind_g, "compute_only": cells_per_step / t_gc,
                                 "ms_per_step": t_g * 1e3, "ms_per_step_compute_only": t_gc * 1e3,
                                 "sample": f"{done_g} steps of the unmodified reference loop in eager PyTorch on this B200 (model and batches "
                                           ".cuda() as scnym/trainer.py:589-599 does); `compute_only` = batches collated beforehand"}
                    tp = reference_predict(65536, device="cuda")
                    if tp is not None:
                        gpu_eager["predict"] = {"value": 65536 / tp, "unit": "cells/s", "cells": 65536,
                                                "sample": "reference Predicter.predict on 65 536 cells with the model on this B200 "
                                                          "(host densify per cell, dense H2D per batch)"}
            except Exception as ex:
                gpu_eager = {"error": repr(ex)[:300]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
                "data": "synthetic", "config": config_dict(world), "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches * args.steps, "launches_per_step": launches, "roofline": roof, "cpu_baseline": cpu,
                "gpu_eager_baseline": gpu_eager, "predict": predict, "sampler_draw": sampler, "dp_parity": dp,
                "single_gpu_same_config": single_same,
                "losses_last_step": {k: (round(float(v), 6) if not isinstance(v, bool) else v) for k, v in losses.items()}}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        # Kernels of the step wait for peers on the device; tearing the process group down underneath a replayed graph can
        # block forever, so synchronise, flush and leave without destroy_process_group().
        torch.cuda.synchronize()
        barrier(world)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
