"""Debug: first buffer (in dataflow order) in which the tc-first-layer engine departs from the spmm engine."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import cuda_harness as CH
from tests import golden_util as GU
kw = dict(G=1024, L=256, U=256, C=6, H0=256, H=256, n_layers=2, opt_name="adam", n_steps=1, seed=5)
d3 = CH.oracle_case(**kw)
def run(fl, wu):
    eng = CH.build_engine(d3, first_layer=fl, w1_update=wu, multi_stream=False)
    CH.inject_step(eng, d3, 0)
    eng.step()
    torch.cuda.synchronize()
    snap = {}
    names = ["b_indptr", "b_val", "Z1", "srcA", "srcB", "gam", "counts", "seg_off", "Ycat", "conf", "logits", "dlogits", "Hd", "dlog", "ddlog", "dHd", "inv3", "coef3", "dZ1"]
    for k in names:
        v = getattr(eng, k, None)
        if v is not None:
            snap[k] = v.detach().float().cpu().clone()
    for a in range(eng.n_app):
        snap[f"Zs{a}"] = eng.Zs[a].detach().cpu().clone(); snap[f"As{a}"] = eng.As[a].detach().cpu().clone()
        snap[f"stats{a}"] = eng.stats[a].detach().cpu().clone(); snap[f"dZs{a}"] = eng.dZs[a].detach().cpu().clone()
    got = CH.collect(eng)
    errs = {n: GU.rel_err(g, d3["step0/grad/" + n]) for n, g in got["grads"].items()}
    return snap, errs, got
A, ea, ga = run("spmm", "simt")
B, eb, gb = run("tc", "simt")
print("spmm engine worst grad err", max(ea.values()), "tc engine", max(eb.values()))
print("tc engine loss errs", {k: GU.rel_err(gb[k], d3["step0/" + k]) for k in ("sup_loss", "unsup_loss", "dan_loss")}, "logits", GU.rel_err(gb["labeled_logits"], d3["step0/labeled_logits"]))
order = ["b_indptr", "b_val", "Z1", "srcA", "srcB", "gam", "counts", "seg_off", "Zs0", "stats0", "As0", "Zs1", "stats1", "As1", "Zs2", "stats2", "As2", "Ycat", "conf", "logits", "dlogits", "Hd", "dlog", "ddlog", "dHd", "dZs2", "dZs1", "dZs0", "inv3", "coef3", "dZ1"]
for k in order:
    if k in A and k in B and A[k].shape == B[k].shape:
        dif = (A[k] - B[k]).abs()
        den = float(A[k].abs().max()) or 1.0
        rows = torch.nonzero(dif.reshape(dif.shape[0], -1).amax(1) > 1e-4 * den).flatten() if dif.dim() >= 2 else torch.nonzero(dif > 1e-4 * den).flatten()
        print(f"{k:8s} rel diff {float(dif.max()) / den:.2e}  rows off {len(rows)} {rows[:6].tolist()}", flush=True)
print({n: f"{e:.1e}" for n, e in eb.items()})
