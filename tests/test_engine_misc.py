"""Engine behaviours around the step: operand planes after external weight writes, optimizer state
ownership, NaN rows in the argmax kernels."""
import os

import numpy as np
import pytest
import torch

from tests import golden_util as GU

pytestmark = pytest.mark.gpu


def _case():
    from tests import cuda_harness as CH
    return CH.oracle_case(G=1024, L=256, U=256, C=6, H0=256, H=128, n_layers=2, opt_name="sgd", lr=0.05, n_steps=1, seed=41)


def test_operand_planes_follow_load_state_dict_and_inplace_writes():
    """Tensor-core path (bf16 operand planes of W1T emitted by the optimizer epilogue): after `load_state_dict`, after an
    in-place torch write to the Parameter, and after `invalidate_planes()` following a raw `.data` write, the next step's
    first layer must see the NEW weights -- compared with the exact-fp32 path that reads W1T directly."""
    from tests import cuda_harness as CH
    d = _case()
    eng_tc = CH.build_engine(d, first_layer="tc", w1_update="tc")
    eng_fp = CH.build_engine(d, first_layer="spmm", w1_update="simt")
    assert eng_tc.w1_emits_planes and not eng_fp.tc_first
    for eng in (eng_tc, eng_fp):
        CH.inject_step(eng, d, 0)
        eng.step()
    torch.cuda.synchronize()
    g = torch.Generator().manual_seed(3)

    def new_state(model):
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        sd["embed.0.weight"] = sd["embed.0.weight"] + 0.05 * torch.randn(sd["embed.0.weight"].shape, generator=g)
        return sd

    def check(tag):
        for eng in (eng_tc, eng_fp):
            CH.inject_step(eng, d, 0)
            eng.step()
        torch.cuda.synchronize()
        e = GU.rel_err(eng_tc.Z1.cpu(), eng_fp.Z1.cpu())
        assert e < 1e-4, f"{tag}: first-layer output of the tensor-core path is stale ({e:.3e})"
        # bring both engines back to identical parameters for the next scenario
        eng_tc.model.load_state_dict({k: v.detach().cpu() for k, v in eng_fp.model.state_dict().items()})

    sd = new_state(eng_fp.model)
    eng_tc.model.load_state_dict(sd); eng_fp.model.load_state_dict(sd)
    check("load_state_dict")
    delta = 0.03 * torch.randn(eng_fp.model.embed[0].weight.shape, generator=g).cuda()
    with torch.no_grad():
        eng_tc.model.embed[0].weight.add_(delta); eng_fp.model.embed[0].weight.add_(delta)
    check("in-place write to the Parameter")
    eng_tc.model.embed[0].weight.data.mul_(1.01); eng_fp.model.embed[0].weight.data.mul_(1.01)
    eng_tc.invalidate_planes()
    check("raw .data write + invalidate_planes()")
    eng_tc.arena.check_aliasing()


def test_optimizer_state_is_owned_by_the_engine_and_round_trips(tmp_path):
    from scnym_b200 import optim as OPT
    from tests import cuda_harness as CH
    d = _case()
    eng = CH.build_engine(d)
    opt = OPT.Adam(eng.model.parameters(), lr=1e-3)
    opt.attach_engine(eng)
    with pytest.raises(RuntimeError):
        opt.step()
    CH.inject_step(eng, d, 0)
    eng.step()
    sd = opt.state_dict()
    assert "engine" in sd and float(sd["engine"]["state1"].abs().max()) > 0 and int(sd["engine"]["state4"][0]) == 1
    eng2 = CH.build_engine(d)
    opt2 = OPT.Adam(eng2.model.parameters(), lr=1e-3)
    opt2.load_state_dict(sd)            # before the engine exists: adopted on attach
    opt2.attach_engine(eng2)
    assert torch.equal(eng2.arena.state1.cpu(), sd["engine"]["state1"]) and int(eng2.state[0]) == 1


def test_argmax_kernels_return_a_valid_class_for_nan_rows():
    from scnym_b200._lib import call, ptr
    C, n = 5, 4
    z = torch.randn(n, C, device="cuda")
    z[1] = float("nan")
    z[2] = float("-inf")
    pred = torch.full((n,), -7, dtype=torch.int64, device="cuda")
    prob = torch.empty(n, C, device="cuda")
    call("scnb_predict_head", ptr(z), 1, n, C, ptr(pred), ptr(prob), None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    p = pred.cpu().numpy()
    assert np.all((p >= 0) & (p < C)), p
    assert p[0] == int(z[0].argmax()) and p[1] == 0 and p[2] == 0
    # pseudolabels with T = 0 (one-hot of the argmax): a NaN row still gets a one-hot, not an all-zero row
    U = n
    q = torch.zeros(U, C, device="cuda")
    conf = torch.zeros(U, device="cuda")
    flag = torch.zeros(U, dtype=torch.uint8, device="cuda")
    scratch = torch.zeros(U, C, device="cuda")
    call("scnb_pseudolabel", ptr(z), 1, U, C, 0.0, 1, 0.0, ptr(q), ptr(conf), ptr(flag), ptr(scratch), None, 0,
         torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert float(q[1].sum()) == 1.0 and float(q[1, 0]) == 1.0


def test_prefetched_prologue_assembles_the_same_batches(monkeypatch):
    """Replayed steps on device-resident sampler tables assemble the batch of step t+1 while step t runs (second set of
    batch buffers, one graph per parity).  Against the same engine with SCNB_PREFETCH=0 (prologue at the head of the
    step): the sampled rows, batch CSR (input dropout included), MixUp plan and gene splits of EVERY step are bit-identical
    -- across a change of tables in the middle (a new epoch invalidates the batch assembled ahead of time) -- and the
    losses of the first steps agree within the fp32 budget."""
    import bench
    from scnym_b200.engine import EngineConfig, TrainEngine
    from scnym_b200.model import CellTypeCLF, DANN
    from scnym_b200.synth import synth_device_csr
    c = dict(bench.CONFIGS[2][0])
    dev = torch.device("cuda", 0)
    lab, y = synth_device_csr(2048, c["G"], c["density"], c["C"], seed=0, device=dev)
    unl, _ = synth_device_csr(2048, c["G"], c["density"], c["C"], seed=1, device=dev)
    fields = ("l_rows", "u_rows", "lab_ids", "base_dom", "b_indptr", "srcA", "srcB", "gam", "w1_splits")
    runs = []
    for pf, gather in (("1", "cache"), ("0", "cache"), ("1", "nocache")):   # (+ the cache-free form of the gather: same bits)
        monkeypatch.setenv("SCNB_PREFETCH", pf)
        monkeypatch.setenv("SCNB_PREFETCH_GATHER", gather)
        torch.manual_seed(0)
        model = CellTypeCLF(n_genes=c["G"], n_cell_types=c["C"], n_hidden=c["H"], n_hidden_init=c["H0"], n_layers=c["n_layers"])
        dann = DANN(model, n_domains=c["D"])
        cfg = EngineConfig(n_augmentations=1, T=0.5, augment_pseudolabels=False, pseudolabel_min_confidence=0.0, mixup_alpha=0.3,
                           use_dan=True, optimizer="sgd", lr=0.01, seed=9, use_cuda_graph=True)
        eng = TrainEngine(model.to(dev), cfg, lab, y.to(torch.int32), c["L"], unlabeled=unl, unlabeled_batch_size=c["U"], dann=dann,
                          mode="mixmatch")
        assert eng.prefetch_ok == (pf == "1")
        eng.set_weights(1.0, 0.1)
        gen = torch.Generator(device=dev)
        gen.manual_seed(5)
        rec = []
        for epoch in range(2):
            eng.sample_epoch_tables(3, generator=gen)
            for _ in range(3 if epoch == 0 else 4):     # the second epoch wraps around its table once
                eng.step()
                torch.cuda.synchronize()
                nnz = int(eng.b_indptr[-1])
                snap = {k: getattr(eng, k).detach().cpu().clone() for k in fields}
                snap["b_idx"] = eng.b_idx[:nnz].cpu().clone()
                snap["b_val"] = eng.b_val[:nnz].cpu().clone()
                snap["losses"] = eng.losses()
                rec.append(snap)
        assert bool(eng._pf_graphs) == (pf == "1")
        runs.append(rec)
    a, b, a2 = runs
    assert len(a) == len(b) == len(a2) == 7
    same = lambda u, v: torch.equal(u, v) or (u.dtype.is_floating_point and torch.equal(torch.nan_to_num(u, nan=-1.0), torch.nan_to_num(v, nan=-1.0)))
    for s, (x, z, x2) in enumerate(zip(a, b, a2)):
        for k in fields + ("b_idx", "b_val"):
            assert same(x[k], z[k]), f"step {s}: {k} differs between the prefetched and the in-step prologue"
            assert same(x2[k], z[k]), f"step {s}: {k} differs between the cache-free and the cached gather"
    for s in range(2):
        for k in ("sup_loss", "unsup_loss", "dan_loss"):
            assert abs(a[s]["losses"][k] - b[s]["losses"][k]) <= 1e-4 * max(1.0, abs(b[s]["losses"][k])), (s, k)


def test_device_sampler_tables_follow_dataloader_semantics():
    """`sample_epoch_tables` (the device-side replacement of DataLoader(shuffle=True, drop_last=True) + `repeater`,
    scnym/main.py:48-68,317-323): inside one pass over a dataset no row repeats, a batch never straddles two passes, every
    pass is a fresh permutation, the unlabeled loader wraps around as often as the epoch needs, MixUp keys are uniform in
    [0, 1) and the Beta(alpha, alpha) draws lie in [0, 1] with the right mean and the U-shape of alpha < 1."""
    from scnym_b200.engine import EngineConfig, TrainEngine
    from scnym_b200.model import CellTypeCLF, DANN
    from scnym_b200.synth import synth_device_csr
    dev = torch.device("cuda", 0)
    n_lab, n_unl, L, U, G = 1000, 700, 64, 96, 512
    lab, y = synth_device_csr(n_lab, G, 0.05, 4, seed=0, device=dev)
    unl, _ = synth_device_csr(n_unl, G, 0.05, 4, seed=1, device=dev)
    torch.manual_seed(0)
    model = CellTypeCLF(n_genes=G, n_cell_types=4, n_hidden=64, n_hidden_init=64, n_layers=1)
    eng = TrainEngine(model.to(dev), EngineConfig(mixup_alpha=0.3, seed=1), lab, y.to(torch.int32), L, unlabeled=unl,
                      unlabeled_batch_size=U, dann=DANN(model, n_domains=2), mode="mixmatch")
    gen = torch.Generator(device=dev)
    gen.manual_seed(11)
    n_steps = 40
    eng.sample_epoch_tables(n_steps, generator=gen)
    t = eng._tabs
    assert t["n"] == n_steps and tuple(t["l"].shape) == (n_steps, L) and tuple(t["u"].shape) == (n_steps, U)
    for tab, n_rows, per in ((t["l"].cpu().numpy(), n_lab, L), (t["u"].cpu().numpy(), n_unl, U)):
        assert tab.min() >= 0 and tab.max() < n_rows
        per_pass = n_rows // per                      # drop_last: batches per pass
        passes = [tab[i:i + per_pass].ravel() for i in range(0, n_steps, per_pass)]
        for p in passes:
            assert len(set(p.tolist())) == p.size, "a row repeats inside one pass over the data"
        full = [p for p in passes if p.size == per_pass * per]
        assert len(full) >= 2 and not np.array_equal(np.sort(full[0]), full[0]), "passes are not shuffled"
        assert not np.array_equal(full[0], full[1]), "two passes used the same permutation"
    keys, gam = t["k"].cpu().numpy(), t["g"].cpu().numpy()
    assert keys.shape == gam.shape == (n_steps, L + U)
    assert keys.min() >= 0.0 and keys.max() < 1.0 and abs(keys.mean() - 0.5) < 0.02
    assert gam.min() >= 0.0 and gam.max() <= 1.0 and abs(gam.mean() - 0.5) < 0.03
    assert ((gam < 0.1) | (gam > 0.9)).mean() > 0.45, "Beta(0.3, 0.3) puts most of its mass near 0 and 1"
    # a second epoch draws new tables into the same device buffers (a captured graph stays valid)
    ptr_before = t["l"].data_ptr()
    first = t["l"].clone()
    eng.sample_epoch_tables(n_steps, generator=gen)
    assert eng._tabs["l"].data_ptr() == ptr_before and not torch.equal(eng._tabs["l"], first)


@pytest.mark.parametrize("H,use_dan", [(256, True), (128, True), (256, False)])
def test_fused_heads_match_the_split_chain_on_device_randomness(monkeypatch, H, use_dan):
    """head_ops.cu (last BatchNorm -> Dropout -> ReLU + head + losses + adjoint of the activation in one launch per branch; the
    adversary branch as a thread-block cluster) against the five-launch chain it replaces (SCNB_HEADS=split), on the
    device's own randomness: the dropout mask of the last application now arrives as keep bytes written by the producer of
    Z[last] instead of a Philox call in scnb_hs_act -- same stream, so the two engines must see the same mask.  Losses, every
    gradient and the BatchNorm statistics of the first step agree within fp32 summation-order noise."""
    from scnym_b200.engine import EngineConfig, TrainEngine
    from scnym_b200.model import CellTypeCLF, DANN
    from scnym_b200.synth import synth_device_csr
    dev = torch.device("cuda", 0)
    G, C, L, U = 3000, 11, 96, 64
    lab, y = synth_device_csr(1024, G, 0.05, C, seed=0, device=dev)
    unl, _ = synth_device_csr(1024, G, 0.05, C, seed=1, device=dev)
    out = []
    for mode in ("fused", "split"):
        monkeypatch.setenv("SCNB_HEADS", mode)
        monkeypatch.setenv("SCNB_BWD_IN", mode)     # likewise scnb_hs_bwd_in_unmix_planes against scnb_hs_bwd_in + scnb_unmix_rows_planes
        monkeypatch.setenv("SCNB_W1_UPDATE", "tc")   # (the tensor-core dW1 path, which reads dZ1 as operand planes, at this small batch)
        torch.manual_seed(3)
        model = CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=256, n_layers=2)
        dann = DANN(model, n_domains=2) if use_dan else None
        cfg = EngineConfig(n_augmentations=1, T=0.5, augment_pseudolabels=False, pseudolabel_min_confidence=0.0, mixup_alpha=0.3,
                           use_dan=use_dan, optimizer="sgd", lr=0.0, seed=11, use_cuda_graph=False)
        eng = TrainEngine(model.to(dev), cfg, lab, y.to(torch.int32), L, unlabeled=unl, unlabeled_batch_size=U, dann=dann,
                          mode="mixmatch")
        assert eng.hs_fused and eng.w1_tc and eng.heads_fused == (mode == "fused")
        eng.set_weights(1.0, 0.1)
        gen = torch.Generator(device=dev)
        gen.manual_seed(5)
        eng.sample_epoch_tables(2, generator=gen)
        eng.step()
        torch.cuda.synchronize()
        last = eng.n_app - 1
        out.append(dict(losses=eng.losses(), grad=eng.arena.grad.detach().clone(), stats=eng.stats[last].detach().clone(),
                        A=eng.As[last].detach().clone(), dY=eng.dYs[last].detach().clone(), logits=eng.logits.detach().clone(),
                        dZ1=eng.dZ1.detach().clone(), dz_hi=eng.dz_hi.detach().clone(), dz_lo=eng.dz_lo.detach().clone()))
    a, b = out
    assert torch.equal(a["A"] > 0, b["A"] > 0), "the two paths applied different dropout masks / ReLU gates"
    for k in ("dz_hi", "dz_lo"):     # bf16 planes: compare as numbers
        a[k], b[k] = a[k].view(torch.bfloat16).float(), b[k].view(torch.bfloat16).float()
    d = (a["dz_hi"] + a["dz_lo"]) - (b["dz_hi"] + b["dz_lo"])
    assert float(d.abs().max()) <= 2e-5 * float((b["dz_hi"] + b["dz_lo"]).abs().max())
    for k in ("stats", "A", "logits", "dY", "grad", "dZ1"):
        den = float(b[k].abs().max()) + 1e-30
        err = float((a[k] - b[k]).abs().max()) / den
        assert err < 2e-5, (k, err)
    for k in ("sup_loss", "unsup_loss") + (("dan_loss",) if use_dan else ()):
        assert abs(a["losses"][k] - b["losses"][k]) <= 2e-6 * max(1.0, abs(b["losses"][k])), (k, a["losses"][k], b["losses"][k])


@pytest.mark.parametrize("weighted", [False, True])
def test_val_epoch_loss_and_accuracy_match_the_oracle(tmp_path, weighted):
    """`Trainer.val_epoch()` (scnym/trainer.py:259-327: eval-mode forward, cross-entropy per batch, the reference's running
    `loss / batch size` bookkeeping, accuracy, best-weights file) on the device-resident fast path -- one fused predict pass
    over the validation CSR + the row-wise CE kernel -- against the oracle's dense eval forward on the same weights:
    epoch loss, accuracy and best loss; 1e-4 on the losses, the accuracy exact."""
    import functools
    from scipy import sparse
    from oracle import scnym_oracle as O
    from scnym_b200.dataprep import DeviceLoader, SingleCellDS
    from scnym_b200.losses import cross_entropy
    from scnym_b200.model import CellTypeCLF
    from scnym_b200.trainer import Trainer
    n, G, C, B = 777, 640, 6, 128                     # a ragged last batch of 9 rows
    ip, ii, dd, y = O.synth_csr(n, G, 0.08, C, seed=5)
    X = sparse.csr_matrix((np.asarray(dd), np.asarray(ii), np.asarray(ip)), shape=(n, G))
    y = np.asarray(y).astype(np.int64)
    torch.manual_seed(9)
    model = CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=128, n_hidden_init=256, n_layers=2).cuda()
    with torch.no_grad():                              # non-trivial running statistics
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0.0, 0.3)
                m.running_var.uniform_(0.5, 1.5)
    cw = torch.linspace(0.5, 1.5, C) if weighted else None
    crit = functools.partial(cross_entropy, class_weight=cw.cuda()) if weighted else cross_entropy
    ds = SingleCellDS(X, y, num_classes=C)
    loaders = {"train": DeviceLoader(ds, B, shuffle=True, drop_last=True), "val": DeviceLoader(ds, B)}
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    tr = Trainer(model, crit, opt, loaders, str(tmp_path), n_epochs=1, verbose=False)
    tr.epoch = 0
    tr.val_epoch()
    # oracle: dense eval forward with the same weights, then the reference's bookkeeping
    om = O.OracleModel({k: v.detach().cpu() for k, v in model.state_dict().items()}, n_layers=2)
    Xd = O.csr_to_dense(ip, ii, dd, G, np.arange(n))
    logits = om.forward(Xd, train=False)
    onehot = torch.nn.functional.one_hot(torch.from_numpy(y), C).float()
    running, n_b = 0.0, 0
    for s in range(0, n, B):
        lb = O.cross_entropy(logits[s:s + B], onehot[s:s + B], class_weight=cw)
        running += float(lb) / min(B, n - s)
        n_b += 1
    want_loss = running / n_b
    want_corrects = int((logits.argmax(1) == torch.from_numpy(y)).sum())
    got_loss = float(tr.best_loss)                     # first validation pass: the best loss IS this epoch's loss
    assert abs(got_loss - want_loss) <= 1e-4 * abs(want_loss), (got_loss, want_loss)
    r_loss, r_nb, corrects, total = tr._eval_loader(loaders["val"])
    assert r_nb == n_b and total == n
    assert abs(r_loss / r_nb - want_loss) <= 1e-4 * abs(want_loss)
    assert int(corrects) == want_corrects
    assert os.path.exists(os.path.join(str(tmp_path), "00_best_model_weights.pkl"))
