"""Reference-surface tests, modelled on the reference's own suite (tests/test_api.py, test_da.py,
test_mixmatch.py, test_model.py, test_utils.py, test_dataprep.py) but on synthetic data:
the reference's fixtures are downloaded datasets, unavailable offline."""
import os

import numpy as np
import pandas as pd
import pytest
import torch
from scipy import sparse

from tests import golden_util as GU


def _synth_adata(n_cells=1200, n_genes=300, n_classes=4, density=0.15, seed=0, unlabeled_frac=0.0, domains=False):
    """Separable synthetic log1p(CPM) data: each class over-expresses its own gene block."""
    from scnym_b200.utils import SimpleAnnData
    rng = np.random.default_rng(seed)
    y = rng.integers(0, n_classes, n_cells)
    mask = rng.random((n_cells, n_genes)) < density
    cnt = (1 + rng.poisson(2.0, (n_cells, n_genes))) * mask
    blk = n_genes // n_classes
    for c in range(n_classes):
        rows = np.where(y == c)[0]
        boost = (1 + rng.poisson(12.0, (len(rows), blk))) * (rng.random((len(rows), blk)) < 0.6)
        cnt[np.ix_(rows, np.arange(c * blk, (c + 1) * blk))] += boost
    cnt[:, 0] += 1
    X = np.log1p(1e6 * cnt / cnt.sum(1, keepdims=True)).astype(np.float32)
    obs = pd.DataFrame({"cell_type": [f"type{c}" for c in y]}, index=[f"cell{i}" for i in range(n_cells)])
    if unlabeled_frac > 0:
        tgt = rng.random(n_cells) < unlabeled_frac
        obs["annot"] = np.where(tgt, "Unlabeled", obs["cell_type"])
        if domains:
            obs["domain"] = np.where(tgt, "batchB", "batchA")
    return SimpleAnnData(sparse.csr_matrix(X), obs, [f"g{i}" for i in range(n_genes)]), y


# ------------------------------------------------------------------ CPU: host logic
def test_api_validation_errors(tmp_path):
    from scnym_b200.api import scnym_api, CONFIGS
    ad, _ = _synth_adata(60, 20)
    with pytest.raises(ValueError):
        scnym_api(ad, task="interpretx", groupby="cell_type", out_path=str(tmp_path))
    with pytest.raises(ValueError):
        scnym_api(ad, task="train", groupby="cell_type", config="nope", out_path=str(tmp_path))
    with pytest.raises(TypeError):
        scnym_api(ad, task="train", groupby="cell_type", config=3, out_path=str(tmp_path))
    with pytest.raises(ValueError):
        scnym_api(ad, task="train", groupby="not_a_column", config="no_new_identity", out_path=str(tmp_path))
    with pytest.raises(ValueError):
        scnym_api(ad, task="predict", trained_model=None, out_path=str(tmp_path))
    with pytest.raises(FileNotFoundError):
        scnym_api(ad, task="predict", trained_model=str(tmp_path / "missing"), out_path=str(tmp_path))
    dup = ad.copy()
    dup.var_names = pd.Index(["g0"] * 2 + list(dup.var_names[2:]))
    with pytest.raises(ValueError, match="Duplicate Genes"):
        scnym_api(dup, task="train", groupby="cell_type", out_path=str(tmp_path))
    raw = ad.copy()
    raw.X = sparse.csr_matrix(np.round(np.expm1(ad.X.toarray())))
    with pytest.raises(ValueError, match="Normalization error"):
        scnym_api(raw, task="train", groupby="cell_type", out_path=str(tmp_path))
    assert CONFIGS["new_identity_discovery"]["ssl_kwargs"]["pseudolabel_min_confidence"] == 0.9
    assert CONFIGS["default"]["optimizer_name"] == "adadelta" and CONFIGS["no_dan"]["ssl_kwargs"]["dan_max_weight"] == 0.0


def test_icl_weight_matches_reference_golden():
    from scnym_b200.trainer import ICLWeight
    d = GU.load("small_functions")
    for i in range(4):
        ramp, burn, mx, sig = d[f"icl{i}_cfg"]
        w = ICLWeight(ramp_epochs=int(ramp), burn_in_epochs=int(burn), max_unsup_weight=float(mx), sigmoid=bool(sig))
        assert np.allclose([w(int(e)) for e in range(40)], d[f"icl{i}"], rtol=1e-12)
    with pytest.raises(TypeError):
        w(1.0)


def test_sharpen_labels_matches_reference_golden():
    from scnym_b200.losses import sharpen_labels
    d = GU.load("small_functions")
    q = torch.from_numpy(d["sharpen_q"])
    for T in (0.0, 0.5, 1.0, 2.0):
        assert GU.rel_err(sharpen_labels(q.clone(), T).float(), d[f"sharpen_T{T}"].astype(np.float32)) < 1e-6


def test_samplemixup_keeps_dominant_observation():
    """reference tests/test_mixmatch.py::test_samplemixup"""
    from scnym_b200.dataprep import SampleMixUp
    torch.manual_seed(1)
    mix = SampleMixUp(alpha=0.3, keep_dominant_obs=True)
    n = 3
    A, B = torch.zeros((n, 2)), torch.zeros((n, 2))
    A[:, 0], B[:, 1] = 1, 1
    C = torch.cat([A, B], 0)
    s = mix({"input": C, "output": C, "some_other_tensor": C})
    for k in ("input", "some_other_tensor"):
        dom = s[k].argmax(1)
        assert torch.all(dom[:n] == 0) and torch.all(dom[n:] == 1)
    d = GU.load("small_functions")
    SampleMixUp.INJECT.append((torch.from_numpy(d["mix_keys"]), torch.from_numpy(d["mix_graw"])))
    s = mix({"input": torch.from_numpy(d["mix_a"]), "output": torch.from_numpy(d["mix_b"])})
    assert GU.rel_err(s["input"], d["mix_in"]) < 1e-6 and GU.rel_err(s["output"], d["mix_out"]) < 1e-6
    assert np.array_equal(s["random_idx"].numpy(), d["mix_ridx"])


def test_build_classification_matrix_dense_and_sparse():
    """reference tests/test_utils.py:13-73"""
    from scnym_b200.utils import build_classification_matrix
    rng = np.random.default_rng(0)
    X = rng.random((50, 30)) * (rng.random((50, 30)) < 0.3)
    sample_genes = np.array([f"g{i}" for i in range(30)])
    model_genes = np.array([f"g{i}" for i in rng.permutation(30)[:20]] + ["absent1", "absent2"])
    for kind in ("dense", "sparse"):
        Xi = X if kind == "dense" else sparse.csr_matrix(X)
        N = build_classification_matrix(Xi, model_genes, sample_genes)
        assert type(N) == type(Xi) and N.shape == (50, 22)
        Nd = N if kind == "dense" else N.toarray()
        for j, g in enumerate(model_genes):
            if g.startswith("absent"):
                assert np.all(Nd[:, j] == 0)
            else:
                assert np.allclose(Nd[:, j], X[:, int(g[1:])])
    assert build_classification_matrix(X, sample_genes, sample_genes) is X
    with pytest.raises(TypeError):
        build_classification_matrix(sparse.csc_matrix(X), model_genes, sample_genes)


def test_single_cell_ds_contract():
    """reference scnym/dataprep.py:12-168"""
    from scnym_b200.dataprep import SingleCellDS
    X = sparse.csr_matrix(np.arange(12, dtype=np.float32).reshape(4, 3))
    ds = SingleCellDS(X, np.array([0, 1, 2, 1]), num_classes=3)
    s = ds[1]
    assert s["input"].tolist() == [3.0, 4.0, 5.0] and s["output"].tolist() == [0.0, 1.0, 0.0]
    assert np.all(np.asarray(s["domain"]) == -1) and len(ds) == 4
    dsd = SingleCellDS(X.toarray(), np.array([0, 1, 2, 1]), domain=np.array([0, 1, 1, 0]), num_classes=3, num_domains=2)
    assert dsd[3]["domain"].tolist() == [1.0, 0.0]
    with pytest.raises(TypeError):
        SingleCellDS(X.tocsc(), np.zeros(4))
    with pytest.raises(TypeError):
        ds[np.int64(1)]
    with pytest.raises(ValueError):
        SingleCellDS(X, np.zeros(5))


def test_count_domains_quirks():
    from scnym_b200.main import _count_domains
    assert _count_domains(None, None, None) == 1 and _count_domains(None, None, np.zeros((2, 2))) == 2
    assert _count_domains(np.array([0, 1, 2]), None, None) == 2          # int(max), no +1 (main.py:269-271)
    assert _count_domains(np.array([0, 1]), np.array([2, 3]), np.zeros((2, 2))) == 4
    with pytest.raises(ValueError):
        _count_domains(None, np.array([0]), None)


# ------------------------------------------------------------------ GPU: behaviour
@pytest.mark.gpu
def test_scnym_api_supervised_train_then_predict(tmp_path):
    """BASELINE configs[0] shape (shrunk): scnym_api train (supervised Trainer, Adadelta) -> predict."""
    from scnym_b200.api import scnym_api
    torch.manual_seed(0); np.random.seed(0)
    ad, y = _synth_adata(1500, 400, 5, seed=1)
    out = str(tmp_path / "sup")
    scnym_api(ad, task="train", groupby="cell_type", out_path=out, config={"n_epochs": 6, "patience": 40, "batch_size": 128})
    res = ad.uns["scNym_train_results"]
    for f in ("00_best_model_weights.pkl", "scnym_train_results.pkl", "train_idx.csv", "test_idx.csv", "val_idx.csv",
              "predictions.csv", "labels.csv", "sup_log.csv"):
        assert os.path.exists(os.path.join(out, f)), f
    assert res["n_genes"] == 400 and res["n_cell_types"] == 5 and res["final_acc"] > 0.9
    sd = torch.load(res["model_path"])
    assert sd["embed.0.weight"].shape == (256, 400) and sd["embed.0.weight"].is_contiguous()
    scnym_api(ad, task="predict", trained_model=out, out_path=out, key_added="scNym", config="no_new_identity")
    acc = np.mean(np.array(ad.obs["scNym"]) == np.array(ad.obs["cell_type"]))
    assert acc > 0.9
    assert ad.obsm["X_scnym"].shape == (1500, 256) and ad.uns["scNym_probabilities"].shape == (1500, 5)
    assert np.allclose(ad.uns["scNym_probabilities"].sum(1), 1.0, atol=1e-4)
    assert np.all(ad.obs["scNym_confidence"] <= 1.0)


@pytest.mark.gpu
def test_scnym_api_mixmatch_dan_train(tmp_path):
    """Semi-supervised + domain adversary through the API (default `new_identity_discovery`)."""
    from scnym_b200.api import scnym_api
    from scnym_b200.api import CONFIGS
    import copy
    torch.manual_seed(0); np.random.seed(0)
    saved = copy.deepcopy(CONFIGS["default"])
    try:
        ad, y = _synth_adata(2000, 300, 4, seed=2, unlabeled_frac=0.4, domains=True)
        out = str(tmp_path / "ssl")
        cfg = copy.deepcopy(CONFIGS["new_identity_discovery"])
        cfg.update(n_epochs=8, batch_size=128)
        cfg["ssl_kwargs"].update(min_epochs=1, ramp_epochs=4, dan_ramp_epochs=2)
        scnym_api(ad, task="train", groupby="annot", domain_groupby="domain", out_path=out, config=cfg)
        assert os.path.exists(os.path.join(out, "02_best_dan_weights.pkl"))
        assert ad.uns["scNym_train_results"]["final_acc"] > 0.85
        log = pd.read_csv(os.path.join(out, "ssl_log.csv"))
        sup = log[log["Mode"] == "train_epoch_sup"]["Running_Loss"].to_numpy()
        assert sup[-1] < sup[0]
        scnym_api(ad, task="predict", trained_model=out, out_path=out, config="new_identity_discovery")
        tgt = np.array(ad.obs["annot"]) == "Unlabeled"
        assert np.mean(np.array(ad.obs["scNym"])[tgt] == np.array(ad.obs["cell_type"])[tgt]) > 0.85
    finally:
        CONFIGS["default"] = saved


@pytest.mark.gpu
def test_dan_gradient_reversal_behaviour():
    """reference tests/test_da.py:259-442 (Gaussian blobs): GRL weight 0 cuts the gradient to the embedding,
    the domain head learns; weight 1 pushes the embedding to confuse it."""
    from scnym_b200 import trainer as T
    from scnym_b200.model import CellTypeCLF
    from scnym_b200.optim import AdamW
    torch.manual_seed(1); np.random.seed(1)
    model = CellTypeCLF(n_genes=3, n_cell_types=2, n_layers=1, n_hidden=16).cuda()
    A, B = torch.randn(48, 3), torch.randn(48, 3) + 5
    dan = T.DANLoss(model=model, dan_criterion=T.cross_entropy, n_domains=3)
    X0, X1, X2 = torch.cat([A[:16], B[:16]]), torch.cat([A[16:32], B[16:32]]) + 1000.0, torch.cat([A[32:], B[32:]]) - 1000.0
    ldom = torch.nn.functional.one_hot(torch.cat([torch.zeros(32), torch.ones(32)]).long(), 3)
    lab = {"input": torch.cat([X0, X2]).cuda(), "output": torch.zeros(64, 2).cuda(), "domain": ldom.cuda()}
    unl = {"input": X1.cuda(), "output": torch.zeros(32, 2).cuda(),
           "domain": torch.nn.functional.one_hot((torch.zeros(32) + 2).long(), 3).cuda()}
    opt = AdamW([{"params": model.parameters()}, {"params": dan.dann.domain_clf.parameters()}], weight_decay=1e-4)
    losses = []
    for i in range(50):
        opt.zero_grad()
        l = dan(labeled_sample=lab, unlabeled_sample=unl, weight=0.0)
        l.backward()
        assert sum(float((p.grad ** 2).sum()) for p in model.embed.parameters() if p.grad is not None) == 0.0
        assert sum(float((p.grad ** 2).sum()) for p in dan.dann.domain_clf.parameters()) > 0.0
        opt.step()
        losses.append(l.item())
    assert losses[0] > losses[-1]
    opt = AdamW([{"params": model.parameters()}], weight_decay=1e-4)
    p0 = [p.detach().clone() for p in dan.dann.domain_clf.parameters()]
    losses = []
    for i in range(50):
        opt.zero_grad()
        l = dan(labeled_sample=lab, unlabeled_sample=unl, weight=1.0)
        l.backward()
        opt.step()
        assert sum(float((p.grad ** 2).sum()) for p in model.embed.parameters() if p.grad is not None) > 0.0
        losses.append(l.item())
    assert losses[0] < losses[-1]
    for a, b in zip(p0, dan.dann.domain_clf.parameters()):
        assert torch.equal(a, b.detach())
    assert dan.x_embed.shape == (96, 16) and 0.0 <= float(dan.dan_acc) <= 1.0


@pytest.mark.gpu
def test_mixmatch_loss_module_level_matches_reference_golden():
    """`MixMatchLoss.__call__` + `DANLoss.__call__` on DENSE tensors (the reference's calling convention) with
    injected randomness, against the reference golden step 0 (losses + labeled logits)."""
    from tests import cuda_harness as CH
    from scnym_b200 import functional as Fn
    from scnym_b200 import trainer as T
    from scnym_b200.dataprep import Log1pDrop, SampleMixUp
    d = GU.load("mixmatch_dan_adam")
    model, dann = CH.build_models(d)
    inp = GU.step_inputs(d, 0)
    mm = T.MixMatchLoss(alpha=0.3, unsup_criterion=torch.nn.MSELoss(reduction="none"), sup_criterion=T.cross_entropy,
                        decay_coef=0.997, mean_teacher=False, augment=Log1pDrop(0.1), n_augmentations=1, T=0.5,
                        augment_pseudolabels=False, pseudolabel_min_confidence=0.0)
    dan = T.DANLoss(model=model, dan_criterion=T.cross_entropy, n_domains=2)
    dan.dann.domain_clf.load_state_dict(dann.domain_clf.state_dict())
    model.train()
    r = inp["rnd"]
    Log1pDrop.KEEP_INJECT.append(r.in_keep_L)
    SampleMixUp.INJECT.append((r.perm_keys, r.gamma_raw))
    Fn.DROPOUT_INJECT.extend(list(r.drop_keep_U) + list(r.drop_keep_L))
    data = {"input": inp["Lx"].cuda(), "output": inp["Ly"].cuda(), "domain": torch.zeros(24, 5).cuda() - 1}
    unl = {"input": inp["Ux"].cuda(), "output": torch.zeros(24).cuda(), "domain": torch.zeros(24, 5).cuda() - 1}
    sup, unsup, logits = mm(model=model, labeled_sample=data, unlabeled_sample=unl)
    Fn.DROPOUT_INJECT.extend(list(r.drop_keep_D))
    dl = dan(labeled_sample=data, unlabeled_sample=unl, weight=float(d["meta_dan_w"]),
             pseudolabel_confidence=mm.running_confidence_scores[-1][0])
    loss = sup + float(d["meta_unsup_w"]) * unsup + dl
    loss.backward()
    assert GU.rel_err(sup.item(), d["step0/sup_loss"]) < 1e-4
    assert GU.rel_err(unsup.item(), d["step0/unsup_loss"]) < 1e-4
    assert GU.rel_err(dl.item(), d["step0/dan_loss"]) < 1e-4
    assert GU.rel_err(logits.detach().cpu(), d["step0/labeled_logits"]) < 1e-4
    for n, p in model.named_parameters():
        assert GU.rel_err(p.grad.cpu(), d["step0/grad/" + n]) < 1e-4, n
    assert mm.step == 1 and len(mm.running_confidence_scores) == 1 and not Fn.DROPOUT_INJECT
    # threshold nobody passes -> zero unsupervised loss (reference tests/test_mixmatch.py:304)
    mm.pseudolabel_min_confidence = 1.0
    data = {"input": inp["Lx"].cuda(), "output": inp["Ly"].cuda()}
    unl = {"input": inp["Ux"].cuda(), "output": torch.zeros(24).cuda()}
    _, unsup0, _ = mm(model=model, labeled_sample=data, unlabeled_sample=unl)
    assert float(unsup0) == 0.0


@pytest.mark.gpu
def test_host_train_pipeline_feeds_the_engine_from_host_memory():
    """HostTrainPipeline: host CSR -> pinned blob -> H2D -> staged batch -> step; losses come back per step."""
    from oracle import scnym_oracle as O
    from scnym_b200.engine import EngineConfig
    from scnym_b200.hostfeed import HostTrainPipeline
    from scnym_b200.model import CellTypeCLF, DANN
    G, C, L, U, n = 600, 4, 32, 48, 300
    lp, li, ld, ly = O.synth_csr(n, G, 0.06, C, seed=21)
    up, ui, ud, _ = O.synth_csr(n, G, 0.06, C, seed=22)
    torch.manual_seed(0)
    model = CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=128, n_hidden_init=128, n_layers=2)
    dann = DANN(model, n_domains=2)
    cfg = EngineConfig(optimizer="adam", lr=1e-3, weight_decay=1e-4, seed=3)
    pipe = HostTrainPipeline(model, cfg, lab=(lp, li, ld, ly), unl=(up, ui, ud), n_genes=G, batch_size=L, unlabeled_batch_size=U,
                             dann=dann, seed=9)
    pipe.engine.set_weights(1.0, 0.1)
    w0 = model.classif[0].weight.detach().clone()
    losses = list(pipe.run(7))
    assert len(losses) == 7 and all(np.isfinite(l["loss"]) for l in losses)
    assert pipe.h2d_bytes_last > 8 * (L + U) and pipe.d2h_bytes_last == 32
    assert not torch.equal(w0, model.classif[0].weight.detach())
    # the staged unlabeled rows of the LAST step are exactly the host rows sampled for step 6 (tables of pre-drawn samples)
    assert pipe._step0 == 7
    lrows, urows, _, _ = pipe._draws(6)
    # shuffled passes without replacement, batches never straddle a pass (DataLoader(shuffle, drop_last) semantics)
    per_pass = n // L
    first_pass = np.concatenate([pipe._draws(s)[0] for s in range(per_pass)])
    assert len(set(first_pass.tolist())) == per_pass * L
    want_idx = np.concatenate([ui[up[r]:up[r + 1]] for r in urows])
    want_val = np.concatenate([ud[up[r]:up[r + 1]] for r in urows])
    eng = pipe.engine
    nnz_u = int(eng.b_indptr[U])
    assert nnz_u == want_idx.size
    assert np.array_equal(eng.b_idx[:nnz_u].cpu().numpy(), want_idx)
    assert np.array_equal(eng.b_val[:nnz_u].cpu().numpy(), want_val)
    assert np.array_equal(eng.lab_ids.cpu().numpy(), ly[lrows].astype(np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["small", "config2"])
def test_host_mapped_pipeline_reads_the_same_batches_as_device_resident_training(shape):
    """HostMappedTrainPipeline: the matrices stay in (page-locked, mapped) HOST memory and the step's own row gather reads
    the sampled rows over PCIe; the loss is stored into mapped host memory.  Against a TrainEngine on device-resident
    copies of the same matrices with the same sampler seed: identical sampled rows, bit-identical batch CSR (input
    dropout included), labels and domains at every step, losses of the first steps within the fp32 budget -- on the
    small shape (batch assembled at the head of the step) and at the config-2 batch shape (tensor-core path: the batch of
    step t+1 is assembled while step t runs, one read-back slot per set of batch buffers)."""
    from oracle import scnym_oracle as O
    from scnym_b200.dataprep import DeviceCSR
    from scnym_b200.engine import EngineConfig, TrainEngine
    from scnym_b200.hostfeed import HostMappedTrainPipeline
    from scnym_b200.model import CellTypeCLF, DANN
    if shape == "small":
        G, C, L, U, n, H, dens = 600, 4, 32, 64, 300, 128, 0.06
    else:
        G, C, L, U, n, H, dens = 4096, 8, 256, 256, 2048, 256, 0.05
    lp, li, ld, ly = O.synth_csr(n, G, dens, C, seed=21)
    up, ui, ud, _ = O.synth_csr(n, G, dens, C, seed=22)
    dl = (np.arange(n) % 3).astype(np.int32)
    du = (np.arange(n) % 3 + 1).astype(np.int32) % 3
    dev = torch.device("cuda", 0)

    def build():
        torch.manual_seed(0)
        model = CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H, n_layers=2)
        return model, DANN(model, n_domains=3), EngineConfig(optimizer="sgd", lr=0.01, seed=3, pseudolabel_min_confidence=0.0)

    model, dann, cfg = build()
    pipe = HostMappedTrainPipeline(model, cfg, lab=(lp, li, ld, ly, dl), unl=(up, ui, ud, du), n_genes=G, batch_size=L,
                                   unlabeled_batch_size=U, dann=dann, seed=9)
    pipe.epoch_steps = 4
    pipe.engine.set_weights(1.0, 0.1)
    assert bool(pipe.engine.prefetch_ok) == (shape == "config2")
    model2, dann2, cfg2 = build()
    t = lambda a, dt: torch.as_tensor(np.asarray(a), dtype=dt).to(dev)
    lab = DeviceCSR(t(lp, torch.int64), t(li, torch.int32), t(ld, torch.float32), G)
    unl = DeviceCSR(t(up, torch.int64), t(ui, torch.int32), t(ud, torch.float32), G)
    eng = TrainEngine(model2.to(dev), cfg2, lab, t(ly, torch.int32), L, unlabeled=unl, unlabeled_batch_size=U, dann=dann2,
                      labeled_domain=t(dl, torch.int32), unlabeled_domain=t(du, torch.int32), mode="mixmatch")
    eng.set_weights(1.0, 0.1)
    gen = torch.Generator(device=dev)
    gen.manual_seed(9)
    it = pipe.run(6)
    fields = ("l_rows", "u_rows", "lab_ids", "base_dom", "b_indptr")
    pe = pipe.engine

    def snapshot(e):
        nnz = int(e.b_indptr[-1])
        d = {k: getattr(e, k).clone() for k in fields}
        d["b_idx"], d["b_val"] = e.b_idx[:nnz].clone(), e.b_val[:nnz].clone()
        return d

    ahead = None     # the pipeline's batch of the step the reference engine runs next (it yields losses one step late)
    compared = 0
    for s in range(6):
        if s % 4 == 0:
            eng.sample_epoch_tables(4, generator=gen)
        eng.step()
        torch.cuda.synchronize()
        mine = snapshot(eng)
        if ahead is not None:
            for k in mine:
                assert torch.equal(ahead[k], mine[k]), (s, k)
            compared += 1
        got = next(it)
        torch.cuda.synchronize()
        if pe._pf_graphs:
            ahead = snapshot(pe) if s + 1 < 6 else None
        else:
            theirs = snapshot(pe)
            for k in mine:
                assert torch.equal(theirs[k], mine[k]), (s, k)
            compared += 1
        want = eng.losses()
        if s < 2:
            for k in ("sup_loss", "unsup_loss", "dan_loss"):
                assert abs(got[k] - want[k]) <= 1e-4 * max(1.0, abs(want[k])), (s, k, got[k], want[k])
        assert np.isfinite(got["loss"])
    assert compared >= 5
    assert list(it) == []
    assert pipe.d2h_bytes_last == 32 and pipe.h2d_bytes_last > 8 * (L + U)
    pipe.close()
