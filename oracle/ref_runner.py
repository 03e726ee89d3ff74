"""Drive the UNMODIFIED reference (calico/scnym v0.4.0) through its own public classes, for timing.

TEST INFRASTRUCTURE ONLY: used by `bench.py`'s `--impl reference` arm, its `cpu_baseline` /
`gpu_eager_baseline` legs and by tests.  Nothing on the product path imports this module.

The reference is pure Python on PyTorch.  `__graft_entry__.build()` installs it, unmodified, with
    python -m pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>
(`oracle/_ref/` is git-ignored: no reference source enters the history; it is NOT gpurun-ignored, so the
installed package travels to the GPU box like our own built .so).  `oracle/ref_shim.py` imports it with the three
absent third-party modules (anndata, scanpy, tensorboardX) stubbed.

What is timed is the reference's own loop, built exactly as `scnym.main.fit_model` builds it
(scnym/main.py:229-232, 317-328, 374-396, 478-599): `SingleCellDS` + `DataLoader(shuffle, drop_last)`,
`repeater` around the unlabeled loader, `MixMatchLoss(augment=AUGMENTATION_SCHEMES["log1p_drop"])`, `DANLoss`,
`torch.optim`, and `SemiSupervisedTrainer.train_epoch()` (scnym/trainer.py:551-702); predict is
`Predicter.predict` (scnym/predict.py:107-216).
"""
import os
import tempfile
import time

import numpy as np
import torch


def available() -> bool:
    from oracle import ref_shim
    return ref_shim.reference_available()


def _ref():
    from oracle import ref_shim
    return ref_shim.load_reference()


class _ListLoader(list):
    """Pre-collated batches in place of a DataLoader (the "compute only" figure): the trainer only iterates it, takes
    `len()` and reads `.batch_size` / `.sampler` for its log."""
    batch_size = 0
    sampler = None


def _csr(ip, ii, dd, G):
    from scipy import sparse
    return sparse.csr_matrix((dd, ii, ip), shape=(len(ip) - 1, G))


def build_trainer(c, lab, lab_y, unl, device, out_path, lab_dom=None, unl_dom=None, optimizer="adam", lr=1e-3, wd=1e-4,
                  ssl_kwargs=None, dense_batches=False, seed=0):
    """The semi-supervised branch of `fit_model` (scnym/main.py:229-232, 317-328, 374-396, 478-599) on CSR matrices
    `lab` / `unl` (scipy).  `c`: dict with G, C, H0, H, n_layers, D, L (batch size).  Returns the trainer."""
    ref = _ref()
    from torch.utils.data import DataLoader
    torch.manual_seed(seed)
    ssl_kwargs = dict(ssl_kwargs or {})
    n_domains = int(c["D"])
    ds = ref.dataprep.SingleCellDS(X=lab, y=np.asarray(lab_y), domain=lab_dom, num_classes=c["C"], num_domains=n_domains)
    uds = ref.dataprep.SingleCellDS(X=unl, y=np.zeros(unl.shape[0]), domain=unl_dom, num_classes=c["C"], num_domains=n_domains)
    B = int(c["L"])
    train_dl = DataLoader(ds, batch_size=B, shuffle=True, drop_last=True)
    unsup_dl = DataLoader(uds, batch_size=B, shuffle=True, drop_last=True)
    if dense_batches:   # collate once, outside any timed region
        train_dl = _ListLoader(list(train_dl))
        unsup_dl = _ListLoader(list(unsup_dl))
        train_dl.batch_size = unsup_dl.batch_size = B
    model = ref.model.CellTypeCLF(n_genes=c["G"], n_cell_types=c["C"], n_hidden=c["H"], n_hidden_init=c["H0"],
                                  n_layers=c["n_layers"])
    if device != "cpu":
        model = model.cuda()
    opt_cls = ref.main.OPTIMIZERS[optimizer]
    opt = opt_cls(model.parameters(), lr=lr, weight_decay=wd) if optimizer != "sgd" else \
        opt_cls(model.parameters(), lr=lr, weight_decay=wd, momentum=0.9)
    from torch import nn
    usl = ref.losses.MixMatchLoss(alpha=0.3, unsup_criterion=nn.MSELoss(reduction="none"), sup_criterion=ref.losses.cross_entropy,
                                  decay_coef=0.997, mean_teacher=False,
                                  augment=ref.dataprep.AUGMENTATION_SCHEMES[ssl_kwargs.get("augment", "log1p_drop")],
                                  n_augmentations=ssl_kwargs.get("n_augmentations", 1), T=ssl_kwargs.get("T", 0.5),
                                  augment_pseudolabels=ssl_kwargs.get("augment_pseudolabels", False),
                                  pseudolabel_min_confidence=ssl_kwargs.get("pseudolabel_min_confidence", 0.0))
    dan = ref.losses.DANLoss(model=model, dan_criterion=ref.losses.cross_entropy, n_domains=n_domains)
    opt.add_param_group({"params": dan.dann.domain_clf.parameters(), "name": "domain_classifier"})
    w_u = ref.trainer.ICLWeight(ramp_epochs=1, max_unsup_weight=1.0, burn_in_epochs=0, sigmoid=False)
    w_d = ref.trainer.ICLWeight(ramp_epochs=1, max_unsup_weight=0.1, burn_in_epochs=0, sigmoid=True)
    T = ref.trainer.SemiSupervisedTrainer(
        unsup_dataloader=unsup_dl if dense_batches else ref.main.repeater(unsup_dl), unsup_criterion=usl, unsup_weight=w_u,
        dan_criterion=dan, dan_weight=w_d, model=model, criterion=ref.losses.cross_entropy, optimizer=opt, scheduler=None,
        dataloaders={"train": train_dl, "val": train_dl}, out_path=out_path, batch_transformers={}, n_epochs=10, min_epochs=0,
        save_freq=100, reg_criterion=None, exp_name="bench", verbose=False, tb_writer=None, patience=None,
        use_gpu=(device != "cpu"))
    T.epoch = 5   # past both ramps: every loss term is active
    return T


def time_train(c, synth, n_steps, warmup, device="cpu", dense_batches=False, threads=None, given_domains=False, seed=0):
    """Seconds per minibatch of `SemiSupervisedTrainer.train_epoch()` on a synthetic dataset of exactly
    (warmup + n_steps) batches (the epoch IS the timed loop; throughput is per step, independent of the cell count).
    `synth(n_cells, seed)` -> (indptr, indices, data, labels) numpy CSR.  Returns (s_per_step, steps)."""
    if threads:
        torch.set_num_threads(int(threads))
    B = int(c["L"])
    rng = np.random.default_rng(seed + 5)
    out_path = tempfile.mkdtemp(prefix="scnym_ref_bench_")

    def make(n_batches, s):
        n = n_batches * B
        lp, li, ld, ly = synth(n, s)
        up, ui, ud, _ = synth(n, s + 1)
        ldom = udom = None
        if given_domains:
            D = int(c["D"])
            ldom = rng.integers(0, max(D // 2, 1), size=n)
            udom = rng.integers(max(D // 2, 1), D, size=n)
        return build_trainer(c, _csr(lp, li, ld, c["G"]), ly, _csr(up, ui, ud, c["G"]), device, out_path, ldom, udom,
                             dense_batches=dense_batches, seed=seed)

    T = make(max(warmup, 1) + n_steps, seed)
    # warm-up: the first `warmup` batches of the same epoch are run through a one-off truncated loader
    full = T.dataloaders["train"]
    if dense_batches:
        warm, timed = _ListLoader(full[:warmup]), _ListLoader(full[warmup:warmup + n_steps])
        warm.batch_size = timed.batch_size = B
        uw = _ListLoader(T.unsup_dataloader[:warmup]); ut = _ListLoader(T.unsup_dataloader[warmup:warmup + n_steps])
        T.dataloaders["train"], T.unsup_dataloader = warm, uw
        T.train_epoch()
        T.dataloaders["train"], T.unsup_dataloader = timed, ut
    else:
        from torch.utils.data import DataLoader, Subset
        ds = full.dataset
        idx = list(range(len(ds)))
        T.dataloaders["train"] = DataLoader(Subset(ds, idx[:max(warmup, 1) * B]), batch_size=B, shuffle=True, drop_last=True)
        T.train_epoch()
        T.dataloaders["train"] = DataLoader(Subset(ds, idx[max(warmup, 1) * B:]), batch_size=B, shuffle=True, drop_last=True)
    if device != "cpu":
        torch.cuda.synchronize()
    steps = len(T.dataloaders["train"])
    t0 = time.perf_counter()
    T.train_epoch()
    if device != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt / max(steps, 1), steps


def time_predict(c, synth, n_cells, device="cpu", threads=None, batch_size=1024, seed=2):
    """Seconds for `Predicter.predict` (scnym/predict.py:107-216) over `n_cells` synthetic cells (CSR input, the
    reference's per-cell densifying loader included) + the embedding pass of `scnym_api` (scnym/api.py:692-708 runs a
    second full pass through `model.embed[:-1]`; counted as 2x the forward when `with_embed`)."""
    ref = _ref()
    if threads:
        torch.set_num_threads(int(threads))
    torch.manual_seed(0)
    model = ref.model.CellTypeCLF(n_genes=c["G"], n_cell_types=c["C"], n_hidden=c["H"], n_hidden_init=c["H0"],
                                  n_layers=c["n_layers"])
    d = tempfile.mkdtemp(prefix="scnym_ref_pred_")
    path = os.path.join(d, "w.pkl")
    torch.save(model.state_dict(), path)
    ip, ii, dd, _ = synth(n_cells, seed)
    X = _csr(ip, ii, dd, c["G"])
    # Predicter moves the models to CUDA whenever a device is visible (scnym/predict.py:95-97); hide it for the CPU figure
    has = torch.cuda.is_available
    if device == "cpu":
        torch.cuda.is_available = lambda: False
    try:
        P = ref.predict.Predicter(model_weights=path, n_genes=c["G"], n_cell_types=c["C"], n_hidden=c["H"],
                                  n_hidden_init=c["H0"], n_layers=c["n_layers"])
        P.predict(X[:min(2048, n_cells)], output="prob", batch_size=batch_size)   # warm-up
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        pred, _, prob = P.predict(X, output="prob", batch_size=batch_size)
        if device != "cpu":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        torch.cuda.is_available = has
    assert len(pred) == n_cells
    return dt
