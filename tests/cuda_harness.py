"""Drive the CUDA engine from a golden/oracle "case" (dict-like, same keys as tests/golden/*.npz)
with the case's injected randomness, and collect what the parity tests compare."""
import numpy as np
import torch

from oracle import scnym_oracle as O
from tests import golden_util as GU


class Case(dict):
    """dict with the `.files` attribute of an NpzFile so golden_util helpers accept both."""

    @property
    def files(self):
        return list(self.keys())


def build_models(d, device="cuda"):
    from scnym_b200.model import CellTypeCLF, DANN
    G, C, H0, H, nl = (int(d[k]) for k in ("meta_G", "meta_C", "meta_H0", "meta_H", "meta_n_layers"))
    model = CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H0, n_layers=nl)
    st = GU.state_from(d, "init/")
    model_st, dan_st = GU.split_state(st)
    model.load_state_dict(model_st)
    model = model.to(device)
    dann = None
    if int(d["meta_use_dan"]) if "meta_use_dan" in d.files else False:
        dann = DANN(model, n_domains=int(d["meta_D"]))
        dann.domain_clf.load_state_dict({k[len("domain_clf."):]: v for k, v in dan_st.items()})
        dann.domain_clf.to(device)
    return model, dann


def build_engine(d, device="cuda", first_layer=None, w1_update=None, multi_stream=None):
    from scnym_b200.dataprep import DeviceCSR
    from scnym_b200.engine import TrainEngine, EngineConfig
    model, dann = build_models(d, device)
    G = int(d["meta_G"])
    lab = DeviceCSR.from_arrays(d["lab_indptr"], d["lab_indices"], d["lab_data"], G, device)
    unl = DeviceCSR.from_arrays(d["unl_indptr"], d["unl_indices"], d["unl_data"], G, device)
    cw = d["class_weight"].tolist() if "class_weight" in d.files else None
    cfg = EngineConfig(n_augmentations=int(d["meta_K"]), T=float(d["meta_T"]), augment_pseudolabels=bool(d["meta_aug_pl"]),
                       pseudolabel_min_confidence=float(d["meta_tau"]), mixup_alpha=0.3,
                       mean_teacher=bool(d["meta_mean_teacher"]), decay_coef=0.997, use_dan=bool(d["meta_use_dan"]),
                       dan_use_conf_pseudolabels=bool(d["meta_dan_conf"]), dan_scale_loss_pseudoconf=bool(d["meta_dan_scale"]),
                       optimizer=str(d["meta_opt"]), lr=float(d["meta_lr"]), weight_decay=float(d["meta_wd"]), class_weight=cw,
                       use_cuda_graph=False, keep_w1_grad=True)
    if first_layer is not None:
        cfg.first_layer = first_layer
    if w1_update is not None:
        cfg.w1_update = w1_update
    if multi_stream is not None:
        cfg.multi_stream = multi_stream
    given = bool(int(d["meta_D_given"]))
    eng = TrainEngine(model, cfg, lab, d["lab_y"], int(d["meta_L"]), unlabeled=unl, unlabeled_batch_size=int(d["meta_U"]),
                      dann=dann, labeled_domain=d["lab_dom"] if given else None,
                      unlabeled_domain=d["unl_dom"] if given else None, mode="mixmatch")
    eng.set_weights(float(d["meta_unsup_w"]), float(d["meta_dan_w"]))
    return eng


def _row_keep(indptr, indices, rows, dense_keep):
    """Per-nnz keep bytes (batch order) from a dense [rows, G] keep mask."""
    out = []
    for i, r in enumerate(rows):
        s, e = indptr[r], indptr[r + 1]
        out.append(dense_keep[i, indices[s:e]])
    return np.concatenate(out).astype(np.uint8) if out else np.zeros(0, np.uint8)


def inject_step(eng, d, s):
    """Stage rows + randomness of step `s` into the engine (injected masks, no Philox)."""
    dev = eng.dev
    pre = f"step{s}/"
    lrow, urow = d[pre + "lrow"], d[pre + "urow"]
    eng.load_batch(lrow, urow, d[pre + "perm_keys"], d[pre + "gamma_raw"])
    # input-dropout mask of the labelled rows, at absolute batch-nnz positions (after the U rows)
    u_len = int(sum(d["unl_indptr"][r + 1] - d["unl_indptr"][r] for r in urow))
    keepL = _row_keep(d["lab_indptr"], d["lab_indices"], lrow, d[pre + "in_keep_L"])
    full = np.ones(eng.cap, dtype=np.uint8)
    full[u_len:u_len + keepL.size] = keepL
    eng.inj_in_keep_L = torch.from_numpy(full).to(dev)
    if (pre + "in_keep_U") in d.files:
        ku = d[pre + "in_keep_U"]
        eng.inj_in_keep_U = []
        for k in range(ku.shape[0]):
            m = _row_keep(d["unl_indptr"], d["unl_indices"], urow, ku[k])
            buf = np.ones(eng.tb_idx.numel(), dtype=np.uint8)
            buf[:m.size] = m
            eng.inj_in_keep_U.append(torch.from_numpy(buf).to(dev))
    hk = []
    for a in range(eng.n_app):
        w = eng.widths[a]
        m = np.ones((eng.Rmax, w), dtype=np.uint8)
        m[:eng.U] = d[pre + f"keep_U{a}"]
        m[eng.U:eng.nb] = d[pre + f"keep_L{a}"]
        if eng.use_dan:
            kd = d[pre + f"keep_D{a}"]
            m[eng.nb:eng.nb + kd.shape[0]] = kd
        hk.append(torch.from_numpy(m).to(dev))
    eng.inj_hidden_keep = hk


def collect(eng):
    torch.cuda.synchronize()
    out = dict(eng.losses())
    U, nb = eng.U, eng.nb
    out["labeled_logits"] = eng.logits[U:nb].cpu()
    out["pseudolabels"] = eng.Ycat[:U].cpu()
    out["confident"] = eng.confident.cpu().numpy().astype(bool)
    out["confidence"] = eng.conf.cpu()
    grads = {}
    for n, p in eng.model.named_parameters():
        grads[n] = p.grad.detach().cpu().clone()
    if eng.use_dan:
        for n, p in eng.dann.domain_clf.named_parameters():
            grads["domain_clf." + n] = p.grad.detach().cpu().clone()
    out["grads"] = grads
    post = {k: v.detach().cpu().clone() for k, v in eng.model.state_dict().items()}
    if eng.use_dan:
        post.update({"domain_clf." + k: v.detach().cpu().clone() for k, v in eng.dann.domain_clf.state_dict().items()})
    out["post"] = post
    return out


def oracle_case(*, G, L, U, C, H0, H, n_layers, K=1, T=0.5, aug_pl=False, tau=0.0, D_given=0, opt_name="adam", lr=1e-3,
                wd=1e-4, unsup_w=1.0, dan_w=0.1, n_steps=1, mean_teacher=False, dan_conf=False, dan_scale=False,
                density=0.05, seed=0, use_dan=True, n_cells=None):
    """Build a case in the golden format whose expected values come from the ORACLE (for sizes
    the golden fixtures do not cover)."""
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed + 100)
    n_cells = n_cells or 3 * max(L, U)
    lp, li, ld, ly = O.synth_csr(n_cells, G, density, C, seed=seed)
    up, ui, ud, _ = O.synth_csr(n_cells, G, density, C, seed=seed + 1)
    n_domains = D_given if D_given else 2
    d = Case(meta_G=G, meta_L=L, meta_U=U, meta_C=C, meta_H0=H0, meta_H=H, meta_n_layers=n_layers, meta_K=K, meta_T=T,
             meta_aug_pl=int(aug_pl), meta_tau=tau, meta_D=n_domains, meta_D_given=int(bool(D_given)), meta_opt=opt_name,
             meta_lr=lr, meta_wd=wd, meta_unsup_w=unsup_w, meta_dan_w=dan_w, meta_n_steps=n_steps,
             meta_mean_teacher=int(mean_teacher), meta_dan_conf=int(dan_conf), meta_dan_scale=int(dan_scale),
             meta_use_dan=int(use_dan), meta_unsup_criterion="mse", lab_indptr=lp, lab_indices=li, lab_data=ld, lab_y=ly,
             unl_indptr=up, unl_indices=ui, unl_data=ud)
    if D_given:
        d["lab_dom"] = rng.integers(0, max(D_given // 2, 1), size=n_cells).astype(np.int64)
        d["unl_dom"] = rng.integers(max(D_given // 2, 1), D_given, size=n_cells).astype(np.int64)
    # initial weights: torch default init of the same architecture (via scnym_b200's constructor on CPU)
    from scnym_b200.model import CellTypeCLF, DANN
    m = CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H0, n_layers=n_layers)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.running_mean.uniform_(-0.3, 0.3)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.uniform_(-0.2, 0.2)
    for k, v in m.state_dict().items():
        d["init/" + k] = v.detach().contiguous().numpy().copy()
    if use_dan:
        dn = DANN(m, n_domains=n_domains)
        for k, v in dn.domain_clf.state_dict().items():
            d["init/domain_clf." + k] = v.detach().numpy().copy()
    n_app = 2 + max(n_layers - 1, 0)
    widths = [H0] + [H] * (n_app - 1)
    model_st, dan_st = GU.split_state(GU.state_from(d, "init/"))
    om = O.OracleModel(model_st, n_layers=n_layers, dan_state=dan_st)
    opt = O.OracleOptimizer(om.parameters(), GU.optim_config(d))
    cfg = GU.step_config(d)
    teacher = None
    for s in range(n_steps):
        pre = f"step{s}/"
        d[pre + "lrow"] = rng.choice(n_cells, size=L, replace=False)
        d[pre + "urow"] = rng.choice(n_cells, size=U, replace=False)
        d[pre + "in_keep_L"] = (rng.random((L, G)) < 0.9).astype(np.uint8)
        if aug_pl:
            d[pre + "in_keep_U"] = (rng.random((K, U, G)) < 0.9).astype(np.uint8)
        d[pre + "perm_keys"] = rng.random(U + L).astype(np.float32)
        d[pre + "gamma_raw"] = torch.distributions.Beta(0.3, 0.3).sample((U + L,)).float().numpy()
        for i, w in enumerate(widths):
            d[pre + f"keep_U{i}"] = (rng.random((U, w)) < 0.5).astype(np.uint8)
            d[pre + f"keep_L{i}"] = (rng.random((L, w)) < 0.5).astype(np.uint8)
            d[pre + f"keep_D{i}"] = (rng.random((L + U, w)) < 0.5).astype(np.uint8)
        inp = GU.step_inputs(d, s)
        out, grads, teacher = O.train_step(om, opt, cfg, inp["Lx"], inp["Ly"], inp["Ux"], inp["rnd"], inp["Ldom"],
                                           inp["Udom"], teacher=teacher, mm_step=s)
        d[pre + "sup_loss"] = out["sup_loss"].detach().numpy()
        d[pre + "unsup_loss"] = out["unsup_loss"].detach().numpy()
        if use_dan:
            d[pre + "dan_loss"] = out["dan_loss"].detach().numpy()
            d[pre + "dan_acc"] = np.float32(out["dan_acc"])
        d[pre + "labeled_logits"] = out["labeled_logits"].detach().numpy()
        d[pre + "pseudolabels"] = out["pseudolabels"].float().numpy()
        d[pre + "confident"] = out["confident"].numpy()
        d[pre + "confidence"] = out["confidence"].numpy()
        for n, g in grads.items():
            d[pre + "grad/" + n] = g.numpy()
        for n, p in om.named_parameters():
            d[pre + "post/" + n] = p.detach().numpy().copy()
        for n, b in om.buffers().items():
            d[pre + "post/" + n] = b.numpy().copy()
    return d
