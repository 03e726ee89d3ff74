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


def build_engine(d, device="cuda", first_layer=None, w1_update=None, multi_stream=None, comm=None):
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
                       use_cuda_graph=False, keep_w1_grad=True,
                       augment=str(d["meta_augment"]) if "meta_augment" in d.files else "log1p_drop")
    if first_layer is not None:
        cfg.first_layer = first_layer
    if w1_update is not None:
        cfg.w1_update = w1_update
    if multi_stream is not None:
        cfg.multi_stream = multi_stream
    given = bool(int(d["meta_D_given"]))
    eng = TrainEngine(model, cfg, lab, d["lab_y"], int(d["meta_L"]), unlabeled=unl, unlabeled_batch_size=int(d["meta_U"]),
                      dann=dann, labeled_domain=d["lab_dom"] if given else None,
                      unlabeled_domain=d["unl_dom"] if given else None, mode="mixmatch", comm=comm)
    eng.set_weights(float(d["meta_unsup_w"]), float(d["meta_dan_w"]))
    return eng


def _row_keep(indptr, indices, rows, dense_keep):
    """Per-nnz values (batch order) from a dense [rows, G] array (keep masks come back as uint8, anything else as is)."""
    out = []
    for i, r in enumerate(rows):
        s, e = indptr[r], indptr[r + 1]
        out.append(dense_keep[i, indices[s:e]])
    if not out:
        return np.zeros(0, np.uint8)
    v = np.concatenate(out)
    return v.astype(np.uint8) if dense_keep.dtype == np.uint8 or dense_keep.dtype == bool else v


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
    if (pre + "pois_L") in d.files:    # injected Poisson samples of the labelled rows, at absolute batch-nnz positions
        pl = _row_keep(d["lab_indptr"], d["lab_indices"], lrow, d[pre + "pois_L"]).astype(np.float32)
        fullp = np.zeros(eng.cap, dtype=np.float32)
        fullp[u_len:u_len + pl.size] = pl
        eng.inj_pois_L = torch.from_numpy(fullp).to(dev)
        eng.inj_rowu_L = torch.zeros(eng.L, dtype=torch.float32, device=dev)   # unused once the samples are injected
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
                density=0.05, seed=0, use_dan=True, n_cells=None, augment="log1p_drop"):
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
             meta_use_dan=int(use_dan), meta_unsup_criterion="mse", meta_augment=augment, lab_indptr=lp, lab_indices=li, lab_data=ld, lab_y=ly,
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
        if augment == "log1p_mask":      # GeneMasking: floor(G * 0.1) genes per row, without replacement (applied batch)
            km = np.ones((L, G), dtype=np.uint8)
            for r in range(L):
                km[r, rng.choice(G, size=int(np.floor(G * 0.1)), replace=False)] = 0
            d[pre + "in_keep_L"] = km
        if augment in ("log1p_poisson", "log1p_poisson_drop"):
            Lx0 = O.csr_to_dense(lp, li, ld, G, d[pre + "lrow"]).numpy().astype(np.float64)
            depth = rng.random(L) * (0.1 - 2.0) + 2.0 if augment == "log1p_poisson_drop" else np.ones(L)
            d[pre + "pois_L"] = rng.poisson(np.expm1(Lx0) * depth[:, None]).astype(np.float32)
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


# ---------------------------------------------------------------------------------------------
# Gate-aware comparison (full-shape parity tests, smoke(), bench.py's data-parallel check)
# ---------------------------------------------------------------------------------------------
def oracle_step0(d, gates=None, trace=None):
    """Re-run step 0 of a case on the ORACLE from the case's initial weights (optionally with the ReLU
    decisions replaced by `gates`, optionally recording every pre-ReLU tensor into `trace`)."""
    model_st, dan_st = GU.split_state(GU.state_from(d, "init/"))
    om = O.OracleModel(model_st, n_layers=int(d["meta_n_layers"]), dan_state=dan_st)
    opt = O.OracleOptimizer(om.parameters(), GU.optim_config(d))
    inp = GU.step_inputs(d, 0)
    out, grads, _ = O.train_step(om, opt, GU.step_config(d), inp["Lx"], inp["Ly"], inp["Ux"], inp["rnd"], inp["Ldom"],
                                 inp["Udom"], teacher=None, mm_step=0, gates=gates, trace=trace)
    post = {n: p.detach().clone() for n, p in om.named_parameters()}
    post.update({n: b.clone() for n, b in om.buffers().items()})
    return out, grads, post


def engine_gates(eng):
    """The ReLU decisions the device took in the last step, in the oracle's pass layout
    ("U" / "L" / "D": per block application; "Dh": the domain classifier's hidden ReLU)."""
    torch.cuda.synchronize()
    U, nb = eng.U, eng.nb
    g = {"U": [(A[:U] > 0).cpu() for A in eng.As], "L": [(A[U:nb] > 0).cpu() for A in eng.As]}
    if eng.use_dan:
        g["D"] = [(A[nb:] > 0).cpu() for A in eng.As]
        g["Dh"] = [(eng.Hd > 0).cpu()]
    return g


def gate_flips(gates, trace, margin=1e-4):
    """Compare the device's ReLU decisions with the oracle's.  Returns (n_flipped, n_total, worst), where `worst` is the
    largest |pre-activation| (oracle value, relative to the tensor's max) among the entries decided differently.  A flip
    is legitimate only where the pre-activation is rounding noise: callers assert `worst <= margin`."""
    flipped = total = 0
    worst = 0.0
    for key, ys in trace.items():
        for i, y in enumerate(ys):
            ge = gates[key][i][: y.shape[0]]
            go = y > 0
            diff = ge != go
            total += y.numel()
            n = int(diff.sum())
            if n:
                flipped += n
                worst = max(worst, float(y[diff].abs().max()) / max(float(y.abs().max()), 1e-30))
    return flipped, total, worst


def adam_safe_mask(gref, p_init, cfg, frac=1e-3):
    """Entries on which Adam's first update lr * g' / (|g'| + eps), g' = g + wd * p (coupled weight decay), is well
    conditioned: |g'| far above eps (1e-8) and above `frac` of the tensor's largest gradient.  Near g' ~ 0 -- which weight
    decay produces wherever it cancels a small gradient -- a 1e-4 gradient difference is amplified a thousandfold; those
    entries are covered by the optimizer replay on the device's own gradient instead."""
    g = gref.double()
    eff = g + (float(cfg.weight_decay) * p_init.double() if str(cfg.name) == "adam" else 0.0)
    gmax = float(g.abs().max())
    if gmax <= 1e-7:
        return torch.zeros_like(g, dtype=torch.bool)
    return (g.abs() > max(frac * gmax, 1e-6)) & (eff.abs() > max(frac * gmax, 1e-6))


def compare_step0(eng, d, tol=1e-4, margin=1e-4, max_flip_frac=2e-5, adam_mask=1e-3):
    """Engine (already stepped once on the case's injected step 0) against the oracle, gate-aware:
      * losses, logits, pseudolabels, confident flags against the plain oracle run;
      * every ReLU decision compared; entries decided differently must be rounding noise (|pre-activation| <= margin
        relative) and rare (<= max_flip_frac of all decisions); they are COUNTED and returned;
      * gradients at `tol` against the oracle re-run with the device's decisions (identical to the plain run when
        nothing flipped);
      * the optimizer rule: post-step parameters == the oracle optimizer applied to the DEVICE gradients (all entries,
        1e-5), and post-step parameters against the oracle's own on entries whose gradient is above `adam_mask` of
        the tensor's largest (below that, Adam's lr*g/(|g|+eps) turns summation noise into O(lr) changes, in the
        reference on other hardware just the same);
      * BatchNorm buffers at `tol`.
    Returns a dict of the measured errors (asserts on failure)."""
    got = collect(eng)
    trace = {}
    out0, grads0, post0 = oracle_step0(d, trace=trace)
    rep = {}
    def chk(name, a, b, t=tol):
        e = GU.rel_err(a, b)
        rep[name] = e
        assert e <= t, f"{name}: rel err {e:.3e} > {t}"
    chk("sup_loss", got["sup_loss"], out0["sup_loss"].detach())
    chk("unsup_loss", got["unsup_loss"], out0["unsup_loss"].detach())
    if eng.use_dan:
        chk("dan_loss", got["dan_loss"], out0["dan_loss"].detach())
    chk("labeled_logits", got["labeled_logits"], out0["labeled_logits"].detach())
    chk("pseudolabels", got["pseudolabels"], out0["pseudolabels"].float())
    assert np.array_equal(got["confident"], out0["confident"].numpy())
    assert np.array_equal(got["labeled_logits"].argmax(1).numpy(), out0["labeled_logits"].detach().argmax(1).numpy())
    gates = engine_gates(eng)
    if not eng.use_dan:
        trace.pop("D", None), trace.pop("Dh", None)
    flipped, total, worst = gate_flips(gates, trace, margin)
    rep["relu_decisions"], rep["relu_flipped"], rep["relu_flip_worst_preact"] = total, flipped, worst
    assert worst <= margin, f"a ReLU decision differs where the pre-activation is not noise: {worst:.3e}"
    assert flipped <= max(max_flip_frac * total, 2), f"{flipped} of {total} ReLU decisions differ"
    grads, post = grads0, post0
    if flipped:
        _, grads, post = oracle_step0(d, gates=gates)
    worst_g = 0.0
    for n, g in got["grads"].items():
        e = GU.rel_err(g, grads[n])
        worst_g = max(worst_g, e)
        assert e <= tol, f"grad/{n}: rel err {e:.3e} > {tol}"
    rep["grad_worst"] = worst_g
    # optimizer rule pinned on the device's own gradients (every entry)
    model_st, dan_st = GU.split_state(GU.state_from(d, "init/"))
    rm = O.OracleModel(model_st, n_layers=int(d["meta_n_layers"]), dan_state=dan_st)
    ocfg = GU.optim_config(d)
    ro = O.OracleOptimizer(rm.parameters(), ocfg)
    init = {n: p.detach().clone() for n, p in rm.named_parameters()}
    ro.step([got["grads"][n].clone() for n, _ in rm.named_parameters()])
    worst_r = worst_p = 0.0
    for n, p in rm.named_parameters():
        e = GU.rel_err(got["post"][n], p.detach())
        worst_r = max(worst_r, e)
        assert e <= 1e-5, f"optimizer replay post/{n}: rel err {e:.3e}"
        m = adam_safe_mask(grads[n], init[n], ocfg, adam_mask)
        if bool(m.any()):
            a, b = got["post"][n].double()[m], post[n].double()[m]
            e = float((a - b).abs().max()) / max(float(post[n].double().abs().max()), 1e-30)
            worst_p = max(worst_p, e)
            assert e <= tol, f"post/{n} (|g| > {adam_mask} max|g|): rel err {e:.3e} > {tol}"
    rep["post_replay_worst"], rep["post_masked_worst"] = worst_r, worst_p
    for n, b in post.items():
        if n.endswith("num_batches_tracked"):
            assert int(got["post"][n]) == int(b), n
        elif "running" in n:
            chk("post/" + n, got["post"][n], b)
    return rep


# ---------------------------------------------------------------------------------------------
# Data-parallel job against the oracle on the CONCATENATED global batch (real ranks: bench.py --gpus N)
# ---------------------------------------------------------------------------------------------
def global_batch_inputs(cases):
    """Oracle inputs of the global batch that `len(cases)` ranks jointly process (rank r steps on cases[r]'s injected
    step 0): batches concatenated rank by rank, MixUp permutation block-diagonal (rank-local), dropout masks stacked
    in the oracle's row order ([U of all ranks ; L of all ranks] for MixUp, [L_aug of all ; U of all] for the DAN pass)."""
    W = len(cases)
    inp = [GU.step_inputs(c, 0) for c in cases]
    cat = lambda f: torch.cat([f(i) for i in inp], 0)
    Lx, Ly, Ux = cat(lambda i: i["Lx"]), cat(lambda i: i["Ly"]), cat(lambda i: i["Ux"])
    Ldom = cat(lambda i: i["Ldom"]) if inp[0]["Ldom"] is not None else None
    Udom = cat(lambda i: i["Udom"]) if inp[0]["Udom"] is not None else None
    r = [i["rnd"] for i in inp]
    l, u = inp[0]["Lx"].shape[0], inp[0]["Ux"].shape[0]
    n_loc, n_g = u + l, W * (u + l)

    def gidx(rank, i):  # local MixUp-concat index -> global concat index
        return rank * u + i if i < u else W * u + rank * l + (i - u)

    ridx_g, gam_g = np.zeros(n_g, dtype=np.int64), np.zeros(n_g, dtype=np.float32)
    for rk in range(W):
        ridx_l, _ = O.mixup_indices(r[rk].perm_keys, r[rk].gamma_raw, n_loc)
        for i in range(n_loc):
            ridx_g[gidx(rk, i)] = gidx(rk, int(ridx_l[i]))
            gam_g[gidx(rk, i)] = float(r[rk].gamma_raw[i])
    keys_g = np.zeros(n_g, dtype=np.float32)
    # argsort(keys_g)[j] must equal ridx_g[j]: rank the global rows so that sorted position j holds row ridx_g[j]
    keys_g[ridx_g] = (np.arange(n_g, dtype=np.float64) / n_g).astype(np.float32)
    na = len(r[0].drop_keep_U)
    rnd_g = O.StepRandomness(
        in_keep_L=torch.cat([x.in_keep_L for x in r]), perm_keys=torch.from_numpy(keys_g), gamma_raw=torch.from_numpy(gam_g),
        drop_keep_U=[torch.cat([x.drop_keep_U[a] for x in r]) for a in range(na)],
        drop_keep_L=[torch.cat([x.drop_keep_L[a] for x in r]) for a in range(na)],
        drop_keep_D=[torch.cat([x.drop_keep_D[a][:l] for x in r] + [x.drop_keep_D[a][l:] for x in r]) for a in range(na)])
    return dict(Lx=Lx, Ly=Ly, Ux=Ux, rnd=rnd_g, Ldom=Ldom, Udom=Udom)


def oracle_global_step(cases, gates=None, trace=None):
    g = global_batch_inputs(cases)
    model_st, dan_st = GU.split_state(GU.state_from(cases[0], "init/"))
    om = O.OracleModel(model_st, n_layers=int(cases[0]["meta_n_layers"]), dan_state=dan_st)
    opt = O.OracleOptimizer(om.parameters(), GU.optim_config(cases[0]))
    out, grads, _ = O.train_step(om, opt, GU.step_config(cases[0]), g["Lx"], g["Ly"], g["Ux"], g["rnd"], g["Ldom"], g["Udom"],
                                 gates=gates, trace=trace)
    post = {n: p.detach().clone() for n, p in om.named_parameters()}
    post.update({n: b.clone() for n, b in om.buffers().items()})
    return out, grads, post


DP_CASE = dict(G=2048, L=256, U=256, C=16, H0=256, H=256, n_layers=2, D_given=8, opt_name="adam", lr=1e-3, wd=1e-4, n_steps=1,
               density=0.05, n_cells=512)


def dp_compare(device, rank, world, comm, tol=1e-4, margin=1e-4):
    """One injected step of the production data-parallel path on `world` REAL ranks (torch.distributed initialised, one GPU
    each); rank 0 compares the job with the oracle on the concatenated global batch: losses (mean over ranks), averaged
    gradients, post-step parameters (identical on every rank), BatchNorm buffers at `tol`; ReLU decisions one by one
    (see compare_step0).  Returns a report dict on rank 0 (None elsewhere); raises AssertionError on a mismatch."""
    import torch.distributed as dist
    cases = [oracle_case(seed=500 + r, **DP_CASE) for r in (range(world) if rank == 0 else [0, rank])]
    d = cases[rank] if rank == 0 else cases[1]
    for k in list(d.keys()):
        if k.startswith("init/"):
            d[k] = cases[0][k]                          # identical replicas
    if rank == 0:
        for c in cases[1:]:
            for k in list(c.keys()):
                if k.startswith("init/"):
                    c[k] = cases[0][k]
    eng = build_engine(d, device=device, comm=comm)
    assert eng.comm is not None and eng.tc_first and eng.w1_tc, "the production data-parallel kernels were not selected"
    inject_step(eng, d, 0)
    eng.step()
    torch.cuda.synchronize()
    got = collect(eng)
    U, nb, l = eng.U, eng.nb, eng.L
    # ---- gather what rank 0 needs ----
    loss_t = torch.tensor([got["sup_loss"], got["unsup_loss"], got["dan_loss"]], device=device, dtype=torch.float64)
    dist.all_reduce(loss_t, op=dist.ReduceOp.SUM)
    loss_t /= world
    names = list(got["grads"].keys())
    avg = {}
    inv_w = torch.tensor(1.0 / world, dtype=torch.float32, device=device)
    for n in names:   # summed in rank order, then scaled: the exchange kernel's arithmetic, bit for bit
        g = got["grads"][n].to(device).contiguous()
        parts_g = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(parts_g, g)
        acc = torch.zeros_like(g)
        for q in range(world):
            acc += parts_g[q]
        avg[n] = (acc * inv_w).cpu()
    parts = [(A > 0).to(torch.uint8).reshape(-1) for A in eng.As] + [(eng.Hd > 0).to(torch.uint8).reshape(-1)]
    flat = torch.cat(parts)
    allg = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(allg, flat)
    flat_p = torch.cat([got["post"][n].reshape(-1).float() for n in sorted(got["post"]) if "num_batches" not in n]).to(device)
    hi, lo = flat_p.clone(), flat_p.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    identical = bool(torch.equal(hi, lo))
    if rank != 0:
        return None
    assert identical, "post-step parameters / BatchNorm buffers differ between ranks"
    # ---- global gates in the oracle's row order ----
    sizes = [A.numel() for A in eng.As] + [eng.Hd.numel()]
    gates = {"U": [], "L": [], "D": [], "Dh": []}
    off = 0
    for a, A in enumerate(eng.As):
        w = A.shape[1]
        per = [allg[r][off:off + sizes[a]].view(-1, w).bool().cpu() for r in range(world)]
        gates["U"].append(torch.cat([x[:U] for x in per]))
        gates["L"].append(torch.cat([x[U:nb] for x in per]))
        gates["D"].append(torch.cat([x[nb:nb + l] for x in per] + [x[nb + l:] for x in per]))
        off += sizes[a]
    per = [allg[r][off:off + sizes[-1]].view(-1, eng.Hd.shape[1]).bool().cpu() for r in range(world)]
    gates["Dh"].append(torch.cat([x[:l] for x in per] + [x[l:] for x in per]))
    trace = {}
    out0, grads, post = oracle_global_step(cases, trace=trace)
    flipped, total, worst = gate_flips(gates, trace, margin)
    assert worst <= margin, f"a ReLU decision differs where the pre-activation is not noise: {worst:.3e}"
    assert flipped <= max(2e-5 * total, 2), f"{flipped} of {total} ReLU decisions differ"
    if flipped:
        _, grads, post = oracle_global_step(cases, gates=gates)
    rep = {"world": world, "relu_decisions": total, "relu_flipped": flipped, "shape": {k: DP_CASE[k] for k in ("G", "L", "U", "H0", "H", "D_given")}}
    for i, k in enumerate(("sup_loss", "unsup_loss", "dan_loss")):
        e = GU.rel_err(loss_t[i].cpu(), out0[k].detach())
        rep[k + "_err"] = e
        assert e <= tol, f"{k}: {e:.3e}"
    worst_g = 0.0
    for n in names:
        e = GU.rel_err(avg[n], grads[n])
        worst_g = max(worst_g, e)
        assert e <= tol, f"grad/{n}: rel err {e:.3e}"
    # the update rule pinned on the job's own averaged gradient, every entry (the exchange kernel's optimizer)
    d0 = cases[0]
    model_st, dan_st = GU.split_state(GU.state_from(d0, "init/"))
    rm = O.OracleModel(model_st, n_layers=int(d0["meta_n_layers"]), dan_state=dan_st)
    ocfg = GU.optim_config(d0)
    ro = O.OracleOptimizer(rm.parameters(), ocfg)
    init = {n: p.detach().clone() for n, p in rm.named_parameters()}
    ro.step([avg[n].clone() for n, _ in rm.named_parameters()])
    worst_r = 0.0
    for n, p in rm.named_parameters():
        e = GU.rel_err(got["post"][n], p.detach())
        worst_r = max(worst_r, e)
        assert e <= 1e-5, f"optimizer replay post/{n}: rel err {e:.3e}"
    rep["post_replay_worst"] = worst_r
    worst_p = 0.0
    for n, p in post.items():
        if n.endswith("num_batches_tracked"):
            assert int(got["post"][n]) == int(p), n
            continue
        if "running" in n:
            e = GU.rel_err(got["post"][n], p)
            assert e <= tol, f"post/{n}: {e:.3e}"
            continue
        if n not in init:
            continue
        # (seen at 4 ranks on the domain classifier, whose averaged gradient is ~1e-5: weight decay cancels it in places)
        m = adam_safe_mask(grads[n], init[n], ocfg)
        if not bool(m.any()):
            continue
        e = float((got["post"][n].double()[m] - p.double()[m]).abs().max()) / max(float(p.double().abs().max()), 1e-30)
        worst_p = max(worst_p, e)
        assert e <= tol, f"post/{n}: {e:.3e}"
    rep.update({"grad_worst": worst_g, "post_masked_worst": worst_p, "replicas_identical": identical, "parity": "green"})
    return rep
