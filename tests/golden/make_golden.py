"""Generate golden vectors by running the UNMODIFIED reference (/root/reference)
with injected randomness.  Run in the build container only:

    python tests/golden/make_golden.py

Writes `tests/golden/*.npz`.  The GPU box has no /root/reference; tests there
read only the committed .npz files.

What is injected (and how, without editing reference code):
  * `nn.Dropout` modules of `CellTypeCLF.embed` are swapped for `QueueDropout`
    (no parameters, so `state_dict` is unchanged), which multiplies by a queued
    0/1 keep-mask divided by (1-p) -- the arithmetic of `torch.nn.Dropout`.
  * the `InputDropout` stage of `AUGMENTATION_SCHEMES["log1p_drop"]` is rebuilt
    from the reference's own `ExpMinusOne` / `LibrarySizeNormalize` classes with
    a queued keep-mask in place of `F.dropout`.
  * `torch.randperm` is patched (only while the criterion runs) to return
    `argsort(perm_keys[:n])`, and `SampleMixUp.beta` is replaced by an object
    whose `.sample((n,))` returns `gamma_raw[:n]`.
Everything else -- MixMatchLoss, DANLoss, DANN, GradReverse, cross_entropy,
sharpen_labels, SampleMixUp, CellTypeCLF, torch.optim -- is the reference's own
code path.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from oracle.scnym_oracle import synth_csr, csr_to_dense  # noqa: E402

ref = load_reference()
from torchvision import transforms  # noqa: E402


class QueueDropout(nn.Module):
    def __init__(self, p, queue):
        super().__init__()
        self.p, self.queue = p, queue

    def forward(self, x):
        if not self.training:
            return x
        keep = self.queue.pop(0)
        assert keep.shape == x.shape, (keep.shape, x.shape)
        return x * (keep.to(x.dtype) / (1.0 - self.p))


class QueueInputDropout(object):
    def __init__(self, p_drop, queue):
        self.p_drop, self.queue = p_drop, queue

    def __call__(self, sample):
        keep = self.queue.pop(0)
        assert keep.shape == sample["input"].shape
        sample["input"] = sample["input"] * (keep.to(sample["input"].dtype) / (1.0 - self.p_drop))
        return sample


class FakeBeta(object):
    def __init__(self, gamma_raw):
        self.g = gamma_raw

    def sample(self, shape):
        return self.g[: shape[0]].clone()


class patched_randperm(object):
    def __init__(self, keys):
        self.keys = keys

    def __enter__(self):
        self.orig = torch.randperm
        keys = self.keys
        torch.randperm = lambda n, *a, **k: torch.argsort(keys[:n], stable=True)

    def __exit__(self, *a):
        torch.randperm = self.orig


def patch_dropouts(model, queue):
    mapping = {}
    for i, m in enumerate(model.embed):
        if isinstance(m, nn.Dropout):
            if id(m) not in mapping:
                mapping[id(m)] = QueueDropout(m.p, queue)
            model.embed[i] = mapping[id(m)]


def sd_np(sd, prefix=""):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def make_optimizer(name, params, lr, wd):
    cls = ref.main.OPTIMIZERS[name]
    if name == "sgd":
        return cls(params, lr=lr, weight_decay=wd, momentum=0.9)
    return cls(params, lr=lr, weight_decay=wd)


def run_mixmatch_case(name, *, G, L, U, C, H0, H, n_layers, K, T, aug_pl, tau, D_given,
                      opt_name, lr, wd, unsup_w, dan_w, n_steps, mean_teacher=False,
                      dan_conf=False, dan_scale=False, density=0.08, seed=0, use_dan=True,
                      class_weight=None, unsup_criterion="mse"):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed + 100)
    n_cells = 4 * max(L, U)
    lp, li, ld, ly = synth_csr(n_cells, G, density, C, seed=seed)
    up, ui, ud, _ = synth_csr(n_cells, G, density, C, seed=seed + 1)

    model = ref.model.CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H0,
                                  n_layers=n_layers)
    # non-trivial BN state so eval-mode teacher is exercised
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.running_mean.uniform_(-0.3, 0.3)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    drop_q, in_q = [], []
    patch_dropouts(model, drop_q)
    augment = transforms.Compose([ref.dataprep.ExpMinusOne(), QueueInputDropout(0.1, in_q),
                                  ref.dataprep.LibrarySizeNormalize(log1p=True)])
    if class_weight is not None:
        from functools import partial
        sup_crit = partial(ref.losses.cross_entropy, class_weight=torch.tensor(class_weight).float())
    else:
        sup_crit = ref.losses.cross_entropy
    if unsup_criterion == "mse":
        unsup_crit = nn.MSELoss(reduction="none")
    else:
        from functools import partial
        unsup_crit = partial(ref.losses.cross_entropy, reduction="none")
    mm = ref.losses.MixMatchLoss(alpha=0.3, unsup_criterion=unsup_crit, sup_criterion=sup_crit,
                                 decay_coef=0.997, mean_teacher=mean_teacher, augment=augment,
                                 n_augmentations=K, T=T, augment_pseudolabels=aug_pl,
                                 pseudolabel_min_confidence=tau)
    n_domains = D_given if D_given else 2
    opt = make_optimizer(opt_name, model.parameters(), lr, wd)
    dan = None
    if use_dan:
        dan = ref.losses.DANLoss(model=model, dan_criterion=ref.losses.cross_entropy,
                                 use_conf_pseudolabels=dan_conf, scale_loss_pseudoconf=dan_scale,
                                 n_domains=n_domains)
        opt.add_param_group({"params": dan.dann.domain_clf.parameters(), "name": "domain_classifier"})
    model.train(True)

    out = {"meta_G": G, "meta_L": L, "meta_U": U, "meta_C": C, "meta_H0": H0, "meta_H": H,
           "meta_n_layers": n_layers, "meta_K": K, "meta_T": T, "meta_aug_pl": int(aug_pl),
           "meta_tau": tau, "meta_D": n_domains, "meta_D_given": int(bool(D_given)),
           "meta_opt": opt_name, "meta_lr": lr, "meta_wd": wd, "meta_unsup_w": unsup_w,
           "meta_dan_w": dan_w, "meta_n_steps": n_steps, "meta_mean_teacher": int(mean_teacher),
           "meta_dan_conf": int(dan_conf), "meta_dan_scale": int(dan_scale),
           "meta_use_dan": int(use_dan), "meta_unsup_criterion": unsup_criterion,
           "lab_indptr": lp, "lab_indices": li, "lab_data": ld, "lab_y": ly,
           "unl_indptr": up, "unl_indices": ui, "unl_data": ud}
    if class_weight is not None:
        out["class_weight"] = np.asarray(class_weight, dtype=np.float32)
    if D_given:
        ldom = rng.integers(0, max(D_given // 2, 1), size=n_cells)
        udom = rng.integers(max(D_given // 2, 1), D_given, size=n_cells)
        out["lab_dom"], out["unl_dom"] = ldom.astype(np.int64), udom.astype(np.int64)
    out.update(sd_np(model.state_dict(), "init/"))
    if dan is not None:
        out.update(sd_np(dan.dann.domain_clf.state_dict(), "init/domain_clf."))

    n_app = 2 + max(n_layers - 1, 0)
    widths = [H0] + [H] * (n_app - 1)
    for s in range(n_steps):
        lrow = rng.choice(n_cells, size=L, replace=False)
        urow = rng.choice(n_cells, size=U, replace=False)
        Lx = csr_to_dense(lp, li, ld, G, lrow)
        Ux = csr_to_dense(up, ui, ud, G, urow)
        Ly = torch.nn.functional.one_hot(torch.from_numpy(ly[lrow]), num_classes=C).float()
        in_keep_L = torch.from_numpy((rng.random((L, G)) < 0.9).astype(np.uint8))
        in_keep_U = torch.from_numpy((rng.random((K, U, G)) < 0.9).astype(np.uint8))
        perm_keys = torch.from_numpy(rng.random(U + L).astype(np.float32))
        gamma_raw = torch.distributions.Beta(0.3, 0.3).sample((U + L,)).float()
        dk = lambda n: [torch.from_numpy((rng.random((n, w)) < 0.5).astype(np.uint8)) for w in widths]
        keep_U, keep_L, keep_D = dk(U), dk(L), dk(L + U)
        pre = f"step{s}/"
        out.update({pre + "lrow": lrow, pre + "urow": urow, pre + "in_keep_L": in_keep_L.numpy(),
                    pre + "perm_keys": perm_keys.numpy(), pre + "gamma_raw": gamma_raw.numpy()})
        if aug_pl:
            out[pre + "in_keep_U"] = in_keep_U.numpy()
        for i in range(n_app):
            out[pre + f"keep_U{i}"] = keep_U[i].numpy()
            out[pre + f"keep_L{i}"] = keep_L[i].numpy()
            out[pre + f"keep_D{i}"] = keep_D[i].numpy()

        data = {"input": Lx.clone(), "output": Ly.clone()}
        unsup_data = {"input": Ux.clone(), "output": torch.zeros(U)}
        if D_given:
            data["domain"] = torch.nn.functional.one_hot(torch.from_numpy(out["lab_dom"][lrow]), num_classes=D_given).float()
            unsup_data["domain"] = torch.nn.functional.one_hot(torch.from_numpy(out["unl_dom"][urow]), num_classes=D_given).float()
        else:
            data["domain"] = torch.zeros(L, C) - 1
            unsup_data["domain"] = torch.zeros(U, C) - 1
        # ---- body of SemiSupervisedTrainer.train_epoch (trainer.py:596-671) ----
        data["input"].requires_grad = True
        unsup_data["input"].requires_grad = True
        opt.zero_grad()
        if aug_pl:
            in_q.extend([in_keep_U[k] for k in range(K)])
        in_q.append(in_keep_L)
        drop_q.extend(keep_U + keep_L)
        mm.mixup_op.beta = FakeBeta(gamma_raw)
        with patched_randperm(perm_keys):
            sup_loss, unsup_loss, sup_outputs = mm(model=model, labeled_sample=data, unlabeled_sample=unsup_data)
        conf_bool = mm.running_confidence_scores[-1][0]
        conf_score = mm.running_confidence_scores[-1][1]
        if dan is not None:
            n_dan_rows = L + (int(conf_bool.sum()) if dan_conf else U)
            drop_q.extend([k[:n_dan_rows] for k in keep_D])
            dan_loss = dan(labeled_sample=data, unlabeled_sample=unsup_data, weight=dan_w,
                           pseudolabel_confidence=conf_bool)
        else:
            dan_loss = torch.zeros(1)
        loss = sup_loss + (unsup_w * unsup_loss) + dan_loss
        loss.backward()
        opt.step()
        assert len(drop_q) == 0 and len(in_q) == 0, (len(drop_q), len(in_q))
        # ---- record ----
        out[pre + "sup_loss"] = sup_loss.detach().numpy()
        out[pre + "unsup_loss"] = unsup_loss.detach().numpy()
        out[pre + "dan_loss"] = dan_loss.detach().numpy().reshape(())
        out[pre + "loss"] = loss.detach().numpy().reshape(())
        out[pre + "labeled_logits"] = sup_outputs.detach().numpy()
        out[pre + "pseudolabels"] = unsup_data["output"].detach().numpy()
        out[pre + "confident"] = conf_bool.numpy()
        out[pre + "confidence"] = conf_score.numpy()
        if dan is not None:
            out[pre + "dan_acc"] = np.float32(dan.dan_acc)
        for n, p in model.named_parameters():
            out[pre + "grad/" + n] = p.grad.detach().numpy().copy()
        if dan is not None:
            for n, p in dan.dann.domain_clf.named_parameters():
                out[pre + "grad/domain_clf." + n] = p.grad.detach().numpy().copy()
            out.update(sd_np(dan.dann.domain_clf.state_dict(), pre + "post/domain_clf."))
        out.update(sd_np(model.state_dict(), pre + "post/"))
        print(f"{name} step{s}: sup={sup_loss.item():.6f} unsup={unsup_loss.item():.6f} "
              f"dan={float(dan_loss):.6f} n_conf={int(conf_bool.sum())}/{U}")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def run_supervised_case(name, *, G, B, C, H0, H, n_layers, opt_name, lr, wd, n_steps, seed=3,
                        density=0.08):
    """Trainer.train_epoch body (trainer.py:189-256), config-1 style."""
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed + 100)
    n_cells = 4 * B
    lp, li, ld, ly = synth_csr(n_cells, G, density, C, seed=seed)
    model = ref.model.CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H0, n_layers=n_layers)
    drop_q = []
    patch_dropouts(model, drop_q)
    opt = make_optimizer(opt_name, model.parameters(), lr, wd)
    model.train(True)
    out = {"meta_G": G, "meta_B": B, "meta_C": C, "meta_H0": H0, "meta_H": H, "meta_n_layers": n_layers,
           "meta_opt": opt_name, "meta_lr": lr, "meta_wd": wd, "meta_n_steps": n_steps,
           "lab_indptr": lp, "lab_indices": li, "lab_data": ld, "lab_y": ly}
    out.update(sd_np(model.state_dict(), "init/"))
    n_app = 2 + max(n_layers - 1, 0)
    widths = [H0] + [H] * (n_app - 1)
    for s in range(n_steps):
        rows = rng.choice(n_cells, size=B, replace=False)
        x = csr_to_dense(lp, li, ld, G, rows)
        y = torch.nn.functional.one_hot(torch.from_numpy(ly[rows]), num_classes=C).float()
        keep = [torch.from_numpy((rng.random((B, w)) < 0.5).astype(np.uint8)) for w in widths]
        pre = f"step{s}/"
        out[pre + "rows"] = rows
        for i in range(n_app):
            out[pre + f"keep{i}"] = keep[i].numpy()
        x.requires_grad_()
        opt.zero_grad()
        drop_q.extend(keep)
        logits = model(x)
        loss = ref.losses.cross_entropy(logits, y)
        loss.backward()
        opt.step()
        out[pre + "loss"] = loss.detach().numpy()
        out[pre + "logits"] = logits.detach().numpy()
        for n, p in model.named_parameters():
            out[pre + "grad/" + n] = p.grad.detach().numpy().copy()
        out.update(sd_np(model.state_dict(), pre + "post/"))
        print(f"{name} step{s}: loss={loss.item():.6f}")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def run_predict_case(name, *, G, N, C, H0, H, n_layers, n_models, seed=5, density=0.08):
    """Predicter.predict (predict.py:107-216) + the api.py:692-708 embedding pass."""
    import tempfile
    torch.manual_seed(seed)
    ip, ii, dd, _ = synth_csr(N, G, density, C, seed=seed)
    from scipy import sparse
    X = sparse.csr_matrix((dd, ii, ip), shape=(N, G))
    out = {"meta_G": G, "meta_N": N, "meta_C": C, "meta_H0": H0, "meta_H": H, "meta_n_layers": n_layers,
           "meta_n_models": n_models, "indptr": ip, "indices": ii, "data": dd}
    paths = []
    tmp = tempfile.mkdtemp()
    for m in range(n_models):
        model = ref.model.CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H0, n_layers=n_layers)
        with torch.no_grad():
            for mod in model.modules():
                if isinstance(mod, nn.BatchNorm1d):
                    mod.running_mean.uniform_(-0.3, 0.3)
                    mod.running_var.uniform_(0.5, 1.5)
                    mod.weight.uniform_(0.5, 1.5)
                    mod.bias.uniform_(-0.2, 0.2)
        p = os.path.join(tmp, f"w{m}.pkl")
        torch.save(model.state_dict(), p)
        paths.append(p)
        out.update(sd_np(model.state_dict(), f"model{m}/"))
    P = ref.predict.Predicter(model_weights=paths, n_genes=G, n_cell_types=C, labels=None,
                              n_hidden=H, n_hidden_init=H0, n_layers=n_layers)
    pred, names, prob = P.predict(X, output="prob", batch_size=64)
    _, _, score = P.predict(X, output="score", batch_size=64)
    model = P.models[0]
    lz_02 = torch.nn.Sequential(*list(list(model.modules())[0].children())[1][:-1])  # api.py:700
    with torch.no_grad():
        Z = lz_02(torch.from_numpy(np.asarray(X.todense(), dtype=np.float32)))
    out.update(pred=pred, prob=prob, score=score, embed=Z.numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: pred hist {np.bincount(pred, minlength=C)}")


def run_small_functions(name):
    """sharpen_labels, ICLWeight, cross_entropy (class weights), SampleMixUp, log1p_drop."""
    torch.manual_seed(11)
    out = {}
    q = torch.softmax(torch.randn(12, 7), dim=1)
    for T in (0.0, 0.5, 1.0, 2.0):
        out[f"sharpen_T{T}"] = ref.losses.sharpen_labels(q.clone(), T=T).numpy()
    out["sharpen_q"] = q.numpy()
    epochs = np.arange(0, 40)
    for i, (ramp, burn, mx, sig) in enumerate([(20, 0, 0.1, True), (100, 0, 1.0, False), (10, 5, 2.0, False), (0, 3, 1.0, True)]):
        w = ref.trainer.ICLWeight(ramp_epochs=ramp, burn_in_epochs=burn, max_unsup_weight=mx, sigmoid=sig)
        out[f"icl{i}_cfg"] = np.array([ramp, burn, mx, float(sig)])
        out[f"icl{i}"] = np.array([w(int(e)) for e in epochs], dtype=np.float64)
    z = torch.randn(9, 5)
    y = torch.softmax(torch.randn(9, 5) * 3, dim=1)
    cw = torch.tensor([1.0, 2.5, 1.2, 4.0, 1.0])
    out.update(ce_z=z.numpy(), ce_y=y.numpy(), ce_cw=cw.numpy())
    for red in ("mean", "sum", "none"):
        out[f"ce_{red}"] = ref.losses.cross_entropy(z, y, reduction=red).numpy()
        out[f"ce_cw_{red}"] = ref.losses.cross_entropy(z, y, class_weight=cw, reduction=red).numpy()
    # log1p_drop through the reference's own classes with an injected mask
    ip, ii, dd, _ = synth_csr(6, 200, 0.1, 3, seed=9)
    x = csr_to_dense(ip, ii, dd, 200)
    keep = (torch.rand(6, 200) < 0.9)
    qd = [keep]
    aug = transforms.Compose([ref.dataprep.ExpMinusOne(), QueueInputDropout(0.1, qd),
                              ref.dataprep.LibrarySizeNormalize(log1p=True)])
    out.update(aug_x=x.numpy(), aug_keep=keep.numpy().astype(np.uint8),
               aug_out=aug({"input": x.clone(), "output": torch.zeros(1)})["input"].numpy())
    # SampleMixUp with injected perm / gamma
    mix = ref.dataprep.SampleMixUp(alpha=0.3, keep_dominant_obs=True)
    keys = torch.rand(10)
    graw = torch.distributions.Beta(0.3, 0.3).sample((10,))
    mix.beta = FakeBeta(graw)
    a, b = torch.randn(10, 4), torch.softmax(torch.randn(10, 3), 1)
    with patched_randperm(keys):
        s = mix({"input": a.clone(), "output": b.clone()})
    out.update(mix_keys=keys.numpy(), mix_graw=graw.numpy(), mix_a=a.numpy(), mix_b=b.numpy(),
               mix_in=s["input"].numpy(), mix_out=s["output"].numpy(), mix_ridx=s["random_idx"].numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok")


if __name__ == "__main__":
    run_small_functions("small_functions")
    # A: api default config shape (K=1, no pseudolabel augmentation, inferred 2 domains, Adam)
    run_mixmatch_case("mixmatch_dan_adam", G=384, L=24, U=24, C=5, H0=64, H=48, n_layers=2, K=1, T=0.5,
                      aug_pl=False, tau=0.0, D_given=0, opt_name="adam", lr=1e-3, wd=1e-4,
                      unsup_w=0.7, dan_w=0.1, n_steps=2, seed=0)
    # B: K=2 augmented views, 3 given domains, Adadelta (api default optimizer), shared third block
    run_mixmatch_case("mixmatch_dan_k2_adadelta", G=320, L=16, U=20, C=4, H0=32, H=40, n_layers=3, K=2, T=0.5,
                      aug_pl=True, tau=0.0, D_given=3, opt_name="adadelta", lr=1.0, wd=1e-4,
                      unsup_w=1.0, dan_w=0.05, n_steps=2, seed=1)
    # C: new_identity_discovery style: confidence threshold + DAN on confident rows, AdamW, mean teacher
    run_mixmatch_case("mixmatch_conf_adamw", G=256, L=20, U=28, C=3, H0=48, H=32, n_layers=2, K=1, T=0.5,
                      aug_pl=False, tau=0.37, D_given=0, opt_name="adamw", lr=2e-3, wd=1e-2,
                      unsup_w=1.0, dan_w=0.1, n_steps=3, mean_teacher=True, dan_conf=True,
                      dan_scale=True, seed=2)
    # C2: edge case -- threshold nobody passes (n_conf = 0; empty MixUp pool on the unlabeled side)
    run_mixmatch_case("mixmatch_noconf_adam", G=128, L=12, U=10, C=3, H0=32, H=32, n_layers=2, K=1, T=0.5,
                      aug_pl=False, tau=0.99, D_given=0, opt_name="adam", lr=1e-3, wd=1e-4,
                      unsup_w=1.0, dan_w=0.1, n_steps=1, dan_conf=True, seed=6)
    # D: MixMatch without adversary, T=0 sharpening, SGD, class weights
    run_mixmatch_case("mixmatch_nodan_sgd", G=256, L=16, U=16, C=4, H0=32, H=32, n_layers=2, K=1, T=0.0,
                      aug_pl=False, tau=0.0, D_given=0, opt_name="sgd", lr=0.05, wd=1e-4,
                      unsup_w=0.5, dan_w=0.0, n_steps=2, use_dan=False, seed=4,
                      class_weight=[1.0, 2.0, 1.5, 3.0])
    # E: supervised Trainer (BASELINE config 1 shape, shrunk)
    run_supervised_case("supervised_adadelta", G=300, B=32, C=5, H0=64, H=64, n_layers=2,
                        opt_name="adadelta", lr=1.0, wd=1e-4, n_steps=3)
    # F: predict (2-model ensemble) + embedding pass
    run_predict_case("predict_ensemble", G=300, N=150, C=6, H0=64, H=48, n_layers=2, n_models=2)
