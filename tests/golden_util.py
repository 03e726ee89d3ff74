"""Helpers shared by the oracle tests (CPU) and the CUDA parity tests (GPU)."""
import os

import numpy as np
import torch

from oracle import scnym_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MIXMATCH_CASES = ["mixmatch_dan_adam", "mixmatch_dan_k2_adadelta", "mixmatch_conf_adamw",
                  "mixmatch_noconf_adam", "mixmatch_nodan_sgd"]


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def state_from(d, prefix):
    out = {}
    for k in d.files:
        if k.startswith(prefix):
            out[k[len(prefix):]] = torch.from_numpy(np.array(d[k]))
    return out


def split_state(st):
    model = {k: v for k, v in st.items() if not k.startswith("domain_clf.")}
    dan = {k: v for k, v in st.items() if k.startswith("domain_clf.")}
    return model, (dan if dan else None)


def rel_err(a, b):
    """norm-relative error used throughout (elementwise relative error
    is meaningless on near-zero activations): max|a-b| / max(|b|)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    if a.shape != b.shape:
        raise AssertionError(f"shape {tuple(a.shape)} vs {tuple(b.shape)}")
    if b.numel() == 0:
        return 0.0
    if torch.isnan(a).any() or torch.isnan(b).any():
        return float("nan") if not torch.equal(torch.isnan(a), torch.isnan(b)) else 0.0
    bmax = float(b.abs().max())
    if bmax < 1e-7:
        # analytically-zero quantities (e.g. the gradient of a Linear bias that feeds
        # BatchNorm): both sides hold only rounding noise; require |a| tiny instead.
        return 0.0 if float(a.abs().max()) < 1e-6 else float("inf")
    return float((a - b).abs().max()) / bmax


def step_config(d):
    cw = torch.from_numpy(d["class_weight"]) if "class_weight" in d.files else None
    return O.StepConfig(n_augmentations=int(d["meta_K"]), T=float(d["meta_T"]),
                        augment_pseudolabels=bool(d["meta_aug_pl"]),
                        pseudolabel_min_confidence=float(d["meta_tau"]),
                        unsup_criterion=str(d["meta_unsup_criterion"]),
                        unsup_weight=float(d["meta_unsup_w"]), use_dan=bool(d["meta_use_dan"]),
                        dan_weight=float(d["meta_dan_w"]),
                        dan_use_conf_pseudolabels=bool(d["meta_dan_conf"]),
                        dan_scale_loss_pseudoconf=bool(d["meta_dan_scale"]),
                        mean_teacher=bool(d["meta_mean_teacher"]), class_weight=cw,
                        augment=str(d["meta_augment"]) if "meta_augment" in d.files else "log1p_drop")


def optim_config(d):
    return O.OptimConfig(name=str(d["meta_opt"]), lr=float(d["meta_lr"]), weight_decay=float(d["meta_wd"]))


def n_app(d):
    return 2 + max(int(d["meta_n_layers"]) - 1, 0)


def step_inputs(d, s):
    """Dense batches + injected randomness of step `s` of a mixmatch golden case."""
    G, C = int(d["meta_G"]), int(d["meta_C"])
    pre = f"step{s}/"
    lrow, urow = d[pre + "lrow"], d[pre + "urow"]
    Lx = O.csr_to_dense(d["lab_indptr"], d["lab_indices"], d["lab_data"], G, lrow)
    Ux = O.csr_to_dense(d["unl_indptr"], d["unl_indices"], d["unl_data"], G, urow)
    Ly = torch.nn.functional.one_hot(torch.from_numpy(d["lab_y"][lrow]), num_classes=C).float()
    Ldom = Udom = None
    if int(d["meta_D_given"]):
        D = int(d["meta_D"])
        Ldom = torch.nn.functional.one_hot(torch.from_numpy(d["lab_dom"][lrow]), num_classes=D).float()
        Udom = torch.nn.functional.one_hot(torch.from_numpy(d["unl_dom"][urow]), num_classes=D).float()
    na = n_app(d)
    t = lambda k: torch.from_numpy(d[pre + k])
    rnd = O.StepRandomness(in_keep_L=t("in_keep_L"), perm_keys=t("perm_keys"), gamma_raw=t("gamma_raw"),
                           drop_keep_U=[t(f"keep_U{i}") for i in range(na)],
                           drop_keep_L=[t(f"keep_L{i}") for i in range(na)],
                           drop_keep_D=[t(f"keep_D{i}") for i in range(na)],
                           in_keep_U=t("in_keep_U") if (pre + "in_keep_U") in d.files else None,
                           pois_L=t("pois_L") if (pre + "pois_L") in d.files else None)
    return dict(Lx=Lx, Ly=Ly, Ux=Ux, Ldom=Ldom, Udom=Udom, rnd=rnd, lrow=lrow, urow=urow)


def is_pre_bn_bias(name):
    """Linear biases that feed a BatchNorm: their true gradient is exactly zero."""
    return name.startswith("embed.") and name.endswith(".bias") and int(name.split(".")[1]) % 4 == 0
