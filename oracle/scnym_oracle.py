"""CPU oracle for the scNym MixMatch(+DAN) train step and batched predict.

TEST INFRASTRUCTURE ONLY.  This file restates, in plain torch-CPU fp32 tensor
algebra on DENSE [rows, genes] batches, the algorithm of the reference hot
path (calico/scnym v0.4.0).  It is imported only by `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs, where it is the checker (or the timed CPU baseline) and never the thing
shipped.  The product path (`scnym_b200`) must never import it.

Why torch-CPU: the reference's arithmetic lives in PyTorch (third-party, pinned
`torch==2.6.*` in the reference's requirements.txt:13; not vendored under
/root/reference).  The oracle therefore uses only elementary torch-CPU tensor
ops (matmul, exp, log, sum, index) plus autograd for derivatives, and restates
every composite the reference takes from the library (BatchNorm1d, Dropout,
softmax, Adam/AdamW/Adadelta/SGD updates, GradReverse) from its published
formula, so that a disagreement with `torch.nn` itself would show up in
`tests/test_oracle.py`.

Parity pin: the reference's own tests hold no numeric golden vectors for this
path.  The oracle is pinned against outputs of the UNMODIFIED
reference run in the build container with injected randomness; those vectors
are committed under `tests/golden/` together with `make_golden.py`.

All randomness is INJECTED (never drawn here):
  * `in_keep`   Bernoulli keep-masks for `InputDropout` (dataprep.py:468-496)
  * `drop_keep` Bernoulli keep-masks for each `nn.Dropout` application
  * `perm_keys` uniform keys; MixUp `ridx = argsort(perm_keys[:n])`
                (stands in for `torch.randperm(n)`, dataprep.py:665)
  * `gamma_raw` Beta(alpha, alpha) draws (dataprep.py:671)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------
# Augmentation  (reference: scnym/dataprep.py:721-729 `log1p_drop`)
# ----------------------------------------------------------------------------
def log1p_drop(x: torch.Tensor, keep: torch.Tensor, p_drop: float = 0.1,
               counts_per_cell_after: float = 1e6) -> torch.Tensor:
    """ExpMinusOne -> InputDropout(p) -> LibrarySizeNormalize(log1p).

    dataprep.py:271 (expm1), :491 (F.dropout == x * keep / (1-p)),
    :245-253 (row sum, proportion, *1e6, log1p).  `keep` is 0/1.
    A row whose entries are all dropped yields 0/0 = NaN, as in the reference.
    """
    e = torch.expm1(x)
    noise = keep.to(x.dtype) / (1.0 - p_drop)
    e = e * noise
    size = torch.sum(e, dim=1).reshape(-1, 1)
    prop = e / size
    return torch.log1p(prop * counts_per_cell_after)


def _lib_norm_log1p(e: torch.Tensor, counts_per_cell_after: float = 1e6) -> torch.Tensor:
    """LibrarySizeNormalize(log1p=True) (dataprep.py:245-253)."""
    size = torch.sum(e, dim=1).reshape(-1, 1)
    return torch.log1p(e / size * counts_per_cell_after)


def augment_rows(x: torch.Tensor, scheme: str, keep: Optional[torch.Tensor] = None, pois: Optional[torch.Tensor] = None,
                 p_drop: float = 0.1) -> torch.Tensor:
    """`AUGMENTATION_SCHEMES[scheme]` (dataprep.py:730-754) on dense rows with every random draw INJECTED.
      log1p_drop          keep = Bernoulli keeps of InputDropout (:491: x * keep / (1-p))
      log1p_mask          keep = 0 on the genes GeneMasking zeroes (:405-465: floor(G*p) genes per row without replacement, the
                          whole batch untouched with probability 1 - p_apply => keep all ones); no rescaling
      log1p_poisson       pois = the Poisson samples of PoissonSample (:277-341: rate = expm1(x) * depth, zero rates stay 0)
      log1p_poisson_drop  pois (depth ~ U(0.1, 2) already inside the samples), then InputDropout keeps
      count_poisson       pois on raw counts, no normalisation"""
    if scheme == "log1p_drop":
        return log1p_drop(x, keep, p_drop)
    if scheme == "count_poisson":
        return torch.where(x == 0.0, torch.zeros_like(x), pois.to(x.dtype))
    e = torch.expm1(x)
    if scheme == "log1p_mask":
        e = e * keep.to(x.dtype)
    elif scheme in ("log1p_poisson", "log1p_poisson_drop"):
        e = torch.where(e == 0.0, torch.zeros_like(e), pois.to(x.dtype))     # PoissonSample casts back to float32 (:340)
        if scheme == "log1p_poisson_drop":
            e = e * (keep.to(x.dtype) / (1.0 - p_drop))
    else:
        raise ValueError(scheme)
    return _lib_norm_log1p(e)


# ----------------------------------------------------------------------------
# Model  (reference: scnym/model.py:85-283 CellTypeCLF, :286-424 GradReverse/DANN)
# ----------------------------------------------------------------------------
@dataclass
class Block:
    """Linear -> BatchNorm1d -> Dropout -> ReLU (model.py:169-182, 210-226)."""
    W: torch.Tensor          # [out, in]
    b: torch.Tensor          # [out]
    g: torch.Tensor          # BN weight [out]
    beta: torch.Tensor       # BN bias [out]
    rm: torch.Tensor         # running mean
    rv: torch.Tensor         # running var
    nbt: torch.Tensor        # num_batches_tracked (int64 scalar)
    p_drop: float = 0.5


def embed_prefixes(n_layers: int) -> List[str]:
    """state_dict prefixes of the Linear layers in `embed`, in application order.

    model.py:207 `hidden_layer * (n_layers-1)` repeats the SAME module objects,
    so for n_layers >= 3 the trailing blocks alias one set of parameters; the
    state_dict still lists them under every index (embed.8 == embed.12 == ...).
    """
    return [f"embed.{4 * i}" for i in range(2 + max(n_layers - 1, 0))]


class OracleModel:
    """Functional restatement of CellTypeCLF (+ optional DANN head)."""

    def __init__(self, state: Dict[str, torch.Tensor], n_layers: int = 2,
                 dan_state: Optional[Dict[str, torch.Tensor]] = None,
                 requires_grad: bool = False):
        self.n_layers = n_layers
        st = {k: v.detach().clone() for k, v in state.items()}
        self.blocks: List[Block] = []
        n_apply = 2 + max(n_layers - 1, 0)
        uniq = min(n_apply, 3)
        made: List[Block] = []
        for i in range(uniq):
            lin, bn = f"embed.{4 * i}", f"embed.{4 * i + 1}"
            made.append(Block(W=st[lin + ".weight"], b=st[lin + ".bias"],
                              g=st[bn + ".weight"], beta=st[bn + ".bias"],
                              rm=st[bn + ".running_mean"], rv=st[bn + ".running_var"],
                              nbt=st[bn + ".num_batches_tracked"]))
        self.unique_blocks = made
        self.blocks = [made[min(i, 2)] for i in range(n_apply)]
        self.Wc = st["classif.0.weight"]
        self.bc = st["classif.0.bias"]
        self.dan = None
        if dan_state is not None:
            ds = {k: v.detach().clone() for k, v in dan_state.items()}
            # DANN(n_layers=1): domain_clf = [Linear(H,H), ReLU, Linear(H,D)] (model.py:376-384)
            self.dan = [ds["domain_clf.0.weight"], ds["domain_clf.0.bias"],
                        ds["domain_clf.2.weight"], ds["domain_clf.2.bias"]]
        if requires_grad:
            for p in self.parameters():
                p.requires_grad_(True)

    # parameter order == reference `model.parameters()` then the DAN head group
    # (main.py:374-381, 582-587); shared modules appear once.
    def named_parameters(self) -> List[Tuple[str, torch.Tensor]]:
        out = []
        for i, blk in enumerate(self.unique_blocks):
            out += [(f"embed.{4 * i}.weight", blk.W), (f"embed.{4 * i}.bias", blk.b),
                    (f"embed.{4 * i + 1}.weight", blk.g), (f"embed.{4 * i + 1}.bias", blk.beta)]
        out += [("classif.0.weight", self.Wc), ("classif.0.bias", self.bc)]
        if self.dan is not None:
            out += [("domain_clf.0.weight", self.dan[0]), ("domain_clf.0.bias", self.dan[1]),
                    ("domain_clf.2.weight", self.dan[2]), ("domain_clf.2.bias", self.dan[3])]
        return out

    def parameters(self) -> List[torch.Tensor]:
        return [p for _, p in self.named_parameters()]

    def buffers(self) -> Dict[str, torch.Tensor]:
        out = {}
        for i, blk in enumerate(self.unique_blocks):
            out[f"embed.{4 * i + 1}.running_mean"] = blk.rm
            out[f"embed.{4 * i + 1}.running_var"] = blk.rv
            out[f"embed.{4 * i + 1}.num_batches_tracked"] = blk.nbt
        return out

    def snapshot_eval(self) -> "OracleModel":
        """`copy.deepcopy(model).eval()` (losses.py:354-359): detached copy."""
        st = {}
        for n, p in self.named_parameters():
            if not n.startswith("domain_clf"):
                st[n] = p.detach().clone()
        st.update({k: v.clone() for k, v in self.buffers().items()})
        return OracleModel(st, n_layers=self.n_layers)

    # -- forward ------------------------------------------------------------
    @staticmethod
    def _bn_train(z, blk: Block, update_running: bool = True):
        """nn.BatchNorm1d training mode: biased batch variance for the
        normalisation, unbiased for the running estimate, momentum 0.1."""
        n = z.shape[0]
        mu = z.mean(dim=0)
        var = ((z - mu) ** 2).mean(dim=0)
        zhat = (z - mu) / torch.sqrt(var + BN_EPS)
        if update_running:
            with torch.no_grad():
                blk.rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mu.detach())
                unb = var.detach() * (n / max(n - 1, 1))
                blk.rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * unb)
                blk.nbt.add_(1)
        return zhat * blk.g + blk.beta

    @staticmethod
    def _bn_eval(z, blk: Block):
        return (z - blk.rm) / torch.sqrt(blk.rv + BN_EPS) * blk.g + blk.beta

    def embed(self, x, train: bool, drop_keep: Optional[Sequence[torch.Tensor]] = None,
              pre_relu_last: bool = False, gates: Optional[Sequence[torch.Tensor]] = None,
              trace: Optional[list] = None):
        """`model.embed(x)` (model.py:210-226,273).  `drop_keep[i]` is the 0/1 keep
        mask of the i-th Dropout application (train mode only).

        Test instrumentation (never changes the arithmetic when left at None):
          trace  list that receives the detached pre-ReLU tensor of every block, so that a parity
                 test can measure how close each pre-activation is to the ReLU threshold;
          gates  per-block boolean [rows, width] tensors that REPLACE the ReLU decision
                 (`a = y * gate`): the parity tests re-run a step with the gates the device
                 took, so that a pre-activation within rounding noise of zero -- whose side of
                 the threshold is decided by summation order, on any hardware -- is counted
                 explicitly instead of widening the tolerance for every entry."""
        a = x
        for i, blk in enumerate(self.blocks):
            z = a @ blk.W.t() + blk.b
            if train:
                y = self._bn_train(z, blk)
                y = y * (drop_keep[i][: y.shape[0]].to(y.dtype) / (1.0 - blk.p_drop))
            else:
                y = self._bn_eval(z, blk)
            if trace is not None:
                trace.append(y.detach().clone())
            if pre_relu_last and i == len(self.blocks) - 1:
                return y
            a = torch.relu(y) if gates is None else y * gates[i][: y.shape[0]].to(y.dtype)
        return a

    def forward(self, x, train: bool, drop_keep=None, return_embed: bool = False, gates=None, trace=None):
        e = self.embed(x, train, drop_keep, gates=gates, trace=trace)
        logits = e @ self.Wc.t() + self.bc
        return (logits, e) if return_embed else logits

    def dann_forward(self, x, train: bool, drop_keep, grl_weight: float, gates=None, trace=None,
                     head_gate=None, head_trace=None):
        """DANN.forward (model.py:395-424)."""
        e = self.embed(x, train, drop_keep, gates=gates, trace=trace)
        r = _GradReverse.apply(e, grl_weight)
        hp = r @ self.dan[0].t() + self.dan[1]
        if head_trace is not None:
            head_trace.append(hp.detach().clone())
        h = torch.relu(hp) if head_gate is None else hp * head_gate[: hp.shape[0]].to(hp.dtype)
        return h @ self.dan[2].t() + self.dan[3], e


class _GradReverse(torch.autograd.Function):
    """model.py:286-338: identity forward, `-weight * grad` backward."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.weight = weight
        return x.view_as(x) * 1.0

    @staticmethod
    def backward(ctx, g):
        return g * -1 * ctx.weight, None


# ----------------------------------------------------------------------------
# Criteria  (reference: scnym/losses.py)
# ----------------------------------------------------------------------------
def log_softmax(z):
    m = z.max(dim=1, keepdim=True).values
    s = z - m
    return s - torch.log(torch.exp(s).sum(dim=1, keepdim=True))


def softmax(z):
    return torch.exp(log_softmax(z))


def cross_entropy(pred, label, class_weight=None, sample_weight=None, reduction="mean"):
    """losses.py:67-151 (soft-target CE with optional class / sample weights)."""
    if pred.size() != label.size():
        raise ValueError("size mismatch")
    lsm = log_softmax(pred)
    loss = -1 * torch.sum(label * lsm, dim=1)
    if sample_weight is not None:
        loss = loss * sample_weight.squeeze()
    if class_weight is not None:
        w, _ = torch.max(class_weight.reshape(1, -1) * label, dim=1)
        loss = loss * w
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def sharpen_labels(q, T: float = 0.5):
    """losses.py:529-573."""
    if T == 0.0:
        idx = torch.argmax(q, dim=1)
        return torch.nn.functional.one_hot(idx, num_classes=q.size(1))
    if T == 1.0:
        return q
    q = torch.pow(q, 1.0 / T)
    return q / q.sum(dim=1).reshape(-1, 1)


def mixup_indices(perm_keys: torch.Tensor, gamma_raw: torch.Tensor, n: int):
    """SampleMixUp (dataprep.py:665-683) with injected randomness.
    `ridx = argsort(perm_keys[:n])` (stable), `gamma = max(g, 1-g)`."""
    ridx = torch.argsort(perm_keys[:n], stable=True)
    g = gamma_raw[:n]
    g = torch.maximum(g, 1 - g).reshape(-1, 1)
    return ridx, g


def icl_weight(epoch: int, ramp_epochs: int, burn_in_epochs: int = 0,
               max_unsup_weight: float = 10.0, sigmoid: bool = False) -> float:
    """trainer.py:1327-1408 ICLWeight."""
    if type(epoch) != int:
        raise TypeError("epoch must be int")
    if epoch < burn_in_epochs:
        return 0.0
    if epoch >= ramp_epochs + burn_in_epochs:
        return max_unsup_weight
    if sigmoid:
        x = (epoch - burn_in_epochs) / ramp_epochs
        return float(math.exp(-5 * (x - 1) ** 2) * max_unsup_weight)
    step = max_unsup_weight if ramp_epochs == 0 else max_unsup_weight / ramp_epochs
    return step * (epoch - burn_in_epochs)


# ----------------------------------------------------------------------------
# Optimizers  (reference: main.py:36-41,374-396 -> torch.optim single-tensor rules)
# ----------------------------------------------------------------------------
@dataclass
class OptimConfig:
    name: str = "adadelta"      # adadelta | adam | adamw | sgd
    lr: float = 1.0
    weight_decay: float = 1e-4
    betas: Tuple[float, float] = (0.9, 0.999)
    eps: Optional[float] = None  # default: adadelta 1e-6, adam 1e-8
    rho: float = 0.9
    momentum: float = 0.9


class OracleOptimizer:
    def __init__(self, params: List[torch.Tensor], cfg: OptimConfig):
        self.params, self.cfg, self.t = params, cfg, 0
        z = lambda: [torch.zeros_like(p) for p in params]
        self.s1, self.s2 = z(), z()

    @torch.no_grad()
    def step(self, grads: List[Optional[torch.Tensor]]):
        c = self.cfg
        self.t += 1
        for p, g, s1, s2 in zip(self.params, grads, self.s1, self.s2):
            if g is None:
                continue
            if c.name == "adadelta":
                eps = 1e-6 if c.eps is None else c.eps
                g = g + c.weight_decay * p
                s1.mul_(c.rho).add_((1 - c.rho) * g * g)            # square_avg
                delta = torch.sqrt(s2 + eps) / torch.sqrt(s1 + eps) * g
                s2.mul_(c.rho).add_((1 - c.rho) * delta * delta)    # acc_delta
                p.sub_(c.lr * delta)
            elif c.name in ("adam", "adamw"):
                eps = 1e-8 if c.eps is None else c.eps
                b1, b2 = c.betas
                if c.name == "adamw":
                    p.mul_(1 - c.lr * c.weight_decay)
                else:
                    g = g + c.weight_decay * p
                s1.add_((1 - b1) * (g - s1))                        # exp_avg.lerp_
                s2.mul_(b2).add_((1 - b2) * g * g)
                bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
                denom = torch.sqrt(s2) / math.sqrt(bc2) + eps
                p.sub_((c.lr / bc1) * s1 / denom)
            elif c.name == "sgd":
                g = g + c.weight_decay * p
                if self.t == 1:
                    s1.copy_(g)
                else:
                    s1.mul_(c.momentum).add_(g)
                p.sub_(c.lr * s1)
            else:
                raise ValueError(c.name)


# ----------------------------------------------------------------------------
# One training step  (reference: trainer.py:575-671 + losses.py:739-927, 1107-1245)
# ----------------------------------------------------------------------------
@dataclass
class StepConfig:
    n_augmentations: int = 1
    T: float = 0.5
    augment_pseudolabels: bool = False
    pseudolabel_min_confidence: float = 0.0
    unsup_criterion: str = "mse"         # mse | ce
    unsup_weight: float = 1.0            # ICLWeight(epoch)
    use_dan: bool = True
    dan_weight: float = 0.1              # GRL weight = ICLWeight(epoch)
    dan_use_conf_pseudolabels: bool = False
    dan_scale_loss_pseudoconf: bool = False
    mean_teacher: bool = False
    decay_coef: float = 0.997
    p_in_drop: float = 0.1
    class_weight: Optional[torch.Tensor] = None
    augment: str = "log1p_drop"          # AUGMENTATION_SCHEMES key (main.py:176)


@dataclass
class StepRandomness:
    in_keep_L: torch.Tensor                       # [L,G] 0/1
    perm_keys: torch.Tensor                       # [U+L]
    gamma_raw: torch.Tensor                       # [U+L]
    drop_keep_U: List[torch.Tensor]               # per Dropout application, [U,H*]
    drop_keep_L: List[torch.Tensor]               # [L,H*]
    drop_keep_D: List[torch.Tensor] = field(default_factory=list)  # [L+U',H*]
    in_keep_U: Optional[torch.Tensor] = None      # [K,U,G] when augment_pseudolabels
    pois_L: Optional[torch.Tensor] = None         # [L,G] injected Poisson samples (log1p_poisson* schemes)
    pois_U: Optional[torch.Tensor] = None         # [K,U,G]


def mixmatch_dan_losses(model: OracleModel, teacher: OracleModel, cfg: StepConfig,
                        Lx, Ly, Ux, rnd: StepRandomness, Ldom=None, Udom=None,
                        gates: Optional[Dict[str, list]] = None, trace: Optional[Dict[str, list]] = None):
    """Forward part of one SemiSupervisedTrainer minibatch (trainer.py:602-656).

    Returns dict with sup/unsup/dan losses, total loss (autograd-attached), the
    labelled logits, pseudolabels and confidence mask.  Order of the three
    student forwards (and hence of BN running-stat updates) follows the
    reference: unlabeled-mixed, labeled-mixed (losses.py:880-886), DAN concat
    (losses.py:1209-1213).

    `gates` / `trace` (test instrumentation, see OracleModel.embed): dicts keyed "U", "L", "D"
    (the three student passes; per-block lists) and "Dh" (the ReLU of the domain classifier).
    """
    out = {}
    gt = (lambda k: None) if gates is None else (lambda k: gates.get(k))
    tr = (lambda k: None) if trace is None else (lambda k: trace.setdefault(k, []))
    U, L = Ux.shape[0], Lx.shape[0]
    # (1) pseudolabels -- losses.py:660-737 (no_grad, teacher in eval mode)
    with torch.no_grad():
        guesses = []
        for k in range(cfg.n_augmentations):
            xk = Ux.clone()
            if cfg.augment_pseudolabels:
                xk = augment_rows(xk, cfg.augment, rnd.in_keep_U[k] if rnd.in_keep_U is not None else None,
                                  rnd.pois_U[k] if rnd.pois_U is not None else None, cfg.p_in_drop)
            guesses.append(softmax(teacher.forward(xk, train=False)))
        qbar = torch.stack(guesses, 0).mean(0)
        conf, _ = qbar.max(dim=1)
        confident = conf >= cfg.pseudolabel_min_confidence
        q = sharpen_labels(qbar, cfg.T) if cfg.T is not None else qbar
        q = q.to(Ly.dtype)
    out["pseudolabels"], out["confident"], out["confidence"] = q, confident, conf
    # (2) augment labelled -- losses.py:801 (replaces the caller's "input")
    Lx_aug = augment_rows(Lx, cfg.augment, rnd.in_keep_L, rnd.pois_L, cfg.p_in_drop)
    # (3) MixUp over [confident unlabelled ; labelled] -- losses.py:811-846
    cidx = torch.where(confident)[0]
    uidx = torch.where(~confident)[0]
    n_conf = cidx.numel()
    cat_x = torch.cat([Ux[cidx], Lx_aug], 0)
    cat_y = torch.cat([q[cidx], Ly], 0)
    ridx, gamma = mixup_indices(rnd.perm_keys, rnd.gamma_raw, n_conf + L)
    mix_x = gamma * cat_x + (1 - gamma) * cat_x[ridx]
    mix_y = gamma * cat_y + (1 - gamma) * cat_y[ridx]
    out["ridx"], out["gamma"] = ridx, gamma
    # (4) forwards -- losses.py:854-886; unconfident rows appended un-mixed
    u_m = torch.cat([mix_x[:n_conf], Ux[uidx]], 0)
    u_y = torch.cat([mix_y[:n_conf], q[uidx]], 0)
    l_m, l_y = mix_x[n_conf:], mix_y[n_conf:]
    u_logits = model.forward(u_m, train=True, drop_keep=rnd.drop_keep_U, gates=gt("U"), trace=tr("U"))
    l_logits = model.forward(l_m, train=True, drop_keep=rnd.drop_keep_L, gates=gt("L"), trace=tr("L"))
    # (5) losses -- losses.py:896-923
    if cfg.unsup_criterion == "mse":
        per = ((softmax(u_logits) - u_y) ** 2).sum(dim=1)
    else:
        per = cross_entropy(softmax(u_logits), u_y, reduction="none")
    scale = torch.zeros_like(per)
    scale[:n_conf] += 1.0
    unsup = (per * scale).mean()
    sup = cross_entropy(l_logits, l_y, class_weight=cfg.class_weight)
    out.update(sup_loss=sup, unsup_loss=unsup, labeled_logits=l_logits,
               unlabeled_logits=u_logits, n_conf=n_conf, mixed_labels_u=u_y, mixed_labels_l=l_y)
    total = sup + cfg.unsup_weight * unsup
    # (6) domain adversary -- losses.py:1107-1245 on cat[L_aug, U_raw(conf subset)]
    if cfg.use_dan:
        if Ldom is None or bool((Ldom == -1).sum() > 0):
            src = torch.nn.functional.one_hot(torch.zeros(L).long(), num_classes=2)
        else:
            src = Ldom
        if Udom is None or bool((Udom == -1).sum() > 0):
            tgt = torch.nn.functional.one_hot(torch.ones(U).long(), num_classes=2)
        else:
            tgt = Udom
        ux = Ux
        if cfg.dan_use_conf_pseudolabels:
            ux, tgt = ux[confident], tgt[confident]
        p_conf = ux.shape[0] / max(U, 1)
        x = torch.cat([Lx_aug, ux], 0)
        dlabel = torch.cat([src, tgt], 0).to(Lx.dtype)
        hg = gt("Dh")
        dlogits, emb = model.dann_forward(x, True, rnd.drop_keep_D, cfg.dan_weight, gates=gt("D"), trace=tr("D"),
                                          head_gate=None if hg is None else hg[0], head_trace=tr("Dh"))
        dan = cross_entropy(dlogits, dlabel)
        if cfg.dan_scale_loss_pseudoconf:
            dan = dan * p_conf
        dan_acc = (dlogits.argmax(1) == dlabel.argmax(1)).float().mean()
        out.update(dan_loss=dan, domain_logits=dlogits, dan_embed=emb.detach(), dan_acc=dan_acc)
        total = total + dan
    out["loss"] = total
    return out


def train_step(model: OracleModel, opt: OracleOptimizer, cfg: StepConfig, Lx, Ly, Ux,
               rnd: StepRandomness, Ldom=None, Udom=None, teacher: Optional[OracleModel] = None,
               mm_step: int = 0, gates=None, trace=None):
    """zero_grad -> criteria -> backward -> optimizer.step (trainer.py:602-671).
    Returns (outputs dict, grads by name, teacher)."""
    params = model.parameters()
    for p in params:
        p.requires_grad_(True)
        p.grad = None
    # losses.py:337-406 teacher refresh
    if cfg.mean_teacher and teacher is not None:
        alpha = min(1 - 1 / (mm_step + 1), cfg.decay_coef)
        with torch.no_grad():
            tp = [p for n, p in teacher.named_parameters()]
            mp = [p for n, p in model.named_parameters() if not n.startswith("domain_clf")]
            for t, m in zip(tp, mp):
                t.mul_(alpha).add_(m.detach(), alpha=1 - alpha)
    else:
        teacher = model.snapshot_eval()
    out = mixmatch_dan_losses(model, teacher, cfg, Lx, Ly, Ux, rnd, Ldom, Udom, gates=gates, trace=trace)
    out["loss"].backward()
    names = [n for n, _ in model.named_parameters()]
    grads = {n: (p.grad.detach().clone() if p.grad is not None else None)
             for n, p in zip(names, params)}
    opt.step([grads[n] for n in names])
    for p in params:
        p.grad = None
    return out, grads, teacher


def supervised_step(model: OracleModel, opt: OracleOptimizer, x, y, drop_keep,
                    class_weight=None):
    """Trainer.train_epoch body (trainer.py:189-256): forward, CE, backward, step."""
    params = model.parameters()
    for p in params:
        p.requires_grad_(True)
        p.grad = None
    logits = model.forward(x, train=True, drop_keep=drop_keep)
    loss = cross_entropy(logits, y, class_weight=class_weight)
    loss.backward()
    names = [n for n, _ in model.named_parameters()]
    grads = {n: (p.grad.detach().clone() if p.grad is not None else None)
             for n, p in zip(names, params)}
    opt.step([grads[n] for n in names])
    for p in params:
        p.grad = None
    return {"loss": loss.detach(), "logits": logits.detach()}, grads


# ----------------------------------------------------------------------------
# Predict  (reference: scnym/predict.py:107-216, scnym/api.py:692-714)
# ----------------------------------------------------------------------------
@torch.no_grad()
def predict(models: Sequence[OracleModel], X: torch.Tensor, batch_size: int = 1024):
    """Ensemble-mean logits -> argmax, softmax, scores; embedding = last BN output
    (pre-ReLU, eval mode) of models[0]."""
    preds, probs, scores, embeds = [], [], [], []
    for s in range(0, X.shape[0], batch_size):
        xb = X[s:s + batch_size]
        outs = torch.stack([m.forward(xb, train=False) for m in models], 0).mean(0)
        scores.append(outs)
        preds.append(outs.argmax(1))
        probs.append(softmax(outs))
        embeds.append(models[0].embed(xb, train=False, pre_relu_last=True))
    return (torch.cat(preds), torch.cat(probs), torch.cat(scores), torch.cat(embeds))


# ----------------------------------------------------------------------------
# Gene re-indexing in front of predict  (reference: scnym/utils.py:155-233)
# ----------------------------------------------------------------------------
def build_classification_matrix(X, model_genes, sample_genes, gene_batch_size: int = 512):
    """Dense numpy restatement of `build_classification_matrix`: returns the [cells, len(model_genes)] matrix
    as a dense float64 array (the reference returns a CSR for CSR input; compare `.toarray()`).
      utils.py:188-192  identical gene lists -> the input itself
      utils.py:209-217  for every sample gene i found in the model: (model index, i), in sample order
      utils.py:222-227  N[:, model batch] = X[:, sample batch], batches of `gene_batch_size` pairs
    A sample gene that occurs twice is assigned twice, so its LAST column wins; a model gene that occurs twice
    makes `.item()` raise ValueError."""
    import numpy as np
    model_genes, sample_genes = np.asarray(model_genes), np.asarray(sample_genes)
    Xd = X.toarray() if hasattr(X, "toarray") else np.asarray(X)
    if len(model_genes) == len(sample_genes) and np.all(model_genes == sample_genes):
        return Xd
    N = np.zeros((Xd.shape[0], len(model_genes)))
    pairs = []
    for i, g in enumerate(sample_genes):
        hits = np.where(g == model_genes)[0]
        if len(hits) > 1:
            raise ValueError("ambiguous model gene %r" % (g,))
        if len(hits) == 1:
            pairs.append((int(hits[0]), i))
    for b in range(0, len(pairs), gene_batch_size):
        for m, i in pairs[b:b + gene_batch_size]:   # in order: a later pair overwrites an earlier one
            N[:, m] = Xd[:, i]
    return N


# ----------------------------------------------------------------------------
# Helpers shared by tests / golden generation
# ----------------------------------------------------------------------------
def synth_csr(n_cells: int, n_genes: int, density: float, n_classes: int, seed: int,
              zipf: bool = False):
    """Synthetic log1p(CPM) CSR matrix per SURVEY.md §8d.  Returns
    (indptr int64, indices int32, data float32, labels int64) as numpy arrays."""
    import numpy as np
    rng = np.random.default_rng(seed)
    lam = rng.gamma(0.5, 2.0, size=(n_classes, n_genes)) + 0.05
    labels = rng.integers(0, n_classes, size=n_cells)
    nnz_row = np.clip(rng.poisson(density * n_genes, size=n_cells), 1, n_genes)
    indptr = np.zeros(n_cells + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(nnz_row)
    indices = np.empty(indptr[-1], dtype=np.int32)
    data = np.empty(indptr[-1], dtype=np.float32)
    if zipf:
        pop = 1.0 / np.arange(1, n_genes + 1)
        pop /= pop.sum()
    for i in range(n_cells):
        k = nnz_row[i]
        if zipf:
            cols = np.sort(rng.choice(n_genes, size=k, replace=False, p=pop))
        else:
            cols = np.sort(rng.choice(n_genes, size=k, replace=False))
        cnt = 1 + rng.poisson(3.0 * lam[labels[i], cols])
        v = np.log1p(1e6 * cnt / cnt.sum())
        indices[indptr[i]:indptr[i + 1]] = cols
        data[indptr[i]:indptr[i + 1]] = v.astype(np.float32)
    return indptr, indices, data, labels.astype(np.int64)


def csr_to_dense(indptr, indices, data, n_genes, rows=None) -> torch.Tensor:
    import numpy as np
    rows = np.arange(len(indptr) - 1) if rows is None else np.asarray(rows)
    out = np.zeros((len(rows), n_genes), dtype=np.float32)
    for o, r in enumerate(rows):
        s, e = indptr[r], indptr[r + 1]
        out[o, indices[s:e]] = data[s:e]
    return torch.from_numpy(out)
