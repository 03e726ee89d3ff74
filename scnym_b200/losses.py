"""Criteria of the scNym hot path (reference surface: scnym/losses.py).

Two ways these classes execute:
  * called directly (`MixMatchLoss(...)(model, labeled_sample, unlabeled_sample)`,
    `DANLoss(...)(labeled_sample, unlabeled_sample, weight)`), they follow the reference's
    sequence on dense CUDA tensors or CSR batches through the autograd functions of
    `scnym_b200.functional` (kernel per op, `loss.backward()` works, any torch optimizer);
  * handed to `scnym_b200.trainer.SemiSupervisedTrainer` together with device loaders, they act
    as the configuration of the fused `TrainEngine` step, which computes the same quantities in
    one CUDA graph and feeds `running_confidence_scores`, `step`, `dan_acc`, ... back.
"""
import copy
import logging
from typing import Callable, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import functional as Fn
from ._lib import call, ptr
from .dataprep import SampleMixUp, CSRBatch
from .model import CellTypeCLF, DANN

logger = logging.getLogger(__name__)


def get_class_weight(y: np.ndarray) -> np.ndarray:
    """Relative class weights, inverse to class frequency, min-normalised (losses.py:37-64)."""
    _, class_counts = np.unique(y, return_counts=True)
    class_weight = 1.0 / (class_counts / len(y))
    return class_weight / class_weight.min()


class _SoftCEFn(torch.autograd.Function):
    """sum/mean/none-reduced soft-target CE with the fused dLogits kernel (scnb_soft_ce_fwd_bwd)."""

    @staticmethod
    def forward(ctx, pred, label, class_weight, reduction):
        pred = Fn._f32c(pred)
        label = Fn._f32c(label.to(pred.device))
        n, C = pred.shape
        cw = None if class_weight is None else Fn._f32c(class_weight.to(pred.device))
        dz = torch.empty_like(pred)
        out = torch.zeros(2, device=pred.device, dtype=torch.float32)
        row = torch.empty(n, device=pred.device, dtype=torch.float32) if reduction == "none" else None
        # unit-scale gradient (n_norm = 1): the reduction's 1/n and the upstream gradient are applied in backward
        call("scnb_soft_ce_fwd_bwd", ptr(pred), n, C, None, ptr(label), None, None, None, 0, ptr(cw), None, 1, 1.0, None, ptr(dz),
             ptr(out), 0, ptr(row), Fn._stream())
        ctx.save_for_backward(dz)
        ctx.reduction, ctx.n = reduction, n
        if reduction == "none":
            return row
        return out[0] / n if reduction == "mean" else out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        if ctx.reduction == "none":
            return dz * g.reshape(-1, 1), None, None, None
        scale = g / ctx.n if ctx.reduction == "mean" else g
        return dz * scale, None, None, None


def cross_entropy(pred_: torch.Tensor, label: torch.Tensor, class_weight: torch.Tensor = None,
                  sample_weight: torch.Tensor = None, reduction: str = "mean") -> torch.Tensor:
    """Cross entropy for possibly soft targets (scnym/losses.py:67-151)."""
    if pred_.size() != label.size():
        raise ValueError(f"pred size {pred_.size()} not compatible with label size {label.size()}\n")
    if reduction.lower() not in ("mean", "sum", "none"):
        raise ValueError(f"{reduction} is not a valid reduction method.")
    if sample_weight is None:
        return _SoftCEFn.apply(pred_, label, class_weight, reduction)
    samplewise = _SoftCEFn.apply(pred_, label, class_weight, "none") * sample_weight.squeeze().to(pred_.device)
    if reduction == "mean":
        return torch.mean(samplewise)
    if reduction == "sum":
        return torch.sum(samplewise)
    return samplewise


class _SoftmaxMSEFn(torch.autograd.Function):
    """mean_i( mask_i * sum_c (softmax(z)_ic - y_ic)^2 ) over ALL rows, mask = [i < n_conf] (losses.py:880-913)."""

    @staticmethod
    def forward(ctx, logits, target, n_conf):
        logits = Fn._f32c(logits)
        target = Fn._f32c(target.to(logits.device))
        n, C = logits.shape
        dz = torch.empty_like(logits)
        out = torch.zeros(1, device=logits.device, dtype=torch.float32)
        nc = torch.tensor([n_conf], dtype=torch.int32, device=logits.device)
        call("scnb_softmax_mse_fwd_bwd", ptr(logits), n, C, ptr(nc), ptr(target), None, None, None, 0, n, None, 1.0, ptr(dz), ptr(out),
             Fn._stream())
        ctx.save_for_backward(dz)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return dz * g, None, None


class scNymCrossEntropy(nn.Module):
    """`cross_entropy` wrapper for `MultiTaskTrainer` (scnym/losses.py:154-239), including the
    reference's softmax-before-CE on the model outputs (:221-231)."""

    def __init__(self, class_weight=None, sample_weight=None, reduction: str = "mean") -> None:
        super().__init__()
        self.class_weight, self.sample_weight, self.reduction = class_weight, sample_weight, reduction

    def __call__(self, labeled_sample: dict, unlabeled_sample: dict, model: nn.Module, weight: float = None) -> torch.Tensor:
        outputs, x_embed = model(labeled_sample["input"], return_embed=True)
        probs = torch.nn.functional.softmax(outputs, dim=-1)
        loss = cross_entropy(pred_=probs, label=labeled_sample["output"], sample_weight=self.sample_weight,
                             class_weight=self.class_weight, reduction=self.reduction)
        labeled_sample["embed"] = x_embed
        if unlabeled_sample is not None:
            _, u_embed = model(unlabeled_sample["input"], return_embed=True)
            unlabeled_sample["embed"] = u_embed
        return loss


def sharpen_labels(q: torch.Tensor, T: float = 0.5) -> torch.Tensor:
    """Temperature sharpening (scnym/losses.py:529-573): T=0 -> one-hot argmax, T=1 -> identity."""
    if T == 0.0:
        _, idx = torch.max(q, dim=1)
        return torch.nn.functional.one_hot(idx, num_classes=q.size(1))
    if T == 1.0:
        return q
    q = torch.pow(q, 1.0 / T)
    return q / torch.sum(q, dim=1).reshape(-1, 1)


def _rows(x, idx):
    """Row selection for dense tensors or CSR batches."""
    if isinstance(x, CSRBatch):
        from .dataprep import DeviceCSR, gather_rows
        ds = DeviceCSR(x.indptr.to(torch.int64), x.indices, x.data, x.n_cols)
        return gather_rows(ds, torch.as_tensor(idx, device=x.device, dtype=torch.int64))
    return x[idx]


def _dense(x):
    return x.to_dense() if isinstance(x, CSRBatch) else x


class InterpolationConsistencyLoss(nn.Module):
    """Teacher management shared by MixMatch (scnym/losses.py:242-406).  Only the parts the
    scnym_api path reaches are implemented: teacher copy / EMA and the MixUp operator."""

    def __init__(self, unsup_criterion: Callable, sup_criterion: Callable, decay_coef: float = 0.9997,
                 mean_teacher: bool = True, augment: Callable = None, teacher_eval: bool = True,
                 teacher_bn_running_stats: bool = None, **kwargs) -> None:
        super().__init__()
        self.unsup_criterion, self.sup_criterion = unsup_criterion, sup_criterion
        self.decay_coef, self.mean_teacher = decay_coef, mean_teacher
        if self.mean_teacher:
            print("IC Loss is using a mean teacher.")
        self.augment, self.teacher_eval = augment, teacher_eval
        self.teacher_bn_running_stats = teacher_bn_running_stats
        self.mixup_op = SampleMixUp(**kwargs)
        self.teacher = None
        self.step = 0

    def _update_teacher(self, model: nn.Module) -> None:
        if self.mean_teacher and self.teacher is not None:
            self._update_teacher_params(model)
        else:
            self.teacher = copy.deepcopy(model)
        if self.teacher_eval:
            self.teacher = self.teacher.eval()
        if self.teacher_bn_running_stats is not None:
            for m in self.teacher.modules():
                if isinstance(m, nn.BatchNorm1d):
                    m.track_running_stats = self.teacher_bn_running_stats

    def _update_teacher_params(self, model: nn.Module) -> None:
        # alpha = min(1 - 1/(step+1), decay) over parameters only (losses.py:397-405)
        alpha = min(1 - 1 / (self.step + 1), self.decay_coef)
        state = torch.tensor([0, 0, 0, self.step], dtype=torch.int64, device=next(model.parameters()).device)
        for tp, mp in zip(self.teacher.parameters(), model.parameters()):
            if tp.data.is_contiguous() and mp.data.is_contiguous():
                call("scnb_ema_update", ptr(tp.data), ptr(mp.data), tp.numel(), float(self.decay_coef), ptr(state), Fn._stream())
            else:  # gene-major first layer: same storage order on both sides
                call("scnb_ema_update", ptr(tp.data.t()), ptr(mp.data.t()), tp.numel(), float(self.decay_coef), ptr(state), Fn._stream())


class MixMatchLoss(InterpolationConsistencyLoss):
    """MixMatch loss on a labeled and an unlabeled batch (scnym/losses.py:576-927)."""

    def __init__(self, n_augmentations: int = 2, T: float = 0.5, augment_pseudolabels: bool = True,
                 pseudolabel_min_confidence: float = 0.0, **kwargs) -> None:
        super().__init__(**kwargs, keep_dominant_obs=True)
        if not callable(self.augment):
            raise TypeError("MixMatch requires a Callable for augment")
        self.n_augmentations = n_augmentations
        self.augment_pseudolabels = augment_pseudolabels
        self.T = T
        self.pseudolabel_min_confidence = pseudolabel_min_confidence
        self.n_batches_to_store = 50
        self.running_confidence_scores = []

    def _store_confidence(self, confident: torch.Tensor, highest_conf: torch.Tensor) -> None:
        if len(self.running_confidence_scores) > self.n_batches_to_store:
            self.running_confidence_scores.pop(0)
        self.running_confidence_scores.append((confident.detach().clone().cpu(), highest_conf.detach().clone().cpu()))

    @torch.no_grad()
    def _generate_labels(self, unlabeled_sample: dict):
        """Teacher guesses averaged over `n_augmentations` views, confidence, sharpening
        (losses.py:660-737) through the fused pseudolabel kernel."""
        x = unlabeled_sample["input"]
        logits = []
        for _ in range(self.n_augmentations):
            xin = x
            if self.augment_pseudolabels:
                xin = self.augment({"input": x.clone() if isinstance(x, torch.Tensor) else x, "output": torch.zeros(1)})["input"]
            logits.append(self.teacher(xin))
        tl = torch.stack(logits, 0).contiguous()
        K, U, C = tl.shape
        q = torch.empty((U, C), device=tl.device)
        conf = torch.empty(U, device=tl.device)
        flag = torch.empty(U, dtype=torch.uint8, device=tl.device)
        scratch = torch.empty((U, C), device=tl.device)
        T = self.T
        call("scnb_pseudolabel", ptr(tl), K, U, C, float(T if T is not None else 1.0), int(T is not None),
             float(self.pseudolabel_min_confidence), ptr(q), ptr(conf), ptr(flag), ptr(scratch), None, 0, Fn._stream())
        confident = flag.bool()
        self._store_confidence(confident, conf)
        return q, confident

    def __call__(self, model: nn.Module, labeled_sample: dict, unlabeled_sample: dict, **kwargs
                 ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        self._update_teacher(model)
        pseudolabels, confident = self._generate_labels(unlabeled_sample)
        pseudolabels = pseudolabels.to(dtype=labeled_sample["output"].dtype)
        labeled_sample = self.augment(labeled_sample)  # replaces the caller's "input" (losses.py:801)
        unlabeled_sample["output"] = pseudolabels
        conf_np = confident.cpu().numpy()
        cidx, uidx = np.where(conf_np)[0], np.where(~conf_np)[0]
        ux = _dense(unlabeled_sample["input"])
        lx = _dense(labeled_sample["input"])
        dev = lx.device
        cat = {"input": torch.cat([ux[cidx], lx], 0),
               "output": torch.cat([pseudolabels[cidx], labeled_sample["output"].to(dev)], 0)}
        mixed = self.mixup_op(cat)
        n_conf = len(cidx)
        u_m = torch.cat([mixed["input"][:n_conf], ux[uidx]])
        u_y = torch.cat([mixed["output"][:n_conf], pseudolabels[uidx]])
        l_m, l_y = mixed["input"][n_conf:], mixed["output"][n_conf:]
        u_logits = model(u_m)
        l_logits = model(l_m)
        if isinstance(self.unsup_criterion, nn.MSELoss):
            unsup = _SoftmaxMSEFn.apply(u_logits, u_y, n_conf)
        else:
            per = self.unsup_criterion(torch.nn.functional.softmax(u_logits, dim=1), u_y)
            if per.dim() > 1:
                per = torch.sum(per, dim=1)
            scale = torch.zeros_like(per)
            scale[:n_conf] += 1.0
            unsup = torch.mean(per * scale)
        sup = self.sup_criterion(l_logits, l_y)
        self.step += 1
        return sup, unsup, l_logits


class MultiTaskMixMatchWrapper(nn.Module):
    """`MixMatchLoss` behind the one-criterion-one-loss API of MultiTaskTrainer (losses.py:930-1034)."""

    def __init__(self, mixmatch_loss: MixMatchLoss, sup_weight: Union[float, Callable] = 1.0,
                 unsup_weight: Union[float, Callable] = 1.0, use_sup_eval: bool = True) -> None:
        super().__init__()
        self.mixmatch_loss, self.sup_weight, self.unsup_weight = mixmatch_loss, sup_weight, unsup_weight
        self.use_sup_eval = use_sup_eval
        self.epoch = 0

    def __call__(self, *, labeled_sample: dict, unlabeled_sample: dict, model: nn.Module, weight: float = None) -> torch.Tensor:
        sup_loss, unsup_loss, _ = self.mixmatch_loss(labeled_sample=labeled_sample, unlabeled_sample=unlabeled_sample, model=model)
        sw = self.sup_weight(self.epoch) if callable(self.sup_weight) else self.sup_weight
        uw = self.unsup_weight(self.epoch) if callable(self.unsup_weight) else self.unsup_weight
        if self.use_sup_eval and not self.training:
            uw = 0.0
        return (sw * sup_loss) + (uw * unsup_loss)


class DANLoss(nn.Module):
    """Domain adversarial loss (scnym/losses.py:1040-1245)."""

    def __init__(self, dan_criterion: Callable, model: CellTypeCLF, use_conf_pseudolabels: bool = False,
                 scale_loss_pseudoconf: bool = False, n_domains: int = 2, **kwargs) -> None:
        super().__init__()
        self.dan_criterion = dan_criterion
        self.dann = DANN(model=model, n_domains=n_domains, **kwargs)
        self.dann.domain_clf = self.dann.domain_clf.to(device=next(iter(model.parameters())).device)
        self.x_embed = torch.zeros((1, 1))
        self.use_conf_pseudolabels = use_conf_pseudolabels
        self.scale_loss_pseudoconf = scale_loss_pseudoconf
        self.no_weight = True  # weighting happens on the reversed gradients (trainer.py:1066-1069)

    def __call__(self, labeled_sample: dict, unlabeled_sample: dict = None, weight: float = 1.0,
                 pseudolabel_confidence: torch.Tensor = None, **kwargs) -> torch.Tensor:
        lx = _dense(labeled_sample["input"])
        dev = lx.device
        if unlabeled_sample is None:
            t = torch.zeros((0, lx.shape[1]), device=dev)
            unlabeled_sample = {"input": t, "domain": torch.zeros((0, self.dann.n_domains), device=dev)}
        # domain labels: given one-hot, or inferred source=[1,0] / target=[0,1] when -1 (losses.py:1165-1186)
        ldom = labeled_sample.get("domain", torch.Tensor([-1]))
        if torch.sum(torch.as_tensor(ldom) == -1) > 0:
            source_label = torch.nn.functional.one_hot(torch.zeros(lx.size(0)).long(), num_classes=2)
        else:
            source_label = ldom
        ux = _dense(unlabeled_sample["input"])
        udom = unlabeled_sample.get("domain", torch.Tensor([-1]))
        if torch.sum(torch.as_tensor(udom) == -1) > 0:
            target_label = torch.nn.functional.one_hot(torch.ones(ux.size(0)).long(), num_classes=2)
        else:
            target_label = udom
        source_label, target_label = torch.as_tensor(source_label).to(dev), torch.as_tensor(target_label).to(dev)
        if self.use_conf_pseudolabels and pseudolabel_confidence is not None:
            sel = pseudolabel_confidence.to(dev)
            ux, target_label = ux[sel], target_label[sel]
        self.n_conf_pseudolabels = ux.size(0)
        self.n_total_unlabeled = _dense(unlabeled_sample["input"]).size(0)
        p_conf = self.n_conf_pseudolabels / max(self.n_total_unlabeled, 1)
        x = torch.cat([lx, ux], 0)
        dlabel = torch.cat([source_label, target_label], 0).float()
        self.dann.set_rev_grad_weight(weight=weight)
        domain_pred, x_embed = self.dann(x)
        if x_embed.size(0) >= self.x_embed.size(0):
            self.x_embed = copy.copy(x_embed.detach().cpu())
            self.dlabel = copy.copy(dlabel.detach().cpu())
        dan_loss = self.dan_criterion(domain_pred, dlabel)
        _, dan_pred = torch.max(domain_pred, dim=1)
        _, dlabel_int = torch.max(dlabel, dim=1)
        self.dan_acc = torch.sum(dan_pred == dlabel_int) / float(dan_pred.size(0))
        if self.scale_loss_pseudoconf:
            dan_loss = dan_loss * p_conf
        return dan_loss


class ICLWeight(object):
    """Loss-weight schedule over epochs (scnym/trainer.py:1327-1408; duplicate at losses.py:1614-1695)."""

    def __init__(self, ramp_epochs: int, burn_in_epochs: int = 0, max_unsup_weight: float = 10.0, sigmoid: bool = False) -> None:
        self.ramp_epochs, self.burn_in_epochs = ramp_epochs, burn_in_epochs
        self.max_unsup_weight, self.sigmoid = max_unsup_weight, sigmoid
        self.step_size = self.max_unsup_weight if self.ramp_epochs == 0.0 else self.max_unsup_weight / self.ramp_epochs
        print("Scaling ICL over %d epochs, %d epochs for burn in." % (self.ramp_epochs, self.burn_in_epochs))

    def _get_weight(self, epoch: int) -> float:
        if epoch >= (self.ramp_epochs + self.burn_in_epochs):
            return self.max_unsup_weight
        if self.sigmoid:
            x = (epoch - self.burn_in_epochs) / self.ramp_epochs
            return np.exp(-5 * (x - 1) ** 2) * self.max_unsup_weight
        return self.step_size * (epoch - self.burn_in_epochs)

    def __call__(self, epoch: int) -> float:
        if type(epoch) != int:
            raise TypeError(f"epoch must be int, you passed a {type(epoch)}")
        return 0.0 if epoch < self.burn_in_epochs else self._get_weight(epoch)


__all__ = ["get_class_weight", "cross_entropy", "scNymCrossEntropy", "sharpen_labels", "InterpolationConsistencyLoss",
           "MixMatchLoss", "MultiTaskMixMatchWrapper", "DANLoss", "ICLWeight"]
