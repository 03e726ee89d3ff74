"""`fit_model` -- orchestration of one training run (reference surface: scnym/main.py:36-690).

Same signature, return value and files written as the reference (`train_idx.csv`, `test_idx.csv`,
`val_idx.csv`, checkpoints, `{exp}_log.csv`, `predictions.csv`, `labels.csv`).  Differences: the
expression matrices are uploaded once to HBM as CSR and sampled on the device
(`DeviceLoader`), the optimizer is a `FusedOptimizer`, and the trainers run the fused engine.
The CLI / cross-validation drivers of scnym/main.py:693-1654 are outside the hot path
(SURVEY.md §2 row 12).
"""
import itertools
import logging
import os
import os.path as osp
from functools import partial
from typing import Tuple, Union

import numpy as np
import torch
import torch.nn as nn
from scipy import sparse

from .dataprep import AUGMENTATION_SCHEMES, DeviceLoader, SampleMixUp, SingleCellDS, balance_classes  # noqa: F401
from .model import CellTypeCLF
from .optim import FusedOptimizer
from .predict import EvalRunner, Predicter  # noqa: F401
from .trainer import (DANLoss, ICLWeight, InterpolationConsistencyLoss, MixMatchLoss, MultiTaskTrainer,  # noqa: F401
                      SemiSupervisedTrainer, Trainer, cross_entropy, get_class_weight)
from .losses import scNymCrossEntropy

logger = logging.getLogger(__name__)

# optimizer map for selection by name (scnym/main.py:36-41)
OPTIMIZERS = {
    "adadelta": partial(FusedOptimizer, kind="adadelta"),
    "adam": partial(FusedOptimizer, kind="adam"),
    "adamw": partial(FusedOptimizer, kind="adamw"),
    "sgd": partial(FusedOptimizer, kind="sgd"),
}


class repeater(object):
    """Infinite iteration over a loader (scnym/main.py:48-68); keeps `.loader` so the fused
    trainer can sample the unlabeled set on the device instead of iterating."""

    def __init__(self, data_loader):
        self.loader = data_loader

    def __iter__(self):
        for loader in itertools.repeat(self.loader):
            for data in loader:
                yield data


def _count_domains(input_domain, unlabeled_domain, unlabeled_counts) -> int:
    """scnym/main.py:256-291, including `int(max)` without +1 when only labeled domains are given."""
    if input_domain is None and unlabeled_domain is None:
        return 2 if unlabeled_counts is not None else 1
    if input_domain is not None and unlabeled_domain is None:
        return int(input_domain.max())
    if input_domain is not None and unlabeled_domain is not None:
        u_max = 0 if len(unlabeled_domain) == 0 else unlabeled_domain.max()
        return int(np.max([input_domain.max(), u_max])) + 1
    raise ValueError("domains supplied for only one set of data")


def fit_model(
    X: Union[np.ndarray, sparse.csr_matrix], y: np.ndarray, traintest_idx: Union[np.ndarray, tuple], val_idx: np.ndarray,
    batch_size: int, n_epochs: int, lr: float, optimizer_name: str, weight_decay: float, ModelClass: nn.Module, out_path: str,
    n_genes: int = None, mixup_alpha: float = None, unlabeled_counts: np.ndarray = None, unsup_max_weight: float = 2.0,
    unsup_mean_teacher: bool = False, ssl_method: str = "mixmatch", ssl_kwargs: dict = {}, weighted_classes: bool = False,
    balanced_classes: bool = False, input_domain: np.ndarray = None, unlabeled_domain: np.ndarray = None, pretrained: str = None,
    patience: int = None, save_freq: int = None, tensorboard: bool = True, **kwargs,
) -> Tuple[float, float]:
    """Fit an scNym model; returns (held-out accuracy, held-out supervised loss)."""
    if not torch.cuda.is_available():
        raise RuntimeError("scnym_b200.fit_model needs a CUDA device (there is no CPU fallback)")
    device = torch.device("cuda", torch.cuda.current_device())
    n_cell_types = len(np.unique(y))
    if n_genes is None:
        n_genes = X.shape[1]
    if type(traintest_idx) != tuple:
        train_idx = np.random.choice(traintest_idx, size=int(np.floor(0.9 * len(traintest_idx))), replace=False).astype("int")
        test_idx = np.setdiff1d(traintest_idx, train_idx).astype("int")
    elif len(traintest_idx) == 2:
        train_idx, test_idx = traintest_idx
    else:
        raise ValueError(f"`traintest_idx` of type {type(traintest_idx)}\nand length {len(traintest_idx)} is invalid.")
    os.makedirs(out_path, exist_ok=True)
    np.savetxt(osp.join(out_path, "train_idx.csv"), train_idx)
    np.savetxt(osp.join(out_path, "test_idx.csv"), test_idx)
    np.savetxt(osp.join(out_path, "val_idx.csv"), val_idx)

    sampler, class_weight = None, None
    if balanced_classes and weighted_classes:
        raise ValueError("balancing AND weighting classes is not useful.\nPick one mode of accounting for class imbalances.")
    elif balanced_classes:
        print("Setting up a stratified sampler...")
        classes, counts = np.unique(y[train_idx], return_counts=True)
        weight_per_example = (1.0 / counts)[y[train_idx]]
        sampler = torch.utils.data.sampler.WeightedRandomSampler(weight_per_example, len(y[train_idx]))
    elif weighted_classes:
        print("Weighting classes for training...")
        class_weight = get_class_weight(y[train_idx])
        print(class_weight)
    else:
        print("Not weighting classes and not balancing classes.")

    n_domains = _count_domains(input_domain, unlabeled_domain, unlabeled_counts)
    print(f"Found {n_domains} unique domains.")
    d_train = input_domain[train_idx] if input_domain is not None else None
    d_test = input_domain[test_idx] if input_domain is not None else None
    n_cls = len(np.unique(y))
    train_ds = SingleCellDS(X=X[train_idx, :], y=y[train_idx], num_classes=n_cls, domain=d_train, num_domains=n_domains)
    test_ds = SingleCellDS(X[test_idx, :], y[test_idx], num_classes=n_cls, domain=d_test, num_domains=n_domains)
    train_dl = DeviceLoader(train_ds, batch_size=batch_size, shuffle=sampler is None, sampler=sampler, drop_last=True, device=device)
    test_dl = DeviceLoader(test_ds, batch_size=batch_size, shuffle=True, device=device)
    dataloaders = {"train": train_dl, "val": test_dl}

    batch_transformers = {}
    if mixup_alpha is not None and ssl_method != "mixmatch":
        print("Using MixUp as a batch transformer.")
        batch_transformers["train"] = SampleMixUp(alpha=mixup_alpha)

    model = ModelClass(n_genes=n_genes, n_cell_types=n_cell_types, **kwargs)
    if pretrained is not None:
        model.load_state_dict(torch.load(pretrained, map_location="cpu"))
    model = model.to(device)

    criterion = cross_entropy if class_weight is None else partial(cross_entropy, class_weight=torch.from_numpy(class_weight).float())
    opt_callable = OPTIMIZERS[optimizer_name.lower()]
    scheduler = None
    if optimizer_name.lower() != "sgd":
        optimizer = opt_callable(model.parameters(), weight_decay=weight_decay, lr=lr)
    else:
        optimizer = opt_callable(model.parameters(), weight_decay=weight_decay, lr=lr, momentum=0.9)
        scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer=optimizer, T_max=n_epochs, eta_min=lr / 10000)

    trainer_kwargs = {
        "model": model, "criterion": criterion, "optimizer": optimizer, "scheduler": scheduler, "dataloaders": dataloaders,
        "out_path": out_path, "batch_transformers": batch_transformers, "n_epochs": n_epochs, "min_epochs": n_epochs // 20,
        "save_freq": max(n_epochs // 5, 1) if save_freq is None else save_freq, "reg_criterion": None,
        "exp_name": osp.basename(out_path), "verbose": False, "tb_writer": osp.join(out_path, "tblog") if tensorboard else None,
        "patience": patience,
    }

    def make_dan():
        dan = DANLoss(model=model, dan_criterion=cross_entropy,
                      use_conf_pseudolabels=ssl_kwargs.get("dan_use_conf_pseudolabels", False),
                      scale_loss_pseudoconf=ssl_kwargs.get("dan_scale_loss_pseudoconf", False), n_domains=n_domains)
        w = ICLWeight(ramp_epochs=ssl_kwargs.get("dan_ramp_epochs", max(n_epochs // 4, 1)),
                      max_unsup_weight=ssl_kwargs.get("dan_max_weight", 1.0), burn_in_epochs=ssl_kwargs.get("dan_burn_in_epochs", 0),
                      sigmoid=ssl_kwargs.get("sigmoid", True))
        optimizer.add_param_group({"params": dan.dann.domain_clf.parameters(), "name": "domain_classifier"})
        return dan, w

    if unlabeled_counts is None and n_domains == 1:
        print("Performing fully supervised training with no domain adaptation.")
        T = Trainer(**trainer_kwargs)
    elif unlabeled_counts is None and n_domains > 1:
        print("Performing supervised training with a domain adversary.")
        dan_criterion, dan_weight = (make_dan() if ssl_kwargs.get("dan_criterion", None) is not None else (None, None))
        criteria = [{"name": "dan", "function": dan_criterion, "weight": dan_weight},
                    {"name": "ce", "function": scNymCrossEntropy(), "weight": 1.0}]
        del trainer_kwargs["criterion"]
        T = MultiTaskTrainer(criteria=criteria, **trainer_kwargs)
    else:
        unsup_dataset = SingleCellDS(X=unlabeled_counts, y=np.zeros(unlabeled_counts.shape[0]), num_classes=n_cls,
                                     domain=unlabeled_domain, num_domains=n_domains)
        unsup_dataloader = repeater(DeviceLoader(unsup_dataset, batch_size=batch_size, shuffle=True, drop_last=True, device=device))
        if ssl_method.lower() == "mixmatch":
            print("Using MixMatch for semi-supervised learning")
            name = ssl_kwargs.get("unsup_criterion", "mse").lower()
            unsup_criterion = nn.MSELoss(reduction="none") if name == "mse" else partial(cross_entropy, reduction="none")
            USL = MixMatchLoss(alpha=mixup_alpha if mixup_alpha is not None else 0.3, unsup_criterion=unsup_criterion,
                               sup_criterion=criterion, decay_coef=ssl_kwargs.get("decay_coef", 0.997),
                               mean_teacher=unsup_mean_teacher, augment=AUGMENTATION_SCHEMES[ssl_kwargs.get("augment", "log1p_drop")],
                               n_augmentations=ssl_kwargs.get("n_augmentations", 1), T=ssl_kwargs.get("T", 0.5),
                               augment_pseudolabels=ssl_kwargs.get("augment_pseudolabels", True),
                               pseudolabel_min_confidence=ssl_kwargs.get("pseudolabel_min_confidence", 0.0))
        elif ssl_method.lower() == "ict":
            raise NotImplementedError("ssl_method='ict' is not reachable from scnym_api (api.py:84 fixes 'mixmatch')")
        else:
            raise ValueError(f'{ssl_method} is not a valid semi-supervised learning method.\nmust be one of {{"ict", "mixmatch"}}')
        weight_schedule = ICLWeight(ramp_epochs=ssl_kwargs.get("ramp_epochs", max(n_epochs // 4, 1)), max_unsup_weight=unsup_max_weight,
                                    burn_in_epochs=ssl_kwargs.get("burn_in_epochs", 20), sigmoid=ssl_kwargs.get("sigmoid", False))
        trainer_kwargs["min_epochs"] = max(trainer_kwargs["min_epochs"], weight_schedule.burn_in_epochs + weight_schedule.ramp_epochs // 5)
        if ssl_kwargs.get("min_epochs", None) is not None:
            trainer_kwargs["min_epochs"] = ssl_kwargs["min_epochs"]
        trainer_kwargs["min_epochs"] = min(trainer_kwargs["min_epochs"], trainer_kwargs["n_epochs"] - 1)
        dan_criterion, dan_weight = (make_dan() if ssl_kwargs.get("dan_criterion", None) is not None else (None, None))
        T = SemiSupervisedTrainer(unsup_dataloader=unsup_dataloader, unsup_criterion=USL, unsup_weight=weight_schedule,
                                  dan_criterion=dan_criterion, dan_weight=dan_weight, **trainer_kwargs)

    print("Training...")
    T.train()
    print("Training complete.\n")

    # held-out evaluation with the best weights (scnym/main.py:606-690)
    print("Evaluating model.")
    model = ModelClass(n_genes=n_genes, n_cell_types=n_cell_types, **kwargs)
    model.load_state_dict(torch.load(osp.join(out_path, "00_best_model_weights.pkl")))
    model = model.to(device).eval()
    val_ds = SingleCellDS(X[val_idx, :], y[val_idx], num_classes=n_cls)
    ev = Trainer.__new__(Trainer)
    ev.model, ev.criterion, ev._device, ev.batch_transformers, ev.reg_criterion = model, criterion, device, {}, None
    running_loss, n_b, corrects, total = ev._eval_loader(DeviceLoader(val_ds, batch_size=batch_size, shuffle=False, device=device))
    # the reference sums per-batch mean losses and divides by the number of batches (main.py:667-677)
    csr = val_ds.device_csr(device)
    run = EvalRunner([model], chunk_rows=min(65536, max(csr.n_rows, 1)))
    pred, _, score, _ = run.run(csr, want_prob=False, want_score=True)
    loss_sum, nb = 0.0, 0
    yv = val_ds.y.to(device)
    for s in range(0, csr.n_rows, batch_size):
        loss_sum += float(criterion(score[s:s + batch_size], yv[s:s + batch_size]).item())
        nb += 1
    norm_loss = loss_sum / max(nb, 1)
    acc = corrects / max(total, 1)
    print("EVAL LOSS: ", norm_loss)
    print("EVAL ACC : ", acc)
    all_predictions = pred.cpu().numpy()
    all_labels = val_ds.y_labels.numpy()
    np.savetxt(osp.join(out_path, "predictions.csv"), all_predictions)
    np.savetxt(osp.join(out_path, "labels.csv"), all_labels)
    print("Predictions | Labels")
    print(np.stack([all_predictions, all_labels], 0).T[:15, :])
    return acc, norm_loss
