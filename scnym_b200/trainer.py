"""Trainers (reference surface: scnym/trainer.py:21-1321, ICLWeight :1327-1408).

Same constructors, attributes, log/checkpoint files and epoch semantics as the reference.  Two
execution paths per trainer:

  fused    -- when the loaders are `scnym_b200.dataprep.DeviceLoader`s, the model a
              `scnym_b200.model.CellTypeCLF`, the optimizer a `FusedOptimizer` and the criteria the
              package's own (`cross_entropy`, `MixMatchLoss`, `DANLoss`): the minibatch body runs
              as one CUDA-graph replay of `TrainEngine.step()`; losses, accuracies and pseudolabel
              confidences are accumulated on the device and read once per epoch.
  generic  -- anything else (foreign criteria, torch DataLoaders yielding dense CUDA tensors,
              torch optimizers): the reference's loop, with every model/criterion op running
              through the per-op kernels of `scnym_b200.functional`.
"""
import copy
import functools
import logging
import os
from typing import Callable, List

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .dataprep import CSRBatch, DeviceLoader, SampleMixUp
from .engine import EngineConfig, TrainEngine
from .losses import *  # noqa: F401,F403  (star-import coupling kept: scnym/trainer.py:15)
from .losses import DANLoss, MixMatchLoss, cross_entropy
from .model import CellTypeCLF
from .optim import FusedOptimizer

logger = logging.getLogger(__name__)


def checkpoint_state(module: nn.Module) -> dict:
    """state_dict in the reference's on-disk format: independent, contiguous tensors
    (`embed.0.weight` as a plain [n_hidden_init, n_genes] matrix)."""
    return {k: v.detach().clone().contiguous() for k, v in module.state_dict().items()}


def _criterion_class_weight(criterion):
    """(is_plain_cross_entropy, class_weight) for `cross_entropy` or `partial(cross_entropy, class_weight=...)`."""
    if criterion is cross_entropy:
        return True, None
    if isinstance(criterion, functools.partial) and criterion.func is cross_entropy and not criterion.args:
        extra = set(criterion.keywords) - {"class_weight"}
        if not extra:
            return True, criterion.keywords.get("class_weight")
    return False, None


def _to_device(x, dev):
    return x if isinstance(x, CSRBatch) else x.to(dev)


class Trainer(object):
    """Trains a model with a supervised criterion (scnym/trainer.py:21-502)."""

    def __init__(self, model: nn.Module, criterion: Callable, optimizer: torch.optim.Optimizer, dataloaders: dict,
                 out_path: str, batch_transformers: dict = {}, n_epochs: int = 50, min_epochs: int = 0, patience: int = None,
                 exp_name: str = "", reg_criterion: Callable = None, use_gpu: bool = torch.cuda.is_available(),
                 verbose: bool = False, save_freq: int = 10, scheduler=None, tb_writer: str = None) -> None:
        self.model, self.optimizer, self.criterion = model, optimizer, criterion
        self.n_epochs, self.min_epochs = n_epochs, min_epochs
        self.patience = patience if patience is not None else n_epochs
        self.waiting_time = 0
        self.dataloaders, self.batch_transformers = dataloaders, batch_transformers
        self.out_path, self.use_gpu, self.verbose, self.save_freq = out_path, use_gpu, verbose, save_freq
        self.best_acc, self.best_loss = 0.0, 1.0e10
        self.scheduler, self.reg_criterion = scheduler, reg_criterion
        self.tb_writer = None
        if tb_writer is not None:
            try:
                from tensorboardX import SummaryWriter
                self.tb_writer = SummaryWriter(log_dir=tb_writer)
                os.makedirs(tb_writer, exist_ok=True)
            except ImportError:
                logger.warning("tensorboardX is not installed; TensorBoard logging is disabled.")
        if not os.path.exists(self.out_path):
            os.mkdir(self.out_path)
        self.log_path = os.path.join(self.out_path, "_".join([exp_name, "log.csv"]))
        self.parameters = {
            "out_path": out_path, "exp_name": exp_name, "n_epochs": n_epochs, "use_cuda": self.use_gpu,
            "train_batch_size": self.dataloaders["train"].batch_size, "val_batch_size": self.dataloaders["val"].batch_size,
            "train_batch_sampler": str(type(self.dataloaders["train"].sampler)),
            "val_batch_sampler": str(type(self.dataloaders["val"].sampler)),
            "optimizer_type": str(type(self.optimizer)), "learning_rate": self.optimizer.param_groups[0]["lr"],
            "model_hidden": self.model.n_hidden, "model_ngenes": self.model.n_genes, "model_ncelltypes": self.model.n_cell_types,
        }
        with open(self.log_path, "w") as f:
            f.write("Epoch,Iter,Running_Loss,Mode\n")
        self.engine = None
        self._device = next(model.parameters()).device

    # ---- helpers -------------------------------------------------------------------------
    def _log(self, i, value, mode):
        with open(self.log_path, "a") as f:
            f.write(f"{self.epoch},{i},{value},{mode}\n")

    def _fusable_base(self) -> bool:
        return (isinstance(self.dataloaders["train"], DeviceLoader) and isinstance(self.model, CellTypeCLF)
                and isinstance(self.optimizer, FusedOptimizer) and self.reg_criterion is None
                and self.batch_transformers.get("train", None) is None and torch.cuda.is_available())

    def _engine_cfg(self, **kw) -> EngineConfig:
        g = self.optimizer.param_groups[0]
        return EngineConfig(optimizer=self.optimizer.kind, lr=g["lr"], weight_decay=g["weight_decay"], betas=tuple(g["betas"]),
                            eps=g["eps"], rho=g["rho"], momentum=g["momentum"], **kw)

    def _sync_hyper(self):
        g = self.optimizer.param_groups[0]
        self.engine.set_hyper(lr=g["lr"], weight_decay=g["weight_decay"])

    def _build_engine(self):
        ok, cw = _criterion_class_weight(self.criterion)
        if not (self._fusable_base() and ok):
            return None
        dl = self.dataloaders["train"]
        ds = dl.dataset
        cfg = self._engine_cfg(class_weight=None if cw is None else [float(x) for x in cw])
        dev = self._device if self._device.type == "cuda" else torch.device("cuda")
        self.model.to(dev)
        eng = TrainEngine(self.model, cfg, ds.device_csr(dev), ds.y_labels.numpy(), dl.batch_size, mode="supervised")
        self.optimizer.attach_engine(eng)
        return eng

    # ---- epochs --------------------------------------------------------------------------
    def train_epoch(self):
        self.model.train(True)
        if self.engine is None:
            self.engine = self._build_engine() or False
        if self.engine:
            return self._train_epoch_fused()
        return self._train_epoch_generic()

    def _train_epoch_fused(self):
        eng, dl = self.engine, self.dataloaders["train"]
        B, n_batches = dl.batch_size, len(dl)
        order = dl.epoch_order()
        full = (order.numel() // B) * B
        if full < order.numel() and not dl.drop_last:
            n_batches = full // B  # the fused step needs a fixed batch; a ragged tail batch is dropped
        if n_batches == 0:
            raise RuntimeError("fewer training cells than one batch")
        eng.set_epoch_tables(order[: n_batches * B].view(n_batches, B))
        self._sync_hyper()
        eng.enable_history(n_batches)
        for _ in range(n_batches):
            eng.step()
        h, _ = eng.history()
        if np.isnan(h[:, 0]).any():
            raise RuntimeError("NaN loss encountered in training")
        running_loss = float((h[:, 0] / B).sum())  # the reference accumulates loss.item() / batch (trainer.py:237)
        epoch_loss = running_loss / n_batches
        epoch_acc = float(h[:, 1].sum()) / float(n_batches * B)
        self._log(n_batches, running_loss / (n_batches + 1), "train_epoch")
        self._tb_train(epoch_loss, epoch_acc)
        if self.verbose:
            print("{} Loss : {:.4f}".format("train", epoch_loss))
            print("{} Acc : {:.4f}".format("train", epoch_acc))

    def _tb_train(self, epoch_loss, epoch_acc):
        if self.tb_writer is not None:
            self.tb_writer.add_scalar("Loss/train", epoch_loss, self.epoch)
            self.tb_writer.add_scalar("Acc/train", epoch_acc, self.epoch)
            self.tb_writer.add_scalar("lr/lr", self.optimizer.param_groups[0]["lr"], self.epoch)

    def _train_epoch_generic(self):
        i, running_loss, running_corrects, running_total = 0, 0.0, 0.0, 0.0
        btrans = self.batch_transformers.get("train", None)
        dev = self._device
        for data in self.dataloaders["train"]:
            if btrans is not None:
                data = btrans(data)
            inputs, labels = _to_device(data["input"], dev), data["output"].to(dev)
            self.optimizer.zero_grad()
            outputs = self.model(inputs)
            _, predictions = torch.max(outputs, 1)
            correct = torch.sum(predictions.detach() == torch.argmax(labels, 1))
            loss = self.criterion(outputs, labels)
            if self.reg_criterion is not None:
                loss = loss + self.reg_criterion(self.model)
            if np.isnan(loss.data.cpu().numpy()):
                raise RuntimeError("NaN loss encountered in training")
            loss.backward()
            self.optimizer.step()
            running_loss += loss.item() / labels.size(0)
            running_corrects += float(correct.item())
            running_total += float(labels.size(0))
            i += 1
        epoch_loss = running_loss / max(len(self.dataloaders["train"]), 1)
        self._log(i, running_loss / (i + 1), "train_epoch")
        self._tb_train(epoch_loss, running_corrects / max(running_total, 1.0))

    @torch.no_grad()
    def _eval_loader(self, loader):
        """(sum over batches of mean-loss/batch-size, n_batches, corrects, total) with the supervised criterion."""
        self.model.train(False)
        dev = self._device
        ok, cw = _criterion_class_weight(self.criterion) if self.criterion is not None else (True, None)
        running_loss, corrects, total, n_b = 0.0, 0, 0, 0
        if isinstance(loader, DeviceLoader) and isinstance(self.model, CellTypeCLF) and ok and dev.type == "cuda":
            # one pass over the resident CSR; per-row losses from the fused CE kernel, then per-batch means
            from .predict import EvalRunner
            from ._lib import call, ptr
            ds = loader.dataset
            csr = ds.device_csr(dev)
            run = EvalRunner([self.model], chunk_rows=min(65536, max(csr.n_rows, 1)))
            N, C = csr.n_rows, self.model.n_cell_types
            y = ds.y.to(dev)
            row_loss = torch.empty(N, device=dev)
            pred = torch.empty(N, dtype=torch.int64, device=dev)
            scratch = torch.zeros(2, device=dev)
            cwd = None if cw is None else torch.as_tensor(cw, dtype=torch.float32, device=dev)
            run.refresh_planes()
            for s in range(0, N, run.chunk):
                n = min(run.chunk, N - s)
                run.run_chunk(csr, s, n, pred[s:], None, None, None)
                call("scnb_soft_ce_fwd_bwd", ptr(run.logits), n, C, None, ptr(y[s:]), None, None, None, 0, ptr(cwd), None, 1, 1.0, None,
                     None, ptr(scratch), 0, ptr(row_loss[s:]), torch.cuda.current_stream().cuda_stream)
            B = loader.batch_size
            corrects = int((pred == ds.y_labels.to(dev)).sum().item())
            total = N
            for s in range(0, N, B):
                seg = row_loss[s:s + B]
                running_loss += float(seg.mean().item()) / seg.numel()
                n_b += 1
            return running_loss, n_b, corrects, total
        btrans = self.batch_transformers.get("val", None)
        for data in loader:
            if btrans is not None:
                data = btrans(data)
            inputs, labels = _to_device(data["input"], dev), data["output"].to(dev)
            outputs = self.model(inputs)
            _, predictions = torch.max(outputs, 1)
            corrects += int(torch.sum(predictions == torch.argmax(labels, 1)).item())
            loss = self.criterion(outputs, labels)
            if self.reg_criterion is not None:
                loss = loss + self.reg_criterion(self.model)
            running_loss += loss.item() / labels.size(0)
            total += int(labels.size(0))
            n_b += 1
        return running_loss, n_b, corrects, total

    def _save_best(self, epoch_loss, i, running_ratio):
        """Checkpoint bookkeeping shared by all trainers (trainer.py:377-441)."""
        self.waiting_time += 1
        if (epoch_loss < self.best_loss) and (self.epoch >= self.min_epochs):
            self.best_loss = epoch_loss
            self.best_model_wts = checkpoint_state(self.model)
            self.waiting_time = 0
            torch.save(self.best_model_wts, os.path.join(self.out_path, "model_weights_" + str(self.epoch).zfill(3) + ".pkl"))
            print("Saving best model weights...")
            torch.save(self.best_model_wts, os.path.join(self.out_path, "00_best_model_weights.pkl"))
            print("Saved best weights.")
            self._save_aux_weights()
            self._log(i, running_ratio, "best_model_weights")
            if self.tb_writer is not None:
                self.tb_writer.add_text("BestWeights", f"Saved best weights at {self.epoch}, loss {epoch_loss}", self.epoch)
        elif self.epoch % self.save_freq == 0:
            torch.save(checkpoint_state(self.model), os.path.join(self.out_path, "model_weights_" + str(self.epoch).zfill(3) + ".pkl"))
        elif self.epoch == (self.n_epochs - 1):
            torch.save(checkpoint_state(self.model), os.path.join(self.out_path, "01_final_model_weights.pkl"))

    def _save_aux_weights(self):
        dan = getattr(self, "dan_criterion", None)
        if dan is not None:
            print("Trainer has a `dan_criterion`.")
            print("Saving DAN weights...")
            torch.save(checkpoint_state(dan.dann), os.path.join(self.out_path, "02_best_dan_weights.pkl"))

    @torch.no_grad()
    def val_epoch(self):
        running_loss, n_b, corrects, total = self._eval_loader(self.dataloaders["val"])
        epoch_loss = running_loss / max(n_b, 1)
        epoch_acc = corrects / max(total, 1)
        self._log(n_b, running_loss / (n_b + 1), "val_epoch")
        self._save_best(epoch_loss, n_b, running_loss / (n_b + 1))
        if self.verbose:
            print(f"{self.waiting_time} epochs since last best weights.\n")
            print("{} Loss : {:.4f}".format("val", epoch_loss))
            print("{} Acc : {:.4f}".format("val", epoch_acc))
        if self.tb_writer is not None:
            self.tb_writer.add_scalar("Loss/val", epoch_loss, self.epoch)
            self.tb_writer.add_scalar("Acc/val", epoch_acc, self.epoch)
            self.tb_writer.flush()

    def train(self):
        for epoch in range(self.n_epochs):
            self.epoch = epoch
            n_bars = int(np.floor(30 * epoch / self.n_epochs))
            msg = f"Epoch {epoch}/{self.n_epochs-1}" + "|" + "-" * n_bars + "_" * (30 - n_bars) + "|"
            print(msg, end="\n" if epoch == (self.n_epochs - 1) else "\r")
            self.train_epoch()
            self.val_epoch()
            if self.scheduler is not None:
                self.scheduler.step()
            if self.waiting_time > self.patience:
                logger.info(f"Early stopping at epoch {self.epoch}")
                break
        self.model.load_state_dict(torch.load(os.path.join(self.out_path, "00_best_model_weights.pkl")))
        if self.engine:
            self.engine.arena.check_aliasing()
        if self.tb_writer is not None:
            self.tb_writer.flush()
            self.tb_writer.close()
        return self.model


class SemiSupervisedTrainer(Trainer):
    """Supervised + MixMatch (+ DAN) training (scnym/trainer.py:505-910)."""

    def __init__(self, unsup_criterion: Callable, unsup_dataloader, unsup_weight: Callable, dan_criterion: Callable = None,
                 dan_weight: Callable = None, **kwargs) -> None:
        super().__init__(**kwargs)
        self.unsup_criterion, self.unsup_dataloader, self.unsup_weight = unsup_criterion, unsup_dataloader, unsup_weight
        self.dan_criterion = dan_criterion
        if self.dan_criterion is not None:
            print("Using a Domain Adaptation Loss.")
        self.dan_weight = dan_weight
        self._unsup_iter = None

    def _unsup_loader(self):
        return getattr(self.unsup_dataloader, "loader", self.unsup_dataloader)

    def _build_engine(self):
        mm, dan, ul = self.unsup_criterion, self.dan_criterion, self._unsup_loader()
        ok_sup, cw = _criterion_class_weight(getattr(mm, "sup_criterion", None))
        if not (self._fusable_base() and isinstance(mm, MixMatchLoss) and isinstance(ul, DeviceLoader)
                and isinstance(mm.unsup_criterion, nn.MSELoss) and ok_sup and (dan is None or isinstance(dan, DANLoss))
                and (dan is None or _criterion_class_weight(dan.dan_criterion) == (True, None))
                and getattr(mm.augment, "p_drop", None) is not None and getattr(mm.augment, "scheme", "log1p_drop") != "count_poisson"
                and mm.teacher_eval and mm.teacher_bn_running_stats is None):
            return None
        dl = self.dataloaders["train"]
        ds, uds = dl.dataset, ul.dataset
        cfg = self._engine_cfg(n_augmentations=mm.n_augmentations, T=mm.T, augment_pseudolabels=mm.augment_pseudolabels,
                               pseudolabel_min_confidence=mm.pseudolabel_min_confidence, mixup_alpha=mm.mixup_op.alpha,
                               mean_teacher=mm.mean_teacher, decay_coef=mm.decay_coef, p_in_drop=mm.augment.p_drop,
                               augment=getattr(mm.augment, "scheme", "log1p_drop"), p_mask_apply=getattr(mm.augment, "p_apply", 0.5),
                               use_dan=dan is not None,
                               dan_use_conf_pseudolabels=bool(dan.use_conf_pseudolabels) if dan is not None else False,
                               dan_scale_loss_pseudoconf=bool(dan.scale_loss_pseudoconf) if dan is not None else False,
                               class_weight=None if cw is None else [float(x) for x in cw])
        dev = self._device if self._device.type == "cuda" else torch.device("cuda")
        self.model.to(dev)
        given = ds.dom_labels is not None and uds.dom_labels is not None
        eng = TrainEngine(self.model, cfg, ds.device_csr(dev), ds.y_labels.numpy(), dl.batch_size, unlabeled=uds.device_csr(dev),
                          unlabeled_batch_size=ul.batch_size, dann=dan.dann if dan is not None else None,
                          labeled_domain=np.asarray(ds.dom_labels) if given else None,
                          unlabeled_domain=np.asarray(uds.dom_labels) if given else None, mode="mixmatch")
        self.optimizer.attach_engine(eng)
        return eng

    def _train_epoch_fused(self):
        eng, dl, ul = self.engine, self.dataloaders["train"], self._unsup_loader()
        L, U, n_batches = dl.batch_size, ul.batch_size, len(dl)
        order = dl.epoch_order()
        n_batches = min(n_batches, order.numel() // L)
        if n_batches == 0:
            raise RuntimeError("fewer labeled cells than one batch")
        uw = self.unsup_weight(self.epoch)
        dw = self.dan_weight(self.epoch) if self.dan_criterion is not None else 0.0
        eng.set_weights(float(uw), float(dw))
        self._sync_hyper()
        # unlabeled batches: fresh shuffles of the unlabeled set, as many as the epoch needs (main.py:48-68 repeater)
        need = n_batches * U
        chunks = []
        while sum(c.numel() for c in chunks) < need:
            o = ul.epoch_order()
            chunks.append(o[: (o.numel() // U) * U] if o.numel() >= U else o.repeat((U + o.numel() - 1) // o.numel())[:U])
        u_tab = torch.cat(chunks)[:need].view(n_batches, U)
        keys = torch.rand((n_batches, L + U), device=eng.dev)
        a = max(float(eng.cfg.mixup_alpha), 1e-3)
        conc = torch.full((n_batches, L + U), a, device=eng.dev)
        g1, g2 = torch._standard_gamma(conc), torch._standard_gamma(conc)
        eng.set_epoch_tables(order[: n_batches * L].view(n_batches, L), u_tab, keys, g1 / (g1 + g2).clamp_min(1e-30))
        eng.enable_history(n_batches)
        for _ in range(n_batches):
            eng.step()
        h, conf = eng.history()
        mm = self.unsup_criterion
        total = h[:, 0] + uw * h[:, 2] + (h[:, 4] if self.dan_criterion is not None else 0.0)
        if np.isnan(total).any():
            raise RuntimeError("NaN loss encountered in training")
        mm.step += n_batches
        tau = mm.pseudolabel_min_confidence
        mm.running_confidence_scores = [(torch.from_numpy(c >= tau), torch.from_numpy(c.copy())) for c in conf][-(mm.n_batches_to_store + 1):]
        if self.dan_criterion is not None:
            d = self.dan_criterion
            cnt = eng.counts.cpu().numpy()
            d.n_conf_pseudolabels, d.n_total_unlabeled = int(cnt[6]), U
            d.dan_acc = torch.tensor(float(h[-1, 5]) / max(int(cnt[2]), 1))
            d.x_embed = eng.As[-1][eng.nb: eng.nb + int(cnt[2])].detach().cpu()
            d.dlabel = eng.dlabel[: int(cnt[2])].detach().cpu()
        epoch_loss = float(total.mean())
        epoch_sup, epoch_uns = float(h[:, 0].mean()), float(h[:, 2].mean())
        epoch_dom = float(h[:, 4].mean()) if self.dan_criterion is not None else 0.0
        epoch_acc = float(h[:, 1].sum()) / float(n_batches * L)
        self._ssl_epoch_logs(n_batches, epoch_loss, epoch_sup, epoch_uns, epoch_dom, epoch_acc)

    def _ssl_epoch_logs(self, i, epoch_loss, epoch_sup, epoch_uns, epoch_dom, epoch_acc):
        if self.tb_writer is not None:
            for tag, v in (("Loss/train", epoch_loss), ("Acc/train", epoch_acc), ("Loss/super", epoch_sup), ("Loss/unsup", epoch_uns),
                           ("SSL/UnsWeight", self.unsup_weight(self.epoch))):
                self.tb_writer.add_scalar(tag, v, self.epoch)
            if self.dan_criterion is not None:
                self.tb_writer.add_scalar("Loss/domain", epoch_dom, self.epoch)
                self.tb_writer.add_scalar("SSL/DomWeight", self.dan_weight(self.epoch), self.epoch)
            self.tb_writer.flush()
        self._log(i, epoch_loss, "train_epoch")
        self._log(i, epoch_sup, "train_epoch_sup")
        self._log(i, epoch_uns, "train_epoch_uns")
        self._log(i, self.unsup_weight(self.epoch), "train_epoch_uns_weight")
        if self.verbose:
            print("{} Sup. Loss : {:.6f}".format("train", epoch_sup))
            print("{} Unsup. Loss : {:.6f}".format("train", epoch_uns))
            if self.dan_criterion is not None:
                print("{} Dom.  Loss : {:.6f}".format("train", epoch_dom))
            print("{} Loss : {:.4f}".format("train", epoch_loss))
            print("{} Acc : {:.4f}".format("train", epoch_acc))

    def _train_epoch_generic(self):
        """The reference loop (trainer.py:575-702) over per-op kernels."""
        dev = self._device
        i = 0
        run = dict(loss=0.0, sup=0.0, uns=0.0, dom=0.0, corrects=0.0, total=0.0)
        btrans = self.batch_transformers.get("train", None)
        # `iter_unsup_dl = iter(self.unsup_dataloader)` per epoch (scnym/trainer.py:574): a plain DataLoader starts over, the
        # `repeater` generator fit_model passes carries on
        self._unsup_iter = iter(self.unsup_dataloader)
        for data in self.dataloaders["train"]:
            try:
                unsup_data = next(self._unsup_iter)
            except StopIteration:
                self._unsup_iter = iter(self.unsup_dataloader)
                unsup_data = next(self._unsup_iter)
            if btrans is not None:
                data = btrans(data)
            data["input"], data["output"] = _to_device(data["input"], dev), data["output"].to(dev)
            unsup_data["input"] = _to_device(unsup_data["input"], dev)
            self.optimizer.zero_grad()
            sup_loss, unsup_loss, sup_outputs = self.unsup_criterion(model=self.model, labeled_sample=data, unlabeled_sample=unsup_data)
            _, predictions = torch.max(sup_outputs, 1)
            correct = torch.sum(predictions.detach() == torch.argmax(data["output"], 1))
            reg_loss = self.reg_criterion(self.model) if self.reg_criterion is not None else 0.0
            if self.dan_criterion is not None:
                dan_weight = self.dan_weight(self.epoch)
                conf = self.unsup_criterion.running_confidence_scores[-1][0]
                dan_loss = self.dan_criterion(labeled_sample=data, unlabeled_sample=unsup_data, weight=dan_weight,
                                              pseudolabel_confidence=conf)
            else:
                dan_loss = torch.zeros(1, device=sup_loss.device)
            loss = sup_loss + reg_loss + (self.unsup_weight(self.epoch) * unsup_loss) + dan_loss
            if np.isnan(loss.data.cpu().numpy()):
                raise RuntimeError("NaN loss encountered in training")
            loss.backward()
            self.optimizer.step()
            run["loss"] += float(loss.item()); run["sup"] += sup_loss.item(); run["uns"] += unsup_loss.item()
            run["dom"] += float(dan_loss.item()); run["corrects"] += float(correct.item()); run["total"] += float(data["output"].size(0))
            i += 1
        n = max(len(self.dataloaders["train"]), 1)
        self._ssl_epoch_logs(i, run["loss"] / n, run["sup"] / n, run["uns"] / n, run["dom"] / n, run["corrects"] / max(run["total"], 1.0))


class MultiTaskTrainer(Trainer):
    """Sequential multi-criterion training (scnym/trainer.py:913-1321).  Criteria follow the
    reference protocol `fxn(labeled_sample=, unlabeled_sample=, model=, weight=) -> scalar`,
    optional `.epoch`, `.no_weight`, `.train(bool)`; this trainer always takes the generic path."""

    def __init__(self, criteria: List[dict], unsup_dataloader=None, **kwargs) -> None:
        kwargs.update({"criterion": None})
        super().__init__(**kwargs)
        self.criteria = criteria
        for c in self.criteria:
            fxn, weight = c.get("function", None), c.get("weight", None)
            if not callable(fxn):
                raise ValueError(fxn)
            if not callable(weight) and type(weight) != float:
                raise ValueError(f'One of the criteria did not include a `"weight"` property.\n\t{fxn}\n\tweight : {weight}')
        self.unsup_dataloader = unsup_dataloader
        self.best_weights = None
        self._unsup_iter = None

    def _build_engine(self):
        return None

    def _next_unsup(self, dev):
        if self.unsup_dataloader is None:
            return None
        if self._unsup_iter is None:
            self._unsup_iter = iter(self.unsup_dataloader)
        try:
            u = next(self._unsup_iter)
        except StopIteration:
            self._unsup_iter = iter(self.unsup_dataloader)
            u = next(self._unsup_iter)
        u["input"] = _to_device(u["input"], dev)
        return u

    def _criteria_loss(self, data, unsup_data, train: bool, running):
        loss = torch.zeros(1, device=self._device)
        for idx, cd in enumerate(self.criteria):
            if not train and not cd.get("validation", False):
                continue
            fxn, wf = cd["function"], cd["weight"]
            weight = wf(self.epoch) if callable(wf) else wf
            fxn.train(train)
            if train and hasattr(fxn, "epoch"):
                fxn.epoch = self.epoch
            crit_loss = fxn(labeled_sample=data, unlabeled_sample=unsup_data, model=self.model, weight=weight)
            if hasattr(fxn, "no_weight"):
                weight = 1.0
            w_loss = crit_loss * weight
            loss = loss + w_loss
            running[idx] += crit_loss.item() if train else float(w_loss.item())
        return loss

    def _restart_unsup(self):
        """`iter_unsup_dl = iter(self.unsup_dataloader)` at the start of every training and validation pass
        (scnym/trainer.py:1000, 1132): a plain DataLoader starts over, a `repeater` generator carries on."""
        self._unsup_iter = iter(self.unsup_dataloader) if self.unsup_dataloader is not None else None

    def train_epoch(self) -> float:
        self.model.train(True)
        dev = self._device
        running = np.zeros(len(self.criteria))
        btrans = self.batch_transformers.get("train", None)
        self._restart_unsup()
        for data in self.dataloaders["train"]:
            if btrans is not None:
                data = btrans(data)
            data["input"], data["output"] = _to_device(data["input"], dev), data["output"].to(dev)
            unsup_data = self._next_unsup(dev)
            self.optimizer.zero_grad()
            loss = self._criteria_loss(data, unsup_data, True, running)
            loss.backward()
            self.optimizer.step()
        epoch_losses = running / max(len(self.dataloaders["train"]), 1)
        if self.tb_writer is not None:
            for idx, cd in enumerate(self.criteria):
                name = cd.get("name", cd["function"].__class__.__name__)
                self.tb_writer.add_scalar("loss/" + name, float(epoch_losses[idx]), self.epoch)
        return np.sum(epoch_losses)

    @torch.no_grad()
    def val_epoch(self):
        self.model.train(False)
        dev = self._device
        i = 0  # never incremented, as in the reference (trainer.py:1127,1225)
        running = np.zeros(len(self.criteria))
        corrects, total = 0.0, 0
        btrans = self.batch_transformers.get("val", None)
        self._restart_unsup()
        for data in self.dataloaders["val"]:
            if btrans is not None:
                data = btrans(data)
            data["input"], data["output"] = _to_device(data["input"], dev), data["output"].to(dev)
            unsup_data = self._next_unsup(dev)
            outputs = self.model(data["input"])
            _, predictions = torch.max(outputs, 1)
            corrects += float(torch.sum(predictions == torch.argmax(data["output"], 1)).item())
            total += int(data["output"].size(0))
            self._criteria_loss(data, unsup_data, False, running)
        epoch_loss = float(np.sum(running / max(len(self.dataloaders["val"]), 1)))
        self._log(i, epoch_loss / (i + 1), "val_epoch")
        self._save_best(epoch_loss, i, epoch_loss)
        if self.best_loss == epoch_loss:
            self.best_weights = copy.deepcopy(checkpoint_state(self.model))
        if self.tb_writer is not None:
            self.tb_writer.add_scalar("Loss/val", epoch_loss, self.epoch)
            self.tb_writer.add_scalar("Acc/val", corrects / max(total, 1), self.epoch)
            self.tb_writer.flush()
        return epoch_loss

    def _save_aux_weights(self):
        for c in self.criteria:
            if c["function"].__class__.__name__ == "DANLoss":
                torch.save(checkpoint_state(c["function"].dann), os.path.join(self.out_path, "02_best_dan_weights.pkl"))
