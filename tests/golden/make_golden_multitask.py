"""Golden vectors for `MultiTaskTrainer` + `scNymCrossEntropy` + `DANLoss` + `MultiTaskMixMatchWrapper`, produced by
running the UNMODIFIED reference (scnym/trainer.py:913-1321, scnym/losses.py:154-239, 930-1034, 1040-1245) in the build
container with injected randomness (same mechanism as make_golden.py: queued dropout / input-dropout masks, patched
randperm, fake Beta).  Run here only:  python tests/golden/make_golden_multitask.py  ->  tests/golden/multitask_*.npz

Two cases:
  multitask_dan_ce   : the supervised + domain-adversary branch of fit_model (main.py:421-463): criteria
                       [DANLoss (weight ICLWeight, no_weight), scNymCrossEntropy (weight 1.0, validation)], SGD.
  multitask_mixmatch : criteria [MultiTaskMixMatchWrapper(MixMatchLoss) (weight 1.0, validation), DANLoss], Adadelta
                       (the reference's own tests/test_multitask.py:13-118 construction).
Each runs 2 epochs x 2 batches of train_epoch() + val_epoch() on list loaders and records the per-call criterion losses,
the values train_epoch()/val_epoch() log, and the parameters / BatchNorm buffers after every epoch."""
import copy
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as MG  # noqa: E402  (helpers only; generation there is under __main__)
from oracle.scnym_oracle import synth_csr, csr_to_dense  # noqa: E402

ref = MG.ref
from torchvision import transforms  # noqa: E402


class ListLoader(object):
    """Deterministic loader: yields deep copies of pre-collated batches (criteria mutate the dicts)."""

    def __init__(self, batches, batch_size):
        self.batches, self.batch_size, self.sampler = batches, batch_size, None

    def __iter__(self):
        for b in self.batches:
            yield {k: v.clone() for k, v in b.items()}

    def __len__(self):
        return len(self.batches)


class Recorder(nn.Module):
    """Wraps a criterion: records every returned loss; forwards attribute protocol (epoch / no_weight / train)."""

    def __init__(self, fxn, log):
        super().__init__()
        self.fxn, self.log = fxn, log
        if hasattr(fxn, "no_weight"):
            self.no_weight = fxn.no_weight
        if hasattr(fxn, "epoch"):
            self.epoch = fxn.epoch

    def train(self, mode=True):
        self.fxn.train(mode)
        return super().train(mode)

    def __call__(self, **kw):
        if hasattr(self.fxn, "epoch"):
            self.fxn.epoch = self.epoch
        out = self.fxn(**kw)
        self.log.append(float(out.detach()))
        return out


def run(name, *, mixmatch, G=96, B=12, C=3, H0=32, H=24, n_layers=2, D=3, opt_name="sgd", lr=0.05, wd=1e-4, n_epochs=2, n_batches=2,
        seed=0):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed + 7)
    n = B * n_batches
    lp, li, ld, ly = synth_csr(n, G, 0.15, C, seed=seed)
    up, ui, ud, _ = synth_csr(n, G, 0.15, C, seed=seed + 1)
    vp, vi, vd, vy = synth_csr(B, G, 0.15, C, seed=seed + 2)
    Lx, Ux, Vx = csr_to_dense(lp, li, ld, G), csr_to_dense(up, ui, ud, G), csr_to_dense(vp, vi, vd, G)
    ldom = rng.integers(0, 2, size=n)
    udom = np.full(n, 2)
    oh = lambda v, k: torch.nn.functional.one_hot(torch.from_numpy(np.asarray(v)).long(), num_classes=k).float()
    lab_batches = [{"input": Lx[i * B:(i + 1) * B], "output": oh(ly[i * B:(i + 1) * B], C), "domain": oh(ldom[i * B:(i + 1) * B], D)}
                   for i in range(n_batches)]
    unl_batches = [{"input": Ux[i * B:(i + 1) * B], "output": torch.zeros(B), "domain": oh(udom[i * B:(i + 1) * B], D)}
                   for i in range(n_batches)]
    val_batches = [{"input": Vx, "output": oh(vy, C), "domain": oh(np.zeros(B, dtype=np.int64), D)}]
    model = ref.model.CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H0, n_layers=n_layers)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.running_mean.uniform_(-0.3, 0.3); m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.2, 0.2)
    drop_q, in_q = [], []
    MG.patch_dropouts(model, drop_q)
    opt = MG.make_optimizer(opt_name, model.parameters(), lr, wd)
    dan = ref.losses.DANLoss(model=model, dan_criterion=ref.losses.cross_entropy, n_domains=D)
    opt.add_param_group({"params": dan.dann.domain_clf.parameters(), "name": "domain_classifier"})
    dan_w = ref.trainer.ICLWeight(ramp_epochs=2, max_unsup_weight=0.3, burn_in_epochs=0, sigmoid=True)
    out = {"meta_G": G, "meta_B": B, "meta_C": C, "meta_H0": H0, "meta_H": H, "meta_n_layers": n_layers, "meta_D": D,
           "meta_opt": opt_name, "meta_lr": lr, "meta_wd": wd, "meta_n_epochs": n_epochs, "meta_n_batches": n_batches,
           "meta_mixmatch": int(mixmatch), "Lx": Lx.numpy(), "Ux": Ux.numpy(), "Vx": Vx.numpy(), "ly": ly, "vy": vy, "ldom": ldom, "udom": udom}
    out.update(MG.sd_np(model.state_dict(), "init/"))
    out.update(MG.sd_np(dan.dann.domain_clf.state_dict(), "init/domain_clf."))
    logs = {"dan": [], "main": []}
    if mixmatch:
        augment = transforms.Compose([ref.dataprep.ExpMinusOne(), MG.QueueInputDropout(0.1, in_q), ref.dataprep.LibrarySizeNormalize(log1p=True)])
        mm = ref.losses.MixMatchLoss(alpha=0.3, unsup_criterion=nn.MSELoss(reduction="none"), sup_criterion=ref.losses.cross_entropy,
                                     decay_coef=0.997, mean_teacher=False, augment=augment, n_augmentations=1, T=0.5,
                                     augment_pseudolabels=False, pseudolabel_min_confidence=0.0)
        uw = ref.trainer.ICLWeight(ramp_epochs=2, max_unsup_weight=1.5, burn_in_epochs=0, sigmoid=False)
        main = ref.losses.MultiTaskMixMatchWrapper(mixmatch_loss=mm, sup_weight=1.0, unsup_weight=uw, use_sup_eval=True)
        criteria = [{"name": "mixmatch", "function": Recorder(main, logs["main"]), "weight": 1.0, "validation": True},
                    {"name": "dan", "function": Recorder(dan, logs["dan"]), "weight": dan_w, "validation": False}]
    else:
        main = ref.losses.scNymCrossEntropy()
        criteria = [{"name": "dan", "function": Recorder(dan, logs["dan"]), "weight": dan_w, "validation": False},
                    {"name": "ce", "function": Recorder(main, logs["main"]), "weight": 1.0, "validation": True}]
    T = ref.trainer.MultiTaskTrainer(criteria=criteria, unsup_dataloader=ListLoader(unl_batches, B), model=model, optimizer=opt,
                                     scheduler=None, dataloaders={"train": ListLoader(lab_batches, B), "val": ListLoader(val_batches, B)},
                                     out_path=tempfile.mkdtemp(), batch_transformers={}, n_epochs=n_epochs, min_epochs=0, save_freq=100,
                                     reg_criterion=None, exp_name="mt", verbose=False, tb_writer=None, patience=None, use_gpu=False)
    n_app = 2 + max(n_layers - 1, 0)
    widths = [H0] + [H] * (n_app - 1)
    for e in range(n_epochs):
        T.epoch = e
        # queue the randomness of this epoch's train batches in the order the reference consumes it
        for b in range(n_batches):
            pre = f"e{e}b{b}/"
            dk = lambda rows: [torch.from_numpy((rng.random((rows, w)) < 0.5).astype(np.uint8)) for w in widths]
            if mixmatch:
                in_keep = torch.from_numpy((rng.random((B, G)) < 0.9).astype(np.uint8))
                keys = torch.from_numpy(rng.random(2 * B).astype(np.float32))
                graw = torch.distributions.Beta(0.3, 0.3).sample((2 * B,)).float()
                kU, kL, kD = dk(B), dk(B), dk(2 * B)
                out.update({pre + "in_keep_L": in_keep.numpy(), pre + "perm_keys": keys.numpy(), pre + "gamma_raw": graw.numpy()})
                for a in range(n_app):
                    out.update({pre + f"keep_U{a}": kU[a].numpy(), pre + f"keep_L{a}": kL[a].numpy(), pre + f"keep_D{a}": kD[a].numpy()})
                in_q.append(in_keep)
                drop_q.extend(kU + kL + kD)          # mixmatch (U pass, L pass) first, then the DAN pass
                out[pre + "_mm"] = 1
            else:
                kD, kL, kU = dk(2 * B), dk(B), dk(B)
                for a in range(n_app):
                    out.update({pre + f"keep_D{a}": kD[a].numpy(), pre + f"keep_L{a}": kL[a].numpy(), pre + f"keep_U{a}": kU[a].numpy()})
                drop_q.extend(kD + kL + kU)          # DAN pass first, then scNymCrossEntropy's labeled and unlabeled forwards
        if mixmatch:
            # MixUp randomness is consumed inside MixMatchLoss: patch per batch through a queue of (keys, gamma)
            seq = [(torch.from_numpy(out[f"e{e}b{b}/perm_keys"]), torch.from_numpy(out[f"e{e}b{b}/gamma_raw"])) for b in range(n_batches)]
            orig_randperm = torch.randperm

            class SeqBeta(object):
                def sample(self, shape):
                    return seq[0][1][: shape[0]].clone()

            def seq_randperm(nrows, *a, **k):
                return torch.argsort(seq[0][0][:nrows], stable=True)
            mm.mixup_op.beta = SeqBeta()
            torch.randperm = seq_randperm
            # advance the queue after every optimizer step
            orig_step = opt.step

            def stepping(*a, **k):
                r = orig_step(*a, **k)
                seq.pop(0)
                return r
            opt.step = stepping
        n0 = (len(logs["dan"]), len(logs["main"]))
        try:
            T.train_epoch()
        finally:
            if mixmatch:
                torch.randperm = orig_randperm
                opt.step = orig_step
        assert not drop_q and not in_q, (len(drop_q), len(in_q))
        out[f"e{e}/train_dan"] = np.array(logs["dan"][n0[0]:], dtype=np.float64)
        out[f"e{e}/train_main"] = np.array(logs["main"][n0[1]:], dtype=np.float64)
        n1 = (len(logs["dan"]), len(logs["main"]))
        if mixmatch:
            # the wrapped MixMatchLoss also runs in validation (use_sup_eval: only its supervised term counts), where it
            # still augments the labeled batch and draws a MixUp permutation
            in_keep = torch.from_numpy((rng.random((B, G)) < 0.9).astype(np.uint8))
            vkeys = torch.from_numpy(rng.random(2 * B).astype(np.float32))
            vgraw = torch.distributions.Beta(0.3, 0.3).sample((2 * B,)).float()
            out.update({f"e{e}val/in_keep_L": in_keep.numpy(), f"e{e}val/perm_keys": vkeys.numpy(), f"e{e}val/gamma_raw": vgraw.numpy()})
            in_q.append(in_keep)
            mm.mixup_op.beta = MG.FakeBeta(vgraw)
            with MG.patched_randperm(vkeys):
                T.val_epoch()
            assert not drop_q and not in_q, (len(drop_q), len(in_q))
        else:
            T.val_epoch()
        out[f"e{e}/val_main"] = np.array(logs["main"][n1[1]:], dtype=np.float64)
        out[f"e{e}/best_loss"] = np.float64(T.best_loss)
        out.update(MG.sd_np(model.state_dict(), f"e{e}/post/"))
        out.update(MG.sd_np(dan.dann.domain_clf.state_dict(), f"e{e}/post/domain_clf."))
        print(name, "epoch", e, "dan", out[f"e{e}/train_dan"], "main", out[f"e{e}/train_main"], "val", out[f"e{e}/val_main"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    run("multitask_dan_ce", mixmatch=False, opt_name="sgd", lr=0.05, seed=21)
    run("multitask_mixmatch", mixmatch=True, opt_name="adadelta", lr=1.0, seed=22)
