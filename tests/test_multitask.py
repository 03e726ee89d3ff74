"""`MultiTaskTrainer` + `scNymCrossEntropy` + `DANLoss` + `MultiTaskMixMatchWrapper` against golden
vectors produced by the UNMODIFIED reference (tests/golden/make_golden_multitask.py; scnym/trainer.py:913-1321,
scnym/losses.py:154-239 incl. the softmax-before-cross-entropy of :221-231, :930-1034, :1040-1245): two epochs of
`train_epoch()` + `val_epoch()` with the reference's dropout / input-dropout / MixUp randomness injected; every
per-batch criterion loss, the validation loss, `best_loss`, and all parameters and BatchNorm buffers after each epoch
within 1e-4 (norm-relative)."""
import numpy as np
import pytest
import torch

from tests import golden_util as GU

TOL = 1e-4


class ListLoader(object):
    def __init__(self, batches, batch_size):
        self.batches, self.batch_size, self.sampler = batches, batch_size, None

    def __iter__(self):
        for b in self.batches:
            yield {k: v.clone() for k, v in b.items()}

    def __len__(self):
        return len(self.batches)


class Recorder(torch.nn.Module):
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


def test_multitask_goldens_are_committed_and_self_consistent():
    """CPU: the fixtures exist and hold what the GPU test reads (and the double-softmax quirk is visible in them:
    scNymCrossEntropy's loss of a C-class problem cannot drop below -log(e/(e+C-1)))."""
    for name in ("multitask_dan_ce", "multitask_mixmatch"):
        d = GU.load(name)
        for e in range(int(d["meta_n_epochs"])):
            assert d[f"e{e}/train_dan"].shape == (int(d["meta_n_batches"]),)
            assert d[f"e{e}/train_main"].shape == (int(d["meta_n_batches"]),)
            assert d[f"e{e}/val_main"].shape == (1,)
    d = GU.load("multitask_dan_ce")
    C = int(d["meta_C"])
    floor = -np.log(np.e / (np.e + C - 1))
    assert float(d["e1/train_main"].min()) >= floor - 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["multitask_dan_ce", "multitask_mixmatch"])
def test_multitask_trainer_matches_reference_golden(name, tmp_path):
    from scnym_b200 import functional as Fn
    from scnym_b200 import trainer as T
    from scnym_b200.dataprep import Log1pDrop, SampleMixUp
    from scnym_b200.model import CellTypeCLF
    from scnym_b200 import optim as OPT
    d = GU.load(name)
    G, B, C, H0, H, nl, D = (int(d[k]) for k in ("meta_G", "meta_B", "meta_C", "meta_H0", "meta_H", "meta_n_layers", "meta_D"))
    n_epochs, n_batches, mixmatch = int(d["meta_n_epochs"]), int(d["meta_n_batches"]), bool(int(d["meta_mixmatch"]))
    oh = lambda v, k: torch.nn.functional.one_hot(torch.from_numpy(np.asarray(v)).long(), num_classes=k).float()
    Lx, Ux, Vx = (torch.from_numpy(d[k]) for k in ("Lx", "Ux", "Vx"))
    lab = [{"input": Lx[i * B:(i + 1) * B], "output": oh(d["ly"][i * B:(i + 1) * B], C), "domain": oh(d["ldom"][i * B:(i + 1) * B], D)}
           for i in range(n_batches)]
    unl = [{"input": Ux[i * B:(i + 1) * B], "output": torch.zeros(B), "domain": oh(d["udom"][i * B:(i + 1) * B], D)} for i in range(n_batches)]
    val = [{"input": Vx, "output": oh(d["vy"], C), "domain": oh(np.zeros(B, dtype=np.int64), D)}]
    st = GU.state_from(d, "init/")
    model_st, dan_st = GU.split_state(st)
    model = CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H0, n_layers=nl)
    model.load_state_dict(model_st)
    model = model.cuda()
    kind, lr, wd = str(d["meta_opt"]), float(d["meta_lr"]), float(d["meta_wd"])
    opt = OPT.SGD(model.parameters(), lr=lr, weight_decay=wd, momentum=0.9) if kind == "sgd" else OPT.Adadelta(model.parameters(), lr=lr, weight_decay=wd)
    dan = T.DANLoss(model=model, dan_criterion=T.cross_entropy, n_domains=D)
    dan.dann.domain_clf.load_state_dict({k[len("domain_clf."):]: v for k, v in dan_st.items()})
    dan.dann.domain_clf.cuda()
    opt.add_param_group({"params": dan.dann.domain_clf.parameters(), "name": "domain_classifier"})
    dan_w = T.ICLWeight(ramp_epochs=2, max_unsup_weight=0.3, burn_in_epochs=0, sigmoid=True)
    logs = {"dan": [], "main": []}
    if mixmatch:
        mm = T.MixMatchLoss(alpha=0.3, unsup_criterion=torch.nn.MSELoss(reduction="none"), sup_criterion=T.cross_entropy,
                            decay_coef=0.997, mean_teacher=False, augment=Log1pDrop(0.1), n_augmentations=1, T=0.5,
                            augment_pseudolabels=False, pseudolabel_min_confidence=0.0)
        uw = T.ICLWeight(ramp_epochs=2, max_unsup_weight=1.5, burn_in_epochs=0, sigmoid=False)
        main = T.MultiTaskMixMatchWrapper(mixmatch_loss=mm, sup_weight=1.0, unsup_weight=uw, use_sup_eval=True)
        criteria = [{"name": "mixmatch", "function": Recorder(main, logs["main"]), "weight": 1.0, "validation": True},
                    {"name": "dan", "function": Recorder(dan, logs["dan"]), "weight": dan_w, "validation": False}]
    else:
        main = T.scNymCrossEntropy()
        criteria = [{"name": "dan", "function": Recorder(dan, logs["dan"]), "weight": dan_w, "validation": False},
                    {"name": "ce", "function": Recorder(main, logs["main"]), "weight": 1.0, "validation": True}]
    trainer = T.MultiTaskTrainer(criteria=criteria, unsup_dataloader=ListLoader(unl, B), model=model, optimizer=opt, scheduler=None,
                                 dataloaders={"train": ListLoader(lab, B), "val": ListLoader(val, B)}, out_path=str(tmp_path),
                                 batch_transformers={}, n_epochs=n_epochs, min_epochs=0, save_freq=100, reg_criterion=None,
                                 exp_name="mt", verbose=False, tb_writer=None, patience=None)
    n_app = 2 + max(nl - 1, 0)
    t = lambda k: torch.from_numpy(d[k])
    for e in range(n_epochs):
        trainer.epoch = e
        for b in range(n_batches):
            pre = f"e{e}b{b}/"
            kU, kL, kD = ([t(pre + f"keep_{s}{a}") for a in range(n_app)] for s in "ULD")
            if mixmatch:
                Log1pDrop.KEEP_INJECT.append(t(pre + "in_keep_L"))
                SampleMixUp.INJECT.append((t(pre + "perm_keys"), t(pre + "gamma_raw")))
                Fn.DROPOUT_INJECT.extend(kU + kL + kD)
            else:
                Fn.DROPOUT_INJECT.extend(kD + kL + kU)
        n0 = (len(logs["dan"]), len(logs["main"]))
        trainer.train_epoch()
        assert not Fn.DROPOUT_INJECT and not Log1pDrop.KEEP_INJECT and not SampleMixUp.INJECT
        got_dan, got_main = np.array(logs["dan"][n0[0]:]), np.array(logs["main"][n0[1]:])
        assert GU.rel_err(got_dan, d[f"e{e}/train_dan"]) <= TOL, (e, got_dan, d[f"e{e}/train_dan"])
        assert GU.rel_err(got_main, d[f"e{e}/train_main"]) <= TOL, (e, got_main, d[f"e{e}/train_main"])
        n1 = len(logs["main"])
        if mixmatch:
            Log1pDrop.KEEP_INJECT.append(t(f"e{e}val/in_keep_L"))
            SampleMixUp.INJECT.append((t(f"e{e}val/perm_keys"), t(f"e{e}val/gamma_raw")))
        trainer.val_epoch()
        assert not Fn.DROPOUT_INJECT and not Log1pDrop.KEEP_INJECT and not SampleMixUp.INJECT
        assert GU.rel_err(np.array(logs["main"][n1:]), d[f"e{e}/val_main"]) <= TOL
        assert GU.rel_err(np.float64(trainer.best_loss), d[f"e{e}/best_loss"]) <= TOL
        post = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        post.update({"domain_clf." + k: v.detach().cpu() for k, v in dan.dann.domain_clf.state_dict().items()})
        for k, v in post.items():
            want = d[f"e{e}/post/" + k]
            if k.endswith("num_batches_tracked"):
                assert int(v) == int(want), k
            else:
                assert GU.rel_err(v, want) <= TOL, (e, k, GU.rel_err(v, want))
