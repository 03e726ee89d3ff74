"""Pin the oracle: CPU restatement (oracle/scnym_oracle.py) vs golden vectors produced by
the UNMODIFIED reference (tests/golden/make_golden.py).  Tolerance: 2e-5 norm-relative
(both sides are fp32 torch-CPU; only summation order differs)."""
import numpy as np
import pytest
import torch

from oracle import scnym_oracle as O
from tests import golden_util as GU

TOL = 2e-5


def cfg_is_adam(d):
    return str(d["meta_opt"]) in ("adam", "adamw")


def _check(name, got, want, tol=TOL):
    e = GU.rel_err(got, want)
    assert e <= tol, f"{name}: rel err {e:.3e} > {tol}"


@pytest.mark.parametrize("case", GU.MIXMATCH_CASES)
def test_mixmatch_step_matches_reference(case):
    d = GU.load(case)
    model_st, dan_st = GU.split_state(GU.state_from(d, "init/"))
    model = O.OracleModel(model_st, n_layers=int(d["meta_n_layers"]), dan_state=dan_st)
    opt = O.OracleOptimizer(model.parameters(), GU.optim_config(d))
    cfg = GU.step_config(d)
    teacher = None
    # Adam divides by sqrt(v): entries whose true gradient is ~0 (e.g. Linear biases feeding
    # BatchNorm) turn summation-order noise into O(lr) updates, in the reference as well.
    # End-to-end post-step parameters are therefore held to a looser bound for adam/adamw;
    # the update rule itself is pinned exactly below by replaying the golden gradients.
    post_tol = 2e-3 if cfg_is_adam(d) else TOL
    replay_model = O.OracleModel(model_st, n_layers=int(d["meta_n_layers"]), dan_state=dan_st)
    replay_opt = O.OracleOptimizer(replay_model.parameters(), GU.optim_config(d))
    for s in range(int(d["meta_n_steps"])):
        inp = GU.step_inputs(d, s)
        out, grads, teacher = O.train_step(model, opt, cfg, inp["Lx"], inp["Ly"], inp["Ux"], inp["rnd"],
                                           inp["Ldom"], inp["Udom"], teacher=teacher, mm_step=s)
        pre = f"step{s}/"
        _check("sup", out["sup_loss"].detach(), d[pre + "sup_loss"])
        _check("unsup", out["unsup_loss"].detach(), d[pre + "unsup_loss"])
        if cfg.use_dan:
            _check("dan", out["dan_loss"].detach(), d[pre + "dan_loss"])
            _check("dan_acc", out["dan_acc"], d[pre + "dan_acc"])
        _check("labeled_logits", out["labeled_logits"].detach(), d[pre + "labeled_logits"])
        _check("pseudolabels", out["pseudolabels"].float(), d[pre + "pseudolabels"])
        assert np.array_equal(out["confident"].numpy(), d[pre + "confident"])
        for n, g in grads.items():
            _check("grad/" + n, g, d[pre + "grad/" + n], tol=5e-5)
        for n, p in model.named_parameters():
            if cfg_is_adam(d) and GU.is_pre_bn_bias(n):
                continue  # zero true gradient: Adam update is pure rounding noise (see above)
            _check("post/" + n, p.detach(), d[pre + "post/" + n], tol=post_tol)
        replay_opt.step([torch.from_numpy(d[pre + "grad/" + n]) for n, _ in replay_model.named_parameters()])
        for n, p in replay_model.named_parameters():
            _check("replay-post/" + n, p.detach(), d[pre + "post/" + n], tol=2e-6)
        for n, b in model.buffers().items():
            if n.endswith("num_batches_tracked"):
                assert int(b) == int(d[pre + "post/" + n])
            else:
                _check("post/" + n, b, d[pre + "post/" + n])
        if cfg_is_adam(d):
            # re-synchronise parameters so that the next step is compared from the same
            # starting point (otherwise the noise-driven pre-BN biases leak into the
            # eval-mode teacher); optimizer state stays the oracle's own.
            with torch.no_grad():
                for n, p in model.named_parameters():
                    p.copy_(torch.from_numpy(d[pre + "post/" + n]))


def test_supervised_step_matches_reference():
    d = GU.load("supervised_adadelta")
    model = O.OracleModel(GU.state_from(d, "init/"), n_layers=int(d["meta_n_layers"]))
    opt = O.OracleOptimizer(model.parameters(), GU.optim_config(d))
    G, C = int(d["meta_G"]), int(d["meta_C"])
    for s in range(int(d["meta_n_steps"])):
        pre = f"step{s}/"
        rows = d[pre + "rows"]
        x = O.csr_to_dense(d["lab_indptr"], d["lab_indices"], d["lab_data"], G, rows)
        y = torch.nn.functional.one_hot(torch.from_numpy(d["lab_y"][rows]), num_classes=C).float()
        keep = [torch.from_numpy(d[pre + f"keep{i}"]) for i in range(GU.n_app(d))]
        out, grads = O.supervised_step(model, opt, x, y, keep)
        _check("loss", out["loss"], d[pre + "loss"])
        _check("logits", out["logits"], d[pre + "logits"])
        for n, g in grads.items():
            _check("grad/" + n, g, d[pre + "grad/" + n], tol=5e-5)
        for n, p in model.named_parameters():
            _check("post/" + n, p.detach(), d[pre + "post/" + n])


def test_predict_matches_reference():
    d = GU.load("predict_ensemble")
    models = [O.OracleModel(GU.state_from(d, f"model{m}/"), n_layers=int(d["meta_n_layers"]))
              for m in range(int(d["meta_n_models"]))]
    X = O.csr_to_dense(d["indptr"], d["indices"], d["data"], int(d["meta_G"]))
    pred, prob, score, emb = O.predict(models, X, batch_size=64)
    assert np.array_equal(pred.numpy(), d["pred"])
    _check("prob", prob, d["prob"])
    _check("score", score, d["score"])
    _check("embed", emb, d["embed"])


def test_small_functions_match_reference():
    d = GU.load("small_functions")
    q = torch.from_numpy(d["sharpen_q"])
    for T in (0.0, 0.5, 1.0, 2.0):
        _check(f"sharpen{T}", O.sharpen_labels(q.clone(), T).float(), d[f"sharpen_T{T}"].astype(np.float32))
    for i in range(4):
        ramp, burn, mx, sig = d[f"icl{i}_cfg"]
        got = [O.icl_weight(int(e), int(ramp), int(burn), float(mx), bool(sig)) for e in range(40)]
        assert np.allclose(got, d[f"icl{i}"], rtol=1e-12, atol=0)
    z, y, cw = (torch.from_numpy(d[k]) for k in ("ce_z", "ce_y", "ce_cw"))
    for red in ("mean", "sum", "none"):
        _check("ce", O.cross_entropy(z, y, reduction=red), d[f"ce_{red}"])
        _check("ce_cw", O.cross_entropy(z, y, class_weight=cw, reduction=red), d[f"ce_cw_{red}"])
    _check("aug", O.log1p_drop(torch.from_numpy(d["aug_x"]), torch.from_numpy(d["aug_keep"])), d["aug_out"])
    ridx, g = O.mixup_indices(torch.from_numpy(d["mix_keys"]), torch.from_numpy(d["mix_graw"]), 10)
    assert np.array_equal(ridx.numpy(), d["mix_ridx"])
    a, b = torch.from_numpy(d["mix_a"]), torch.from_numpy(d["mix_b"])
    _check("mix_in", g * a + (1 - g) * a[ridx], d["mix_in"])
    _check("mix_out", g * b + (1 - g) * b[ridx], d["mix_out"])


def test_icl_weight_rejects_non_int():
    with pytest.raises(TypeError):
        O.icl_weight(1.0, 10)


@pytest.mark.parametrize("name", ["adadelta", "adam", "adamw", "sgd"])
def test_oracle_optimizer_matches_torch_optim(name):
    """The restated update rules against torch.optim itself."""
    torch.manual_seed(0)
    p0 = [torch.randn(7, 5), torch.randn(5)]
    ref_p = [p.clone().requires_grad_(True) for p in p0]
    kw = dict(lr=0.05 if name != "adadelta" else 1.0, weight_decay=1e-2)
    cls = {"adadelta": torch.optim.Adadelta, "adam": torch.optim.Adam, "adamw": torch.optim.AdamW,
           "sgd": torch.optim.SGD}[name]
    ropt = cls(ref_p, momentum=0.9, **kw) if name == "sgd" else cls(ref_p, **kw)
    mine = [p.clone() for p in p0]
    oopt = O.OracleOptimizer(mine, O.OptimConfig(name=name, lr=kw["lr"], weight_decay=kw["weight_decay"]))
    for t in range(5):
        gs = [torch.randn_like(p) for p in p0]
        for p, g in zip(ref_p, gs):
            p.grad = g.clone()
        ropt.step()
        oopt.step(gs)
        for a, b in zip(mine, ref_p):
            assert GU.rel_err(a, b.detach()) < 1e-6


def test_oracle_bn_matches_torch_batchnorm():
    torch.manual_seed(1)
    bn = torch.nn.BatchNorm1d(6)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-1, 1)
    blk = O.Block(W=None, b=None, g=bn.weight.detach().clone(), beta=bn.bias.detach().clone(),
                  rm=bn.running_mean.clone(), rv=bn.running_var.clone(), nbt=bn.num_batches_tracked.clone())
    z = torch.randn(13, 6) * 3 + 1
    bn.train()
    want = bn(z)
    got = O.OracleModel._bn_train(z, blk)
    assert GU.rel_err(got, want.detach()) < 1e-6
    assert GU.rel_err(blk.rm, bn.running_mean) < 1e-6 and GU.rel_err(blk.rv, bn.running_var) < 1e-6
    bn.eval()
    assert GU.rel_err(O.OracleModel._bn_eval(z, blk), bn(z).detach()) < 1e-6
