"""GPU parity tests: the CUDA engine (through the C ABI) against
  (a) golden vectors produced by the unmodified reference (tests/golden/*.npz), and
  (b) the CPU oracle on freshly seeded inputs at other sizes,
with identical injected randomness.  Tolerance (BASELINE.json north_star): logits, losses and
gradients within 1e-4 norm-relative in fp32; argmax / confident flags exact."""
import numpy as np
import pytest
import torch

from oracle import scnym_oracle as O
from tests import golden_util as GU

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _chk(name, got, want, tol=TOL):
    e = GU.rel_err(got, want)
    assert e <= tol, f"{name}: rel err {e:.3e} > {tol}"


def _is_adam(d):
    return str(d["meta_opt"]) in ("adam", "adamw")


def _run_and_compare(d):
    from tests import cuda_harness as CH
    eng = CH.build_engine(d)
    use_dan = bool(int(d["meta_use_dan"]))
    for s in range(int(d["meta_n_steps"])):
        CH.inject_step(eng, d, s)
        eng.step()
        got = CH.collect(eng)
        pre = f"step{s}/"
        _chk("sup_loss", got["sup_loss"], d[pre + "sup_loss"])
        _chk("unsup_loss", got["unsup_loss"], d[pre + "unsup_loss"])
        if use_dan:
            _chk("dan_loss", got["dan_loss"], d[pre + "dan_loss"])
            _chk("dan_acc", got["dan_acc"], d[pre + "dan_acc"])
        _chk("labeled_logits", got["labeled_logits"], d[pre + "labeled_logits"])
        assert np.array_equal(got["labeled_logits"].argmax(1).numpy(), d[pre + "labeled_logits"].argmax(1))
        _chk("pseudolabels", got["pseudolabels"], d[pre + "pseudolabels"])
        assert np.array_equal(got["confident"], d[pre + "confident"])
        for n, g in got["grads"].items():
            _chk("grad/" + n, g, d[pre + "grad/" + n])
        # Post-step parameters.  Adadelta / SGD: every entry at 1e-4.  Adam / AdamW: the first update is
        # lr*g/(|g|+eps), so an entry whose gradient is summation noise moves by O(lr) in either direction (in the
        # reference on other hardware just the same).  Instead of widening the tolerance for every entry, (i) the
        # update RULE is pinned on all entries by replaying the oracle optimizer on the device's own gradients, and
        # (ii) the parameters are compared at 1e-4 wherever the reference gradient is above 1e-3 of the tensor's
        # largest (tests/cuda_harness.py::compare_step0 derives the threshold).
        for n, v in got["post"].items():
            key = pre + "post/" + n
            if key not in d.files:
                continue
            if n.endswith("num_batches_tracked"):
                assert int(v) == int(d[key]), n
                continue
            if "running" in n or not _is_adam(d):
                _chk("post/" + n, v, d[key])
                continue
            gref = torch.from_numpy(np.array(d[pre + "grad/" + n])).double().abs()
            if float(gref.max()) < 1e-7:
                continue
            m = gref > 1e-3 * float(gref.max())
            want = torch.from_numpy(np.array(d[key])).double()
            e = float((v.double()[m] - want[m]).abs().max()) / float(want.abs().max())
            assert e <= TOL, f"post/{n} on |g| > 1e-3 max|g|: rel err {e:.3e}"
        if _is_adam(d) and s == 0:
            model_st, dan_st = GU.split_state(GU.state_from(d, "init/"))
            rm = O.OracleModel(model_st, n_layers=int(d["meta_n_layers"]), dan_state=dan_st)
            ro = O.OracleOptimizer(rm.parameters(), GU.optim_config(d))
            ro.step([got["grads"][n].clone() for n, _ in rm.named_parameters()])
            for n, p in rm.named_parameters():
                _chk("optimizer replay post/" + n, got["post"][n], p.detach(), tol=1e-5)
        if _is_adam(d):
            sd = {k: torch.from_numpy(np.array(d[pre + "post/" + k])) for k in eng.model.state_dict().keys()
                  if "running" not in k and "num_batches" not in k}
            eng.model.load_state_dict(sd, strict=False)
            if use_dan:
                eng.dann.domain_clf.load_state_dict(
                    {k: torch.from_numpy(np.array(d[pre + "post/domain_clf." + k])) for k in eng.dann.domain_clf.state_dict().keys()})
            eng.arena.check_aliasing()


@pytest.mark.parametrize("case", GU.MIXMATCH_CASES)
def test_engine_step_matches_reference_golden(case):
    _run_and_compare(GU.load(case))


@pytest.mark.parametrize("kw", [
    dict(G=2000, L=64, U=64, C=8, H0=128, H=128, n_layers=2, opt_name="adam", n_steps=2, seed=11),
    dict(G=1500, L=48, U=80, C=6, H0=256, H=64, n_layers=3, K=2, aug_pl=True, D_given=4, opt_name="adadelta", lr=1.0, seed=12),
    dict(G=1000, L=40, U=56, C=5, H0=96, H=72, n_layers=2, tau=0.22, dan_conf=True, dan_scale=True, opt_name="sgd", lr=0.05,
         mean_teacher=True, n_steps=2, seed=13),
    dict(G=777, L=33, U=17, C=3, H0=40, H=24, n_layers=1, opt_name="adamw", lr=1e-3, wd=1e-2, seed=14),
    # fused heads (H % 128 == 0) with thresholded pseudolabels, confident-only DAN rows and given domains
    dict(G=1200, L=40, U=48, C=7, H0=128, H=256, n_layers=2, tau=0.2, dan_conf=True, dan_scale=True, D_given=4, opt_name="adam",
         n_steps=2, seed=15),
    dict(G=900, L=24, U=40, C=4, H0=64, H=128, n_layers=3, K=2, aug_pl=True, opt_name="adadelta", lr=1.0, seed=16),
])
def test_engine_step_matches_oracle(kw):
    from tests import cuda_harness as CH
    _run_and_compare(CH.oracle_case(**kw))


def test_supervised_engine_matches_reference_golden():
    from scnym_b200.dataprep import DeviceCSR
    from scnym_b200.engine import TrainEngine, EngineConfig
    from scnym_b200.model import CellTypeCLF
    d = GU.load("supervised_adadelta")
    G, C, H0, H, nl, B = (int(d[k]) for k in ("meta_G", "meta_C", "meta_H0", "meta_H", "meta_n_layers", "meta_B"))
    model = CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H0, n_layers=nl)
    model.load_state_dict(GU.state_from(d, "init/"))
    model = model.cuda()
    lab = DeviceCSR.from_arrays(d["lab_indptr"], d["lab_indices"], d["lab_data"], G)
    cfg = EngineConfig(optimizer="adadelta", lr=float(d["meta_lr"]), weight_decay=float(d["meta_wd"]), use_cuda_graph=False, keep_w1_grad=True)
    eng = TrainEngine(model, cfg, lab, d["lab_y"], B, mode="supervised")
    for s in range(int(d["meta_n_steps"])):
        pre = f"step{s}/"
        eng.load_batch(d[pre + "rows"])
        eng.inj_hidden_keep = [torch.from_numpy(d[pre + f"keep{a}"]).cuda() for a in range(eng.n_app)]
        eng.step()
        torch.cuda.synchronize()
        _chk("loss", eng.losses()["loss"], d[pre + "loss"])
        _chk("logits", eng.logits.cpu(), d[pre + "logits"])
        for n, p in model.named_parameters():
            _chk("grad/" + n, p.grad.cpu(), d[pre + "grad/" + n])
        for n, v in model.state_dict().items():
            if n.endswith("num_batches_tracked"):
                assert int(v) == int(d[pre + "post/" + n])
            else:
                _chk("post/" + n, v.cpu(), d[pre + "post/" + n])


def test_predict_matches_reference_golden(tmp_path):
    from scnym_b200.model import CellTypeCLF
    from scnym_b200.predict import Predicter
    from scipy import sparse
    d = GU.load("predict_ensemble")
    G, C, H0, H, nl, N = (int(d[k]) for k in ("meta_G", "meta_C", "meta_H0", "meta_H", "meta_n_layers", "meta_N"))
    paths = []
    for m in range(int(d["meta_n_models"])):
        # checkpoints in the REFERENCE's on-disk format (plain state_dict, contiguous [H0, G] first weight)
        p = str(tmp_path / f"w{m}.pkl")
        torch.save(GU.state_from(d, f"model{m}/"), p)
        paths.append(p)
    P = Predicter(model_weights=paths, n_genes=G, n_cell_types=C, labels=[f"c{i}" for i in range(C)], n_hidden=H,
                  n_hidden_init=H0, n_layers=nl)
    X = sparse.csr_matrix((d["data"], d["indices"], d["indptr"]), shape=(N, G))
    pred, names, prob = P.predict(X, output="prob")
    assert np.array_equal(pred, d["pred"])
    assert names == [f"c{i}" for i in d["pred"]]
    _chk("prob", prob, d["prob"])
    _, _, score = P.predict(X, output="score")
    _chk("score", score, d["score"])
    _, _, _, emb = P.predict_device(X, want_embed=True)
    _chk("embed", emb.cpu(), d["embed"])
    # dense ndarray input gives the same answer (SingleCellDS accepts both, dataprep.py:67-72)
    pred2, _ = P.predict(np.asarray(X.todense(), dtype=np.float32))
    assert np.array_equal(pred2, d["pred"])
    with pytest.raises(ValueError):
        P.predict(X[:, :-1])


def test_module_level_autograd_path_matches_oracle():
    """CellTypeCLF.forward / DANN.forward composed with torch autograd (the path reference-style
    criteria and `torch.optim` use) against the oracle, dense and CSR inputs."""
    from oracle import scnym_oracle as O
    from scnym_b200 import functional as Fn
    from scnym_b200.dataprep import DeviceCSR, gather_rows
    from scnym_b200.model import CellTypeCLF, DANN
    torch.manual_seed(3)
    G, C, H0, H, n = 300, 4, 48, 32, 20
    ip, ii, dd, y = O.synth_csr(n, G, 0.1, C, seed=5)
    X = O.csr_to_dense(ip, ii, dd, G)
    model = CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H0, n_layers=2)
    dann = DANN(model, n_domains=2)
    st = {k: v.clone() for k, v in model.state_dict().items()}
    dst = {"domain_clf." + k: v.clone() for k, v in dann.domain_clf.state_dict().items()}
    om = O.OracleModel(st, n_layers=2, dan_state=dst, requires_grad=True)
    keep = [(torch.rand(n, w) < 0.5).to(torch.uint8) for w in (H0, H, H)]
    Y = torch.nn.functional.one_hot(torch.from_numpy(y), C).float()
    dl = torch.nn.functional.one_hot(torch.arange(n) % 2, 2).float()
    # oracle
    lo = om.forward(X, train=True, drop_keep=keep)
    dlo, _ = om.dann_forward(X, True, keep, 0.3)
    (O.cross_entropy(lo, Y) + O.cross_entropy(dlo, dl)).backward()
    want = {nme: p.grad.clone() for nme, p in om.named_parameters()}
    # CUDA, dense input then CSR input
    model.cuda(); dann.domain_clf.cuda()
    ds = DeviceCSR.from_arrays(ip, ii, dd, G)
    for kind in ("dense", "csr"):
        model.zero_grad(); dann.zero_grad()
        model.train()
        xin = X.cuda() if kind == "dense" else gather_rows(ds)
        Fn.DROPOUT_INJECT.extend([k.clone() for k in keep])
        logits = model(xin)
        Fn.DROPOUT_INJECT.extend([k.clone() for k in keep])
        dann.set_rev_grad_weight(0.3)
        dlogits, _ = dann(xin)
        lsm = torch.log_softmax(logits, 1)
        dsm = torch.log_softmax(dlogits, 1)
        loss = -(Y.cuda() * lsm).sum(1).mean() - (dl.cuda() * dsm).sum(1).mean()
        loss.backward()
        _chk(kind + "/logits", logits.detach().cpu(), lo.detach())
        _chk(kind + "/dlogits", dlogits.detach().cpu(), dlo.detach())
        for nme, p in model.named_parameters():
            _chk(kind + "/grad/" + nme, p.grad.cpu(), want[nme])
        for nme, p in dann.domain_clf.named_parameters():
            _chk(kind + "/grad/dan/" + nme, p.grad.cpu(), want["domain_clf." + nme])
    # eval mode + embedding
    model.eval()
    with torch.no_grad():
        le = model(X.cuda())
        emb = Fn.embed_forward(model, X.cuda(), pre_relu_last=True)
    om2 = O.OracleModel({k: v.cpu() for k, v in model.state_dict().items()}, n_layers=2)
    _chk("eval/logits", le.cpu(), om2.forward(X, train=False))
    _chk("eval/embed", emb.cpu(), om2.embed(X, train=False, pre_relu_last=True))


def test_cuda_graph_step_equals_eager_step():
    """Graph replay and eager launches of the same step (Philox dropout, same seed/counters) agree bit for bit
    except for the atomically accumulated dW1 / loss scalars (fp32 atomics: order-dependent)."""
    from tests import cuda_harness as CH
    d = CH.oracle_case(G=600, L=32, U=32, C=4, H0=64, H=64, n_layers=2, opt_name="adam", n_steps=1, seed=21)
    outs = []
    for graph in (False, True):
        eng = CH.build_engine(d)
        eng.cfg.use_cuda_graph = graph
        for s in range(3):
            eng.load_batch(d["step0/lrow"], d["step0/urow"], d["step0/perm_keys"], d["step0/gamma_raw"])
            eng.step()
        torch.cuda.synchronize()
        outs.append({k: v.detach().cpu().clone() for k, v in eng.model.state_dict().items()})
        assert int(eng.state[0]) == 3 and eng.kernels_per_step > 10
    for k in outs[0]:
        if "num_batches" in k:
            assert int(outs[0][k]) == int(outs[1][k]) == 9
        else:
            # fp32 atomics in the dW1 scatter make two runs differ in the last bits of the first-step gradient;
            # Adam then amplifies that on near-zero-gradient entries (see tests/test_oracle.py), so compare loosely.
            _chk(k, outs[1][k], outs[0][k], tol=2e-3)


def test_engine_step_with_tensor_core_first_layer_matches_reference_golden():
    """first_layer="tc" (densify-then-tcgen05 GEMM, bf16x3): losses, logits and every gradient within 1e-4 of the
    reference.  Post-step parameters are not compared here: Adam's first update is lr*sign(g), so the ~1e-5 relative
    noise of the split product flips the update of near-zero gradient entries (see tests/test_oracle.py)."""
    from tests import cuda_harness as CH
    d = GU.load("mixmatch_dan_adam")
    eng = CH.build_engine(d)
    if int(d["meta_H0"]) % 32 != 0:
        pytest.skip("golden case H0 not a multiple of 32")
    import scnym_b200.engine as E
    eng2_cfg = eng.cfg
    eng2_cfg.first_layer = "tc"
    eng2_cfg.w1_update = "tc"      # tcgen05 dW1 + optimizer kernel as well
    model, dann = CH.build_models(d)
    from scnym_b200.dataprep import DeviceCSR
    G = int(d["meta_G"])
    lab = DeviceCSR.from_arrays(d["lab_indptr"], d["lab_indices"], d["lab_data"], G)
    unl = DeviceCSR.from_arrays(d["unl_indptr"], d["unl_indices"], d["unl_data"], G)
    given = bool(int(d["meta_D_given"]))
    eng = E.TrainEngine(model, eng2_cfg, lab, d["lab_y"], int(d["meta_L"]), unlabeled=unl, unlabeled_batch_size=int(d["meta_U"]),
                        dann=dann, labeled_domain=d["lab_dom"] if given else None, unlabeled_domain=d["unl_dom"] if given else None,
                        mode="mixmatch")
    assert eng.tc_first and eng.w1_tc
    eng.set_weights(float(d["meta_unsup_w"]), float(d["meta_dan_w"]))
    CH.inject_step(eng, d, 0)
    eng.step()
    got = CH.collect(eng)
    pre = "step0/"
    for k in ("sup_loss", "unsup_loss", "dan_loss"):
        _chk(k, got[k], d[pre + k])
    _chk("labeled_logits", got["labeled_logits"], d[pre + "labeled_logits"])
    _chk("pseudolabels", got["pseudolabels"], d[pre + "pseudolabels"])
    for n, g in got["grads"].items():
        _chk("grad/" + n, g, d[pre + "grad/" + n])


@pytest.mark.parametrize("opt_name,lr,n_steps", [("sgd", 0.05, 2), ("adadelta", 1.0, 1)])
def test_full_size_tensor_core_step_agrees_with_fp32_step(opt_name, lr, n_steps):
    """BASELINE config-2 shape (G = 20 000 genes, L = U = 256, H0 = H = 256, 5 % CSR, MixMatch + DAN), too large for the CPU
    oracle to finish in seconds: the production path (tcgen05 first layer, tcgen05 dW1 + optimizer with fused operand
    planes, graph replay) against the exact-fp32 path of the same engine (warp-per-row SpMM, SIMT dW1 + optimizer, eager)
    on identical batches and Philox streams.  Two steps, so that the second forward runs on the operand planes the first
    step's optimizer epilogue emitted: losses, logits, first-layer outputs and gradients, every parameter and BatchNorm
    buffer agree within the 1e-4 budget (measured: <= 1e-5) and the predicted classes of the labeled rows are identical.
    (Longer runs are not comparable element-wise: a ~1e-5 difference in a pre-activation that lies within 1e-5 of zero
    eventually flips a ReLU gate in one of the two runs -- observed at the third SGD step and, with Adadelta's large
    first update, at the second -- and the trajectories part; Adadelta is therefore compared after one step.)"""
    import bench
    from scnym_b200.engine import EngineConfig, TrainEngine
    from scnym_b200.model import CellTypeCLF, DANN
    from scnym_b200.synth import synth_device_csr
    c = bench.CONFIGS[2][0]
    dev = torch.device("cuda", 0)
    lab, y = synth_device_csr(4096, c["G"], c["density"], c["C"], seed=0, device=dev)
    unl, _ = synth_device_csr(4096, c["G"], c["density"], c["C"], seed=1, device=dev)
    runs = []
    for fl, wu, graph in (("spmm", "simt", False), ("tc", "tc", True)):
        torch.manual_seed(0)
        model = CellTypeCLF(n_genes=c["G"], n_cell_types=c["C"], n_hidden=c["H"], n_hidden_init=c["H0"], n_layers=c["n_layers"])
        dann = DANN(model, n_domains=c["D"])
        cfg = EngineConfig(n_augmentations=1, T=0.5, augment_pseudolabels=False, pseudolabel_min_confidence=0.0, mixup_alpha=0.3,
                           use_dan=True, optimizer=opt_name, lr=lr, weight_decay=1e-4, seed=77, first_layer=fl, w1_update=wu,
                           use_cuda_graph=graph)
        eng = TrainEngine(model.to(dev), cfg, lab, y.to(torch.int32), c["L"], unlabeled=unl, unlabeled_batch_size=c["U"], dann=dann,
                          mode="mixmatch")
        assert eng.tc_first == (fl == "tc") and eng.w1_tc == (wu == "tc") and eng.w1_emits_planes == (wu == "tc")
        eng.set_weights(1.0, 0.1)
        gen = torch.Generator(device=dev)
        gen.manual_seed(5)
        eng.sample_epoch_tables(n_steps, generator=gen)
        losses = []
        for _ in range(n_steps):
            eng.step()
            losses.append(eng.losses())
        torch.cuda.synchronize()
        runs.append(dict(losses=losses, Z1=eng.Z1.detach().cpu().clone(), dZ1=eng.dZ1.detach().cpu().clone(),
                         logits=eng.logits.detach().cpu().clone(),
                         params={n: p.detach().cpu().clone() for n, p in model.named_parameters()},
                         bn={n: b.detach().cpu().clone() for n, b in model.named_buffers() if "running" in n}))
    a, b = runs
    for s in range(n_steps):
        for k in ("sup_loss", "unsup_loss", "dan_loss"):
            _chk(f"step{s}/{k}", torch.tensor(b["losses"][s][k]), torch.tensor(a["losses"][s][k]))
    if float(b["Z1"].abs().max()) > 0:   # (replayed steps on prefetched batches clear Z1 for their successor at the end)
        _chk("Z1", b["Z1"], a["Z1"])
    _chk("dZ1", b["dZ1"], a["dZ1"])
    _chk("logits", b["logits"], a["logits"])
    assert torch.equal(b["logits"][c["U"]:].argmax(1), a["logits"][c["U"]:].argmax(1))
    for n in a["params"]:
        _chk("param/" + n, b["params"][n], a["params"][n])
    for n in a["bn"]:
        _chk("bn/" + n, b["bn"][n], a["bn"][n])
