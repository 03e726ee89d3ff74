"""GPU parity at the BASELINE.json shapes, against the CPU ORACLE (not against another path of this package):
the production kernels (tcgen05 first layer, tcgen05 dW1 + optimizer, fused hidden stack, tcgen05 hidden GEMMs where the
shape selects them) run one injected step and are compared with `oracle.train_step` on the same dense batch, the same
weights and the same dropout / MixUp randomness.

Tolerance (north_star): 1e-4 norm-relative on losses, logits, pseudolabels, every gradient, BatchNorm buffers and
post-step parameters; predicted classes exact.  ReLU decisions are compared one by one: an entry may be decided
differently only where the oracle's pre-activation is below 1e-4 of the tensor's largest (rounding noise -- its sign
depends on summation order on any hardware); such entries are counted, must stay under 2e-5 of all decisions, and the
gradients are then compared with the oracle re-run on the device's decisions (tests/cuda_harness.py::compare_step0)."""
import numpy as np
import pytest
import torch

from tests import golden_util as GU

pytestmark = pytest.mark.gpu


def _run(kw, expect_tc=True, **eng_kw):
    from tests import cuda_harness as CH
    d = CH.oracle_case(**kw)
    eng = CH.build_engine(d, **eng_kw)
    if expect_tc:
        assert eng.tc_first and eng.w1_tc, "the production tensor-core kernels were not selected for this shape"
    CH.inject_step(eng, d, 0)
    eng.step()
    rep = CH.compare_step0(eng, d)
    print({k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in rep.items()})
    return eng, rep


def test_config2_full_shape_step_matches_oracle():
    """BASELINE config 2: G = 20 000, L = U = 256, H0 = H = 256, n_layers = 2, 5 % CSR, MixMatch + DAN (2 inferred domains), Adam."""
    eng, _ = _run(dict(G=20000, L=256, U=256, C=16, H0=256, H=256, n_layers=2, opt_name="adam", lr=1e-3, wd=1e-4, seed=31,
                       density=0.05, n_cells=768))
    assert eng.w1_emits_planes and eng.hs_fused and eng.heads_fused


def test_config2_full_shape_split_heads_match_oracle(monkeypatch):
    """The same shape with the heads as separate launches behind scnb_hs_act (SCNB_HEADS=split: the chain head_ops.cu replaces,
    still taken for widths other than 128 / 256 or more than 32 classes / domains)."""
    monkeypatch.setenv("SCNB_HEADS", "split")
    eng, _ = _run(dict(G=20000, L=256, U=256, C=16, H0=256, H=256, n_layers=2, opt_name="adam", lr=1e-3, wd=1e-4, seed=37,
                       density=0.05, n_cells=768))
    assert eng.hs_fused and not eng.heads_fused


def test_config2_shape_unfused_hidden_stack_matches_oracle():
    """The same shape through the unfused BatchNorm / GEMM kernels (the path taken under pseudolabel thresholding or data
    parallelism)."""
    from tests import cuda_harness as CH
    d = CH.oracle_case(G=20000, L=256, U=256, C=16, H0=256, H=256, n_layers=2, opt_name="adam", lr=1e-3, wd=1e-4, seed=36,
                       density=0.05, n_cells=768)
    import os
    os.environ["SCNB_HIDDEN"] = "unfused"
    try:
        eng = CH.build_engine(d)
    finally:
        del os.environ["SCNB_HIDDEN"]
    assert not eng.hs_fused and eng.tc_first
    CH.inject_step(eng, d, 0)
    eng.step()
    CH.compare_step0(eng, d)


@pytest.mark.parametrize("kw", [
    dict(G=1500, L=32, U=64, C=6, H0=128, H=64, n_layers=3, D_given=4, opt_name="adadelta", lr=1.0, seed=51),   # shared third block
    dict(G=900, L=64, U=32, C=4, H0=64, H=128, n_layers=1, opt_name="sgd", lr=0.05, seed=52),                     # one hidden Linear
    dict(G=1200, L=96, U=96, C=5, H0=192, H=192, n_layers=2, use_dan=False, opt_name="adamw", lr=1e-3, wd=1e-2, seed=53),  # no adversary
    dict(G=800, L=32, U=32, C=3, H0=512, H=256, n_layers=2, opt_name="adam", seed=54),                             # K = 512 operand panel
    dict(G=1000, L=64, U=96, C=32, H0=128, H=128, n_layers=2, D_given=5, opt_name="adam", seed=55),                # fused heads, cluster of 2, 32 classes
    dict(G=1100, L=96, U=32, C=7, H0=256, H=256, n_layers=2, use_dan=False, opt_name="adadelta", lr=1.0, seed=56),  # fused classifier head alone
])
def test_fused_hidden_stack_shapes_match_oracle(kw):
    """Other shapes of the fused hidden stack (hidden_ops.cu): shared modules, 2-4 applications, 2 / 3 segments,
    operand panels of 64..512 columns, ragged numbers of 32-row tiles per segment."""
    eng, _ = _run(kw, expect_tc=False)
    assert eng.hs_fused


def test_config2_full_shape_adadelta_matches_oracle():
    """Same shape with the reference's default optimizer (Adadelta, lr 1.0)."""
    _run(dict(G=20000, L=256, U=256, C=16, H0=256, H=256, n_layers=2, opt_name="adadelta", lr=1.0, wd=1e-4, seed=32,
              density=0.05, n_cells=768))


def test_config3_shape_step_matches_oracle():
    """BASELINE config 3 (per-rank shape): G = 30 000, 8 GIVEN domains, L = U = 256."""
    eng, _ = _run(dict(G=30000, L=256, U=256, C=16, H0=256, H=256, n_layers=2, D_given=8, opt_name="adam", lr=1e-3, wd=1e-4,
                       seed=33, density=0.05, n_cells=768))
    assert eng.D == 8 and eng.hs_fused and eng.heads_fused


def test_config5_shape_step_matches_oracle():
    """BASELINE config 5 shape with the rows reduced (L = U = 512 so that the oracle stays in seconds): H = 1024,
    n_layers = 3 (the last two hidden blocks share one Linear / BatchNorm module), 20 % density, 8 given domains, and
    the tcgen05 hidden-GEMM kernel (tc_gemm.cu) forced on for every product of the stack."""
    from scnym_b200 import _lib
    _lib.call("scnb_gemm_tc_min_flops", 1)
    try:
        eng, _ = _run(dict(G=20000, L=512, U=512, C=16, H0=256, H=1024, n_layers=3, D_given=8, opt_name="adam", lr=1e-3,
                           wd=1e-4, seed=34, density=0.20, n_cells=1536))
        assert eng.n_app == 4 and eng.H == 1024
        import ctypes
        n = ctypes.c_longlong(0)
        _lib.call("scnb_gemm_tc_planes_count", ctypes.byref(n))
        assert eng._gemm_ws and n.value > 0      # the planes form of the tcgen05 products ran inside the engine
    finally:
        _lib.call("scnb_gemm_tc_min_flops", int(2e9))


def test_config5_h0_1024_shape_step_matches_oracle():
    """Config 5 with n_hidden_init = 1024 (SURVEY §8d "report both")."""
    from scnym_b200 import _lib
    _lib.call("scnb_gemm_tc_min_flops", 1)
    try:
        eng, _ = _run(dict(G=8000, L=512, U=512, C=16, H0=1024, H=1024, n_layers=3, D_given=8, opt_name="adam", lr=1e-3,
                           wd=1e-4, seed=35, density=0.20, n_cells=1536), expect_tc=False)
        assert eng.widths[0] == 1024
    finally:
        _lib.call("scnb_gemm_tc_min_flops", int(2e9))


def test_config4_shape_predict_matches_oracle():
    """BASELINE config 4 shape (G = 20 000, H0 = H = 256, 5 %) through the production predict path (tcgen05 first layer
    on bf16x3 operand planes, fused eval BatchNorm, tcgen05 hidden layer): classes exact, probabilities / scores /
    embedding within 1e-4 of the oracle on 4096 cells."""
    from oracle import scnym_oracle as O
    from scnym_b200.dataprep import DeviceCSR
    from scnym_b200.model import CellTypeCLF
    from scnym_b200.predict import EvalRunner
    G, C, H, N = 20000, 16, 256, 4096
    ip, ii, dd, _ = O.synth_csr(N, G, 0.05, C, seed=41)
    torch.manual_seed(4)
    model = CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H, n_layers=2)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.running_mean.uniform_(-0.3, 0.3)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.uniform_(-0.2, 0.2)
    om = O.OracleModel({k: v.detach().clone() for k, v in model.state_dict().items()}, n_layers=2)
    X = O.csr_to_dense(ip, ii, dd, G)
    pred, prob, score, emb = O.predict([om], X, batch_size=1024)
    model = model.cuda().eval()
    run = EvalRunner([model], chunk_rows=2048)    # two chunks: the chunk loop is part of the path
    assert run.use_tc and all(run.tc_hidden[1:]), "the tensor-core predict path was not selected"
    ds = DeviceCSR.from_arrays(ip, ii, dd, G)
    got_pred, got_prob, got_score, got_emb = run.run(ds, want_prob=True, want_score=True, want_embed=True)
    torch.cuda.synchronize()
    # classes: exact wherever the oracle's top-2 margin is not rounding noise (counted)
    top2 = score.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]) / score.abs().max()
    diff = got_pred.cpu().numpy() != pred.numpy()
    assert not np.any(diff & (margin.numpy() > 1e-4)), "predicted class differs where the margin is not noise"
    assert diff.sum() <= 1
    for name, a, b in (("prob", got_prob, prob), ("score", got_score, score), ("embed", got_emb, emb)):
        e = GU.rel_err(a.cpu(), b)
        assert e <= 1e-4, f"{name}: rel err {e:.3e}"
