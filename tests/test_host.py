"""CPU tests of the host-side logic: model surface (state_dict / init parity), parameter arena."""
import numpy as np
import pytest
import torch

from tests import golden_util as GU


def test_state_dict_keys_match_reference_golden():
    from scnym_b200.model import CellTypeCLF, DANN
    for case in ("mixmatch_dan_adam", "mixmatch_dan_k2_adadelta"):
        d = GU.load(case)
        G, C, H0, H, nl = (int(d[k]) for k in ("meta_G", "meta_C", "meta_H0", "meta_H", "meta_n_layers"))
        m = CellTypeCLF(n_genes=G, n_cell_types=C, n_hidden=H, n_hidden_init=H0, n_layers=nl)
        ref_st, dan_st = GU.split_state(GU.state_from(d, "init/"))
        assert list(m.state_dict().keys()) == list(ref_st.keys())
        for k, v in m.state_dict().items():
            assert tuple(v.shape) == tuple(ref_st[k].shape), k
        m.load_state_dict(ref_st)  # reference checkpoint loads
        w = m.embed[0].weight
        assert w.t().is_contiguous() and torch.equal(w.detach(), ref_st["embed.0.weight"])
        dn = DANN(m, n_domains=int(d["meta_D"]))
        assert ["domain_clf." + k for k in dn.domain_clf.state_dict().keys()] == list(dan_st.keys())


def test_checkpoint_roundtrip_through_torch_save(tmp_path):
    from scnym_b200.model import CellTypeCLF
    torch.manual_seed(0)
    m = CellTypeCLF(n_genes=50, n_cell_types=4, n_hidden=16, n_hidden_init=24, n_layers=3)
    p = str(tmp_path / "w.pkl")
    torch.save(m.state_dict(), p)
    m2 = CellTypeCLF(n_genes=50, n_cell_types=4, n_hidden=16, n_hidden_init=24, n_layers=3)
    m2.load_state_dict(torch.load(p))
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    assert m2.embed[0].weight.t().is_contiguous()
    # shared trailing block (model.py:207): embed.8 and embed.12 alias one module
    assert m.embed[8] is m.embed[12] and m.embed[9] is m.embed[13]


def test_unsupported_variants_raise():
    from scnym_b200.model import CellTypeCLF
    for kw in (dict(residual=True), dict(batch_norm=False), dict(use_raw_counts=True), dict(n_decoder_layers=2)):
        with pytest.raises(NotImplementedError):
            CellTypeCLF(n_genes=10, n_cell_types=2, **kw)


def test_init_matches_reference_constructor_when_available():
    from oracle.ref_shim import reference_available, load_reference
    if not reference_available():
        pytest.skip("reference tree not present (GPU box)")
    ref = load_reference(modules=("model",))
    from scnym_b200.model import CellTypeCLF, DANN
    for nl in (1, 2, 3):
        torch.manual_seed(7)
        a = ref.model.CellTypeCLF(n_genes=120, n_cell_types=5, n_hidden=32, n_hidden_init=48, n_layers=nl)
        da = ref.model.DANN(a, n_domains=3)
        torch.manual_seed(7)
        b = CellTypeCLF(n_genes=120, n_cell_types=5, n_hidden=32, n_hidden_init=48, n_layers=nl)
        db = DANN(b, n_domains=3)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k
        assert [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()]
        for (k, x), (_, y) in zip(da.domain_clf.state_dict().items(), db.domain_clf.state_dict().items()):
            assert torch.equal(x, y), k


def test_param_arena_aliases_module_parameters():
    from scnym_b200.engine import ParamArena
    from scnym_b200.model import CellTypeCLF
    torch.manual_seed(0)
    m = CellTypeCLF(n_genes=30, n_cell_types=3, n_hidden=8, n_hidden_init=12, n_layers=3)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    arena = ParamArena([("model." + n, p) for n, p in m.named_parameters()], torch.device("cpu"))
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert arena.total % 4 == 0 and all(o % 4 == 0 for o in arena.offsets.values())
    # the first-layer weight is stored gene-major inside the arena
    w = arena.w("model.embed.0.weight")
    assert tuple(w.shape) == (30, 12) and w.is_contiguous()
    assert torch.equal(w.t(), m.embed[0].weight.detach())
    # writes through the arena are visible in the module and vice versa; grads alias too
    arena.flat.mul_(2.0)
    assert torch.equal(m.classif[0].weight.detach(), before["classif.0.weight"] * 2)
    arena.grad.fill_(1.5)
    assert float(m.embed[0].weight.grad.sum()) == 1.5 * 30 * 12
    m.load_state_dict(before)
    assert torch.equal(arena.w("model.classif.0.bias"), before["classif.0.bias"])
    arena.check_aliasing()
    m.classif[0].weight.data = torch.zeros(3, 8)
    with pytest.raises(RuntimeError):
        arena.check_aliasing()


def test_host_gather_rows_matches_numpy():
    """CPU entry points of the library (no GPU needed): row gather of host CSR matrices into staging buffers, one matrix
    (`scnb_host_gather_rows`) and the labeled+unlabeled pair in one call (`scnb_host_gather_rows2`); ragged and empty rows."""
    import numpy as np
    from scnym_b200 import _lib
    h = _lib.lib()
    rng = np.random.default_rng(0)

    def make(n, g):
        nnz = rng.integers(0, 40, n)
        nnz[::7] = 0
        indptr = np.concatenate([[0], np.cumsum(nnz)]).astype(np.int64)
        return indptr, rng.integers(0, g, indptr[-1]).astype(np.int32), rng.random(indptr[-1], dtype=np.float32)

    def want(m, rows):
        ip, ix, vv = m
        seg = [np.arange(ip[r], ip[r + 1]) for r in rows]
        cat = np.concatenate(seg) if seg else np.zeros(0, np.int64)
        return np.concatenate([[0], np.cumsum([len(s) for s in seg])]).astype(np.int64), ix[cat], vv[cat]

    A, B = make(300, 1000), make(200, 1000)
    ra, rb = rng.integers(0, 300, 64).astype(np.int64), rng.integers(0, 200, 48).astype(np.int64)
    P = lambda x: x.ctypes.data
    for threads in (1, 3, 8):
        oa = (np.zeros(65, np.int64), np.zeros(64 * 40, np.int32), np.zeros(64 * 40, np.float32))
        ob = (np.zeros(49, np.int64), np.zeros(48 * 40, np.int32), np.zeros(48 * 40, np.float32))
        assert h.scnb_host_gather_rows2(P(A[0]), P(A[1]), P(A[2]), P(ra), 64, P(oa[0]), P(oa[1]), P(oa[2]), 64 * 40,
                                        P(B[0]), P(B[1]), P(B[2]), P(rb), 48, P(ob[0]), P(ob[1]), P(ob[2]), 48 * 40, threads) == 0
        o1 = (np.zeros(65, np.int64), np.zeros(64 * 40, np.int32), np.zeros(64 * 40, np.float32))
        assert h.scnb_host_gather_rows(P(A[0]), P(A[1]), P(A[2]), P(ra), 64, P(o1[0]), P(o1[1]), P(o1[2]), 64 * 40, threads) == 0
        for got, m, rows in ((oa, A, ra), (ob, B, rb), (o1, A, ra)):
            wip, wix, wv = want(m, rows)
            assert np.array_equal(got[0], wip)
            assert np.array_equal(got[1][: wip[-1]], wix) and np.array_equal(got[2][: wip[-1]], wv)
    # capacity check
    assert h.scnb_host_gather_rows2(P(A[0]), P(A[1]), P(A[2]), P(ra), 64, P(oa[0]), P(oa[1]), P(oa[2]), 3,
                                    P(B[0]), P(B[1]), P(B[2]), P(rb), 48, P(ob[0]), P(ob[1]), P(ob[2]), 48 * 40, 2) != 0
    assert b"capacity" in h.scnb_last_error()


def test_checkpoint_file_loads_into_the_reference_model_when_available(tmp_path):
    """The weights file the trainers write (`00_best_model_weights.pkl`, scnym/trainer.py:404-417) is a plain state_dict in the
    reference's own layout: the UNMODIFIED reference `CellTypeCLF` loads it with strict=True and its eval-mode forward equals
    the oracle's on those weights (the gene-major storage of the first layer never shows in the file)."""
    from oracle.ref_shim import reference_available, load_reference
    if not reference_available():
        pytest.skip("reference tree not present (GPU box)")
    ref = load_reference(modules=("model",))
    from oracle import scnym_oracle as O
    from scnym_b200.model import CellTypeCLF
    from scnym_b200.trainer import checkpoint_state
    for nl in (1, 2, 3):
        torch.manual_seed(11 + nl)
        ours = CellTypeCLF(n_genes=90, n_cell_types=4, n_hidden=24, n_hidden_init=40, n_layers=nl)
        with torch.no_grad():
            for m in ours.modules():
                if isinstance(m, torch.nn.BatchNorm1d):
                    m.running_mean.normal_(0.0, 0.3)
                    m.running_var.uniform_(0.5, 1.5)
        path = str(tmp_path / f"00_best_model_weights_{nl}.pkl")
        torch.save(checkpoint_state(ours), path)
        sd = torch.load(path)
        assert all(v.is_contiguous() for v in sd.values())
        theirs = ref.model.CellTypeCLF(n_genes=90, n_cell_types=4, n_hidden=24, n_hidden_init=40, n_layers=nl)
        theirs.load_state_dict(sd, strict=True)
        theirs.eval()
        X = torch.rand(17, 90) * 3.0
        with torch.no_grad():
            want = theirs(X)
        got = O.OracleModel(sd, n_layers=nl).forward(X, train=False)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
