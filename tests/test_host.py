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
