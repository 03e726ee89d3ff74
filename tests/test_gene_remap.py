"""Gene re-indexing in front of predict (reference scnym/utils.py:155-233).

CPU: the oracle restatement and the host forms against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden_genes.py).  GPU: the device kernels (scnb_csr_remap_count / scnb_csr_remap_fill through
`dataprep.remap_columns`) against the same golden vectors and, on seeded inputs, against the oracle.  Bar: bit-exact
(index and byte work: the values are moved, never recomputed)."""
import os

import numpy as np
import pytest
import torch
from scipy import sparse

from oracle import scnym_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ("subset", "superset", "disjoint")


def _golden(case):
    d = np.load(os.path.join(HERE, "golden", "gene_remap.npz"))
    g = {k[len(case) + 1:]: d[k] for k in d.files if k.startswith(case + "_")}
    X = sparse.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=(len(g["indptr"]) - 1, len(g["sample_genes"])))
    return X, g["model_genes"], g["sample_genes"], g["N_from_csr"], g["N_from_dense"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case):
    X, mg, sg, n_csr, n_dense = _golden(case)
    assert np.array_equal(n_csr, n_dense)
    assert np.array_equal(O.build_classification_matrix(X, mg, sg, gene_batch_size=7).astype(np.float32), n_csr)
    assert np.array_equal(O.build_classification_matrix(X.toarray(), mg, sg).astype(np.float32), n_csr)


@pytest.mark.parametrize("case", CASES)
def test_host_forms_match_reference_golden(case):
    from scnym_b200.utils import build_classification_matrix
    X, mg, sg, want, _ = _golden(case)
    N = build_classification_matrix(X, mg, sg)
    assert sparse.isspmatrix_csr(N) and N.has_sorted_indices
    assert np.array_equal(N.toarray().astype(np.float32), want)
    Nd = build_classification_matrix(X.toarray(), mg, sg)
    assert isinstance(Nd, np.ndarray) and np.array_equal(Nd.astype(np.float32), want)


def test_gene_column_map_semantics():
    from scnym_b200.utils import gene_column_map
    model = np.array(["a", "b", "c", "d"])
    sample = np.array(["d", "x", "a", "d", "b"])
    cm = gene_column_map(model, sample)
    assert cm.dtype == np.int32
    # "d" occurs twice in the sample: the reference assigns both columns, the last one wins
    assert cm.tolist() == [-1, -1, 0, 3, 1]
    X = np.arange(15, dtype=np.float64).reshape(3, 5)
    want = O.build_classification_matrix(X, model, sample)
    got = np.zeros((3, 4))
    got[:, cm[cm >= 0]] = X[:, cm >= 0]
    assert np.array_equal(got, want)
    # a model gene listed twice is ambiguous in the reference (`.item()` on two hits raises ValueError) -- when the sample has it
    with pytest.raises(ValueError):
        gene_column_map(np.array(["a", "b", "a"]), np.array(["a", "q"]))
    with pytest.raises(ValueError):
        O.build_classification_matrix(X[:, :2], np.array(["a", "b", "a"]), np.array(["a", "q"]))
    assert gene_column_map(np.array(["a", "b", "a"]), np.array(["b", "q"])).tolist() == [1, -1]


# ------------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------------
def _device_dense(ds):
    ip = ds.indptr.cpu().numpy()
    n = len(ip) - 1
    return sparse.csr_matrix((ds.data.cpu().numpy()[:ip[-1]], ds.indices.cpu().numpy()[:ip[-1]], ip), shape=(n, ds.n_cols))


def _check_sorted_unique(M):
    for r in range(M.shape[0]):
        c = M.indices[M.indptr[r]:M.indptr[r + 1]]
        assert np.all(np.diff(c) > 0), f"row {r} not strictly increasing"


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_device_remap_matches_reference_golden(case):
    from scnym_b200.dataprep import DeviceCSR
    from scnym_b200.utils import build_classification_matrix
    X, mg, sg, want, _ = _golden(case)
    N = build_classification_matrix(DeviceCSR.from_scipy(X), mg, sg)
    assert isinstance(N, DeviceCSR) and N.shape == want.shape
    M = _device_dense(N)
    _check_sorted_unique(M)
    assert np.array_equal(M.toarray(), want)


def _random_case(n_rows, n_src, n_dst, density, frac_kept, seed, empty_rows=True):
    rng = np.random.default_rng(seed)
    X = sparse.random(n_rows, n_src, density=density, format="csr", dtype=np.float32, random_state=seed)
    X.data = (rng.random(X.nnz, dtype=np.float32) * 13 + 0.01).astype(np.float32)
    X.sort_indices()
    if empty_rows and n_rows > 4:
        X = sparse.csr_matrix(sparse.vstack([X[:2], sparse.csr_matrix((1, n_src), dtype=np.float32), X[2:]]))
    n_keep = min(int(frac_kept * n_src), n_dst)
    col_map = np.full(n_src, -1, dtype=np.int32)
    col_map[rng.choice(n_src, size=n_keep, replace=False)] = rng.choice(n_dst, size=n_keep, replace=False).astype(np.int32)
    return X, col_map


def _host_remap(X, col_map, n_dst):
    keep = col_map[X.indices] >= 0
    rows = np.repeat(np.arange(X.shape[0]), np.diff(X.indptr))
    N = sparse.csr_matrix((X.data[keep], (rows[keep], col_map[X.indices[keep]])), shape=(X.shape[0], n_dst))
    N.sort_indices()
    return N


@pytest.mark.gpu
@pytest.mark.parametrize("n_rows,n_src,n_dst,density,frac", [
    (1, 40, 40, 0.5, 1.0),            # one row, pure permutation
    (63, 300, 200, 0.1, 0.5),         # fewer rows than one counting CTA
    (64 * 3 + 5, 2000, 2500, 0.05, 0.8),
    (5000, 500, 333, 0.02, 0.6),      # many short rows; several counting CTAs
    (70000, 64, 50, 0.05, 0.7),       # > 1024 counting CTAs: the scan of their totals takes two rounds
    (40, 20000, 20000, 0.3, 0.9),     # long rows (6000 stored entries), config-2 gene count
    (17, 3000, 450000, 0.2, 1.0),     # 450 k destination genes: bitmap + prefix above the 48 KB default
])
def test_device_remap_matches_host_remap(n_rows, n_src, n_dst, density, frac):
    from scnym_b200.dataprep import DeviceCSR, remap_columns
    X, col_map = _random_case(n_rows, n_src, n_dst, density, frac, seed=n_rows + n_src)
    want = _host_remap(X, col_map, n_dst)
    ds = DeviceCSR.from_scipy(X)
    N = remap_columns(ds, torch.as_tensor(col_map).cuda(), n_dst)
    ip = N.indptr.cpu().numpy()
    assert np.array_equal(ip, want.indptr.astype(np.int64))
    assert np.array_equal(N.indices.cpu().numpy()[:ip[-1]], want.indices)
    assert np.array_equal(N.data.cpu().numpy()[:ip[-1]], want.data)          # values are moved bit for bit
    # the oracle (dense restatement of the reference) on the small cases
    if n_rows * n_dst <= 2_000_000:
        sg = np.array([f"g{i}" for i in range(n_src)])
        mg = np.array([f"absent{j}" for j in range(n_dst)], dtype=object)
        mg[col_map[col_map >= 0]] = sg[col_map >= 0]
        assert np.array_equal(O.build_classification_matrix(X, mg.astype(str), sg).astype(np.float32), want.toarray())


@pytest.mark.gpu
def test_device_remap_row_range_and_caller_buffers():
    """A row range of a larger matrix into caller-provided buffers (the form the predict loop uses: no read-back)."""
    from scnym_b200 import _lib
    from scnym_b200.dataprep import DeviceCSR, remap_columns
    X, col_map = _random_case(1000, 700, 600, 0.05, 0.75, seed=5)
    ds = DeviceCSR.from_scipy(X)
    cm = torch.as_tensor(col_map).cuda()
    row0, n = 333, 500
    cap = int(X.indptr[row0 + n] - X.indptr[row0])
    out = (torch.empty(n + 1, dtype=torch.int64, device="cuda"), torch.full((cap,), -7, dtype=torch.int32, device="cuda"),
           torch.zeros(cap, dtype=torch.float32, device="cuda"),
           torch.empty(_lib.lib().scnb_csr_remap_workspace_len(n), dtype=torch.int64, device="cuda"))
    N = remap_columns(ds, cm, 600, row0=row0, n_rows=n, out=out)
    want = _host_remap(X[row0:row0 + n], col_map, 600)
    ip = N.indptr.cpu().numpy()
    assert np.array_equal(ip, want.indptr.astype(np.int64))
    assert np.array_equal(N.indices.cpu().numpy()[:ip[-1]], want.indices)
    assert np.array_equal(N.data.cpu().numpy()[:ip[-1]], want.data)
    assert (out[1][ip[-1]:] == -7).all()      # nothing written past the surviving entries
    with pytest.raises(TypeError):
        remap_columns(ds, cm.to(torch.int64), 600)


@pytest.mark.gpu
def test_predict_with_col_map_equals_predict_on_reindexed_matrix():
    """`Predicter.predict_device(X, col_map=...)` (chunk-wise device re-indexing inside the predict loop) against predict on
    the matrix re-indexed beforehand by the host form, across several chunks.  The re-indexed rows are bit-identical (tests
    above); the first layer adds its gene-axis partial products with atomics, so logits agree to rounding (1e-5), the
    predicted classes exactly."""
    import tempfile
    from scnym_b200.model import CellTypeCLF
    from scnym_b200.predict import EvalRunner, Predicter
    from scnym_b200.utils import build_classification_matrix, gene_column_map
    torch.manual_seed(3)
    n_model, n_sample, C = 192, 260, 5
    sg = np.array([f"g{i}" for i in range(n_sample)])
    rng = np.random.default_rng(3)
    mg = np.array([f"only_model{j}" for j in range(n_model)], dtype=object)
    common = rng.choice(n_sample, size=150, replace=False)
    mg[rng.choice(n_model, size=150, replace=False)] = sg[common]
    mg = mg.astype(str)
    X, _ = _random_case(900, n_sample, n_model, 0.1, 0.5, seed=9)
    m = CellTypeCLF(n_genes=n_model, n_cell_types=C, n_hidden=64, n_hidden_init=64)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "w.pkl")
        torch.save(m.state_dict(), path)
        P = Predicter(path, n_genes=n_model, n_cell_types=C, n_hidden=64, n_hidden_init=64)
    from scnym_b200.dataprep import DeviceCSR
    cm = gene_column_map(mg, sg)
    Xm = build_classification_matrix(X, mg, sg)
    runner = EvalRunner(P.models, chunk_rows=256)     # 4 chunks, the last one ragged
    a = runner.run(DeviceCSR.from_scipy(X), True, True, True, col_map=torch.as_tensor(cm).cuda())
    b = runner.run(DeviceCSR.from_scipy(Xm), True, True, True)
    def same(u, v):
        assert torch.equal(u[0], v[0])                                        # argmax exact
        for x, y in zip(u[1:], v[1:]):
            assert float((x - y).abs().max()) <= 1e-5 * float(y.abs().max())

    same(a, b)
    c = P.predict_device(X, want_prob=True, want_score=True, want_embed=True, col_map=cm)     # the public entry, one chunk
    d = P.predict_device(Xm, want_prob=True, want_score=True, want_embed=True)
    same(c, d)
    same(a, c)
    with pytest.raises(ValueError):
        P.predict_device(X)            # sample gene count without a column map
