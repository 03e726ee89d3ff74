"""Augmentation schemes other than log1p_drop (scnym/dataprep.py:730-754: log1p_mask, log1p_poisson, log1p_poisson_drop,
count_poisson).  Golden vectors come from the UNMODIFIED reference's own AUGMENTATION_SCHEMES objects with every random
draw recorded (tests/golden/make_golden_augment.py).
  CPU: the oracle restatement == golden.
  GPU: the device kernel (`scnb_csr_gather_rows_aug` through `dataprep.AUGMENTATION_SCHEMES`) with the recorded draws
       injected == golden (1e-6: elementwise fp32), on dense and on CSR input; the in-kernel samplers (Philox) have the
       reference's distributions (exact masked-gene count per row, batch-level apply probability, Poisson moments and
       probability mass); the MixMatch engine step with each scheme == oracle."""
import numpy as np
import pytest
import torch

from oracle import scnym_oracle as O
from tests import golden_util as GU

SCHEMES = ("log1p_mask", "log1p_poisson", "log1p_poisson_drop", "count_poisson")


def _golden():
    return GU.load("augment_schemes")


def _inputs(d, name):
    t = lambda k: torch.from_numpy(d[k]) if k in d.files else None
    x = t("counts") if name == "count_poisson" else t("x")
    return x, t(name + "/keep"), t(name + "/pois"), t(name + "/row_u"), t(name + "/out")


@pytest.mark.parametrize("name", SCHEMES)
def test_oracle_augment_matches_reference(name):
    d = _golden()
    x, keep, pois, _, want = _inputs(d, name)
    got = O.augment_rows(x, name, keep=keep, pois=pois)
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert GU.rel_err(torch.nan_to_num(got), torch.nan_to_num(want)) <= 1e-6
    if name == "log1p_mask":
        assert np.all(d["log1p_mask/n_masked"] == int(np.floor(int(d["meta_G"]) * 0.1)))
        noop = O.augment_rows(x, name, keep=torch.ones_like(x))
        assert GU.rel_err(noop, d["log1p_mask_noop/out"]) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCHEMES)
@pytest.mark.parametrize("kind", ["dense", "csr"])
def test_device_augment_matches_reference_golden(name, kind):
    from scnym_b200.dataprep import AUGMENTATION_SCHEMES, CSRBatch, SparseAugment
    d = _golden()
    x, keep, pois, row_u, want = _inputs(d, name)
    G = x.shape[1]
    aug = AUGMENTATION_SCHEMES[name]
    if kind == "dense":
        if keep is not None:
            SparseAugment.KEEP_INJECT.append(keep)
        if pois is not None:
            SparseAugment.POIS_INJECT.append(pois)
        if row_u is not None:
            SparseAugment.ROWU_INJECT.append(row_u)
        got = aug({"input": x.cuda(), "output": torch.zeros(x.shape[0])})["input"].cpu()
    else:
        nz = x != 0
        ip = torch.zeros(x.shape[0] + 1, dtype=torch.int32)
        ip[1:] = torch.cumsum(nz.sum(1), 0)
        rows, cols = torch.nonzero(nz, as_tuple=True)
        batch = CSRBatch(ip.cuda(), cols.to(torch.int32).cuda(), x[nz].cuda(), x.shape[0], G)
        if keep is not None:
            SparseAugment.KEEP_INJECT.append(keep[nz])
        if pois is not None:
            SparseAugment.POIS_INJECT.append(pois[nz])
        if row_u is not None:
            SparseAugment.ROWU_INJECT.append(row_u)
        out = aug({"input": batch, "output": torch.zeros(x.shape[0])})["input"]
        got = out.to_dense().cpu()
    assert not SparseAugment.KEEP_INJECT and not SparseAugment.POIS_INJECT and not SparseAugment.ROWU_INJECT
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert GU.rel_err(torch.nan_to_num(got), torch.nan_to_num(want)) <= 2e-6, name


@pytest.mark.gpu
def test_gene_masking_sampler_masks_exactly_floor_pG_genes_per_row_half_of_the_time():
    """GeneMasking (dataprep.py:405-465): one batch-level decision with probability p_apply, then exactly floor(G * p) genes
    per row, drawn without replacement from all G genes (dense rows store every gene: the count is exact)."""
    from scnym_b200.dataprep import SparseAugment
    G, B = 1000, 64
    x = (torch.rand(B, G) * 5 + 0.5).cuda()
    aug = SparseAugment("log1p_mask", p_drop=0.1, p_apply=0.5)
    applied, firsts = 0, []
    for _ in range(60):
        y = aug({"input": x.clone(), "output": torch.zeros(B)})["input"]
        zeros = (y == 0).sum(1).cpu().numpy()
        assert np.all(zeros == zeros[0])
        assert zeros[0] in (0, int(np.floor(G * 0.1))), zeros[0]
        applied += int(zeros[0] > 0)
        if zeros[0] > 0:
            firsts.append((y == 0).float().mean(0).cpu())
    assert 15 <= applied <= 45, applied            # Binomial(60, 0.5): +-4 sigma
    freq = torch.stack(firsts).mean(0)              # every gene is masked about 10 % of the time
    assert abs(float(freq.mean()) - 0.1) < 1e-3 and float(freq.max()) < 0.2
    aug_never = SparseAugment("log1p_mask", p_drop=0.1, p_apply=0.0)
    y = aug_never({"input": x.clone(), "output": torch.zeros(B)})["input"]
    assert int((y == 0).sum()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("lam", [0.3, 3.0, 9.5, 10.5, 60.0, 2500.0, 3.0e5])
def test_poisson_sampler_moments_and_mass(lam):
    """In-kernel Poisson draws (inversion below 10, PTRS above): mean and variance equal the rate, the probability mass of
    small rates matches the exact pmf."""
    from scnym_b200.dataprep import SparseAugment
    B, G = 256, 2048
    n = B * G
    x = torch.full((B, G), float(lam)).cuda()
    y = SparseAugment("count_poisson")({"input": x, "output": torch.zeros(B)})["input"].double().cpu()
    assert float(y.min()) >= 0 and bool(torch.all(y == torch.round(y)))
    se_mean = np.sqrt(lam / n)
    assert abs(float(y.mean()) - lam) < 5 * se_mean + 1e-7 * lam, (float(y.mean()), lam)
    var = float(y.var())
    se_var = lam * np.sqrt(2.0 / n + 1.0 / (lam * n))
    assert abs(var - lam) < 6 * se_var + 1e-6 * lam, (var, lam)
    if lam < 20:
        from scipy import stats
        ks = np.arange(0, int(lam + 8 * np.sqrt(lam) + 8))
        emp = np.array([float((y == k).double().mean()) for k in ks])
        pmf = stats.poisson.pmf(ks, lam)
        assert np.max(np.abs(emp - pmf)) < 5 * np.sqrt(0.25 / n) + 1e-4
    # zero rates stay zero (PoissonSample's rate-0 guard, dataprep.py:326-335)
    z = SparseAugment("count_poisson")({"input": torch.zeros(4, 64).cuda(), "output": torch.zeros(4)})["input"]
    assert int((z != 0).sum()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("aug", ["log1p_mask", "log1p_poisson", "log1p_poisson_drop"])
def test_engine_step_with_other_augmentation_schemes_matches_oracle(aug):
    from tests import cuda_harness as CH
    d = CH.oracle_case(G=1500, L=64, U=64, C=6, H0=128, H=64, n_layers=2, opt_name="adadelta", lr=1.0, seed=61, augment=aug)
    eng = CH.build_engine(d)
    assert eng.aug_code != 1
    CH.inject_step(eng, d, 0)
    eng.step()
    CH.compare_step0(eng, d)
    # production randomness (Philox): the step runs and stays finite
    eng2 = CH.build_engine(d)
    eng2.cfg.use_cuda_graph = True
    for _ in range(3):
        eng2.load_batch(d["step0/lrow"], d["step0/urow"], d["step0/perm_keys"], d["step0/gamma_raw"])
        eng2.step()
    torch.cuda.synchronize()
    assert np.isfinite(eng2.losses()["loss"])
