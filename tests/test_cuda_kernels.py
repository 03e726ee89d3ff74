"""GPU unit tests of individual C-ABI entry points against fp64 torch-CPU restatements."""
import numpy as np
import pytest
import torch

from tests import golden_util as GU

pytestmark = pytest.mark.gpu


def _st():
    return torch.cuda.current_stream().cuda_stream


_keep = []


def D(t):
    """Move to the GPU and keep the tensor alive until the test ends (raw pointers are borrowed)."""
    t = t.cuda()
    _keep.append(t)
    return t


@pytest.fixture(autouse=True)
def _release_device_tensors():
    yield
    torch.cuda.synchronize()
    _keep.clear()


def _chk(name, got, want, tol=1e-5):
    e = GU.rel_err(got, want)
    assert e <= tol, f"{name}: rel err {e:.3e} > {tol}"


@pytest.mark.parametrize("tA", [0, 1])
@pytest.mark.parametrize("tB", [0, 1])
@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (130, 70, 50), (33, 5, 17), (1, 256, 256), (512, 16, 256), (7, 3, 1)])
def test_sgemm(tA, tB, M, N, K):
    from scnym_b200._lib import call, ptr
    torch.manual_seed(M * 7 + N * 3 + K + tA * 2 + tB)
    A = torch.randn((K, M) if tA else (M, K))
    B = torch.randn((N, K) if tB else (K, N))
    C0 = torch.randn(M, N)
    bias = torch.randn(N)
    gate = torch.randn(M, N)
    opA = A.t() if tA else A
    opB = B.t() if tB else B
    for (alpha, beta, use_bias, relu, use_gate) in [(1.0, 0.0, False, 0, False), (-0.3, 1.0, True, 0, False),
                                                     (1.0, 0.0, True, 1, False), (1.0, 0.0, False, 0, True)]:
        want = alpha * (opA.double() @ opB.double()) + beta * C0.double()
        if use_bias:
            want = want + bias.double()
        if relu:
            want = want.clamp_min(0)
        if use_gate:
            want = want * (gate > 0)
        Cd = C0.cuda().clone()
        Ad, Bd = A.cuda(), B.cuda()
        bd, gd = bias.cuda(), gate.cuda()
        call("scnb_sgemm", tA, tB, M, N, K, ptr(Ad), Ad.shape[1], ptr(Bd), Bd.shape[1], ptr(Cd), N, alpha, beta,
             ptr(bd) if use_bias else None, relu, ptr(gd) if use_gate else None, _st())
        _chk(f"sgemm {tA}{tB} {M}x{N}x{K}", Cd.cpu(), want.float(), tol=2e-5)


def test_colsum_and_relu_bwd():
    from scnym_b200._lib import call, ptr
    A = torch.randn(300, 70)
    out = torch.ones(70).cuda()
    call("scnb_colsum", ptr(D(A)), 300, 70, 70, ptr(out), 1, _st())
    _chk("colsum", out.cpu(), 1 + A.double().sum(0), tol=1e-5)
    g, y = torch.randn(1000), torch.randn(1000)
    o = torch.empty(1000).cuda()
    call("scnb_relu_bwd", ptr(D(g)), ptr(D(y)), ptr(o), 1000, _st())
    assert torch.equal(o.cpu(), g * (y > 0))


@pytest.mark.parametrize("H0", [128, 256, 512, 96, 33])
def test_csr_linear_fwd_bwd(H0):
    from oracle import scnym_oracle as O
    from scnym_b200._lib import call, ptr
    from scnym_b200.dataprep import DeviceCSR, gather_rows
    G, n = 500, 37
    ip, ii, dd, _ = O.synth_csr(n + 10, G, 0.08, 3, seed=H0)
    rows = torch.randperm(n + 10)[:n]
    X = O.csr_to_dense(ip, ii, dd, G, rows.numpy())
    ds = DeviceCSR.from_arrays(ip, ii, dd, G)
    b = gather_rows(ds, rows.cuda())
    torch.manual_seed(H0)
    WT, b1 = torch.randn(G, H0) * 0.1, torch.randn(H0)
    Z = torch.empty(n, H0).cuda()
    call("scnb_csr_linear_fwd", ptr(b.indptr), ptr(b.indices), ptr(b.data), n, ptr(D(WT)), ptr(D(b1)), H0, ptr(Z), _st())
    _chk("csr fwd", Z.cpu(), X.double() @ WT.double() + b1.double())
    _chk("to_dense", b.to_dense().cpu(), X, tol=0)
    dZ = torch.randn(n, H0)
    dW = torch.zeros(G, H0).cuda()
    call("scnb_csr_linear_bwd_dw", ptr(b.indptr), ptr(b.indices), ptr(b.data), n, ptr(D(dZ)), H0, ptr(dW), _st())
    _chk("csr bwd", dW.cpu(), X.double().t() @ dZ.double())
    # dataset-CSR (int64 indptr) variant on a contiguous row range
    Z2 = torch.empty(9, H0).cuda()
    call("scnb_csr_linear_fwd_i64", ptr(ds.indptr[5:]), ptr(ds.indices), ptr(ds.data), 9, ptr(D(WT)), None, H0, ptr(Z2), _st())
    _chk("csr fwd i64", Z2.cpu(), O.csr_to_dense(ip, ii, dd, G, np.arange(5, 14)).double() @ WT.double())


def test_gather_rows_log1p_drop_matches_oracle():
    from oracle import scnym_oracle as O
    from scnym_b200.dataprep import DeviceCSR, gather_rows
    G, n = 400, 25
    ip, ii, dd, _ = O.synth_csr(n, G, 0.1, 3, seed=4)
    rows = np.random.default_rng(0).permutation(n)
    X = O.csr_to_dense(ip, ii, dd, G, rows)
    keep = (torch.rand(n, G) < 0.9)
    want = O.log1p_drop(X, keep)
    ds = DeviceCSR.from_arrays(ip, ii, dd, G)
    km = np.concatenate([keep[i, ii[ip[r]:ip[r + 1]]].numpy() for i, r in enumerate(rows)]).astype(np.uint8)
    buf = np.ones(n * ds.max_row_nnz, dtype=np.uint8)
    buf[:km.size] = km
    b = gather_rows(ds, torch.from_numpy(rows).cuda(), augment=True, keep_mask=torch.from_numpy(buf).cuda())
    _chk("log1p_drop", b.to_dense().cpu(), want, tol=2e-6)
    # Philox path: ~10% dropped, rows renormalised to 1e6
    rng = torch.tensor([123, 7], dtype=torch.int64).cuda()
    b2 = gather_rows(ds, torch.from_numpy(rows).cuda(), augment=True, rng=rng, stream_id=3)
    D2 = b2.to_dense().cpu()
    frac = float(((D2 == 0) & (X > 0)).sum()) / float((X > 0).sum())
    assert 0.05 < frac < 0.15
    _chk("row sums", torch.expm1(D2.double()).sum(1), torch.full((n,), 1e6, dtype=torch.float64), tol=1e-4)
    b3 = gather_rows(ds, torch.from_numpy(rows).cuda(), augment=True, rng=rng, stream_id=3)
    assert torch.equal(b3.to_dense(), b2.to_dense())  # counter-based: same (seed, counter, stream) -> same mask


def _bn_ref(Z, segs, g, b, keep, p):
    """fp64 autograd restatement of segmented BN -> dropout -> relu."""
    outs = []
    for (r0, r1) in segs:
        z = Z[r0:r1]
        mu = z.mean(0)
        var = ((z - mu) ** 2).mean(0)
        y = (z - mu) / torch.sqrt(var + 1e-5) * g + b
        y = y * keep[r0:r1].double() / (1 - p)
        outs.append(torch.relu(y))
    return torch.cat(outs, 0)


@pytest.mark.parametrize("H,segs,maxr", [(64, [(0, 40), (40, 70), (70, 130)], 150), (37, [(0, 5), (5, 300)], 300),
                                         (256, [(0, 256), (256, 512), (512, 1024)], 1024),
                                         # large batches: the rows of a segment are split over a thread-block cluster
                                         (256, [(0, 3000), (3000, 9001), (9001, 12000)], 12400),
                                         (64, [(0, 4100), (4100, 5000)], 5003)])
def test_bn_act_fwd_bwd(H, segs, maxr):
    from scnym_b200._lib import call, ptr
    torch.manual_seed(H)
    Z = (torch.randn(maxr, H) * 2 + 0.5)
    g, b = torch.rand(H) + 0.5, torch.randn(H) * 0.1
    keep = (torch.rand(maxr, H) < 0.5).to(torch.uint8)
    dA = torch.randn(maxr, H)
    Zd = Z.double().requires_grad_(True)
    gd, bd = g.double().requires_grad_(True), b.double().requires_grad_(True)
    A_ref = _bn_ref(Zd, segs, gd, bd, keep, 0.5)
    n_act = segs[-1][1]
    (A_ref * dA[:n_act].double()).sum().backward()
    seg = torch.tensor([s[0] for s in segs] + [n_act], dtype=torch.int32).cuda()
    A = torch.full((maxr, H), 7.0).cuda()
    stats = torch.zeros(len(segs), 3, H).cuda()
    Zc, gc, bc, kc = Z.cuda(), g.cuda(), b.cuda(), keep.cuda()
    call("scnb_bn_act_fwd", ptr(Zc), ptr(A), None, H, len(segs), ptr(seg), maxr, 1, ptr(gc), ptr(bc), None, None, ptr(stats),
         ptr(kc), None, 0, 0.5, 1, None, None, _st())
    _chk("bn fwd", A[:n_act].cpu(), A_ref.detach())
    assert float(A[n_act:].abs().sum()) == 0.0
    dZ = torch.full((maxr, H), 7.0).cuda()
    dg, db = torch.zeros(H).cuda(), torch.zeros(H).cuda()
    call("scnb_bn_act_bwd", ptr(D(dA)), ptr(Zc), ptr(A), H, len(segs), ptr(seg), maxr, ptr(gc), ptr(stats), ptr(kc), None, 0,
         0.5, 1, ptr(dZ), ptr(dg), ptr(db), None, None, _st())
    _chk("bn dZ", dZ[:n_act].cpu(), Zd.grad[:n_act], tol=2e-5)
    _chk("bn dgamma", dg.cpu(), gd.grad, tol=2e-5)
    _chk("bn dbeta", db.cpu(), bd.grad, tol=2e-5)
    assert float(dZ[n_act:].abs().sum()) == 0.0
    # running statistics, segment order, unbiased variance
    rm, rv = torch.zeros(H), torch.ones(H)
    for (r0, r1) in segs:
        z = Z[r0:r1].double()
        rm = 0.9 * rm + 0.1 * z.mean(0)
        rv = 0.9 * rv + 0.1 * z.var(0, unbiased=True)
    rmc, rvc = torch.zeros(H).cuda(), torch.ones(H).cuda()
    nbt = torch.zeros((), dtype=torch.int64).cuda()
    call("scnb_bn_running_update", ptr(rmc), ptr(rvc), ptr(nbt), ptr(stats), None, None, None, ptr(seg), len(segs), H, 0.1, None, _st())
    _chk("running_mean", rmc.cpu(), rm)
    _chk("running_var", rvc.cpu(), rv)
    assert int(nbt) == len(segs)
    # eval mode
    Ae = torch.empty(maxr, H).cuda()
    pre = torch.empty(maxr, H).cuda()
    call("scnb_bn_act_fwd", ptr(Zc), ptr(Ae), ptr(pre), H, 1, ptr(D(torch.tensor([0, maxr], dtype=torch.int32))), maxr, 0,
         ptr(gc), ptr(bc), ptr(rmc), ptr(rvc), None, None, None, 0, 0.0, 1, None, None, _st())
    want = (Z.double() - rmc.cpu().double()) / torch.sqrt(rvc.cpu().double() + 1e-5) * g.double() + b.double()
    _chk("bn eval pre", pre.cpu(), want)
    _chk("bn eval", Ae.cpu(), want.clamp_min(0))


def test_bn_act_philox_dropout_consistent_between_fwd_and_bwd():
    from scnym_b200._lib import call, ptr
    H, n = 128, 512
    Z = torch.randn(n, H).cuda()
    g, b = torch.ones(H).cuda(), torch.full((H,), 3.0).cuda()   # beta=3 -> almost all pre-dropout values positive
    seg = torch.tensor([0, n], dtype=torch.int32).cuda()
    rng = torch.tensor([42, 5], dtype=torch.int64).cuda()
    A, stats = torch.empty(n, H).cuda(), torch.zeros(1, 3, H).cuda()
    call("scnb_bn_act_fwd", ptr(Z), ptr(A), None, H, 1, ptr(seg), n, 1, ptr(g), ptr(b), None, None, ptr(stats), None, ptr(rng), 17,
         0.5, 1, None, None, _st())
    frac = float((A == 0).float().mean())
    assert 0.47 < frac < 0.53
    dZ, dg, db = torch.empty(n, H).cuda(), torch.zeros(H).cuda(), torch.zeros(H).cuda()
    ones = torch.ones(n, H).cuda()
    call("scnb_bn_act_bwd", ptr(ones), ptr(Z), ptr(A), H, 1, ptr(seg), n, ptr(g), ptr(stats), None, ptr(rng), 17, 0.5, 1, ptr(dZ),
         ptr(dg), ptr(db), None, None, _st())
    # dbeta = sum of dy = 2 * (#kept and positive) per column
    _chk("dbeta", db.cpu(), 2.0 * (A > 0).float().sum(0).cpu(), tol=1e-6)
    A2 = torch.empty(n, H).cuda()
    rng2 = torch.tensor([42, 6], dtype=torch.int64).cuda()
    call("scnb_bn_act_fwd", ptr(Z), ptr(A2), None, H, 1, ptr(seg), n, 1, ptr(g), ptr(b), None, None, ptr(stats), None, ptr(rng2),
         17, 0.5, 1, None, None, _st())
    assert not torch.equal(A2 == 0, A == 0)  # next step counter -> fresh mask


@pytest.mark.parametrize("K,T,C", [(1, 0.5, 7), (3, 0.5, 40), (2, 0.0, 5), (1, 1.0, 5), (2, 2.0, 33)])
def test_pseudolabel(K, T, C):
    from oracle import scnym_oracle as O
    from scnym_b200._lib import call, ptr
    torch.manual_seed(K * 10 + C)
    U = 50
    logits = torch.randn(K, U, C) * 2
    qbar = torch.stack([O.softmax(logits[k]) for k in range(K)]).mean(0)
    conf = qbar.max(1).values
    tau = float(conf.median())
    want_q = O.sharpen_labels(qbar.clone(), T).float()
    q, cf = torch.empty(U, C).cuda(), torch.empty(U).cuda()
    flag = torch.empty(U, dtype=torch.uint8).cuda()
    call("scnb_pseudolabel", ptr(D(logits)), K, U, C, T, 1, tau, ptr(q), ptr(cf), ptr(flag), ptr(D(torch.empty(U, C))), None, 0, _st())
    _chk("q", q.cpu(), want_q)
    _chk("conf", cf.cpu(), conf)
    exact = (conf - tau).abs() > 1e-6
    assert torch.equal(flag.cpu().bool()[exact], (conf >= tau)[exact])


def _plan(confident, keys, graw, U, L, use_dan, dan_conf, mixup_on=1):
    from scnym_b200._lib import call, ptr
    nb = U + L
    R = nb + (nb if use_dan else 0)
    i32 = lambda *s: torch.full(s, -7, dtype=torch.int32).cuda()
    out = dict(counts=i32(8), seg=i32(4), work=i32(U + 3 * nb), srcA=i32(R), srcB=i32(R), gam=torch.zeros(R).cuda(),
               inv3=i32(3, nb), coef3=torch.zeros(3, nb).cuda(), pconf=torch.zeros(1).cuda())
    call("scnb_mixmatch_plan", ptr(D(confident)), ptr(D(keys)), ptr(D(graw)), U, L, use_dan, dan_conf, mixup_on,
         ptr(out["counts"]), ptr(out["seg"]), ptr(out["work"]), ptr(out["srcA"]), ptr(out["srcB"]), ptr(out["gam"]), R,
         ptr(out["inv3"]), ptr(out["coef3"]), ptr(out["pconf"]), _st())
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in out.items()}, R


@pytest.mark.parametrize("U,L,frac,use_dan,dan_conf", [(40, 24, 0.5, 1, 1), (40, 24, 0.5, 1, 0), (33, 17, 0.0, 1, 1),
                                                       (16, 16, 1.0, 0, 0), (300, 260, 0.3, 1, 1), (1500, 1100, 0.7, 1, 0)])
def test_mixmatch_plan_and_mix_adjoint(U, L, frac, use_dan, dan_conf):
    from oracle import scnym_oracle as O
    from scnym_b200._lib import call, ptr
    torch.manual_seed(U + L)
    confident = (torch.rand(U) < frac).to(torch.uint8)
    keys, graw = torch.rand(U + L), torch.distributions.Beta(0.3, 0.3).sample((U + L,))
    out, R = _plan(confident, keys, graw, U, L, use_dan, dan_conf)
    nb = U + L
    cidx = torch.where(confident.bool())[0]
    uidx = torch.where(~confident.bool())[0]
    n_conf = cidx.numel()
    ridx, gam = O.mixup_indices(keys, graw, n_conf + L)
    cat2base = torch.cat([cidx, U + torch.arange(L)])
    n_dan_unl = (n_conf if dan_conf else U) if use_dan else 0
    n_dan = L + n_dan_unl if use_dan else 0
    assert out["counts"][:4].tolist() == [n_conf, n_conf + L, n_dan, nb + n_dan]
    assert out["seg"].tolist() == [0, U, nb, nb + n_dan]
    assert abs(float(out["pconf"]) - n_dan_unl / max(U, 1)) < 1e-7
    # expected forward maps
    eA = torch.cat([cat2base[:n_conf], uidx, U + torch.arange(L)])
    eB = torch.cat([cat2base[ridx[:n_conf]], uidx, cat2base[ridx[n_conf:]]])
    eG = torch.cat([gam[:n_conf, 0], torch.ones(uidx.numel()), gam[n_conf:, 0]])
    if use_dan:
        dan_rows = torch.cat([U + torch.arange(L), cidx if dan_conf else torch.arange(U)])
        eA, eB, eG = torch.cat([eA, dan_rows]), torch.cat([eB, dan_rows]), torch.cat([eG, torch.ones(dan_rows.numel())])
    na = nb + n_dan
    assert torch.equal(out["srcA"][:na].long(), eA)
    assert torch.equal(out["srcB"][:na].long(), eB)
    assert torch.allclose(out["gam"][:na], eG.float(), rtol=0, atol=0)
    # mix / unmix are adjoint:  <mix(Z), dOut> == <Z, unmix(dOut)>
    H = 24
    Z, dOut = torch.randn(nb, H), torch.randn(R, H)
    dOut[na:] = 0
    mixed = torch.empty(R, H).cuda()
    call("scnb_mix_rows", ptr(D(Z)), H, ptr(D(out["srcA"])), ptr(D(out["srcB"])), ptr(D(out["gam"])),
         ptr(D(out["counts"][3:])), R, ptr(mixed), _st())
    want = eG[:, None].double() * Z[eA].double() + (1 - eG[:, None].double()) * Z[eB].double()
    _chk("mix", mixed[:na].cpu(), want)
    assert float(mixed[na:].abs().sum()) == 0
    dZ = torch.empty(nb, H).cuda()
    call("scnb_unmix_rows", ptr(D(dOut)), H, ptr(D(out["inv3"])), ptr(D(out["coef3"])), nb, ptr(dZ), _st())
    lhs = float((want * dOut[:na].double()).sum())
    rhs = float((Z.double() * dZ.cpu().double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)
    Zd = Z.double().requires_grad_(True)
    ((eG[:, None].double() * Zd[eA] + (1 - eG[:, None].double()) * Zd[eB]) * dOut[:na].double()).sum().backward()
    _chk("unmix", dZ.cpu(), Zd.grad, tol=2e-5)


def test_losses_match_oracle_autograd():
    from oracle import scnym_oracle as O
    from scnym_b200._lib import call, ptr
    torch.manual_seed(0)
    n, C = 45, 9
    z = (torch.randn(n, C) * 2).double().requires_grad_(True)
    Y = torch.softmax(torch.randn(n, C) * 2, 1)
    cw = torch.rand(C) + 0.5
    loss = O.cross_entropy(z, Y.double(), class_weight=cw.double())
    loss.backward()
    dz, lo = torch.empty(n, C).cuda(), torch.zeros(2).cuda()
    call("scnb_soft_ce_fwd_bwd", ptr(D(z.detach().float())), n, C, None, ptr(D(Y)), None, None, None, 0, ptr(D(cw)),
         None, n, 1.0, None, ptr(dz), ptr(lo), 1, None, _st())
    _chk("ce loss", lo[0].cpu(), loss.detach())
    _chk("ce grad", dz.cpu(), z.grad, tol=2e-5)
    assert int(lo[1]) == int((z.argmax(1) == Y.argmax(1)).sum())
    # masked MSE on probabilities with mixed labels, weight 0.7, n_conf < n
    z2 = (torch.randn(n, C) * 2).double().requires_grad_(True)
    a, b = torch.randint(0, n, (n,)), torch.randint(0, n, (n,))
    g = torch.rand(n) * 0.5 + 0.5
    n_conf = 30
    Ym = g[:, None].double() * Y[a].double() + (1 - g[:, None].double()) * Y[b].double()
    per = ((O.softmax(z2) - Ym) ** 2).sum(1)
    mask = torch.zeros(n, dtype=torch.float64)
    mask[:n_conf] = 1
    l2 = (per * mask).mean()
    (0.7 * l2).backward()
    dz2, lo2 = torch.empty(n, C).cuda(), torch.zeros(1).cuda()
    call("scnb_softmax_mse_fwd_bwd", ptr(D(z2.detach().float())), n, C, ptr(D(torch.tensor([n_conf], dtype=torch.int32))),
         ptr(D(Y)), ptr(D(a.int())), ptr(D(b.int())), ptr(D(g)), 0, n, None, 0.7, ptr(dz2), ptr(lo2), _st())
    _chk("mse loss", lo2[0].cpu(), l2.detach())
    _chk("mse grad", dz2.cpu(), z2.grad, tol=2e-5)


@pytest.mark.parametrize("name", ["adadelta", "adam", "adamw", "sgd"])
def test_optim_step_matches_torch_optim(name):
    from scnym_b200._lib import call, ptr, OPTIMIZER_CODES
    torch.manual_seed(1)
    n = 5000
    p0 = torch.randn(n)
    rp = p0.clone().requires_grad_(True)
    kw = dict(lr=1.0 if name == "adadelta" else 0.01, weight_decay=1e-2)
    cls = {"adadelta": torch.optim.Adadelta, "adam": torch.optim.Adam, "adamw": torch.optim.AdamW, "sgd": torch.optim.SGD}[name]
    ropt = cls([rp], momentum=0.9, **kw) if name == "sgd" else cls([rp], **kw)
    p, s1, s2 = p0.cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
    eps = 1e-6 if name == "adadelta" else 1e-8
    hp = torch.tensor([kw["lr"], kw["weight_decay"], 0.9, 0.999, eps, 0.9, 0.9, 0.0]).cuda()
    state = torch.zeros(4, dtype=torch.int64).cuda()
    for t in range(4):
        g = torch.randn(n)
        rp.grad = g.clone()
        ropt.step()
        call("scnb_step_begin", ptr(state), _st())
        call("scnb_optim_step", OPTIMIZER_CODES[name], ptr(p), ptr(D(g)), ptr(s1), ptr(s2), n, ptr(hp), ptr(state), _st())
        _chk(f"{name} t={t}", p.cpu(), rp.detach(), tol=2e-6)


def test_ema_update_and_counters():
    from scnym_b200._lib import call, ptr
    t, p = torch.randn(1000), torch.randn(1000)
    td, state = t.cuda(), torch.zeros(4, dtype=torch.int64).cuda()
    call("scnb_ema_update", ptr(td), ptr(D(p)), 1000, 0.997, ptr(state), _st())
    _chk("ema step0 == copy", td.cpu(), p, tol=0)
    for _ in range(5):
        call("scnb_mm_step_inc", ptr(state), _st())
    td = t.cuda()
    call("scnb_ema_update", ptr(td), ptr(D(p)), 1000, 0.997, ptr(state), _st())
    a = min(1 - 1 / 6, 0.997)
    _chk("ema step5", td.cpu(), a * t + (1 - a) * p, tol=1e-6)


def test_bad_arguments_raise():
    from scnym_b200._lib import call, ScnbError
    with pytest.raises(ScnbError):
        call("scnb_sgemm", 0, 0, 4, 4, 4, None, 4, None, 4, None, 4, 1.0, 0.0, None, 0, None, _st())


@pytest.mark.parametrize("n,H0", [(512, 256), (37, 128), (300, 256), (2500, 128)])
def test_csr_linear_fwd_split_rows(n, H0):
    """Warps-per-row split (small batches) and the one-warp-per-row form agree with the dense product."""
    from oracle import scnym_oracle as O
    from scnym_b200._lib import call, ptr
    from scnym_b200.dataprep import DeviceCSR, gather_rows
    G = 3000
    ip, ii, dd, _ = O.synth_csr(n, G, 0.05, 3, seed=n)
    X = O.csr_to_dense(ip, ii, dd, G, np.arange(n))
    ds = DeviceCSR.from_arrays(ip, ii, dd, G)
    b = gather_rows(ds, torch.arange(n).cuda())
    torch.manual_seed(n)
    WT, b1 = torch.randn(G, H0) * 0.1, torch.randn(H0)
    Z = torch.empty(n, H0).cuda()
    call("scnb_csr_linear_fwd", ptr(b.indptr), ptr(b.indices), ptr(b.data), n, ptr(D(WT)), ptr(D(b1)), H0, ptr(Z), _st())
    _chk("csr fwd split", Z.cpu(), X.double() @ WT.double() + b1.double())


@pytest.mark.parametrize("opt", ["none", "adadelta", "adam", "adamw", "sgd"])
@pytest.mark.parametrize("n,H0", [(96, 128), (1100, 64)])   # dZ slice staged in shared memory / read through L2
def test_csr_w1_bwd_optim_matches_dense_and_torch_optim(opt, n, H0):
    from oracle import scnym_oracle as O
    from scnym_b200._lib import call, ptr, OPTIMIZER_CODES
    from scnym_b200.dataprep import DeviceCSR, gather_rows
    G = 700
    ip, ii, dd, _ = O.synth_csr(n, G, 0.06, 3, seed=5)
    dd = dd.copy()
    dd[::7] = 0.0                                   # stored zeros (dropped entries) must be skipped
    X = O.csr_to_dense(ip, ii, dd, G, np.arange(n))
    ds = DeviceCSR.from_arrays(ip, ii, dd, G)
    b = gather_rows(ds, torch.arange(n).cuda())
    nnz = int(ip[n])
    cnt = torch.zeros(G, dtype=torch.int32).cuda()
    c_ptr, cur = torch.zeros(G + 1, dtype=torch.int32).cuda(), torch.zeros(G, dtype=torch.int32).cuda()
    ent = torch.zeros(nnz, 2, dtype=torch.int32).cuda()
    torch.manual_seed(3)
    W0 = torch.randn(G, H0) * 0.1
    W, s1, s2 = W0.cuda(), torch.zeros(G, H0).cuda(), torch.zeros(G, H0).cuda()
    rp = W0.clone().requires_grad_(True)
    kw = dict(lr=1.0 if opt == "adadelta" else 0.01, weight_decay=1e-2)
    cls = {"adadelta": torch.optim.Adadelta, "adam": torch.optim.Adam, "adamw": torch.optim.AdamW, "sgd": torch.optim.SGD}.get(opt)
    ropt = None if cls is None else (cls([rp], momentum=0.9, **kw) if opt == "sgd" else cls([rp], **kw))
    eps = 1e-6 if opt == "adadelta" else 1e-8
    hp = torch.tensor([kw["lr"], kw["weight_decay"], 0.9, 0.999, eps, 0.9, 0.9, 0.0]).cuda()
    state = torch.zeros(4, dtype=torch.int64).cuda()
    for t in range(3):
        dZ = torch.randn(n, H0)
        want_g = (X.double().t() @ dZ.double()).float()
        call("scnb_csc_count", ptr(b.indices), ptr(b.data), nnz, ptr(cnt), _st())
        call("scnb_csr_to_csc", ptr(b.indptr), ptr(b.indices), ptr(b.data), n, G, ptr(cnt), ptr(c_ptr), ptr(cur), ptr(ent), _st())
        assert int(cnt.abs().sum()) == 0                      # histogram reset for the next step
        assert int(c_ptr[G]) == int((dd[:nnz] != 0).sum())
        grad = torch.full((G, H0), 7.0).cuda()                # every gene row must be overwritten
        call("scnb_step_begin", ptr(state), _st())
        if opt == "none":
            call("scnb_csr_w1_bwd_optim", -1, ptr(c_ptr), ptr(ent), ptr(D(dZ)), n, H0, G, None, ptr(grad), None, None, None, None, _st())
        else:
            call("scnb_csr_w1_bwd_optim", OPTIMIZER_CODES[opt], ptr(c_ptr), ptr(ent), ptr(D(dZ)), n, H0, G, ptr(W), ptr(grad),
                 ptr(s1), ptr(s2), ptr(hp), ptr(state), _st())
            rp.grad = want_g.clone()
            ropt.step()
            _chk(f"{opt} W t={t}", W.cpu(), rp.detach(), tol=2e-5)  # summation order of dW1 varies between runs
        _chk("dW1", grad.cpu(), want_g, tol=1e-5)
    if opt == "adam":   # gradient not materialised: same update with grad_out = NULL
        W2, a2, b2 = W0.cuda(), torch.zeros(G, H0).cuda(), torch.zeros(G, H0).cuda()
        st2 = torch.zeros(4, dtype=torch.int64).cuda()
        call("scnb_step_begin", ptr(st2), _st())
        call("scnb_csc_count", ptr(b.indices), ptr(b.data), nnz, ptr(cnt), _st())
        call("scnb_csr_to_csc", ptr(b.indptr), ptr(b.indices), ptr(b.data), n, G, ptr(cnt), ptr(c_ptr), ptr(cur), ptr(ent), _st())
        call("scnb_csr_w1_bwd_optim", 1, ptr(c_ptr), ptr(ent), ptr(D(dZ)), n, H0, G, ptr(W2), None, ptr(a2), ptr(b2), ptr(hp),
             ptr(st2), _st())
        r2 = W0.clone().requires_grad_(True)
        o2 = torch.optim.Adam([r2], **kw)
        r2.grad = want_g.clone()
        o2.step()
        _chk("adam no-grad-out", W2.cpu(), r2.detach(), tol=5e-6)


@pytest.mark.parametrize("M,N,K,tA,tB,beta", [(1024, 256, 256, 0, 1, 0.0), (512, 256, 256, 0, 0, 0.0), (256, 256, 1024, 1, 0, 1.0),
                                               (16, 256, 512, 1, 0, 1.0), (1024, 2, 256, 0, 1, 0.0), (100, 40, 300, 1, 0, 1.0)])
def test_sgemm_small_tiles_and_split_k(M, N, K, tA, tB, beta):
    from scnym_b200._lib import call, ptr
    torch.manual_seed(M + N + K)
    A = torch.randn((K, M) if tA else (M, K))
    B = torch.randn((N, K) if tB else (K, N))
    C0 = torch.randn(M, N)
    want = (A.t() if tA else A).double() @ (B.t() if tB else B).double() + beta * C0.double()
    Cd = C0.cuda().clone()
    Ad, Bd = A.cuda(), B.cuda()
    call("scnb_sgemm", tA, tB, M, N, K, ptr(Ad), Ad.shape[1], ptr(Bd), Bd.shape[1], ptr(Cd), N, 1.0, beta, None, 0, None, _st())
    _chk(f"sgemm {M}x{N}x{K}", Cd.cpu(), want.float(), tol=2e-5)


def test_bn_eval_fwd_large_rows():
    from scnym_b200._lib import call, ptr
    n, H = 5000, 96
    Z = torch.randn(n, H)
    g, b, rm, rv = torch.rand(H) + 0.5, torch.randn(H), torch.randn(H), torch.rand(H) + 0.5
    A, pre = torch.empty(n + 8, H).cuda(), torch.empty(n + 8, H).cuda()
    Zd = torch.cat([Z, torch.ones(8, H)]).cuda()
    seg = torch.tensor([0, n], dtype=torch.int32).cuda()
    call("scnb_bn_act_fwd", ptr(Zd), ptr(A), ptr(pre), H, 1, ptr(seg), n + 8, 0, ptr(D(g)), ptr(D(b)), ptr(D(rm)), ptr(D(rv)),
         None, None, None, 0, 0.0, 1, None, None, _st())
    y = (Z.double() - rm.double()) / torch.sqrt(rv.double() + 1e-5) * g.double() + b.double()
    _chk("bn eval pre", pre[:n].cpu(), y, tol=2e-6)
    _chk("bn eval relu", A[:n].cpu(), y.clamp_min(0), tol=2e-6)
    assert float(A[n:].abs().max()) == 0.0 and float(pre[n:].abs().max()) == 0.0


@pytest.mark.parametrize("n,G,H0,ksplits", [(300, 1000, 128, 1), (300, 1000, 256, 0), (130, 777, 64, 3), (512, 2048, 256, 0),
                                            (1, 64, 32, 1)])
def test_csr_linear_fwd_tensor_core_matches_dense(n, G, H0, ksplits):
    """densify-then-tcgen05 first layer (bf16x3 split, fp32 TMEM accumulation) against the fp64 dense product."""
    from oracle import scnym_oracle as O
    from scnym_b200._lib import call, lib, ptr
    from scnym_b200.dataprep import DeviceCSR, gather_rows
    ip, ii, dd, _ = O.synth_csr(n, G, 0.07, 3, seed=n + H0)
    X = O.csr_to_dense(ip, ii, dd, G, np.arange(n))
    ds = DeviceCSR.from_arrays(ip, ii, dd, G)
    torch.manual_seed(G)
    WT, b1 = torch.randn(G, H0) * 0.1, torch.randn(H0)
    n_kt = (G + 63) // 64
    hi = torch.zeros(n_kt * H0 * 64, dtype=torch.int16).cuda()
    lo = torch.zeros(n_kt * H0 * 64, dtype=torch.int16).cuda()
    WTd = D(WT)
    call("scnb_w1_planes", ptr(WTd), G, H0, ptr(hi), ptr(lo), _st())
    want = X.double() @ WT.double() + b1.double()
    # dataset CSR (int64 indptr), rows [0, n)
    Z = torch.full((n, H0), 7.0).cuda()
    call("scnb_csr_linear_fwd_tc_i64", ptr(ds.indptr), ptr(ds.indices), ptr(ds.data), n, G, H0, ptr(hi), ptr(lo), ptr(D(b1)), ptr(Z),
         ksplits, _st())
    torch.cuda.synchronize()
    _chk("tc fwd i64", Z.cpu(), want, tol=3e-5)
    # batch CSR (int32 indptr), permuted rows, no bias
    rows = torch.randperm(n)
    b = gather_rows(ds, rows.cuda())
    Z2 = torch.full((n, H0), -3.0).cuda()
    call("scnb_csr_linear_fwd_tc", ptr(b.indptr), ptr(b.indices), ptr(b.data), n, G, H0, ptr(hi), ptr(lo), None, ptr(Z2), ksplits, _st())
    torch.cuda.synchronize()
    _chk("tc fwd batch", Z2.cpu(), X[rows].double() @ WT.double(), tol=3e-5)
    if ksplits == 0:
        # the same product with every CTA's starting entry precomputed (scnb_row_fwd_splits; one launch with the dW1 gene
        # blocks through scnb_row_gene_splits2): the form the engine uses when the batch was assembled one step ahead
        import ctypes
        ks, gp = ctypes.c_int(0), ctypes.c_int(0)
        call("scnb_csr_linear_fwd_tc_geometry", n, G, H0, ctypes.addressof(ks), ctypes.addressof(gp))
        assert ks.value >= 1 and gp.value % 32 == 0 and ks.value * gp.value >= G
        sp = torch.full(((ks.value + 1) * n,), -1, dtype=torch.int32).cuda()
        call("scnb_row_fwd_splits", ptr(b.indptr), ptr(b.indices), n, G, H0, ptr(sp), _st())
        ipb, idb = b.indptr.cpu().numpy(), b.indices.cpu().numpy()
        want_sp = np.stack([[ipb[r] + np.searchsorted(idb[ipb[r]:ipb[r + 1]], s * gp.value) for r in range(n)]
                            for s in range(ks.value + 1)]).reshape(-1)
        assert np.array_equal(sp.cpu().numpy(), want_sp)
        gt = lib().scnb_w1tc_gene_block(G)
        nblk = (G + gt - 1) // gt
        s1, s1b = torch.zeros((nblk + 1) * n, dtype=torch.int32).cuda(), torch.zeros((nblk + 1) * n, dtype=torch.int32).cuda()
        sp2 = torch.zeros_like(sp)
        call("scnb_row_gene_splits", ptr(b.indptr), ptr(b.indices), n, G, ptr(s1), _st())
        call("scnb_row_gene_splits2", ptr(b.indptr), ptr(b.indices), n, G, H0, ptr(s1b), ptr(sp2), _st())
        assert torch.equal(s1, s1b) and torch.equal(sp, sp2)
        for prezeroed in (0, -1):
            Z3 = torch.zeros(n, H0).cuda() if prezeroed else torch.full((n, H0), 5.0).cuda()
            call("scnb_csr_linear_fwd_tc_splits", ptr(b.indptr), ptr(b.indices), ptr(b.data), n, G, H0, ptr(hi), ptr(lo), ptr(D(b1)),
                 ptr(Z3), prezeroed, ptr(sp), _st())
            torch.cuda.synchronize()
            _chk("tc fwd batch, split table", Z3.cpu(), X[rows].double() @ WT.double() + b1.double(), tol=3e-5)


@pytest.mark.parametrize("n,K,N", [(300, 256, 256), (1000, 64, 128), (5, 32, 32)])
def test_tensor_core_hidden_layer_from_planes(n, K, N):
    """BatchNorm(eval)->ReLU emitting bf16 operand planes, nn.Linear weight planes, Y = A W^T + b on tcgen05."""
    from scnym_b200._lib import call, ptr
    torch.manual_seed(n + K)
    Z = torch.randn(n, K)
    g, b, rm, rv = torch.rand(K) + 0.5, torch.randn(K) * 0.3, torch.randn(K) * 0.2, torch.rand(K) + 0.5
    W, bias = torch.randn(N, K) * 0.1, torch.randn(N)
    rows_pad = (n + 255) // 256 * 256
    ahi, alo = torch.zeros(rows_pad * K, dtype=torch.int16).cuda(), torch.zeros(rows_pad * K, dtype=torch.int16).cuda()
    whi, wlo = torch.zeros(N * K, dtype=torch.int16).cuda(), torch.zeros(N * K, dtype=torch.int16).cuda()
    A, pre = torch.empty(n, K).cuda(), torch.empty(n, K).cuda()
    call("scnb_bn_eval_planes", ptr(D(Z)), n, K, ptr(D(g)), ptr(D(b)), ptr(D(rm)), ptr(D(rv)), 1, ptr(A), ptr(pre), ptr(ahi), ptr(alo), _st())
    y = (Z.double() - rm.double()) / torch.sqrt(rv.double() + 1e-5) * g.double() + b.double()
    _chk("bn planes pre", pre.cpu(), y, tol=2e-6)
    _chk("bn planes act", A.cpu(), y.clamp_min(0), tol=2e-6)
    call("scnb_linear_planes", ptr(D(W)), N, K, ptr(whi), ptr(wlo), _st())
    Y = torch.full((n, N), 9.0).cuda()
    call("scnb_linear_fwd_tc_planes", ptr(ahi), ptr(alo), n, K, N, ptr(whi), ptr(wlo), ptr(D(bias)), ptr(Y), _st())
    torch.cuda.synchronize()
    _chk("tc linear", Y.cpu(), y.clamp_min(0) @ W.double().t() + bias.double(), tol=3e-5)


def test_head_logits_small_output_linear():
    from scnym_b200._lib import call, ptr
    n, H, C = 777, 256, 11
    A, W, b = torch.randn(n, H), torch.randn(C, H) * 0.1, torch.randn(C)
    out = torch.ones(n, C).cuda()
    call("scnb_head_logits", ptr(D(A)), ptr(D(W)), ptr(D(b)), H, C, n, ptr(out), 0, _st())
    want = A.double() @ W.double().t() + b.double()
    _chk("head logits", out.cpu(), want)
    call("scnb_head_logits", ptr(D(A)), ptr(D(W)), ptr(D(b)), H, C, n, ptr(out), 1, _st())
    _chk("head logits acc", out.cpu(), 2 * want)


@pytest.mark.parametrize("opt", ["none", "adam", "adadelta", "sgd"])
@pytest.mark.parametrize("n,G,H0", [(96, 700, 128), (512, 3000, 256), (70, 300, 64), (512, 20000, 256)])
def test_csr_w1_bwd_optim_tensor_core(opt, n, G, H0):
    """tcgen05 dW1 + optimizer kernel (X^T dZ1 with the update in the epilogue) against the dense product / torch.optim."""
    from oracle import scnym_oracle as O
    from scnym_b200._lib import call, lib, ptr, OPTIMIZER_CODES
    from scnym_b200.dataprep import DeviceCSR, gather_rows
    ip, ii, dd, _ = O.synth_csr(n, G, 0.06, 3, seed=n)
    dd = dd.copy()
    dd[::9] = 0.0
    X = O.csr_to_dense(ip, ii, dd, G, np.arange(n))
    ds = DeviceCSR.from_arrays(ip, ii, dd, G)
    b = gather_rows(ds, torch.arange(n).cuda())
    GT = lib().scnb_w1tc_gene_block(G)
    n_blk = (G + GT - 1) // GT
    splits = torch.zeros((n_blk + 1) * n, dtype=torch.int32).cuda()
    n_pad = (n + 31) // 32 * 32
    dz_hi, dz_lo = torch.zeros(n_pad * H0, dtype=torch.int16).cuda(), torch.zeros(n_pad * H0, dtype=torch.int16).cuda()
    torch.manual_seed(3)
    W0 = torch.randn(G, H0) * 0.1
    W, s1, s2 = W0.cuda(), torch.zeros(G, H0).cuda(), torch.zeros(G, H0).cuda()
    rp = W0.clone().requires_grad_(True)
    kw = dict(lr=1.0 if opt == "adadelta" else 0.01, weight_decay=1e-2)
    cls = {"adadelta": torch.optim.Adadelta, "adam": torch.optim.Adam, "sgd": torch.optim.SGD}.get(opt)
    ropt = None if cls is None else (cls([rp], momentum=0.9, **kw) if opt == "sgd" else cls([rp], **kw))
    eps = 1e-6 if opt == "adadelta" else 1e-8
    hp = torch.tensor([kw["lr"], kw["weight_decay"], 0.9, 0.999, eps, 0.9, 0.9, 0.0]).cuda()
    state = torch.zeros(4, dtype=torch.int64).cuda()
    n_kt = (G + 63) // 64
    pl_hi, pl_lo = (torch.zeros(n_kt * H0 * 64, dtype=torch.int16).cuda() for _ in range(2))
    want_hi, want_lo = (torch.zeros(n_kt * H0 * 64, dtype=torch.int16).cuda() for _ in range(2))
    call("scnb_row_gene_splits", ptr(b.indptr), ptr(b.indices), n, G, ptr(splits), _st())
    for t in range(2):
        dZ = torch.randn(n, H0)
        want_g = (X.double().t() @ dZ.double()).float()
        dZd = D(dZ)
        call("scnb_w1_planes", ptr(dZd), n, H0, ptr(dz_hi), ptr(dz_lo), _st())
        grad = torch.full((G, H0), 7.0).cuda()
        call("scnb_step_begin", ptr(state), _st())
        if opt == "none":
            call("scnb_csr_w1_bwd_optim_tc", -1, ptr(splits), ptr(b.indices), ptr(b.data), n, H0, G, ptr(dz_hi), ptr(dz_lo), None,
                 ptr(grad), None, None, None, None, None, None, _st())
        else:
            call("scnb_csr_w1_bwd_optim_tc", OPTIMIZER_CODES[opt], ptr(splits), ptr(b.indices), ptr(b.data), n, H0, G, ptr(dz_hi),
                 ptr(dz_lo), ptr(W), ptr(grad), ptr(s1), ptr(s2), ptr(hp), ptr(state), ptr(pl_hi), ptr(pl_lo), _st())
            # the operand planes emitted by the epilogue are those scnb_w1_planes builds from the updated W1T
            call("scnb_w1_planes", ptr(W), G, H0, ptr(want_hi), ptr(want_lo), _st())
            torch.cuda.synchronize()
            assert torch.equal(pl_hi, want_hi) and torch.equal(pl_lo, want_lo), "operand planes of the updated W1T"
        torch.cuda.synchronize()
        _chk("tc dW1", grad.cpu(), want_g, tol=3e-5)
        if opt != "none":
            # the reference optimizer is fed the kernel's own gradient: the update rule is what is checked here
            rp.grad = grad.cpu().clone()
            ropt.step()
            _chk(f"tc {opt} W t={t}", W.cpu(), rp.detach(), tol=5e-6)


@pytest.mark.parametrize("opt", ["sgd", "adam", "adadelta"])
@pytest.mark.parametrize("world,total", [(2, 4096 + 8), (3, 40), (4, 1 << 18)])
def test_dp_reduce_optim_bcast_emulated_ranks(opt, world, total):
    """Fused gradient reduce-scatter + optimizer + parameter all-gather over peer memory: `world` emulated ranks on one
    GPU (one stream each, their kernels wait for each other on the device) against torch.optim on the averaged gradient."""
    from scnym_b200._lib import call, ptr, OPTIMIZER_CODES
    rnd = lambda b: (b + 255) // 256 * 256
    off_flags, off_grad = 256, 256 + 1024
    off_flat = off_grad + rnd(4 * total)
    nbytes = off_flat + rnd(4 * total)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(world)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    view = lambda b, off: b[off: off + 4 * total].view(torch.float32)
    torch.manual_seed(5)
    p0 = torch.randn(total)
    kw = dict(lr=1.0 if opt == "adadelta" else 0.01, weight_decay=1e-2)
    cls = {"adadelta": torch.optim.Adadelta, "adam": torch.optim.Adam, "sgd": torch.optim.SGD}[opt]
    rp = p0.clone().requires_grad_(True)
    ropt = cls([rp], momentum=0.9, **kw) if opt == "sgd" else cls([rp], **kw)
    eps = 1e-6 if opt == "adadelta" else 1e-8
    hp = torch.tensor([kw["lr"], kw["weight_decay"], 0.9, 0.999, eps, 0.9, 0.9, 0.0]).cuda()
    state = torch.zeros(4, dtype=torch.int64).cuda()
    s1 = [torch.zeros(total).cuda() for _ in range(world)]
    s2 = [torch.zeros(total).cuda() for _ in range(world)]
    for r in range(world):
        view(bufs[r], off_flat).copy_(p0)
    streams = [torch.cuda.Stream() for _ in range(world)]
    for step in range(3):
        grads = [torch.randn(total) for _ in range(world)]
        for r in range(world):
            view(bufs[r], off_grad).copy_(grads[r])
        call("scnb_step_begin", ptr(state), _st())
        torch.cuda.synchronize()
        for r in range(world):
            call("scnb_dp_reduce_optim_bcast", OPTIMIZER_CODES[opt], ptr(ptrs), r, world, off_flags, off_grad, off_flat, total,
                 ptr(s1[r]), ptr(s2[r]), ptr(hp), ptr(state), ptr(state), streams[r].cuda_stream)
        torch.cuda.synchronize()
        rp.grad = torch.stack(grads).sum(0) / world
        ropt.step()
        for r in range(world):
            _chk(f"{opt} rank {r} step {step}", view(bufs[r], off_flat).cpu(), rp.detach(), tol=2e-5)
        assert all(torch.equal(view(bufs[0], off_flat), view(bufs[r], off_flat)) for r in range(world)), "replicas differ"


@pytest.mark.parametrize("n,H", [(512, 256), (70, 64), (33, 128)])
def test_unmix_rows_planes_equals_unmix_then_planes(n, H):
    """Fused MixUp adjoint + bf16 operand planes == scnb_unmix_rows followed by scnb_w1_planes, bit for bit."""
    from scnym_b200._lib import call, ptr
    torch.manual_seed(n)
    R = 2 * n
    dOut = torch.randn(R, H).cuda()
    inv = torch.randint(-1, R, (3, n), dtype=torch.int32).cuda()
    coef = torch.rand(3, n).cuda()
    n_pad = (n + 31) // 32 * 32
    dz_a, dz_b = torch.zeros(n, H).cuda(), torch.zeros(n, H).cuda()
    hi_a, lo_a, hi_b, lo_b = (torch.zeros(n_pad * H, dtype=torch.int16).cuda() for _ in range(4))
    call("scnb_unmix_rows", ptr(dOut), H, ptr(inv), ptr(coef), n, ptr(dz_a), _st())
    call("scnb_w1_planes", ptr(dz_a), n, H, ptr(hi_a), ptr(lo_a), _st())
    call("scnb_unmix_rows_planes", ptr(dOut), H, ptr(inv), ptr(coef), n, ptr(dz_b), ptr(hi_b), ptr(lo_b), _st())
    torch.cuda.synchronize()
    assert torch.equal(dz_a, dz_b)
    assert torch.equal(hi_a, hi_b) and torch.equal(lo_a, lo_b)


@pytest.mark.parametrize("lay,M,N,K,extras", [
    (0, 1024, 320, 1000, "bias_relu"), (0, 300, 256, 64, "plain"), (0, 2048, 1024, 1024, "bias"),
    (1, 1024, 320, 1000, "gate"), (1, 384, 64, 200, "alpha"),
    (2, 256, 512, 8192, "beta1"), (2, 1024, 1024, 4096, "beta1"), (2, 128, 72, 260, "plain")])
@pytest.mark.parametrize("planes", [False, True])
def test_sgemm_tcgen05_path_matches_fp64(lay, M, N, K, extras, planes):
    """scnb_sgemm's tcgen05 kernels (bf16x3) for the three nn.Linear layouts (forward, input gradient, weight gradient incl.
    the K split), ragged M / N / K and every epilogue option: operands split from fp32 inside the kernel, and -- with a
    workspace registered for the stream -- the planes form (operand planes written once, 256 x 256 tiles)."""
    import ctypes
    from scnym_b200._lib import call, ptr
    call("scnb_gemm_tc_min_flops", 1)
    ws = None
    try:
        if planes:
            need = ctypes.c_longlong(0)
            call("scnb_gemm_tc_workspace_bytes", M, N, K, ctypes.byref(need))
            ws = torch.full((need.value + 1024,), 0x7f, dtype=torch.uint8).cuda()   # garbage: every tile byte must be written
            call("scnb_gemm_tc_workspace", _st(), ws.data_ptr() + (-ws.data_ptr()) % 1024, need.value)
        n0 = ctypes.c_longlong(0)
        call("scnb_gemm_tc_planes_count", ctypes.byref(n0))
        torch.manual_seed(lay * 100 + M)
        tA, tB = (lay == 2), (lay == 0)
        A = torch.randn((K, M) if tA else (M, K)).cuda()
        B = torch.randn((N, K) if tB else (K, N)).cuda()
        C0 = torch.randn(M, N).cuda()
        C = C0.clone()
        bias = torch.randn(N).cuda() if "bias" in extras else None
        gate = (torch.rand(M, N) > 0.3).float().cuda() if extras == "gate" else None
        alpha = -0.37 if extras == "alpha" else 1.0
        beta = 1.0 if extras == "beta1" else 0.0
        relu = 1 if "relu" in extras else 0
        call("scnb_sgemm", int(tA), int(tB), M, N, K, ptr(A), A.shape[1], ptr(B), B.shape[1], ptr(C), N, alpha, beta, ptr(bias), relu,
             ptr(gate), _st())
        torch.cuda.synchronize()
        opA = A.double().t() if tA else A.double()
        opB = B.double().t() if tB else B.double()
        want = alpha * (opA @ opB) + beta * C0.double()
        if bias is not None:
            want = want + bias.double()
        if relu:
            want = want.clamp_min(0)
        if gate is not None:
            want = want * (gate.double() > 0)
        _chk(f"tc gemm lay {lay} {extras} planes={planes}", C.cpu(), want.float().cpu(), tol=3e-5)
        n1 = ctypes.c_longlong(0)
        call("scnb_gemm_tc_planes_count", ctypes.byref(n1))
        assert n1.value - n0.value == (1 if planes else 0)
    finally:
        call("scnb_gemm_tc_min_flops", 2_000_000_000)
        if ws is not None:
            call("scnb_gemm_tc_workspace", _st(), None, 0)


# ---- operand planes written by the producer of a tensor (tc_gemm.cu planes form, dense_ops.cu PlaneSink) ----------------
def _planes_encode(X):
    """fp32 [R, K] -> (hi, lo) bf16 planes tiled [R/256][K/32], each tile a SWIZZLE_64B K-major image of 256 x 32 (the layout
    of k_gemm_operand_planes / PlaneSink): 16-byte chunk c of row r sits at chunk position c ^ ((r >> 1) & 3)."""
    R, K = X.shape
    Rp, Kp = (R + 255) // 256 * 256, (K + 31) // 32 * 32
    Xp = torch.zeros(Rp, Kp, dtype=torch.float32, device=X.device)
    Xp[:R, :K] = X
    hi = Xp.to(torch.bfloat16)
    lo = (Xp - hi.float()).to(torch.bfloat16)
    rr = torch.arange(256, device=X.device)
    pc = torch.arange(4, device=X.device)
    src = (pc[None, :] ^ ((rr[:, None] >> 1) & 3))           # logical chunk stored at physical position pc of row rr

    def tile(P):
        T = P.view(Rp // 256, 256, Kp // 32, 4, 8).permute(0, 2, 1, 3, 4)          # [rt, kt, rr, c, j]
        idx = src[None, None, :, :, None].expand(T.shape[0], T.shape[1], 256, 4, 8)
        return torch.gather(T, 3, idx).contiguous().view(-1)
    return tile(hi), tile(lo)


def _aligned(n_elems, dtype):
    buf = torch.zeros(n_elems + 1024, dtype=dtype).cuda()
    off = ((-buf.data_ptr()) % 1024) // buf.element_size()
    return buf, buf[off:off + n_elems]


@pytest.mark.parametrize("H,segs,maxr", [(256, [(0, 3000), (3000, 9001), (9001, 12000)], 12400), (64, [(0, 300), (300, 520)], 600)])
def test_bn_act_kernels_emit_operand_planes_of_bound_outputs(H, segs, maxr):
    """scnb_bn_act_fwd / scnb_bn_act_bwd with planes bound to their outputs: the planes are the bf16 hi / lo split of the fp32
    output, in the tiled layout, zero in the rows past the tensor's -- bit for bit."""
    from scnym_b200._lib import call, ptr
    torch.manual_seed(H + maxr)
    Z = (torch.randn(maxr, H) * 2 + 0.5).cuda()
    g, b = (torch.rand(H) + 0.5).cuda(), (torch.randn(H) * 0.1).cuda()
    keep = (torch.rand(maxr, H) < 0.5).to(torch.uint8).cuda()
    dA = torch.randn(maxr, H).cuda()
    n_act = segs[-1][1]
    seg = torch.tensor([s[0] for s in segs] + [n_act], dtype=torch.int32).cuda()
    A, dZ = torch.full((maxr, H), 7.0).cuda(), torch.full((maxr, H), 7.0).cuda()
    stats = torch.zeros(len(segs), 3, H).cuda()
    Rp = (maxr + 255) // 256 * 256
    bufs = [_aligned(Rp * H, torch.int16) for _ in range(4)]
    for _, v in bufs:
        v.fill_(0x7f7f)
    try:
        call("scnb_gemm_tc_bind_planes", ptr(A), maxr, H, ptr(bufs[0][1]), ptr(bufs[1][1]))
        call("scnb_gemm_tc_bind_planes", ptr(dZ), maxr, H, ptr(bufs[2][1]), ptr(bufs[3][1]))
        call("scnb_bn_act_fwd", ptr(Z), ptr(A), None, H, len(segs), ptr(seg), maxr, 1, ptr(g), ptr(b), None, None, ptr(stats),
             ptr(keep), None, 0, 0.5, 1, None, None, _st())
        dg, db = torch.zeros(H).cuda(), torch.zeros(H).cuda()
        call("scnb_bn_act_bwd", ptr(dA), ptr(Z), ptr(A), H, len(segs), ptr(seg), maxr, ptr(g), ptr(stats), ptr(keep), None, 0,
             0.5, 1, ptr(dZ), ptr(dg), ptr(db), None, None, _st())
        torch.cuda.synchronize()
        for name, X, (hi, lo) in (("A", A, (bufs[0][1], bufs[1][1])), ("dZ", dZ, (bufs[2][1], bufs[3][1]))):
            assert float(X[n_act:].abs().sum()) == 0.0
            want_hi, want_lo = _planes_encode(X)
            assert torch.equal(hi.view(torch.bfloat16), want_hi), name + " hi plane"
            assert torch.equal(lo.view(torch.bfloat16), want_lo), name + " lo plane"
    finally:
        call("scnb_gemm_tc_bind_planes", ptr(A), maxr, H, None, None)
        call("scnb_gemm_tc_bind_planes", ptr(dZ), maxr, H, None, None)


@pytest.mark.parametrize("lay,M,N,K", [(0, 1000, 320, 512), (1, 600, 96, 256), (2, 256, 512, 8192), (2, 1024, 1024, 4000),
                                       (2, 128, 160, 700)])
def test_sgemm_tcgen05_takes_bound_operand_planes(lay, M, N, K):
    """Products whose operands already exist as planes (bound by scnb_gemm_tc_bind_planes): the row-major A operand of the
    forward / input-gradient layouts is taken as it is, and the weight-gradient layout reads BOTH operands' row-major planes as
    MN-major operands (k_gemm_planes_mn).  The bound planes are the only copy of the data that is right: the fp32 tensors handed
    to scnb_sgemm for the bound operands hold garbage."""
    import ctypes
    from scnym_b200._lib import call, ptr
    call("scnb_gemm_tc_min_flops", 1)
    torch.manual_seed(lay * 1000 + M + K)
    tA, tB = (lay == 2), (lay == 0)
    A = torch.randn((K, M) if tA else (M, K)).cuda()
    B = torch.randn((N, K) if tB else (K, N)).cuda()
    C0 = torch.randn(M, N).cuda()
    C = C0.clone()
    beta = 1.0 if lay == 2 else 0.0
    need = ctypes.c_longlong(0)
    call("scnb_gemm_tc_workspace_bytes", M, N, K, ctypes.byref(need))
    ws, wsv = _aligned(need.value, torch.uint8)
    bound = [A, B] if lay == 2 else [A]
    keepalive, fakes = [], []
    try:
        call("scnb_gemm_tc_workspace", _st(), ptr(wsv), need.value)
        for X in bound:
            hi, lo = _planes_encode(X)
            bh, vh = _aligned(hi.numel(), torch.bfloat16)
            bl, vl = _aligned(lo.numel(), torch.bfloat16)
            vh.copy_(hi)
            vl.copy_(lo)
            keepalive += [bh, bl]
            fake = torch.full_like(X, 1e30)          # what scnb_sgemm is handed: must never be read
            fakes.append(fake)
            call("scnb_gemm_tc_bind_planes", ptr(fake), X.shape[0], X.shape[1], ptr(vh), ptr(vl))
        n0, n1 = ctypes.c_longlong(0), ctypes.c_longlong(0)
        call("scnb_gemm_tc_bound_count", ctypes.byref(n0))
        fa = fakes[0]
        fb = fakes[1] if lay == 2 else B
        call("scnb_sgemm", int(tA), int(tB), M, N, K, ptr(fa), fa.shape[1], ptr(fb), fb.shape[1], ptr(C), N, 1.0, beta, None, 0, None, _st())
        torch.cuda.synchronize()
        call("scnb_gemm_tc_bound_count", ctypes.byref(n1))
        assert n1.value - n0.value == 1
        opA = A.double().t() if tA else A.double()
        opB = B.double().t() if tB else B.double()
        want = opA @ opB + beta * C0.double()
        _chk(f"tc gemm lay {lay} bound planes", C.cpu(), want.float().cpu(), tol=3e-5)
    finally:
        call("scnb_gemm_tc_min_flops", 2_000_000_000)
        call("scnb_gemm_tc_workspace", _st(), None, 0)
        for f in fakes:
            call("scnb_gemm_tc_bind_planes", ptr(f), f.shape[0], f.shape[1], None, None)


@pytest.mark.parametrize("n,G,H0", [(300, 1000, 128), (700, 2048, 256), (256, 640, 64), (1, 64, 32)])
def test_first_layer_with_fused_bn_eval_planes_equals_the_two_launches(n, G, H0):
    """scnb_csr_linear_fwd_tc_bn_planes_i64 (first layer -> eval BatchNorm -> ReLU -> operand planes, one launch) writes the
    planes scnb_csr_linear_fwd_tc_i64 + scnb_bn_eval_planes write, bit for bit, zeros in the padding rows included."""
    from oracle import scnym_oracle as O
    from scnym_b200._lib import call, ptr
    from scnym_b200.dataprep import DeviceCSR
    ip, ii, dd, _ = O.synth_csr(n, G, 0.07, 3, seed=n + H0 + 1)
    ds = DeviceCSR.from_arrays(ip, ii, dd, G)
    torch.manual_seed(G + n)
    WT, b1 = D(torch.randn(G, H0) * 0.1), D(torch.randn(H0))
    g, b = D(torch.rand(H0) + 0.5), D(torch.randn(H0) * 0.1)
    rm, rv = D(torch.randn(H0) * 0.2), D(torch.rand(H0) + 0.3)
    n_kt = (G + 63) // 64
    hi = torch.zeros(n_kt * H0 * 64, dtype=torch.int16).cuda()
    lo = torch.zeros(n_kt * H0 * 64, dtype=torch.int16).cuda()
    call("scnb_w1_planes", ptr(WT), G, H0, ptr(hi), ptr(lo), _st())
    rows_pad = (n + 255) // 256 * 256
    Z = torch.empty(n, H0).cuda()
    a = [torch.full((rows_pad * H0,), 0x7f7f, dtype=torch.int16).cuda() for _ in range(4)]
    call("scnb_csr_linear_fwd_tc_i64", ptr(ds.indptr), ptr(ds.indices), ptr(ds.data), n, G, H0, ptr(hi), ptr(lo), ptr(b1), ptr(Z), 1, _st())
    call("scnb_bn_eval_planes", ptr(Z), n, H0, ptr(g), ptr(b), ptr(rm), ptr(rv), 1, None, None, ptr(a[0]), ptr(a[1]), _st())
    call("scnb_csr_linear_fwd_tc_bn_planes_i64", ptr(ds.indptr), ptr(ds.indices), ptr(ds.data), n, G, H0, ptr(hi), ptr(lo), ptr(b1),
         ptr(g), ptr(b), ptr(rm), ptr(rv), 1, ptr(a[2]), ptr(a[3]), _st())
    torch.cuda.synchronize()
    assert torch.equal(a[0], a[2]), "hi planes differ"
    assert torch.equal(a[1], a[3]), "lo planes differ"


@pytest.mark.parametrize("n,K,N", [(300, 128, 128), (700, 256, 256), (256, 64, 96)])
def test_hidden_linear_with_fused_bn_eval_planes_equals_the_two_launches(n, K, N):
    """scnb_linear_fwd_tc_planes_bn_planes == scnb_linear_fwd_tc_planes + scnb_bn_eval_planes, bit for bit."""
    from scnym_b200._lib import call, ptr
    torch.manual_seed(n + K + N)
    X = D(torch.randn(n, K))
    W, bias = D(torch.randn(N, K) * 0.1), D(torch.randn(N))
    g, b = D(torch.rand(N) + 0.5), D(torch.randn(N) * 0.1)
    rm, rv = D(torch.randn(N) * 0.2), D(torch.rand(N) + 0.3)
    ones, zeros = D(torch.ones(K)), D(torch.zeros(K))
    rows_pad = (n + 255) // 256 * 256
    i16 = lambda m: torch.full((m,), 0x7f7f, dtype=torch.int16).cuda()
    a_hi, a_lo = i16(rows_pad * K), i16(rows_pad * K)
    # input planes: identity "BatchNorm" (mean 0, var 1 - eps is not exactly representable: use the kernel on Z with gamma = sqrt(1 + eps))
    call("scnb_bn_eval_planes", ptr(X), n, K, ptr(ones), ptr(zeros), ptr(zeros), ptr(ones), 0, None, None, ptr(a_hi), ptr(a_lo), _st())
    w_hi, w_lo = i16(N * K), i16(N * K)
    call("scnb_linear_planes", ptr(W), N, K, ptr(w_hi), ptr(w_lo), _st())
    Y = torch.empty(n, N).cuda()
    o = [i16(rows_pad * N) for _ in range(4)]
    call("scnb_linear_fwd_tc_planes", ptr(a_hi), ptr(a_lo), n, K, N, ptr(w_hi), ptr(w_lo), ptr(bias), ptr(Y), _st())
    call("scnb_bn_eval_planes", ptr(Y), n, N, ptr(g), ptr(b), ptr(rm), ptr(rv), 1, None, None, ptr(o[0]), ptr(o[1]), _st())
    call("scnb_linear_fwd_tc_planes_bn_planes", ptr(a_hi), ptr(a_lo), n, K, N, ptr(w_hi), ptr(w_lo), ptr(bias), ptr(g), ptr(b), ptr(rm),
         ptr(rv), 1, ptr(o[2]), ptr(o[3]), _st())
    torch.cuda.synchronize()
    assert torch.equal(o[0], o[2]), "hi planes differ"
    assert torch.equal(o[1], o[3]), "lo planes differ"


@pytest.mark.parametrize("n,K,N,C,acc,want_pre", [(300, 128, 128, 5, 0, True), (700, 256, 256, 16, 1, False), (256, 64, 96, 1, 0, True)])
def test_last_hidden_linear_with_fused_bn_and_classifier(n, K, N, C, acc, want_pre):
    """scnb_linear_fwd_tc_planes_bn_head == scnb_linear_fwd_tc_planes -> eval BatchNorm (pre-ReLU output kept) -> ReLU ->
    scnb_sgemm classifier: logits (also accumulated into, for ensembles) and embedding."""
    from scnym_b200._lib import call, ptr
    torch.manual_seed(n + K + N + C)
    X = D(torch.randn(n, K))
    W, bias = D(torch.randn(N, K) * 0.1), D(torch.randn(N))
    g, b = D(torch.rand(N) + 0.5), D(torch.randn(N) * 0.1)
    rm, rv = D(torch.randn(N) * 0.2), D(torch.rand(N) + 0.3)
    Wc, bc = D(torch.randn(C, N) * 0.2), D(torch.randn(C))
    ones, zeros = D(torch.ones(K)), D(torch.zeros(K))
    rows_pad = (n + 255) // 256 * 256
    i16 = lambda m: torch.zeros(m, dtype=torch.int16).cuda()
    a_hi, a_lo, w_hi, w_lo = i16(rows_pad * K), i16(rows_pad * K), i16(N * K), i16(N * K)
    call("scnb_bn_eval_planes", ptr(X), n, K, ptr(ones), ptr(zeros), ptr(zeros), ptr(ones), 0, None, None, ptr(a_hi), ptr(a_lo), _st())
    call("scnb_linear_planes", ptr(W), N, K, ptr(w_hi), ptr(w_lo), _st())
    # reference chain (fp32 kernels of this package; the tcgen05 product is shared)
    Y = torch.empty(n, N).cuda()
    call("scnb_linear_fwd_tc_planes", ptr(a_hi), ptr(a_lo), n, K, N, ptr(w_hi), ptr(w_lo), ptr(bias), ptr(Y), _st())
    pre_ref = (Y.double() - rm.double()) / torch.sqrt(rv.double() + 1e-5) * g.double() + b.double()
    init = torch.randn(n, C).cuda()
    want = pre_ref.clamp_min(0) @ Wc.double().t() + bc.double() + (init.double() if acc else 0.0)
    logits = init.clone()
    pre = torch.full((n, N), 7.0).cuda() if want_pre else None
    call("scnb_linear_fwd_tc_planes_bn_head", ptr(a_hi), ptr(a_lo), n, K, N, ptr(w_hi), ptr(w_lo), ptr(bias), ptr(g), ptr(b), ptr(rm),
         ptr(rv), 1, ptr(Wc), ptr(bc), C, ptr(logits), acc, ptr(pre), _st())
    torch.cuda.synchronize()
    _chk("fused head logits", logits.cpu(), want.float().cpu(), tol=2e-6)
    if want_pre:
        _chk("fused head embedding", pre.cpu(), pre_ref.float().cpu(), tol=2e-6)


@pytest.mark.parametrize("R,H,na", [(5000, 256, 4500), (4096, 1024, 4096), (300, 64, 250)])
def test_mix_rows_small_and_tall(R, H, na):
    """scnb_mix_rows (hidden-space MixUp gather, losses.py:811-875 folded behind the first layer): the CTA-per-row form and,
    from 4096 rows, the warp-per-row form; rows past the active count are zeros."""
    from scnym_b200._lib import call, ptr
    torch.manual_seed(R + H)
    n_base = R // 2
    Z = torch.randn(n_base, H).cuda()
    a = torch.randint(0, n_base, (R,), dtype=torch.int32).cuda()
    b = torch.randint(0, n_base, (R,), dtype=torch.int32).cuda()
    b[::7] = a[::7]
    g = torch.rand(R).cuda()
    g[::5] = 1.0
    cnt = torch.tensor([na], dtype=torch.int32).cuda()
    out = torch.full((R, H), 7.0).cuda()
    call("scnb_mix_rows", ptr(Z), H, ptr(a), ptr(b), ptr(g), ptr(cnt), R, ptr(out), _st())
    torch.cuda.synchronize()
    al, bl = a.long(), b.long()
    want = g[:, None] * Z[al] + (1.0 - g)[:, None] * Z[bl]
    same = (al == bl) | (g == 1.0)
    want[same] = Z[al][same]
    want[na:] = 0.0
    _chk("mix_rows", out.cpu(), want.cpu(), tol=1e-6)
    assert float(out[na:].abs().sum()) == 0.0
