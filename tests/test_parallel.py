"""Data-parallel path.
CPU: world_size-2 `gloo` tests of the collective plumbing and of the cross-rank BatchNorm exchange protocol.
GPU: two engine replicas (emulated ranks: two threads, one GPU, host-side rendezvous -- no kernel waits on
another) against the oracle on the CONCATENATED global batch with a block-diagonal MixUp permutation."""
import os
import threading

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import golden_util as GU


def test_shard_bounds_cover_rows_exactly():
    from scnym_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from scnym_b200.parallel import TorchDistComm, merge_bn_sums, shard_bounds
        comm = TorchDistComm()
        assert comm.world == world and comm.rank == rank
        torch.manual_seed(0)
        n_seg, H, rows = 3, 16, [40, 24, 64]
        Z = [torch.randn(r, H) * 2 + 0.7 for r in rows]              # identical on both ranks (same seed)
        # each rank owns a (ragged) shard of every segment
        buf = torch.zeros(n_seg * 2 * H + n_seg)
        for s, z in enumerate(Z):
            a, b = shard_bounds(z.shape[0], rank, world)
            zl = z[a:b]
            buf[(s * 2 + 0) * H:(s * 2 + 1) * H] = zl.sum(0)
            buf[(s * 2 + 1) * H:(s * 2 + 2) * H] = (zl * zl).sum(0)
            buf[n_seg * 2 * H + s] = b - a
        comm.all_reduce_sum(buf)
        mean, var, invstd, cnt = merge_bn_sums(buf, n_seg, H)
        for s, z in enumerate(Z):
            assert int(cnt[s]) == z.shape[0]
            assert GU.rel_err(mean[s], z.mean(0)) < 1e-5
            assert GU.rel_err(var[s], z.var(0, unbiased=False)) < 1e-4
        # gradient averaging over a flat arena + broadcast of the initial replica
        g = torch.full((1000,), float(rank + 1))
        comm.all_reduce_avg(g)
        assert torch.allclose(g, torch.full((1000,), (1 + world) / 2))
        p = torch.full((10,), float(rank))
        comm.broadcast(p)
        assert torch.all(p == 0)
        m = torch.tensor([float(rank)])
        comm.all_reduce_max(m)
        assert float(m) == world - 1
        comm.barrier()
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_gloo_world2_collectives_and_bn_exchange():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gloo_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


# --------------------------------------------------------------------------------------------------
class ThreadComm(object):
    """Emulated ranks inside one process: rendezvous on the host, reduce with torch ops."""

    def __init__(self, world, shared, rank):
        self.world, self.rank, self.sh = world, rank, shared

    def _all(self, t, op):
        torch.cuda.synchronize()
        self.sh["slots"][self.rank] = t
        self.sh["bar"].wait()
        acc = torch.stack([x.clone() for x in self.sh["slots"]]).sum(0)
        self.sh["bar"].wait()
        t.copy_(acc / self.world if op == "avg" else acc)
        torch.cuda.synchronize()
        self.sh["bar"].wait()

    def all_reduce_sum(self, t):
        self._all(t, "sum")

    def all_reduce_avg(self, t):
        self._all(t, "avg")

    def broadcast(self, t, src=0):
        torch.cuda.synchronize()
        self.sh["slots"][self.rank] = t
        self.sh["bar"].wait()
        v = self.sh["slots"][src].clone()
        self.sh["bar"].wait()
        t.copy_(v)
        torch.cuda.synchronize()
        self.sh["bar"].wait()


class PeerThreadComm(ThreadComm):
    """Same, plus peer-addressable exchange buffers: the BatchNorm kernels of the two emulated ranks then wait for each
    other ON THE DEVICE (they run concurrently on different streams of the one GPU), as they do across GPUs."""

    def make_peer_exchange(self, nbytes, device):
        from scnym_b200.parallel import PeerExchange
        self.sh["peer"][self.rank] = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        torch.cuda.synchronize()
        self.sh["bar"].wait()
        px = PeerExchange.local(list(self.sh["peer"]), self.rank)
        self.sh["bar"].wait()
        return px


@pytest.mark.gpu
@pytest.mark.parametrize("opt_name,lr,peer,fused", [("sgd", 0.05, False, False), ("adadelta", 1.0, False, False), ("sgd", 0.05, True, False),
                                                     ("adadelta", 1.0, True, False), ("sgd", 0.05, True, True), ("adadelta", 1.0, True, True),
                                                     ("sgd", 0.05, True, "tc"), ("adadelta", 1.0, True, "tc")])
def test_data_parallel_step_equals_global_batch_oracle(opt_name, lr, peer, fused, monkeypatch):
    """peer=False: host-side rendezvous for every exchange.  peer=True: the product path of a one-node NCCL job -- the
    BatchNorm statistics and the gradient reduce-scatter + optimizer + parameter all-gather run inside kernels that wait
    for the other rank ON THE DEVICE.  Two such ranks share one GPU here, so each emulated rank drains its stream after
    every call: the ranks then advance kernel by kernel and a waiting kernel never starves the other rank's work of the
    GPU front end (across real GPUs nothing of the sort is needed; bench.py --gpus 2 runs the asynchronous path)."""
    monkeypatch.setenv("SCNB_PEER_SPIN_LIMIT", str(1 << 22))   # a missing peer traps after a few seconds
    if peer:
        import scnym_b200.engine as E
        import scnym_b200._lib as LIB
        raw_call = LIB.call

        def stepwise_call(name, *a):
            rc = raw_call(name, *a)
            torch.cuda.current_stream().synchronize()
            return rc
        monkeypatch.setattr(E, "call", stepwise_call)
    from oracle import scnym_oracle as O
    from tests import cuda_harness as CH
    kw = dict(G=600, L=24, U=32, C=5, H0=64, H=48, n_layers=2, opt_name=opt_name, lr=lr, n_steps=1, D_given=0)
    tc = fused == "tc"   # + tensor-core first layer and dW1: every rank's gradient rows are PUSHED into the owners' inboxes
    fused = bool(fused)
    if fused:   # shapes the fused hidden stack takes (hidden_ops.cu): statistics exchanged by scnb_hs_xchg between its launches
        kw.update(L=32, U=64, H0=128, H=64)
    cases = [CH.oracle_case(seed=31, **kw), CH.oracle_case(seed=32, **kw)]
    for k in list(cases[1].keys()):
        if k.startswith("init/"):
            cases[1][k] = cases[0][k]                      # identical replicas
    l, u = kw["L"], kw["U"]
    # ---- the global batch the two ranks jointly process, fed to the oracle ----
    inp = [GU.step_inputs(c, 0) for c in cases]
    cat = lambda f: torch.cat([f(inp[0]), f(inp[1])], 0)
    Lx, Ly, Ux = cat(lambda i: i["Lx"]), cat(lambda i: i["Ly"]), cat(lambda i: i["Ux"])
    r = [i["rnd"] for i in inp]
    n_loc, n_g = u + l, 2 * (u + l)

    def gidx(rank, i):  # local MixUp-concat index -> global concat index
        return rank * u + i if i < u else 2 * u + rank * l + (i - u)

    ridx_g, gam_g = np.zeros(n_g, dtype=np.int64), np.zeros(n_g, dtype=np.float32)
    for rk in range(2):
        ridx_l, _ = O.mixup_indices(r[rk].perm_keys, r[rk].gamma_raw, n_loc)
        for i in range(n_loc):
            ridx_g[gidx(rk, i)] = gidx(rk, int(ridx_l[i]))
            gam_g[gidx(rk, i)] = float(r[rk].gamma_raw[i])
    keys_g = np.zeros(n_g, dtype=np.float32)
    keys_g[ridx_g] = np.arange(n_g, dtype=np.float32) / n_g   # argsort(keys_g) == ridx_g
    na = len(r[0].drop_keep_U)
    rnd_g = O.StepRandomness(
        in_keep_L=torch.cat([r[0].in_keep_L, r[1].in_keep_L]), perm_keys=torch.from_numpy(keys_g), gamma_raw=torch.from_numpy(gam_g),
        drop_keep_U=[torch.cat([r[0].drop_keep_U[a], r[1].drop_keep_U[a]]) for a in range(na)],
        drop_keep_L=[torch.cat([r[0].drop_keep_L[a], r[1].drop_keep_L[a]]) for a in range(na)],
        drop_keep_D=[torch.cat([r[0].drop_keep_D[a][:l], r[1].drop_keep_D[a][:l], r[0].drop_keep_D[a][l:], r[1].drop_keep_D[a][l:]])
                     for a in range(na)])
    model_st, dan_st = GU.split_state(GU.state_from(cases[0], "init/"))
    om = O.OracleModel(model_st, n_layers=kw["n_layers"], dan_state=dan_st)
    oopt = O.OracleOptimizer(om.parameters(), GU.optim_config(cases[0]))
    out, grads, _ = O.train_step(om, oopt, GU.step_config(cases[0]), Lx, Ly, Ux, rnd_g)
    # ---- two emulated ranks ----
    shared = {"slots": [None, None], "peer": [None, None], "bar": threading.Barrier(2)}
    engines, errors = [None, None], []

    def worker(rk):
        try:
            from scnym_b200.dataprep import DeviceCSR
            from scnym_b200.engine import TrainEngine
            d = cases[rk]
            model, dann = CH.build_models(d)
            ref = CH.build_engine  # same config as the single-rank harness, plus a communicator
            G = int(d["meta_G"])
            lab = DeviceCSR.from_arrays(d["lab_indptr"], d["lab_indices"], d["lab_data"], G)
            unl = DeviceCSR.from_arrays(d["unl_indptr"], d["unl_indices"], d["unl_data"], G)
            from scnym_b200.engine import EngineConfig
            cfg = EngineConfig(optimizer=opt_name, lr=lr, weight_decay=float(d["meta_wd"]), use_cuda_graph=False, keep_w1_grad=True,
                               first_layer="tc" if tc else "auto", w1_update="tc" if tc else "auto")
            with torch.cuda.stream(torch.cuda.Stream()):       # one stream per emulated rank: their kernels overlap
                eng = TrainEngine(model, cfg, lab, d["lab_y"], l, unlabeled=unl, unlabeled_batch_size=u, dann=dann, mode="mixmatch",
                                  comm=(PeerThreadComm if peer else ThreadComm)(2, shared, rk))
                assert (eng.peer is not None) == peer and eng.hs_fused == fused and eng.w1_tc == tc
                eng.set_weights(float(d["meta_unsup_w"]), float(d["meta_dan_w"]))
                CH.inject_step(eng, d, 0)
                eng.step()
                torch.cuda.synchronize()
            engines[rk] = eng
        except Exception as e:  # pragma: no cover
            errors.append(e)
            shared["bar"].abort()

    th = [threading.Thread(target=worker, args=(rk,)) for rk in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
    if peer:   # the fused exchange never materialises the averaged gradient: average the ranks' local gradients here
        for n, p in engines[0].model.named_parameters():
            avg = 0.5 * (p.grad + dict(engines[1].model.named_parameters())[n].grad)
            assert GU.rel_err(avg.cpu(), grads[n]) < 1e-4, ("grad", n)
        for n, p in engines[0].dann.domain_clf.named_parameters():
            avg = 0.5 * (p.grad + dict(engines[1].dann.domain_clf.named_parameters())[n].grad)
            assert GU.rel_err(avg.cpu(), grads["domain_clf." + n]) < 1e-4, ("grad dan", n)
        assert torch.equal(engines[0].arena.flat, engines[1].arena.flat), "replicas differ after the step"
    for rk in range(2):
        eng = engines[rk]
        for n, p in (() if peer else eng.model.named_parameters()):
            assert GU.rel_err(p.grad.cpu(), grads[n]) < 1e-4, (rk, "grad", n)
        for n, p in (() if peer else eng.dann.domain_clf.named_parameters()):
            assert GU.rel_err(p.grad.cpu(), grads["domain_clf." + n]) < 1e-4, (rk, "grad dan", n)
        for n, p in om.named_parameters():
            got = dict(eng.model.named_parameters()).get(n)
            if got is None:
                got = dict(eng.dann.domain_clf.named_parameters())[n[len("domain_clf."):]]
            assert GU.rel_err(got.detach().cpu(), p.detach()) < 1e-4, (rk, "post", n)
        for n, b in om.buffers().items():
            v = eng.model.state_dict()[n].cpu()
            if n.endswith("num_batches_tracked"):
                assert int(v) == int(b)
            else:
                assert GU.rel_err(v, b) < 1e-4, (rk, n)
    # losses: mean over the global batch == average of the ranks' local means
    for key in ("sup_loss", "unsup_loss", "dan_loss"):
        avg = 0.5 * (engines[0].losses()[key] + engines[1].losses()[key])
        assert GU.rel_err(avg, out[key].detach()) < 1e-4, key
