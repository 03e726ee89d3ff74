"""Training fed from HOST-resident expression matrices (the end-to-end / out-of-HBM entry).

The reference assembles every minibatch on the host (`SingleCellDS.__getitem__` + collate,
scnym/dataprep.py:114-168; `DataLoader` iteration scnym/trainer.py:574-582) and copies a dense
[B, G] tensor to the device each step (scnym/trainer.py:589-599).  `HostTrainPipeline` keeps that
contract -- inputs start in host memory every step, the loss comes back to the host every step --
but overlaps the three stages:

  producer threads  batches i+1, i+2, ...: sample rows, gather their CSR rows into a pinned blob (C threads,
                    `scnb_host_gather_rows`), draw the MixUp keys / Beta draws; `n_producers` of them work on
                    consecutive steps in parallel (the gather is a latency-bound random read of host memory and
                    releases the GIL), the consumer takes their batches in step order
  copy stream       H2D of blob i+1 into one of two device landing buffers
  compute stream    device-to-device stash of landing buffer i into the engine's staging CSR,
                    replay of the captured step graph, D2H of the 8 loss scalars

so a step costs max(host gather, PCIe copy, device step) instead of their sum.  The loss of step i
is read on the host after step i+1 has been enqueued.
"""
import queue
import threading
from typing import Optional

import numpy as np
import torch

from . import _lib
from .dataprep import DeviceCSR


class _HostCSR(object):
    def __init__(self, indptr, indices, data, labels=None, domains=None):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float32)
        self.labels = None if labels is None else np.ascontiguousarray(labels, dtype=np.int32)
        self.domains = None if domains is None else np.ascontiguousarray(domains, dtype=np.int32)
        self.n_rows = self.indptr.shape[0] - 1
        self.max_row_nnz = int(np.diff(self.indptr).max()) if self.n_rows > 0 else 1


class _BlobLayout(object):
    """Byte layout of one batch blob: fixed-size header regions, then the four variable CSR regions."""

    def __init__(self, L, U, cap_l, cap_u):
        self.L, self.U, self.cap_l, self.cap_u = L, U, cap_l, cap_u
        off = 0
        self.regions = {}

        def add(name, count, dtype):
            nonlocal off
            nbytes = count * np.dtype(dtype).itemsize
            self.regions[name] = (off, count, dtype)
            off += (nbytes + 15) // 16 * 16

        add("indptr_l", L + 1, np.int64)
        add("indptr_u", U + 1, np.int64)
        add("y_l", L, np.int32)
        add("dom_l", L, np.int32)
        add("dom_u", max(U, 1), np.int32)
        add("keys", L + U, np.float32)
        add("gam", L + U, np.float32)
        self.header_bytes = off
        add("idx_l", cap_l, np.int32)
        add("val_l", cap_l, np.float32)
        add("idx_u", max(cap_u, 1), np.int32)
        add("val_u", max(cap_u, 1), np.float32)
        self.total_bytes = off

    def views(self, blob: torch.Tensor):
        """name -> typed 1-D torch view into a uint8 blob (host or device)."""
        tdt = {np.int64: torch.int64, np.int32: torch.int32, np.float32: torch.float32}
        out = {}
        for name, (o, count, dt) in self.regions.items():
            nbytes = count * np.dtype(dt).itemsize
            out[name] = blob[o:o + nbytes].view(tdt[dt])
        return out


class HostTrainPipeline(object):
    """Drive a `TrainEngine` (mode "mixmatch") from host-resident labeled / unlabeled CSR matrices.

    pipe = HostTrainPipeline(model, cfg, lab=(indptr, indices, data, y[, domain]), unl=(indptr, indices, data[, domain]),
                             n_genes=G, batch_size=256, dann=dann)
    for losses in pipe.run(n_steps): ...        # dict of the step's loss scalars, one per step
    """

    def __init__(self, model, cfg, lab, unl, n_genes: int, batch_size: int, unlabeled_batch_size: Optional[int] = None, dann=None,
                 device="cuda", n_threads: int = 4, seed: int = 0, prefetch: int = 3, comm=None, n_producers: int = 1) -> None:
        from .engine import TrainEngine
        if not torch.cuda.is_available():
            raise _lib.ScnbError("HostTrainPipeline needs a CUDA device (no CPU fallback)")
        self.dev = torch.device(device)
        self.L = int(batch_size)
        self.U = int(unlabeled_batch_size if unlabeled_batch_size is not None else batch_size)
        self.lab = _HostCSR(*lab[:3], labels=lab[3], domains=lab[4] if len(lab) > 4 else None)
        self.unl = _HostCSR(*unl[:3], domains=unl[3] if len(unl) > 3 else None)
        self.n_threads = int(n_threads)
        self.alpha = max(float(cfg.mixup_alpha), 1e-3)
        cap_l, cap_u = self.L * self.lab.max_row_nnz, self.U * self.unl.max_row_nnz
        self.layout = _BlobLayout(self.L, self.U, cap_l, cap_u)
        lay = self.layout
        # pinned blobs (producer <-> copy stream) and two device landing buffers
        self.n_producers = max(int(n_producers), 1)
        self.n_blobs = max(int(prefetch) + 2, 3 * self.n_producers + 1)
        self.blobs = [torch.empty(lay.total_bytes, dtype=torch.uint8).pin_memory() for _ in range(self.n_blobs)]
        self.blob_np = [{k: v.numpy() for k, v in lay.views(b).items()} for b in self.blobs]
        self.blob_t = [lay.views(b) for b in self.blobs]
        self.n_land = 3     # uploads run two steps ahead of the step that consumes them
        self.land = [torch.zeros(lay.total_bytes, dtype=torch.uint8, device=self.dev) for _ in range(self.n_land)]
        self.land_t = [lay.views(b) for b in self.land]
        # the engine's staging batch: one device blob with the same layout (rows 0..B-1 of its CSRs are the batch)
        self.stage = torch.zeros(lay.total_bytes, dtype=torch.uint8, device=self.dev)
        self.s = lay.views(self.stage)
        csr_l = DeviceCSR(self.s["indptr_l"], self.s["idx_l"], self.s["val_l"], int(n_genes), self.lab.max_row_nnz)
        csr_l.n_rows = self.L
        csr_u = DeviceCSR(self.s["indptr_u"], self.s["idx_u"], self.s["val_u"], int(n_genes), self.unl.max_row_nnz)
        csr_u.n_rows = self.U
        given = self.lab.domains is not None and self.unl.domains is not None
        self.engine = TrainEngine(model.to(self.dev), cfg, csr_l, self.s["y_l"], self.L, unlabeled=csr_u, unlabeled_batch_size=self.U,
                                  dann=dann, labeled_domain=self.s["dom_l"] if given else None,
                                  unlabeled_domain=self.s["dom_u"] if given else None, mode="mixmatch", comm=comm)
        self.engine.l_rows.copy_(torch.arange(self.L))
        self.engine.u_rows.copy_(torch.arange(self.U))
        # MixUp randomness is read straight from the staging blob (set before the step graph is captured)
        self.engine.perm_keys, self.engine.gamma_raw = self.s["keys"], self.s["gam"]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.ev_h2d = [torch.cuda.Event() for _ in range(self.n_land)]
        self.ev_stash = [torch.cuda.Event() for _ in range(self.n_land)]
        self.ev_done = [torch.cuda.Event() for _ in range(self.n_land)]
        # loss scalars come back through host memory the device can store into: a device-to-host memcpy is queued by the
        # copy engine behind the upload of a later batch (150 us), a 32-byte store is not
        import ctypes
        hp, dp = ctypes.c_void_p(), ctypes.c_void_p()
        _lib.call("scnb_host_alloc_mapped", 64 * self.n_land, ctypes.byref(hp), ctypes.byref(dp))
        self._loss_host_ptr, self._loss_dev_ptr = hp.value, dp.value
        self.loss_host = np.ctypeslib.as_array(ctypes.cast(hp.value, ctypes.POINTER(ctypes.c_float)), shape=(16 * self.n_land,))
        self._graphs = {}
        self._tab_rng = np.random.default_rng(seed)
        self._tab_lock = threading.Lock()
        self._blocks, self._perm_state, self._step0 = {}, {}, 0
        self.h2d_bytes_last = 0
        self.d2h_bytes_last = 8 * 4       # stored by the publish kernel into mapped host memory

    # ---- producer (runs in a thread; the C gather releases the GIL) ----------------------------
    _TAB = 128   # steps of random draws made at once (amortises numpy call overhead held under the GIL)

    def _perm_rows(self, which, n_rows, per, n_steps):
        """`n_steps` consecutive minibatches of `per` rows from a running shuffled pass over [0, n_rows): every pass is a
        fresh permutation, batches never straddle passes (`DataLoader(shuffle=True, drop_last=True)`, scnym/main.py:317-323;
        the unlabeled loader wraps around through `repeater`, main.py:48-68).  Same semantics as
        `TrainEngine.sample_epoch_tables`."""
        st = self._perm_state.setdefault(which, {"perm": None, "pos": 0})
        out = np.empty((n_steps, per), dtype=np.int64)
        for i in range(n_steps):
            if st["perm"] is None or st["pos"] + per > n_rows:
                st["perm"], st["pos"] = self._tab_rng.permutation(n_rows), 0
                if per > n_rows:      # fewer rows than one batch: sample with replacement (degenerate, tests only)
                    st["perm"] = self._tab_rng.integers(0, n_rows, per)
            out[i] = st["perm"][st["pos"]:st["pos"] + per]
            st["pos"] += per
        return out

    def _block(self, b):
        """Random draws of steps [b*_TAB, (b+1)*_TAB), generated once (by whichever producer needs them first)."""
        with self._tab_lock:
            while b not in self._blocks:
                nb = (max(self._blocks) + 1) if self._blocks else 0
                T, L, U, rng = self._TAB, self.L, self.U, self._tab_rng
                self._blocks[nb] = {"l": self._perm_rows("l", self.lab.n_rows, L, T), "u": self._perm_rows("u", self.unl.n_rows, U, T),
                                    "k": rng.random((T, L + U), dtype=np.float32),
                                    "g": rng.beta(self.alpha, self.alpha, (T, L + U)).astype(np.float32)}
                for old in [k for k in self._blocks if k < nb - 2]:
                    del self._blocks[old]
            return self._blocks[b]

    def _draws(self, step):
        """Sampled rows / MixUp keys / Beta draws of global step `step`."""
        tab = self._block(step // self._TAB)
        i = step % self._TAB
        return tab["l"][i], tab["u"][i], tab["k"][i], tab["g"][i]

    def _produce(self, k: int, step: int):
        h = _lib.lib()
        v, L, U = self.blob_np[k], self.L, self.U
        lrows, urows, keys, gam = self._draws(step)
        a, b = self.lab, self.unl
        rc = h.scnb_host_gather_rows2(a.indptr.ctypes.data, a.indices.ctypes.data, a.data.ctypes.data, lrows.ctypes.data, L,
                                      v["indptr_l"].ctypes.data, v["idx_l"].ctypes.data, v["val_l"].ctypes.data, self.layout.cap_l,
                                      b.indptr.ctypes.data, b.indices.ctypes.data, b.data.ctypes.data, urows.ctypes.data, U,
                                      v["indptr_u"].ctypes.data, v["idx_u"].ctypes.data, v["val_u"].ctypes.data, self.layout.cap_u,
                                      self.n_threads)
        if rc != 0:
            raise _lib.ScnbError(h.scnb_last_error().decode())
        v["y_l"][:] = self.lab.labels[lrows]
        if self.lab.domains is not None:
            v["dom_l"][:] = self.lab.domains[lrows]
        if self.unl.domains is not None:
            v["dom_u"][:U] = self.unl.domains[urows]
        v["keys"][:] = keys
        v["gam"][:] = gam
        return int(v["indptr_l"][L]), int(v["indptr_u"][U])

    def _producer_loop(self, j, n_steps, free_q, ready_q):
        """Producer j assembles steps j, j + P, j + 2P, ... and hands them over through its own queue."""
        try:
            for i in range(j, n_steps, self.n_producers):
                k = free_q.get()
                if k is None:
                    return
                nnz = self._produce(k, self._step0 + i)
                ready_q.put((k, nnz))
        except BaseException as e:  # surface producer failures in the consumer
            ready_q.put(e)

    # ---- consumer ------------------------------------------------------------------------------
    def _enqueue_h2d(self, k, nnz, slot):
        """Copy blob k -> landing buffer `slot` on the copy stream (exact bytes: header + used CSR prefixes)."""
        src, dst, hb = self.blobs[k], self.land[slot], self.layout.header_bytes
        nl, nu = nnz
        exact = hb + 8 * (nl + nu)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_stash[slot])     # landing buffer no longer read by an earlier stash
            if self._whole_blob(exact):
                dst.copy_(src, non_blocking=True)                # one transfer: the unused capacity is small
                exact = self.layout.total_bytes
            else:
                dst[:hb].copy_(src[:hb], non_blocking=True)
                for name, n in (("idx_l", nl), ("val_l", nl), ("idx_u", nu), ("val_u", nu)):
                    if n > 0:
                        self.land_t[slot][name][:n].copy_(self.blob_t[k][name][:n], non_blocking=True)
            self.ev_h2d[slot].record(self.copy_stream)
        return exact

    def _whole_blob(self, exact_bytes):
        """Move the whole blob in one copy when its unused capacity (rows shorter than the longest row) is below 25 %:
        five enqueues and their launch gaps cost more than the extra bytes."""
        return self.layout.total_bytes <= 1.25 * exact_bytes

    def _step_graph(self, slot):
        """[landing buffer `slot` -> staging batch] + the step + [loss scalars -> mapped host memory] as ONE graph launch
        (the whole blob is stashed: a device-to-device copy of its unused capacity costs less than separate launches)."""
        eng = self.engine
        key = (slot, eng.unsup_weight, eng.dan_weight)
        if self._graphs and next(iter(self._graphs))[1:] != key[1:]:
            self._graphs = {}             # the loss weights are launch arguments: new weights retire the old graphs
        g = self._graphs.get(key)
        if g is None:
            dst = self._loss_dev_ptr + 64 * slot
            # the read-back is launched by the engine as soon as the losses are final (mid-step, on an idle branch)
            eng.post_launch = lambda e: _lib.call("scnb_publish_f32", _lib.ptr(e.loss_buf), 8, dst, torch.cuda.current_stream().cuda_stream)
            try:
                g = eng.capture_step(pre=lambda: self.stage.copy_(self.land[slot], non_blocking=True))
            finally:
                eng.post_launch = None
            self._graphs[key] = g
        return g

    def _enqueue_step(self, nnz, slot):
        eng = self.engine
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ev_h2d[slot])
        if eng.cfg.use_cuda_graph:
            g = self._step_graph(slot)
            eng.begin_step()
            g.replay()
            eng.end_step()
        else:
            self.stage.copy_(self.land[slot], non_blocking=True)
            eng.step()
            _lib.call("scnb_publish_f32", _lib.ptr(eng.loss_buf), 8, self._loss_dev_ptr + 64 * slot, cur.cuda_stream)
        self.ev_stash[slot].record(cur)
        self.ev_done[slot].record(cur)

    def _read_loss(self, slot):
        self.ev_done[slot].synchronize()
        v = self.loss_host[16 * slot: 16 * slot + 8].tolist()
        eng = self.engine
        out = {"sup_loss": v[0], "sup_acc": v[1] / max(self.L, 1), "unsup_loss": v[2]}
        if eng.use_dan:
            out["dan_loss"] = v[4]
        out["loss"] = v[0] + eng.unsup_weight * v[2] + (v[4] if eng.use_dan else 0.0)
        return out

    def close(self):
        if getattr(self, "_loss_host_ptr", None):
            torch.cuda.synchronize()
            self.loss_host = None
            _lib.call("scnb_host_free_mapped", self._loss_host_ptr)
            self._loss_host_ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, n_steps: int):
        """Generator over `n_steps` training steps; yields the loss dict of each step (in order)."""
        if self.engine.cfg.use_cuda_graph:
            for slot in range(self.n_land):      # capture before copies are in flight on other streams
                self._step_graph(slot)
        torch.cuda.synchronize()
        P = self.n_producers
        # every producer owns its own pinned blobs (blob k belongs to producer k mod P): a producer that runs ahead can
        # then never starve the one whose batch the consumer is waiting for
        free_qs, ready_qs = [queue.Queue() for _ in range(P)], [queue.Queue() for _ in range(P)]
        for k in range(self.n_blobs):
            free_qs[k % P].put(k)
        ths = [threading.Thread(target=self._producer_loop, args=(j, n_steps, free_qs[j], ready_qs[j]), daemon=True) for j in range(P)]
        for th in ths:
            th.start()
        taken = [0]

        def next_ready():
            item = ready_qs[taken[0] % P].get()      # step order: producer (step mod P) made it
            taken[0] += 1
            if isinstance(item, BaseException):
                raise item
            return item

        try:
            NL = self.n_land
            pending = None           # (slot, blob) of the step whose loss has not been read yet
            ahead = []               # uploaded (or uploading) batches not yet stepped: (blob, nnz), in step order
            fetched = 0
            while fetched < min(2, n_steps):     # uploads run two steps ahead of the compute stream
                item = next_ready()
                b = self._enqueue_h2d(item[0], item[1], fetched % NL)
                if fetched == 0:
                    self.h2d_bytes_last = b
                ahead.append(item)
                fetched += 1
            for i in range(n_steps):
                slot = i % NL
                k, nnz = ahead.pop(0)
                if fetched < n_steps:
                    item = next_ready()
                    self._enqueue_h2d(item[0], item[1], fetched % NL)
                    ahead.append(item)
                    fetched += 1
                self._enqueue_step(nnz, slot)
                if pending is not None:
                    yield self._read_loss(pending[0])
                    free_qs[pending[1] % P].put(pending[1])     # its H2D finished long ago: the pinned blob can be refilled
                pending = (slot, k)
            if pending is not None:
                yield self._read_loss(pending[0])
                free_qs[pending[1] % P].put(pending[1])
        finally:
            for q in free_qs:
                q.put(None)
            for th in ths:
                th.join(timeout=5.0)
            self._step0 += n_steps


class MappedArray(object):
    """A host numpy array, page-locked in place and mapped into the device's address space: `data_ptr()` is what a kernel
    dereferences (reads travel over PCIe when the kernel runs).  Quacks like the tensors the engine borrows pointers from."""
    is_mapped_host = True
    is_cuda = True

    def __init__(self, arr: np.ndarray) -> None:
        import ctypes
        self.arr = np.ascontiguousarray(arr)
        self._host_ptr = None
        if self.arr.size == 0:
            self._dev_ptr = 0
            return
        dp = ctypes.c_void_p()
        _lib.call("scnb_host_register_mapped", self.arr.ctypes.data, int(self.arr.nbytes), ctypes.byref(dp))
        self._host_ptr, self._dev_ptr = self.arr.ctypes.data, dp.value

    def data_ptr(self):
        return self._dev_ptr

    def numel(self):
        return int(self.arr.size)

    def close(self):
        if self._host_ptr is not None:
            _lib.call("scnb_host_unregister", self._host_ptr)
            self._host_ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MappedHostCSR(object):
    """Host-resident expression matrix whose arrays the device reads in place (same attributes as `DeviceCSR`)."""

    def __init__(self, indptr, indices, data, n_cols: int, device) -> None:
        ip = np.ascontiguousarray(indptr, dtype=np.int64)
        self.n_rows = int(ip.shape[0] - 1)
        self.n_cols = int(n_cols)
        self.max_row_nnz = int(np.diff(ip).max()) if self.n_rows > 0 else 0
        self.indptr = MappedArray(ip)
        self.indices = MappedArray(np.ascontiguousarray(indices, dtype=np.int32))
        self.data = MappedArray(np.ascontiguousarray(data, dtype=np.float32))
        self.device = torch.device(device)

    @property
    def shape(self):
        return (self.n_rows, self.n_cols)

    @property
    def nnz(self):
        return self.data.numel()

    def close(self):
        for a in (self.indptr, self.indices, self.data):
            a.close()


class HostMappedTrainPipeline(object):
    """Training straight out of HOST-resident matrices with no host work per step: the labeled / unlabeled CSR arrays (and
    labels, domains) are page-locked in place and mapped into the device; the step's own prologue -- sampler hand-off and
    row gather, `scnb_step_prep` + `scnb_csr_gather_rows2` -- reads the sampled rows over PCIe while the previous step is
    still running (the engine assembles the batch of step t+1 during step t), and the loss scalars are stored into mapped
    host memory by the last launch of the step's graph.  Same contract as `HostTrainPipeline` (inputs in host memory at
    every step, loss on the host after every step; replaces the DataLoader + per-step `.to(device)` of
    scnym/trainer.py:574-599), without producer threads, staging blobs or copy-engine transfers.

    pipe = HostMappedTrainPipeline(model, cfg, lab=(indptr, indices, data, y[, domain]), unl=(indptr, indices, data[, domain]),
                                   n_genes=G, batch_size=256, dann=dann)
    for losses in pipe.run(n_steps): ...
    """

    def __init__(self, model, cfg, lab, unl, n_genes: int, batch_size: int, unlabeled_batch_size: Optional[int] = None, dann=None,
                 device="cuda", seed: int = 0, comm=None) -> None:
        import ctypes
        from .engine import TrainEngine
        if not torch.cuda.is_available():
            raise _lib.ScnbError("HostMappedTrainPipeline needs a CUDA device (no CPU fallback)")
        self.dev = torch.device(device)
        self.L = int(batch_size)
        self.U = int(unlabeled_batch_size if unlabeled_batch_size is not None else batch_size)
        self.lab = MappedHostCSR(lab[0], lab[1], lab[2], n_genes, self.dev)
        self.unl = MappedHostCSR(unl[0], unl[1], unl[2], n_genes, self.dev)
        i32 = lambda x: MappedArray(np.ascontiguousarray(x, dtype=np.int32))
        self.y = i32(lab[3])
        given = len(lab) > 4 and lab[4] is not None and len(unl) > 3 and unl[3] is not None
        self.dom_l = i32(lab[4]) if given else None
        self.dom_u = i32(unl[3]) if given else None
        self.engine = TrainEngine(model.to(self.dev), cfg, self.lab, self.y, self.L, unlabeled=self.unl, unlabeled_batch_size=self.U,
                                  dann=dann, labeled_domain=self.dom_l, unlabeled_domain=self.dom_u, mode="mixmatch", comm=comm)
        hp, dp = ctypes.c_void_p(), ctypes.c_void_p()
        _lib.call("scnb_host_alloc_mapped", 128, ctypes.byref(hp), ctypes.byref(dp))
        self._loss_host_ptr, self._loss_dev_ptr = hp.value, dp.value
        self.loss_host = np.ctypeslib.as_array(ctypes.cast(hp.value, ctypes.POINTER(ctypes.c_float)), shape=(32,))
        # one read-back slot per set of batch buffers (the engine replays one graph per set when it assembles batches ahead)
        self.engine.post_launch = lambda eng: _lib.call(
            "scnb_publish_f32", _lib.ptr(eng.loss_buf), 8, self._loss_dev_ptr + 64 * eng.__dict__["_cur"],
            torch.cuda.current_stream().cuda_stream)
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(int(seed))
        self.epoch_steps = 256
        self.h2d_bytes_last = 0
        self.d2h_bytes_last = 8 * 4

    def _loss(self, slot):
        self.ev_done[slot].synchronize()
        v = self.loss_host[16 * slot: 16 * slot + 8].tolist()
        eng = self.engine
        out = {"sup_loss": v[0], "sup_acc": v[1] / max(self.L, 1), "unsup_loss": v[2]}
        if eng.use_dan:
            out["dan_loss"] = v[4]
        out["loss"] = v[0] + eng.unsup_weight * v[2] + (v[4] if eng.use_dan else 0.0)
        return out

    def batch_bytes(self):
        """Host bytes the step that just ran pulled over PCIe: its rows' index / value entries, their row offsets, labels."""
        eng = self.engine
        nnz = int(eng.b_indptr[-1].item())
        per_row = 16 + 4 + (8 if self.dom_l is not None else 0)
        return 8 * nnz + per_row * (self.L + self.U)

    def run(self, n_steps: int):
        """Generator over `n_steps` training steps; yields the loss dict of each step (in order).  The loss of step i is
        read after step i+1 has been enqueued when the engine alternates between two sets of batch buffers (two read-back
        slots); otherwise after every step."""
        eng = self.engine
        cur = torch.cuda.current_stream()
        pending = None
        done = 0
        while done < n_steps:
            n = min(self.epoch_steps, n_steps - done)
            if eng._tabs is None or eng._tabs["n"] != self.epoch_steps or self._left == 0:
                eng.sample_epoch_tables(self.epoch_steps, generator=self.gen)
                self._left = self.epoch_steps
            n = min(n, self._left)
            for _ in range(n):
                eng.step()
                slot = eng.__dict__["_cur"]
                lag = bool(eng._pf_graphs)
                self.ev_done[slot].record(cur)
                if pending is not None:
                    yield self._loss(pending)
                    pending = None
                if lag:
                    pending = slot
                else:
                    yield self._loss(slot)
            self._left -= n
            done += n
        self.h2d_bytes_last = self.batch_bytes()
        if pending is not None:
            yield self._loss(pending)

    _left = 0

    def close(self):
        torch.cuda.synchronize()
        for a in (self.lab, self.unl, self.y, self.dom_l, self.dom_u):
            if a is not None:
                a.close()
        if self._loss_host_ptr:
            self.loss_host = None
            _lib.call("scnb_host_free_mapped", self._loss_host_ptr)
            self._loss_host_ptr = None
