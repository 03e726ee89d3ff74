"""`Predicter` -- batched inference (reference surface: scnym/predict.py:12-216).

The reference densifies each cell on the host, copies [1024, G] dense batches to the GPU and
runs a second full pass for the embedding (scnym/api.py:692-708).  Here the atlas stays in HBM
as CSR and one pass of kernels per row chunk emits argmax, probabilities/scores and the
`X_scnym` embedding (last BatchNorm output, eval mode, pre-ReLU).
"""
import os
from typing import Optional, Union

# first layer of batched inference: "tc" = densify-then-tcgen05 GEMM (bf16x3), "spmm" = warp-per-row CSR SpMM
FIRST_LAYER = os.environ.get("SCNB_PREDICT_FIRST_LAYER", "tc")

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr
from .dataprep import DeviceCSR, remap_columns
from .model import CellTypeCLF

CHUNK_ROWS = 148 * 2 * 256  # rows per kernel launch (two full waves of 256-row tensor-core tiles on 148 SMs); results do not depend on it


class EvalRunner(object):
    """Eval-mode forward of `CellTypeCLF` over row ranges of a DeviceCSR (no autograd)."""

    def __init__(self, models, chunk_rows: int = CHUNK_ROWS) -> None:
        self.models = list(models)
        m0 = self.models[0]
        self.dev = next(m0.parameters()).device
        if self.dev.type != "cuda":
            raise _lib.ScnbError("scnym_b200 inference needs the models on a CUDA device (no CPU fallback)")
        self.widths = [b[0].out_features for b in m0.block_modules()]
        if any(b[1].running_mean is None for m in self.models for b in m.block_modules()):
            raise NotImplementedError("batched predict needs BatchNorm running statistics (track_running_stats=True)")
        self.C = m0.n_cell_types
        self.chunk = int(chunk_rows)
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.dev)
        self.Z = [f32(self.chunk, w) for w in self.widths]
        self.A = [f32(self.chunk, w) for w in self.widths]
        self.logits = f32(self.chunk, self.C)
        self.seg = torch.zeros(2, dtype=torch.int32, device=self.dev)
        self._seg_n = -1
        H0, G = self.widths[0], m0.n_genes
        self.use_tc = FIRST_LAYER == "tc" and H0 % 32 == 0 and 32 <= H0 <= 256
        self.G = G
        i16 = lambda n: torch.zeros(n, dtype=torch.int16, device=self.dev)
        if self.use_tc:
            n_kt = (G + 63) // 64
            self.planes = [(i16(n_kt * H0 * 64), i16(n_kt * H0 * 64)) for _ in self.models]
        # hidden layers a >= 1 on tcgen05 when their widths allow it: A planes come from the previous BN kernel
        ok = lambda a: (self.widths[a] % 32 == 0 and 32 <= self.widths[a] <= 256 and self.widths[a - 1] % 32 == 0)
        self.tc_hidden = [FIRST_LAYER == "tc" and a >= 1 and ok(a) for a in range(len(self.widths))]
        rows_pad = (self.chunk + 255) // 256 * 256
        self.a_planes = [((i16(rows_pad * self.widths[a - 1]), i16(rows_pad * self.widths[a - 1])) if self.tc_hidden[a] else None)
                         for a in range(len(self.widths))]
        self.w_planes = [[((i16(self.widths[a] * self.widths[a - 1]), i16(self.widths[a] * self.widths[a - 1]))
                           if self.tc_hidden[a] else None) for a in range(len(self.widths))] for _ in self.models]
        H = self.widths[-1]
        self.fuse_head = (H % 128 == 0 and H // 128 in (1, 2, 4, 8) and self.C * H * 4 <= 96 * 1024)

    def refresh_planes(self):
        """bf16 hi/lo planes of every model's Linear weights (call after the weights change)."""
        st = torch.cuda.current_stream().cuda_stream
        for mi, model in enumerate(self.models):
            blocks = model.block_modules()
            if self.use_tc:
                hi, lo = self.planes[mi]
                call("scnb_w1_planes", ptr(blocks[0][0].weight_t()), self.G, self.widths[0], ptr(hi), ptr(lo), st)
            for a in range(1, len(blocks)):
                if self.tc_hidden[a]:
                    hi, lo = self.w_planes[mi][a]
                    call("scnb_linear_planes", ptr(blocks[a][0].weight), self.widths[a], self.widths[a - 1], ptr(hi), ptr(lo), st)

    def _set_seg(self, n):
        if n != self._seg_n:
            self.seg.copy_(torch.tensor([0, n], dtype=torch.int32))
            self._seg_n = n

    def run_chunk(self, ds: DeviceCSR, row0: int, n: int, pred, prob, score, embed):
        """Rows [row0, row0+n) -> writes pred[n], prob[n,C], score[n,C], embed[n,H] (any may be None)."""
        st = torch.cuda.current_stream().cuda_stream
        self._set_seg(n)
        for mi, model in enumerate(self.models):
            blocks = model.block_modules()
            n_blk = len(blocks)
            head_done = False
            for a, (lin, bn, _) in enumerate(blocks):
                w = self.widths[a]
                if (a == 0 and self.use_tc and n_blk > 1 and self.tc_hidden[1] and w % 32 == 0
                        and os.environ.get("SCNB_PREDICT_FUSE_BN", "1") != "0"):
                    # first layer + eval BatchNorm + ReLU in one launch, ending in the next layer's operand planes
                    hi, lo = self.planes[mi]
                    ahi, alo = self.a_planes[1]
                    call("scnb_csr_linear_fwd_tc_bn_planes_i64", ptr(ds.indptr[row0:]), ptr(ds.indices), ptr(ds.data), n, self.G, w,
                         ptr(hi), ptr(lo), ptr(lin.bias), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var), 1,
                         ptr(ahi), ptr(alo), st)
                    continue
                if a == 0 and self.use_tc:
                    hi, lo = self.planes[mi]
                    call("scnb_csr_linear_fwd_tc_i64", ptr(ds.indptr[row0:]), ptr(ds.indices), ptr(ds.data), n, self.G, w,
                         ptr(hi), ptr(lo), ptr(lin.bias), ptr(self.Z[0]), 0, st)
                elif a == 0:
                    wt = lin.weight_t()
                    call("scnb_csr_linear_fwd_i64", ptr(ds.indptr[row0:]), ptr(ds.indices), ptr(ds.data), n, ptr(wt),
                         ptr(lin.bias), w, ptr(self.Z[0]), st)
                elif (self.tc_hidden[a] and a + 1 < n_blk and self.tc_hidden[a + 1] and w % 32 == 0
                      and os.environ.get("SCNB_PREDICT_FUSE_BN", "1") != "0"):
                    # hidden Linear + eval BatchNorm + ReLU in one launch, planes in, planes out
                    (ahi, alo), (whi, wlo), (ohi, olo) = self.a_planes[a], self.w_planes[mi][a], self.a_planes[a + 1]
                    call("scnb_linear_fwd_tc_planes_bn_planes", ptr(ahi), ptr(alo), n, self.widths[a - 1], w, ptr(whi), ptr(wlo),
                         ptr(lin.bias), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var), 1, ptr(ohi), ptr(olo), st)
                    continue
                elif (self.tc_hidden[a] and a == n_blk - 1 and self.C <= 16
                      and os.environ.get("SCNB_PREDICT_FUSE_BN", "1") != "0"):
                    # last hidden Linear + eval BatchNorm (+ embedding) + ReLU + classifier in one launch: planes in, logits out
                    (ahi, alo), (whi, wlo) = self.a_planes[a], self.w_planes[mi][a]
                    clf = model.classif[0]
                    pre = embed if (mi == 0 and embed is not None) else None
                    call("scnb_linear_fwd_tc_planes_bn_head", ptr(ahi), ptr(alo), n, self.widths[a - 1], w, ptr(whi), ptr(wlo),
                         ptr(lin.bias), ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var), 1, ptr(clf.weight),
                         ptr(clf.bias), self.C, ptr(self.logits), 0 if mi == 0 else 1, ptr(pre), st)
                    head_done = True
                    continue
                elif self.tc_hidden[a]:
                    (ahi, alo), (whi, wlo) = self.a_planes[a], self.w_planes[mi][a]
                    call("scnb_linear_fwd_tc_planes", ptr(ahi), ptr(alo), n, self.widths[a - 1], w, ptr(whi), ptr(wlo),
                         ptr(lin.bias), ptr(self.Z[a]), st)
                else:
                    wp = self.widths[a - 1]
                    call("scnb_sgemm", 0, 1, n, w, wp, ptr(self.A[a - 1]), wp, ptr(lin.weight), wp, ptr(self.Z[a]), w, 1.0, 0.0,
                         ptr(lin.bias), 0, None, st)
                last = a == n_blk - 1
                pre = embed if (last and mi == 0 and embed is not None) else None
                nxt_tc = (not last) and self.tc_hidden[a + 1]
                if nxt_tc and w % 32 == 0:
                    # BatchNorm -> ReLU whose output goes straight into the next layer's tensor-core operand planes
                    ahi, alo = self.a_planes[a + 1]
                    call("scnb_bn_eval_planes", ptr(self.Z[a]), n, w, ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean),
                         ptr(bn.running_var), 1, None, ptr(pre), ptr(ahi), ptr(alo), st)
                else:
                    call("scnb_bn_act_fwd", ptr(self.Z[a]), ptr(self.A[a]), ptr(pre), w, 1, ptr(self.seg), n, 0, ptr(bn.weight),
                         ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var), None, None, None, 0, 0.0, 1, None, None, st)
            clf = model.classif[0]
            H = self.widths[-1]
            if head_done:
                pass
            elif self.fuse_head:
                call("scnb_head_logits", ptr(self.A[-1]), ptr(clf.weight), ptr(clf.bias), H, self.C, n, ptr(self.logits),
                     0 if mi == 0 else 1, st)
            else:
                call("scnb_sgemm", 0, 1, n, self.C, H, ptr(self.A[-1]), H, ptr(clf.weight), H, ptr(self.logits), self.C, 1.0,
                     0.0 if mi == 0 else 1.0, ptr(clf.bias), 0, None, st)
        call("scnb_predict_head", ptr(self.logits), len(self.models), n, self.C, ptr(pred), ptr(prob), ptr(score), st)

    def run(self, ds: DeviceCSR, want_prob=True, want_score=False, want_embed=False, row0: int = 0, n_rows: Optional[int] = None,
            col_map: Optional[torch.Tensor] = None):
        """`col_map` (int32 [ds.n_cols] on the device, model column of every column of `ds` or -1): the atlas is in its
        own gene order and every row chunk is re-indexed to the model's genes in HBM right before its first layer
        (scnym/utils.py:155-233 folded into the predict loop: the re-indexed atlas is never materialised)."""
        N = ds.n_rows - row0 if n_rows is None else n_rows
        dev = self.dev
        pred = torch.empty(N, dtype=torch.int64, device=dev)
        prob = torch.empty((N, self.C), dtype=torch.float32, device=dev) if want_prob else None
        score = torch.empty((N, self.C), dtype=torch.float32, device=dev) if want_score else None
        embed = torch.empty((N, self.widths[-1]), dtype=torch.float32, device=dev) if want_embed else None
        self.refresh_planes()
        scratch = None
        if col_map is not None:
            if N > 0:
                starts = torch.arange(row0, row0 + N + self.chunk, self.chunk, device=ds.indptr.device).clamp_(max=row0 + N)
                bounds = ds.indptr[starts]
                cap = max(int((bounds[1:] - bounds[:-1]).max().item()), 1)   # stored entries of the fullest chunk (one read-back)
            else:
                cap = 1
            n_ws = _lib.lib().scnb_csr_remap_workspace_len(self.chunk)
            scratch = (torch.empty(self.chunk + 1, dtype=torch.int64, device=dev), torch.empty(cap, dtype=torch.int32, device=dev),
                       torch.empty(cap, dtype=torch.float32, device=dev), torch.empty(n_ws, dtype=torch.int64, device=dev))
        elif ds.n_cols != self.G:
            raise ValueError("%d genes in X, %d genes in model." % (ds.n_cols, self.G))
        for s in range(0, N, self.chunk):
            n = min(self.chunk, N - s)
            part, r0 = ds, row0 + s
            if col_map is not None:
                part, r0 = remap_columns(ds, col_map, self.G, row0 + s, n, out=scratch), 0
            self.run_chunk(part, r0, n, pred[s:], prob[s:] if prob is not None else None,
                           score[s:] if score is not None else None, embed[s:] if embed is not None else None)
        return pred, prob, score, embed


def _pin(a: np.ndarray):
    """Page-lock a host array IN PLACE for asynchronous copies (cudaHostRegister; nothing is copied).  Returns
    (torch view, unpin callable).  When registration is refused the data go through one pinned copy instead."""
    t = torch.from_numpy(a)
    if t.numel() == 0 or t.is_pinned():
        return t, (lambda: None)
    rt = torch.cuda.cudart()
    try:
        rc = rt.cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
        if int(rc) == 0:
            return t, (lambda: rt.cudaHostUnregister(t.data_ptr()))
    except Exception:
        pass
    return t.pin_memory(), (lambda: None)


class HostAtlasStream(object):
    """Predict over an expression matrix that lives in HOST memory (scipy CSR), end to end:
    host CSR -> device -> predictions / probabilities / embedding back in host arrays.

    The reference's path (scnym/api.py:656-714, scnym/predict.py:160-205) densifies every cell on the host, copies a
    dense [1024, G] batch per step and runs a second full pass for the embedding.  Here the CSR arrays are page-locked
    in place and streamed in row chunks through two device staging slots: the H2D copy of chunk i+1, the kernels of
    chunk i and the D2H copy of chunk i-1's results run concurrently on three streams (both DMA directions busy), so
    the wall time is max(PCIe in, compute, PCIe out) instead of their sum."""

    def __init__(self, runner: EvalRunner, cap_nnz: int, want_prob=True, want_score=False, want_embed=True) -> None:
        self.r = runner
        dev, ch = runner.dev, runner.chunk
        self.cap = int(max(cap_nnz, 1))
        mk = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)
        self.slots = []
        for _ in range(2):
            self.slots.append(dict(
                ip=mk(ch + 1, dt=torch.int64), ix=mk(self.cap, dt=torch.int32), dv=mk(self.cap),
                pred=mk(ch, dt=torch.int64), prob=mk(ch, runner.C) if want_prob else None,
                score=mk(ch, runner.C) if want_score else None, emb=mk(ch, runner.widths[-1]) if want_embed else None,
                ev_in=torch.cuda.Event(), ev_comp=torch.cuda.Event(), ev_out=torch.cuda.Event()))
        self.ip_pin = [torch.empty(ch + 1, dtype=torch.int64).pin_memory() for _ in range(2)]
        self.s_in, self.s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self.h2d_bytes = self.d2h_bytes = 0

    def run(self, indptr: np.ndarray, indices: np.ndarray, data: np.ndarray, out: Optional[dict] = None):
        """indptr int64 [N+1], indices int32 [nnz], data fp32 [nnz] (host, C-contiguous).  Returns a dict of pinned host
        tensors: pred int64 [N], prob / score fp32 [N, C], embed fp32 [N, H] (those requested)."""
        r, ch = self.r, self.r.chunk
        N = int(indptr.shape[0] - 1)
        ix_t, un1 = _pin(indices)
        dv_t, un2 = _pin(data)
        want = self.slots[0]
        if out is None:
            out = dict(pred=torch.empty(N, dtype=torch.int64).pin_memory())
            for k, w in (("prob", r.C), ("score", r.C), ("emb", r.widths[-1])):
                if want[k] is not None:
                    out[k] = torch.empty((N, w), dtype=torch.float32).pin_memory()
        main = torch.cuda.current_stream()
        self.h2d_bytes = self.d2h_bytes = 0
        r.refresh_planes()
        try:
            starts = list(range(0, N, ch))

            def upload(i):
                s = starts[i]
                n = min(ch, N - s)
                a, b = int(indptr[s]), int(indptr[s + n])
                if b - a > self.cap:
                    raise ValueError("HostAtlasStream: a row chunk holds more stored entries than the staging capacity")
                sl, pin = self.slots[i % 2], self.ip_pin[i % 2]
                with torch.cuda.stream(self.s_in):
                    self.s_in.wait_event(sl["ev_comp"])          # the slot's previous chunk has been consumed
                    sl["ev_in"].synchronize() if i >= 2 else None  # its pinned indptr copy has left the host buffer
                    np.subtract(indptr[s:s + n + 1], a, out=pin.numpy()[:n + 1])
                    sl["ip"][:n + 1].copy_(pin[:n + 1], non_blocking=True)
                    sl["ix"][:b - a].copy_(ix_t[a:b], non_blocking=True)
                    sl["dv"][:b - a].copy_(dv_t[a:b], non_blocking=True)
                    sl["ev_in"].record(self.s_in)
                self.h2d_bytes += 8 * (n + 1) + 8 * (b - a)
                return n

            n_next = upload(0) if starts else 0
            for i, s in enumerate(starts):
                n = n_next
                if i + 1 < len(starts):
                    n_next = upload(i + 1)
                sl = self.slots[i % 2]
                main.wait_event(sl["ev_in"])
                main.wait_event(sl["ev_out"])                    # the slot's previous results have left the device
                ds = DeviceCSR(sl["ip"][:n + 1], sl["ix"], sl["dv"], r.G, max_row_nnz=r.G)
                r.run_chunk(ds, 0, n, sl["pred"], sl["prob"], sl["score"], sl["emb"])
                sl["ev_comp"].record(main)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(sl["ev_comp"])
                    out["pred"][s:s + n].copy_(sl["pred"][:n], non_blocking=True)
                    self.d2h_bytes += 8 * n
                    for k in ("prob", "score", "emb"):
                        if sl[k] is not None:
                            out[k][s:s + n].copy_(sl[k][:n], non_blocking=True)
                            self.d2h_bytes += 4 * n * sl[k].shape[1]
                    sl["ev_out"].record(self.s_out)
            self.s_out.synchronize()
            main.synchronize()
        finally:
            torch.cuda.synchronize()
            un1(), un2()
        return out


class Predicter(object):
    """Predict cell types from expression data using `CellTypeCLF` (scnym/predict.py:12-216)."""

    def __init__(self, model_weights: Union[str, list, tuple], n_genes: int = None, n_cell_types: int = None,
                 labels: list = None, **kwargs) -> None:
        self.model_weights = [model_weights] if type(model_weights) == str else model_weights
        self.labels = labels
        if n_cell_types is None:
            print("Assuming `n_cell_types` is the same as in the pretrained model weights.")
            params = torch.load(self.model_weights[0], map_location="cpu")
            fkey = list(params.keys())[-1]
            self.n_cell_types = len(params[fkey])
        else:
            self.n_cell_types = n_cell_types
        for weights in self.model_weights:
            if not os.path.exists(weights):
                raise FileNotFoundError()
        if n_genes is None:
            print("Assuming `n_genes` is the same as in the pretrained model weights.")
            params = torch.load(self.model_weights[0], map_location="cpu")
            fkey = list(params.keys())[0]
            self.n_genes = params[fkey].shape[1]
        else:
            self.n_genes = n_genes
        if not torch.cuda.is_available():
            raise _lib.ScnbError("scnym_b200.Predicter needs a CUDA device (no CPU fallback)")
        self.models = []
        for weights in self.model_weights:
            model = CellTypeCLF(n_genes=self.n_genes, n_cell_types=self.n_cell_types, **kwargs)
            model.load_state_dict(torch.load(weights, map_location="cpu"))
            self.models.append(model.cuda().eval())
        self._runner = None

    def predict(self, X, output: str = None, batch_size: int = 1024, **kwargs):
        """Returns (predictions int[N], names list|None[, prob-or-score float32 [N, C]])."""
        if not X.shape[1] == self.n_genes:
            gs = (X.shape[1], self.n_genes)
            raise ValueError("%d genes in X, %d genes in model." % gs)
        return_prob = kwargs.get("return_prob", None)
        if output not in ["prob", "score"] and output is not None:
            raise ValueError(f"{output} is not a valid additional output.")
        from scipy import sparse
        if sparse.isspmatrix_csr(X) and X.shape[0] > 0:
            # host-resident CSR: streamed through the device in row chunks, results land in host arrays
            res = self.predict_host(X, want_prob=True, want_score=(output == "score"), want_embed=False)
            predictions = res["pred"].numpy()
            prob_h = res["prob"].numpy()
            score_h = res["score"].numpy() if "score" in res else None
        else:
            pred, prob, score, _ = self.predict_device(X, want_prob=True, want_score=(output == "score"), want_embed=False)
            predictions = pred.cpu().numpy()
            prob_h = prob.cpu().numpy()
            score_h = score.cpu().numpy() if score is not None else None
        names = [self.labels[p] for p in predictions] if self.labels is not None else None
        if return_prob is True:
            return predictions, names, prob_h
        elif output is not None:
            if output == "prob":
                return predictions, names, prob_h
            return predictions, names, score_h
        return predictions, names

    def predict_host(self, X, want_prob=True, want_score=False, want_embed=True, chunk_rows: Optional[int] = None):
        """scipy CSR in host memory -> dict of host (pinned) tensors `pred`, `prob`, `score`, `emb` (see HostAtlasStream)."""
        from scipy import sparse
        if not sparse.isspmatrix_csr(X):
            X = sparse.csr_matrix(X)
        if not X.has_sorted_indices:
            X = X.sorted_indices()
        if X.shape[1] != self.n_genes:
            raise ValueError("%d genes in X, %d genes in model." % (X.shape[1], self.n_genes))
        ip = np.ascontiguousarray(X.indptr, dtype=np.int64)
        ix = np.ascontiguousarray(X.indices, dtype=np.int32)
        dv = np.ascontiguousarray(X.data, dtype=np.float32)
        N = X.shape[0]
        ch = int(chunk_rows or min(CHUNK_ROWS, max(N, 1)))
        if self._runner is None or self._runner.chunk != ch:
            self._runner = EvalRunner(self.models, chunk_rows=ch)
            self._stream = None
        cap = int(max(ip[min(s + ch, N)] - ip[s] for s in range(0, N, ch)))
        key = (want_prob, want_score, want_embed)
        if getattr(self, "_stream", None) is None or self._stream.cap < cap or self._stream_key != key:
            self._stream = HostAtlasStream(self._runner, cap, want_prob, want_score, want_embed)
            self._stream_key = key
        return self._stream.run(ip, ix, dv)

    def predict_device(self, X, want_prob=True, want_score=False, want_embed=False, col_map=None):
        """Device-resident results (torch CUDA tensors); X may be ndarray, scipy CSR, dense torch or DeviceCSR.
        `col_map` (see `utils.gene_column_map`): X is in its own gene order and is re-indexed chunk by chunk on the device."""
        if isinstance(X, torch.Tensor):
            X = X.detach().cpu().numpy()
        dev = next(self.models[0].parameters()).device
        ds = DeviceCSR.from_any(X, device=dev)
        if col_map is not None:
            col_map = torch.as_tensor(np.asarray(col_map, dtype=np.int32) if not isinstance(col_map, torch.Tensor) else col_map)
            col_map = col_map.to(device=dev, dtype=torch.int32)
        if self._runner is None:
            self._runner = EvalRunner(self.models, chunk_rows=min(CHUNK_ROWS, max(ds.n_rows, 1)))
        elif self._runner.chunk < min(CHUNK_ROWS, ds.n_rows):
            self._runner = EvalRunner(self.models, chunk_rows=min(CHUNK_ROWS, ds.n_rows))
        return self._runner.run(ds, want_prob, want_score, want_embed, col_map=col_map)
