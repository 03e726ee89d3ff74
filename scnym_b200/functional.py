"""Module-level (autograd) path over the libscnym_b200 kernels.

These `torch.autograd.Function`s let `CellTypeCLF` / `DANN` be composed freely with
reference-style criteria (`loss.backward()`, any `torch.optim` optimizer) exactly like the
reference's nn.Modules.  PyTorch is used only to allocate outputs and to carry the autograd
graph; every arithmetic op is a C-ABI call.  The fused training engine
(`scnym_b200.engine`) bypasses this layer and drives the same kernels directly.
"""
from typing import Optional

import torch

from . import _lib
from ._lib import call, ptr

# Test hook: when non-empty, each training-mode Dropout application pops the next 0/1 keep
# mask from this list instead of drawing one (parity tests inject the reference's masks).
DROPOUT_INJECT = []

_seg_cache = {}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _seg_off(n: int, device) -> torch.Tensor:
    key = (device, n)
    t = _seg_cache.get(key)
    if t is None:
        t = torch.tensor([0, n], dtype=torch.int32, device=device)
        _seg_cache[key] = t
    return t


def sgemm(tA, tB, M, N, K, A, lda, B, ldb, C, ldc, alpha=1.0, beta=0.0, bias=None, relu=False, gate=None):
    call("scnb_sgemm", int(tA), int(tB), M, N, K, ptr(A), lda, ptr(B), ldb, ptr(C), ldc, float(alpha), float(beta),
         ptr(bias), int(relu), ptr(gate), _stream())


def colsum(A, M, N, lda, out, accumulate=False):
    call("scnb_colsum", ptr(A), M, N, lda, ptr(out), int(accumulate), _stream())


def _f32c(t: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(t)
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b (optionally ReLU-fused).  W is [out, in] row-major."""

    @staticmethod
    def forward(ctx, x, W, b, relu):
        x, W = _f32c(x), _f32c(W)
        n, K = x.shape
        N = W.shape[0]
        y = torch.empty((n, N), device=x.device, dtype=torch.float32)
        sgemm(0, 1, n, N, K, x, K, W, K, y, N, bias=b, relu=relu)
        ctx.relu = relu
        ctx.save_for_backward(x, W, y if relu else None)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, W, y = ctx.saved_tensors
        g = _f32c(g)
        n, K = x.shape
        N = W.shape[0]
        if ctx.relu:
            g2 = torch.empty_like(g)
            call("scnb_relu_bwd", ptr(g), ptr(y), ptr(g2), g.numel(), _stream())
            g = g2
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            sgemm(0, 0, n, K, N, g, N, W, K, dx, K)
        if ctx.needs_input_grad[1]:
            dW = torch.empty_like(W)
            sgemm(1, 0, N, K, n, g, N, x, K, dW, K)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty((N,), device=x.device, dtype=torch.float32)
            colsum(g, n, N, N, db)
        return dx, dW, db, None


def linear(x, W, b=None, relu: bool = False):
    return _LinearFn.apply(x, W, b, relu)


class _GeneLinearDenseFn(torch.autograd.Function):
    """First layer on a DENSE batch: y = x W1T + b with W1T [G, H0] gene-major."""

    @staticmethod
    def forward(ctx, x, W, b):
        x = _f32c(x)
        WT = W.t()  # [G, H0]; contiguous by construction (GeneLinear)
        assert WT.is_contiguous()
        n, G = x.shape
        H0 = WT.shape[1]
        y = torch.empty((n, H0), device=x.device, dtype=torch.float32)
        sgemm(0, 0, n, H0, G, x, G, WT, H0, y, H0, bias=b)
        ctx.save_for_backward(x, WT)
        return y

    @staticmethod
    def backward(ctx, g):
        x, WT = ctx.saved_tensors
        g = _f32c(g)
        n, G = x.shape
        H0 = WT.shape[1]
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            sgemm(0, 1, n, G, H0, g, H0, WT, H0, dx, G)
        if ctx.needs_input_grad[1]:
            dWT = torch.empty_like(WT)
            sgemm(1, 0, G, H0, n, x, G, g, H0, dWT, H0)
            dW = dWT.t()
        if ctx.needs_input_grad[2]:
            db = torch.empty((H0,), device=x.device, dtype=torch.float32)
            colsum(g, n, H0, H0, db)
        return dx, dW, db


class _GeneLinearCSRFn(torch.autograd.Function):
    """First layer on a CSR batch (warp-per-row SpMM fwd, transposed scatter bwd)."""

    @staticmethod
    def forward(ctx, batch, W, b):
        WT = W.t()
        assert WT.is_contiguous()
        H0 = WT.shape[1]
        y = torch.empty((batch.n_rows, H0), device=WT.device, dtype=torch.float32)
        call("scnb_csr_linear_fwd", ptr(batch.indptr), ptr(batch.indices), ptr(batch.data), batch.n_rows, ptr(WT), ptr(b),
             H0, ptr(y), _stream())
        ctx.batch = batch
        ctx.shape = WT.shape
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        batch = ctx.batch
        G, H0 = ctx.shape
        dW = db = None
        if ctx.needs_input_grad[1]:
            dWT = torch.zeros((G, H0), device=g.device, dtype=torch.float32)
            call("scnb_csr_linear_bwd_dw", ptr(batch.indptr), ptr(batch.indices), ptr(batch.data), batch.n_rows, ptr(g), H0,
                 ptr(dWT), _stream())
            dW = dWT.t()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty((H0,), device=g.device, dtype=torch.float32)
            colsum(g, batch.n_rows, H0, H0, db)
        return None, dW, db


def gene_linear(x, lin):
    from .dataprep import CSRBatch
    lin.weight_t()  # restores the gene-major layout if something replaced .data
    if isinstance(x, CSRBatch):
        return _GeneLinearCSRFn.apply(x, lin.weight, lin.bias)
    return _GeneLinearDenseFn.apply(x, lin.weight, lin.bias)


class _BnActFn(torch.autograd.Function):
    """BatchNorm1d -> Dropout -> ReLU on one segment (all rows)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, bn, training, p_drop, relu, keep):
        z = _f32c(z)
        n, H = z.shape
        a = torch.empty_like(z)
        seg = _seg_off(n, z.device)
        tracked = bn.running_mean is not None
        batch_stats = training or not tracked      # BatchNorm1d(track_running_stats=False) normalises with batch statistics always
        stats = torch.empty((1, 3, H), device=z.device, dtype=torch.float32) if batch_stats else None
        if batch_stats and n < 2:
            raise ValueError("Expected more than 1 value per channel when training")
        call("scnb_bn_act_fwd", ptr(z), ptr(a), None, H, 1, ptr(seg), n, int(batch_stats), ptr(gamma), ptr(beta),
             ptr(bn.running_mean), ptr(bn.running_var), ptr(stats), ptr(keep), None, 0,
             float(p_drop if training else 0.0), int(relu), None, None, _stream())
        if training and tracked:
            call("scnb_bn_running_update", ptr(bn.running_mean), ptr(bn.running_var), ptr(bn.num_batches_tracked), ptr(stats),
                 None, None, None, ptr(seg), 1, H, float(bn.momentum), None, _stream())
        ctx.training, ctx.p_drop, ctx.relu, ctx.bn = batch_stats, float(p_drop if training else 0.0), relu, bn
        ctx.save_for_backward(z, a, gamma, stats, keep, seg)
        return a

    @staticmethod
    def backward(ctx, g):
        z, a, gamma, stats, keep, seg = ctx.saved_tensors
        g = _f32c(g)
        n, H = z.shape
        dz = torch.empty_like(z)
        if ctx.training:
            dgamma = torch.zeros((H,), device=z.device, dtype=torch.float32)
            dbeta = torch.zeros((H,), device=z.device, dtype=torch.float32)
            call("scnb_bn_act_bwd", ptr(g), ptr(z), ptr(a), H, 1, ptr(seg), n, ptr(gamma), ptr(stats), ptr(keep), None, 0,
                 ctx.p_drop, int(ctx.relu), ptr(dz), ptr(dgamma), ptr(dbeta), None, None, _stream())
        else:
            call("scnb_bn_eval_bwd", ptr(g), ptr(a), int(ctx.relu), n, H, ptr(gamma), ptr(ctx.bn.running_var), ptr(dz), _stream())
            dgamma = dbeta = None
        return dz, dgamma, dbeta, None, None, None, None, None


def bn_act(z, bn, dropout, training: bool, relu: bool = True):
    keep = None
    p = float(dropout.p) if dropout is not None else 0.0
    if training and p > 0.0:
        if DROPOUT_INJECT:
            keep = DROPOUT_INJECT.pop(0).to(device=z.device, dtype=torch.uint8).contiguous()
            if tuple(keep.shape) != tuple(z.shape):
                raise ValueError(f"injected dropout mask {tuple(keep.shape)} != activation {tuple(z.shape)}")
        else:
            keep = (torch.rand(z.shape, device=z.device) >= p).to(torch.uint8)
    return _BnActFn.apply(z, bn.weight, bn.bias, bn, training, p, relu, keep)


def embed_forward(model, x, pre_relu_last: bool = False):
    """`model.embed(x)` through the kernels; `pre_relu_last` returns the last BN output
    (the `X_scnym` embedding of scnym/api.py:700-708)."""
    from .dataprep import CSRBatch
    if not isinstance(x, CSRBatch):
        _lib.require_cuda(x)
    blocks = model.block_modules()
    a = x
    for i, (lin, bn, drop) in enumerate(blocks):
        z = gene_linear(a, lin) if i == 0 else linear(a, lin.weight, lin.bias)
        last = i == len(blocks) - 1
        a = bn_act(z, bn, drop, model.training and bn.training, relu=not (pre_relu_last and last))
    return a
