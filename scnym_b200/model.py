"""`CellTypeCLF`, `GradReverse`, `DANN` -- the reference's model surface (scnym/model.py:85-424)
backed by libscnym_b200 kernels.

The classes keep the reference's constructor arguments, attribute names, `state_dict` keys
(`embed.{0,1,4,5,8,9,...}.*`, `classif.0.*`, `domain_clf.{0,2}.*`) and parameter order, and they
consume the torch RNG in the same order as the reference constructor, so
`torch.manual_seed(s); CellTypeCLF(...)` yields the same initial weights and checkpoints
interchange in both directions.

Differences that matter to a maintainer:
  * `embed[0].weight` has the reference's shape [n_hidden_init, n_genes] but is stored
    gene-major (a transposed view of a contiguous [n_genes, n_hidden_init] buffer), which is the
    layout the CSR first-layer kernels read with coalesced 512-byte rows;
  * `forward` runs hand-written CUDA through the C ABI; there is no CPU or eager fallback
    (calling it on CPU tensors raises).
"""
from typing import Tuple

import torch
import torch.nn as nn

from . import _lib
from . import functional as Fn


class GeneLinear(nn.Linear):
    """nn.Linear(n_genes, n_hidden_init) whose weight storage is gene-major.

    `weight.shape == (out, in)` as usual; `weight.t()` is contiguous [in, out].
    """

    def __init__(self, in_features: int, out_features: int) -> None:
        super().__init__(in_features, out_features)  # same RNG draws as the reference (model.py:211)
        w = self.weight.data
        self.weight = nn.Parameter(w.t().contiguous().t())

    def weight_t(self) -> torch.Tensor:
        """[n_genes, n_hidden_init] contiguous view."""
        wt = self.weight.t()
        if not wt.is_contiguous():
            # e.g. after an external `.data = ...` assignment: restore the layout once.
            self.weight.data = self.weight.data.t().contiguous().t()
            wt = self.weight.t()
        return wt


class CellTypeCLF(nn.Module):
    """Cell type classifier from expression data (scnym/model.py:85-283)."""

    def __init__(
        self,
        n_genes: int,
        n_cell_types: int,
        n_hidden: int = 256,
        n_hidden_init: int = 256,
        n_layers: int = 2,
        init_dropout: float = 0.0,
        residual: bool = False,
        batch_norm: bool = True,
        track_running_stats: bool = True,
        n_decoder_layers: int = 0,
        use_raw_counts: bool = False,
    ) -> None:
        super().__init__()
        if residual or not batch_norm or n_decoder_layers > 1 or use_raw_counts:
            # non-default variants are outside the hot path; no shipped
            # CONFIGS entry reaches them (scnym/api.py:98-103).
            raise NotImplementedError(
                "scnym_b200.CellTypeCLF implements the configuration scnym_api ships "
                "(residual=False, batch_norm=True, n_decoder_layers<=1, use_raw_counts=False)")
        self.n_genes = n_genes
        self.n_cell_types = n_cell_types
        self.n_hidden = n_hidden
        self.n_hidden_init = n_hidden_init
        self.n_decoder_layers = n_decoder_layers
        self.n_layers = n_layers
        self.residual = residual
        self.batch_norm = batch_norm
        self.track_running_stats = track_running_stats
        self.use_raw_counts = use_raw_counts
        self.init_dropout = nn.Dropout(p=init_dropout)

        # RNG-order parity with the reference constructor (model.py:169-233): the shared hidden
        # Linear is created first, then the (unused here) ResBlock's two Linears, then the
        # first two embed Linears, then the classifier.
        # track_running_stats=False (model.py:150-153): batch statistics in training AND evaluation, no running buffers;
        # served by the module-level path (functional.bn_act); the fused engine / batched predict need the running buffers
        BN = lambda w: nn.BatchNorm1d(w, track_running_stats=track_running_stats)
        vanilla_layer = [nn.Linear(n_hidden, n_hidden), BN(n_hidden), nn.Dropout(), nn.ReLU(inplace=True)]
        nn.Linear(n_hidden, n_hidden)  # ResBlock.linear00 draw
        nn.Linear(n_hidden, n_hidden)  # ResBlock.linear01 draw
        hidden_layers = vanilla_layer * (n_layers - 1)  # the SAME module objects repeated (model.py:207)
        self.embed = nn.Sequential(
            GeneLinear(n_genes, n_hidden_init),
            BN(n_hidden_init),
            nn.Dropout(),
            nn.ReLU(inplace=True),
            nn.Linear(n_hidden_init, n_hidden),
            BN(n_hidden),
            nn.Dropout(),
            nn.ReLU(inplace=True),
            *hidden_layers,
        )
        self.classif = nn.Sequential(nn.Linear(n_hidden, n_cell_types))

    # ---- structure helpers used by the kernels / engine --------------------------------
    def block_modules(self):
        """[(linear, bn, dropout)] per block APPLICATION (shared modules repeat)."""
        mods = list(self.embed)
        return [(mods[i], mods[i + 1], mods[i + 2]) for i in range(0, len(mods), 4)]

    def forward(self, x, return_embed: bool = False):
        """x: dense CUDA tensor [B, n_genes] or a `scnym_b200.dataprep.CSRBatch`."""
        if self.init_dropout.p > 0:
            raise NotImplementedError("init_dropout > 0 is outside the scnym_api configurations (api.py:101)")
        x_embed = Fn.embed_forward(self, x)
        lin = self.classif[0]
        pred = Fn.linear(x_embed, lin.weight, lin.bias)
        return (pred, x_embed) if return_embed else pred


class GradReverse(torch.autograd.Function):
    """Identity forward, `-weight * grad` backward (scnym/model.py:286-338)."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.weight = weight
        return x.view_as(x) * 1.0

    @staticmethod
    def backward(ctx, grad_output):
        return (grad_output * -1 * ctx.weight), None


class DANN(nn.Module):
    """Domain adversary sharing `model.embed` (scnym/model.py:341-424)."""

    def __init__(self, model: CellTypeCLF, n_domains: int = 2, weight: float = 1.0, n_layers: int = 1) -> None:
        super().__init__()
        self.model = model
        self.n_domains = n_domains
        self.embed = model.embed
        hidden_layers = [nn.Linear(self.model.n_hidden, self.model.n_hidden), nn.ReLU()] * n_layers
        self.domain_clf = nn.Sequential(*hidden_layers, nn.Linear(self.model.n_hidden, self.n_domains))
        # NOTE: as in the reference, the constructor's `weight` is not stored; `.weight` exists only
        # after `set_rev_grad_weight`.

    def set_rev_grad_weight(self, weight: float) -> None:
        self.weight = weight

    def forward(self, x) -> Tuple[torch.Tensor, torch.Tensor]:
        x_embed = Fn.embed_forward(self.model, x)
        x_rev = GradReverse.apply(x_embed, self.weight)
        h = x_rev
        mods = list(self.domain_clf)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                fuse_relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                h = Fn.linear(h, m.weight, m.bias, relu=fuse_relu)
                i += 2 if fuse_relu else 1
            else:
                raise RuntimeError(f"unexpected module in domain_clf: {type(m)}")
        return h, x_embed
