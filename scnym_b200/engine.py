"""Fused MixMatch(+DAN) / supervised training engine.

One `step()` is the body of the reference's minibatch loop
(scnym/trainer.py:575-671 `SemiSupervisedTrainer.train_epoch`, :189-256 `Trainer.train_epoch`):
batch assembly, `log1p_drop`, teacher pseudolabels, MixUp, the three student forwards, the
three losses, backward and `optimizer.step()` -- as a fixed sequence of libscnym_b200 kernels
on device-resident CSR data, with no host synchronisation, replayable as one CUDA graph.

Algebra exploited:
  * the first Linear is applied ONCE per step to the rows [U_raw ; L_aug]; MixUp is applied to
    its output (linearity), and the same product feeds the eval-mode teacher (when the teacher
    is the current student and pseudolabels are not augmented) and the DAN pass;
  * the three student forwards run as one "super-batch" [U-mixed | L-mixed | DAN] whose
    BatchNorm layers keep per-segment statistics; Linear layers and all weight gradients see
    one tall matrix;
  * dW1 is one transposed sparse product from the un-mixed, summed dZ1.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import call, ptr, OPTIMIZER_CODES
from .dataprep import DeviceCSR
from .model import CellTypeCLF, DANN

# Philox stream ids (one logical random stream per use inside a step)
SID_IN_L = 0          # input dropout, labelled rows
SID_IN_U = 1          # + k : input dropout of teacher view k
SID_HID = 16          # + application index : hidden dropout of the student super-batch


def _release_gemm_ws(handles, bound=()):
    try:
        for h in handles:
            call("scnb_gemm_tc_workspace", h, None, 0)
        for (p, r, w) in bound:
            call("scnb_gemm_tc_bind_planes", p, r, w, None, None)
    except Exception:
        pass


@dataclass
class EngineConfig:
    # MixMatchLoss (losses.py:603-610) / fit_model defaults (main.py:511-524, api.py:71-105)
    n_augmentations: int = 1
    T: float = 0.5
    augment_pseudolabels: bool = False
    pseudolabel_min_confidence: float = 0.0
    mixup_alpha: float = 0.3
    mean_teacher: bool = False
    decay_coef: float = 0.997
    p_in_drop: float = 0.1
    augment: str = "log1p_drop"      # AUGMENTATION_SCHEMES key (dataprep.py:730-754): log1p_drop | log1p_mask | log1p_poisson | log1p_poisson_drop
    p_mask_apply: float = 0.5        # GeneMasking p_apply (log1p_mask)
    # DANLoss (losses.py:1043-1105)
    use_dan: bool = True
    dan_use_conf_pseudolabels: bool = False
    dan_scale_loss_pseudoconf: bool = False
    # optimizer (main.py:36-41, 374-396)
    optimizer: str = "adadelta"
    lr: float = 1.0
    weight_decay: float = 1e-4
    betas: tuple = (0.9, 0.999)
    eps: Optional[float] = None
    rho: float = 0.9
    momentum: float = 0.9
    class_weight: Optional[List[float]] = None
    seed: int = 0
    use_cuda_graph: bool = True
    first_layer: str = "auto"   # "spmm": warp-per-row CSR SpMM; "tc": densify-then-tcgen05 GEMM (bf16x3); "auto": tc from 512 rows per step
    w1_update: str = "auto"     # first-layer gradient+optimizer kernel: "simt" (batch CSC), "tc" (tcgen05 GEMM), "auto": tc from 512 rows
    hidden: str = "auto"        # hidden stack: "fused" (BatchNorm folded into the tile GEMMs, hidden_ops.cu), "unfused", "auto"
    fuse_heads: bool = True      # one-warp-per-row classifier/DAN-tail/teacher-head kernels (Linear + loss + backward)
    multi_stream: bool = True    # run independent kernels of the step on side streams (parallel graph branches)
    keep_w1_grad: bool = False   # also write the first-layer gradient to HBM (parity tests; implied by data parallelism)


class ParamArena(object):
    """All trainable parameters in one flat fp32 buffer (+ grads and optimizer state).

    Parameters keep their nn.Module identity: each `p.data` becomes a view into the arena and
    `p.grad` a view into the gradient arena, so `state_dict()` / `load_state_dict()` /
    `torch.save` behave as before while the optimizer is one elementwise kernel.
    """

    @staticmethod
    def padded_total(named_params: List) -> int:
        seen, total = set(), 0
        for _, p in named_params:
            if id(p) not in seen:
                seen.add(id(p))
                total += (p.numel() + 3) // 4 * 4
        return total

    def __init__(self, named_params: List, device, flat=None, grad=None) -> None:
        """flat / grad: optional externally allocated fp32 buffers of `padded_total` elements (data parallelism keeps
        the parameter and gradient arenas in peer-addressable memory)."""
        self.names, self.params, self.offsets, self.sizes = [], [], {}, {}
        total = 0
        seen = set()
        for name, p in named_params:
            if id(p) in seen:
                continue
            seen.add(id(p))
            n = p.numel()
            self.names.append(name)
            self.params.append(p)
            self.offsets[name] = total
            self.sizes[name] = n
            total += (n + 3) // 4 * 4  # keep every tensor 16-byte aligned
        self.total = total
        if flat is not None and (flat.numel() != total or grad is None or grad.numel() != total):
            raise ValueError("external arena buffers must hold padded_total() floats")
        self.flat = torch.zeros(total, dtype=torch.float32, device=device) if flat is None else flat.zero_()
        self.grad = torch.zeros(total, dtype=torch.float32, device=device) if grad is None else grad.zero_()
        self.state1 = torch.zeros(total, dtype=torch.float32, device=device)
        self.state2 = torch.zeros(total, dtype=torch.float32, device=device)
        self._views, self._gviews = {}, {}
        for name, p in zip(self.names, self.params):
            o, n = self.offsets[name], self.sizes[name]
            if p.dim() == 2 and not p.data.is_contiguous() and p.data.t().is_contiguous():
                shp = p.data.t().shape  # gene-major first layer: arena holds [G, H0]
                v = self.flat[o:o + n].view(shp)
                v.copy_(p.data.t())
                gv = self.grad[o:o + n].view(shp)
                p.data = v.t()
                p.grad = gv.t()
            else:
                v = self.flat[o:o + n].view(p.shape)
                v.copy_(p.data)
                gv = self.grad[o:o + n].view(p.shape)
                p.data = v
                p.grad = gv
            self._views[name], self._gviews[name] = v, gv

    def w(self, name):  # storage-layout view of a parameter
        return self._views[name]

    def g(self, name):
        return self._gviews[name]

    def check_aliasing(self):
        """Raise if something re-pointed a parameter away from the arena (e.g. `.cuda()`)."""
        for name, p in zip(self.names, self.params):
            base = self._views[name]
            if p.data.untyped_storage().data_ptr() != base.untyped_storage().data_ptr():
                raise RuntimeError(f"parameter {name} no longer aliases the engine arena; rebuild the engine")


# Everything the step's prologue (sampler hand-off, row gather, MixUp plan, gene-block splits) writes and the rest of the step
# only reads.  Held twice: while step t runs on one set, the prologue of step t+1 fills the other (see TrainEngine.step).
_BATCH_FIELDS = ("b_indptr", "b_idx", "b_val", "l_rows", "u_rows", "perm_keys", "gamma_raw", "lab_ids", "base_dom", "counts",
                 "seg_off", "plan_work", "srcA", "srcB", "gam", "inv3", "coef3", "pconf", "w1_splits", "fwd_splits")


def _batch_field(name):
    def fget(self):
        return self.__dict__["_sets"][self.__dict__["_cur"]][name]

    def fset(self, v):
        self.__dict__["_sets"][self.__dict__["_cur"]][name] = v
    return property(fget, fset)


class TrainEngine(object):
    """Device-resident training step for `CellTypeCLF` (+ `DANN` head).

    mode "mixmatch": SemiSupervisedTrainer step (MixMatchLoss + optional DANLoss)
    mode "supervised": Trainer step (cross_entropy on raw batches)
    """

    def __init__(self, model: CellTypeCLF, cfg: EngineConfig, labeled: DeviceCSR, labeled_y, batch_size: int,
                 unlabeled: Optional[DeviceCSR] = None, unlabeled_batch_size: Optional[int] = None,
                 dann: Optional[DANN] = None, labeled_domain=None, unlabeled_domain=None, mode: str = "mixmatch",
                 extra_named_params: Optional[List] = None, comm=None) -> None:
        if not torch.cuda.is_available():
            raise _lib.ScnbError("TrainEngine needs a CUDA device (no CPU fallback)")
        _lib.lib()
        self.__dict__["_sets"], self.__dict__["_cur"] = [{}, {}], 0
        self.post_launch = None
        self.model, self.cfg, self.mode = model, cfg, mode
        self.comm = comm if (comm is not None and comm.world > 1) else None
        self.dev = labeled.device
        dev = self.dev
        self.lab, self.unl = labeled, unlabeled
        self.L = int(batch_size)
        self.U = int(unlabeled_batch_size if unlabeled_batch_size is not None else batch_size) if mode == "mixmatch" else 0
        if mode == "mixmatch" and unlabeled is None:
            raise ValueError("mixmatch mode needs unlabeled data")
        self.use_dan = bool(cfg.use_dan and dann is not None and mode == "mixmatch")
        self.dann = dann if self.use_dan else None
        self.G, self.C = model.n_genes, model.n_cell_types
        blocks = model.block_modules()
        if any(b[1].running_mean is None for b in blocks):
            raise NotImplementedError("TrainEngine needs BatchNorm running statistics (track_running_stats=True, the shipped default)")
        self.n_app = len(blocks)
        self.widths = [blk[0].out_features for blk in blocks]
        self.H = self.widths[-1]
        self.D = int(dann.n_domains) if self.use_dan else 0
        L, U = self.L, self.U
        self.nb = U + L
        self.Rd = (U + L) if self.use_dan else 0
        self.Rmax = self.nb + self.Rd
        self.K_eff = cfg.n_augmentations if cfg.augment_pseudolabels else 1
        self.teacher_own_spmm = bool(cfg.mean_teacher or cfg.augment_pseudolabels)

        # ---- parameters -> arena (optimizer order: model.parameters(), then the DAN head group,
        #      main.py:374-381, 582-587) ----
        if next(model.parameters()).device != dev:
            model.to(dev)
        named = [("model." + n, p) for n, p in model.named_parameters()]
        if self.use_dan:
            self.dann.domain_clf.to(dev)
            named += [("dan." + n, p) for n, p in self.dann.domain_clf.named_parameters()]
        if extra_named_params:
            named += list(extra_named_params)
        # Data parallelism over peer memory: one exchange buffer per rank holds
        #   [BatchNorm packets | flag block | gradient arena | parameter arena]
        # so that the BatchNorm kernels and the fused gradient-reduce + optimizer kernel of every rank can address them.
        self.peer = None
        flat = grad = None
        mk = getattr(self.comm, "make_peer_exchange", None) if self.comm is not None else None
        if mk is not None and all(w % 8 == 0 for w in self.widths):
            n_seg = (3 if self.use_dan else 2) if mode == "mixmatch" else 1
            self.peer_slots = 2 * len(self.widths)
            bn_bytes = max(_lib.lib().scnb_bn_peer_bytes(w, n_seg, self.peer_slots, self.comm.world) for w in self.widths)
            bn_bytes = max(bn_bytes, _lib.lib().scnb_hs_peer_bytes(max(self.widths), n_seg, self.peer_slots, self.comm.world))
            total = ParamArena.padded_total(named)
            rnd = lambda b: (b + 255) // 256 * 256
            self.peer_bn_bytes = rnd(bn_bytes)
            self.peer_off_flags = self.peer_bn_bytes
            self.peer_off_grad = self.peer_off_flags + 1024
            self.peer_off_flat = self.peer_off_grad + rnd(4 * total)
            # inbox: `world` regions of one ownership slice each -- the dW1 kernel of every rank stores its gradient rows
            # straight into the owner's inbox (scnb_csr_w1_bwd_push_tc)
            import ctypes as _ct
            sl = _ct.c_longlong()
            call("scnb_dp_slice_floats", int(total), int(self.comm.world), _ct.byref(sl))
            self.peer_off_inbox = self.peer_off_flat + rnd(4 * total)
            self.peer = mk(self.peer_off_inbox + rnd(4 * sl.value * self.comm.world), dev)
            if self.peer is not None:
                grad = self.peer.tensor(self.peer_off_grad, total)
                flat = self.peer.tensor(self.peer_off_flat, total)
        self.arena = ParamArena(named, dev, flat=flat, grad=grad)
        self.n_model_params = sum(((p.numel() + 3) // 4 * 4) for n, p in zip(self.arena.names, self.arena.params)
                                  if n.startswith("model."))
        # block application -> parameter names
        mods = list(model.embed)
        self._blk = []
        for a in range(self.n_app):
            lin, bn = mods[4 * a], mods[4 * a + 1]
            li = [i for i, m in enumerate(mods) if m is lin][0]
            bi = [i for i, m in enumerate(mods) if m is bn][0]
            self._blk.append(dict(lin=lin, bn=bn, drop=mods[4 * a + 2], W=f"model.embed.{li}.weight", b=f"model.embed.{li}.bias",
                                  g=f"model.embed.{bi}.weight", beta=f"model.embed.{bi}.bias", bn_key=bi))
        self._dan_lin = [m for m in self.dann.domain_clf if isinstance(m, nn.Linear)] if self.use_dan else []
        if self.use_dan and len(self._dan_lin) != 2:
            raise NotImplementedError("engine supports DANN(n_layers=1) (the only configuration fit_model builds)")

        # ---- labels / domains ----
        def ids(x):
            # device int32 tensors are borrowed as they are (host-batch staging buffers are rewritten every step)
            if x is None:
                return None
            if getattr(x, "is_mapped_host", False):      # page-locked host array mapped into the device (hostfeed.MappedArray)
                return x
            if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.int32:
                return x
            if isinstance(x, torch.Tensor):
                return x.to(device=dev, dtype=torch.int32)
            return torch.as_tensor(np.asarray(x), dtype=torch.int32).to(dev)
        self.lab_y, self.lab_dom, self.unl_dom = ids(labeled_y), ids(labeled_domain), ids(unlabeled_domain)
        self.class_weight = None if cfg.class_weight is None else torch.tensor(cfg.class_weight, dtype=torch.float32, device=dev)

        f32 = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        i32 = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
        # ---- per-step inputs (host writes, graph reads) ----
        self.l_rows = torch.zeros(L, dtype=torch.int64, device=dev)
        self.u_rows = torch.zeros(max(U, 1), dtype=torch.int64, device=dev)
        self.perm_keys = f32(max(self.nb, 1))
        self.gamma_raw = f32(max(self.nb, 1)) + 1.0
        self.hp = f32(8)
        self.state = torch.zeros(4, dtype=torch.int64, device=dev)  # t, seed, rng counter, mm step
        self.state[1] = int(cfg.seed)
        self.set_hyper(lr=cfg.lr)
        self.unsup_weight, self.dan_weight = 1.0, 1.0
        # ---- batch CSR ----
        maxnnz = max(labeled.max_row_nnz, unlabeled.max_row_nnz if unlabeled is not None else 0, 1)
        cap = self.nb * maxnnz
        self.cap = cap
        self.b_indptr, self.b_idx, self.b_val = i32(self.nb + 1), i32(cap), f32(cap)
        # matrices that stay in (page-locked, mapped) HOST memory: the sampled rows are first fetched verbatim over PCIe into
        # a device staging CSR by a light persistent kernel, and the augmenting gather then reads the staged rows
        self.host_mapped = bool(getattr(labeled.indices, "is_mapped_host", False))
        if self.host_mapped:
            self.st_indptr = torch.zeros(self.nb + 1, dtype=torch.int64, device=dev)
            self.st_idx, self.st_val = i32(cap), f32(cap)
        H0 = self.widths[0]
        self.Z1 = f32(self.nb, H0)
        # gene-major copy of the batch for the fused dW1 + optimizer kernel (needs W1T first in the arena)
        self.P1 = self.G * H0
        self.fused_w1 = (H0 % 64 == 0) and self.arena.names[0] == self._blk[0]["W"] and self.arena.offsets[self._blk[0]["W"]] == 0
        wu = cfg.w1_update
        if wu == "auto":
            wu = os.environ.get("SCNB_W1_UPDATE", "tc" if self.nb >= 512 else "simt")
        self.w1_tc = bool(self.fused_w1 and wu == "tc" and H0 % 32 == 0 and 32 <= H0 <= 256)
        self.fwd_splits = None
        if self.w1_tc:
            # tensor-core dW1 + optimizer: per-row split points of the gene blocks, bf16 planes of dZ1
            gt = _lib.lib().scnb_w1tc_gene_block(self.G)
            n_blk = (self.G + gt - 1) // gt
            self.w1_splits = i32((n_blk + 1) * self.nb)
            # where every gene split of the tensor-core forward starts in every batch row (same launch as w1_splits; used by
            # steps whose batch was assembled one step ahead, so the forward's CTAs do not search for their first entry)
            ks, gp = ctypes.c_int(0), ctypes.c_int(0)
            call("scnb_csr_linear_fwd_tc_geometry", self.nb, self.G, H0, ctypes.addressof(ks), ctypes.addressof(gp))
            self.fwd_splits = i32((ks.value + 1) * self.nb)
            n_pad = (self.nb + 31) // 32 * 32
            self.dz_hi = torch.zeros(n_pad * H0, dtype=torch.int16, device=dev)
            self.dz_lo = torch.zeros(n_pad * H0, dtype=torch.int16, device=dev)
        elif self.fused_w1:
            self.gene_cnt, self.c_ptr, self.c_cursor = i32(self.G), i32(self.G + 1), i32(self.G)
            self.c_ent = i32(cap, 2)
        # injected randomness (tests); None -> Philox in-kernel
        from .dataprep import AUG_CODES
        if cfg.augment not in AUG_CODES or cfg.augment == "count_poisson":
            raise ValueError(f"augmentation scheme {cfg.augment!r} is not a log1p(CPM) scheme the MixMatch step can use")
        self.aug_code = AUG_CODES[cfg.augment]
        self.inj_pois_L = None        # fp32 [cap] injected Poisson samples at batch nnz positions (tests)
        self.inj_rowu_L = None        # fp32 [L] injected per-row depth uniforms (tests)
        self.inj_pois_U = None        # list of K fp32 [U*maxnnz] (augmented teacher views)
        self.inj_rowu_U = None
        self.inj_in_keep_L = None     # uint8 [cap] aligned with batch nnz positions
        self.inj_in_keep_U = None     # list of K uint8 [U*maxnnz]
        self.inj_hidden_keep = None   # list over applications of uint8 [Rmax, H_a]
        if mode == "mixmatch":
            K = self.K_eff
            if self.teacher_own_spmm:
                self.tb_indptr, self.tb_idx, self.tb_val = i32(U + 1), i32(U * maxnnz), f32(U * maxnnz)
                self.TZ1 = f32(K * U, H0)
            self.TA = [f32(K * U, w) for w in self.widths]
            self.TZ = [None] + [f32(K * U, w) for w in self.widths[1:]]
            self.t_logits = f32(K * U, self.C)
            self.Ycat = f32(self.nb, self.C)
            self.conf = f32(U)
            self.confident = torch.zeros(U, dtype=torch.uint8, device=dev)
            self._all_confident = torch.ones(max(U, 1), dtype=torch.uint8, device=dev)
            self.pl_scratch = f32(U, self.C)
            self.lab_ids = i32(L)
            self.base_dom = i32(self.nb)
            self.counts, self.seg_off = i32(8), i32(4)
            self.plan_work = i32(U + 3 * self.nb)
            self.srcA, self.srcB, self.gam = i32(self.Rmax), i32(self.Rmax), f32(self.Rmax)
            self.inv3, self.coef3 = i32(3, self.nb), f32(3, self.nb)
            self.pconf = f32(1)
            self.n_seg = 3 if self.use_dan else 2
            if cfg.mean_teacher:
                self.teacher_flat = self.arena.flat[: self.n_model_params].clone()
                self.teacher_bn = None  # snapshot of BN buffers taken at the first step (losses.py:343-348)
        else:
            self.Rmax = L
            self.n_seg = 1
            self.seg_off = torch.tensor([0, L], dtype=torch.int32, device=dev)
            self.Ycat = f32(L, self.C)
            self.lab_ids = i32(L)
        R = self.Rmax
        self.Zs = [f32(R, w) for w in self.widths]
        self.As = [f32(R, w) for w in self.widths]
        self.stats = [f32(3, 3, w) for w in self.widths]
        if self.comm is not None:
            # cross-rank exchange buffers per application: [n_seg][2][w] sums + [n_seg] row counts
            self.dp_fwd = [f32(self.n_seg * 2 * w + self.n_seg) for w in self.widths]
            self.dp_bwd = [f32(self.n_seg * 2 * w + self.n_seg) for w in self.widths]
            self.comm.broadcast(self.arena.flat)  # identical replicas to start from
            for b in self._blk:
                self.comm.broadcast(b["bn"].running_mean)
                self.comm.broadcast(b["bn"].running_var)
        self.logits = f32(self.nb if mode == "mixmatch" else L, self.C)
        self.dlogits = torch.zeros_like(self.logits)
        wmax = max(self.widths)
        self.dA = f32(R, wmax)
        self.dZs = [f32(R, w) for w in self.widths]   # one per application: weight-gradient GEMMs overlap the dA chain
        # fused hidden stack (hidden_ops.cu): BatchNorm -> Dropout -> ReLU folded into 32 x 64 tile GEMMs, statistics carried
        # between launches as per-tile partials.  Needs static segment boundaries on 32-row tiles (no pseudolabel
        # thresholding), 64-column tiles and operand panels that fit shared memory; single process for now.
        seg_host = [0, U, U + L, U + L + self.Rd][: self.n_seg + 1] if mode == "mixmatch" else [0, L]
        hs_ok = ((self.comm is None or self.peer is not None) and all(w % 64 == 0 for w in self.widths) and all(w <= 256 for w in self.widths[1:])
                 and self.widths[0] <= 512 and all(o % 32 == 0 for o in seg_host) and R <= 4096
                 and (mode != "mixmatch" or cfg.pseudolabel_min_confidence <= 0.0))
        hm = cfg.hidden if cfg.hidden != "auto" else os.environ.get("SCNB_HIDDEN", "fused")
        if hm == "fused" and not hs_ok and cfg.hidden == "fused":
            raise ValueError("hidden='fused' needs widths % 64 == 0 (<= 256 after the first), segment sizes % 32 == 0, "
                             "no pseudolabel thresholding and a single process")
        self.hs_fused = bool(hm == "fused" and hs_ok)
        if self.hs_fused:
            self._seg_host = np.asarray(seg_host + [seg_host[-1]] * (4 - len(seg_host)), dtype=np.int32)
            ku = self.K_eff * U if mode == "mixmatch" else 0
            self._seg_teacher = np.asarray([0, ku, ku, ku], dtype=np.int32)
            self.dYs = [f32(R, w) for w in self.widths]
            self.pf = [f32(R // 32, 2, w) for w in self.widths]
            self.pb = [f32(R // 32, 2, w) for w in self.widths]
            # dropout keep bytes of the applications consumed by a tile GEMM: drawn once by the kernel that produces Z[a]
            self.keepbuf = [torch.ones((R, w), dtype=torch.uint8, device=dev) for w in self.widths]
            if self.comm is not None:   # data parallel: global backward means per application, written by scnb_hs_xchg
                self.bmean = [f32(self.n_seg, 2, w) for w in self.widths]
        # "auto": the tensor-core form pays off once the batch fills at least two 256-row tiles (measured: 512 rows,
        # 5 % density: 22 us vs 40 us for the SpMM; 8192 rows, 20 %: ~10x); tiny batches keep the exact-fp32 SpMM
        fl = cfg.first_layer
        if fl == "auto":
            fl = os.environ.get("SCNB_FIRST_LAYER", "tc" if self.nb >= 512 else "spmm")
        self.tc_first = (fl == "tc") and (H0 % 32 == 0) and (32 <= H0 <= 256)
        if self.tc_first:
            n_kt = (self.G + 63) // 64
            self.w1_hi = torch.zeros(n_kt * H0 * 64, dtype=torch.int16, device=dev)
            self.w1_lo = torch.zeros(n_kt * H0 * 64, dtype=torch.int16, device=dev)
            self._ev_planes = torch.cuda.Event()
        # the tensor-core dW1 + optimizer kernel writes the operand planes of the updated W1T itself (single process
        # only: under data parallelism the update happens in the arena-wide optimizer pass)
        self.w1_emits_planes = bool(self.tc_first and self.w1_tc and comm is None)
        self._planes_version = None
        if hasattr(model, "register_load_state_dict_post_hook"):
            import weakref
            me = weakref.ref(self)

            def _on_load(module, incompatible_keys):   # must return None
                eng = me()
                if eng is not None:
                    eng.invalidate_planes()
            model.register_load_state_dict_post_hook(_on_load)
        self._bn_running_done = False
        self._dz_planes_ready = False
        self._side = None
        self._gemm_ws = None
        fus = lambda c: (self.H % 128 == 0 and self.H // 128 in (1, 2, 4, 8) and c * self.H * 4 <= 96 * 1024)
        self.fuse_classif = bool(cfg.fuse_heads and fus(self.C))
        self.fuse_dan = bool(cfg.fuse_heads and self.use_dan and fus(self.D))
        # heads fused with the ends of the hidden stack (head_ops.cu): last BatchNorm -> Dropout -> ReLU, head, losses, dA and the
        # adjoint of the last activation in one launch per branch (SCNB_HEADS=split keeps the five-launch chain)
        self.heads_fused = bool(self.hs_fused and mode == "mixmatch" and cfg.fuse_heads and self.H in (128, 256) and self.C <= 32
                                and (not self.use_dan or self.D <= 32) and os.environ.get("SCNB_HEADS", "fused") != "split")
        self._ev_grl, self._ev_csc, self._ev_zero = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self._ev_teacher, self._ev_plan, self._ev_bnrun, self._ev_z1 = (torch.cuda.Event() for _ in range(4))
        self._ev_dz = [torch.cuda.Event() for _ in self.widths]
        self._ev_prep = torch.cuda.Event()
        self.dZ1 = f32(self.nb, H0) if mode == "mixmatch" else None
        if self.use_dan:
            self.Hd, self.dHd = f32(self.Rd, self.H), f32(self.Rd, self.H)
            self.dlog, self.ddlog = f32(self.Rd, self.D), f32(self.Rd, self.D)
            self.dlabel = f32(self.Rd, self.D)
        if mode == "mixmatch":
            self._const_seg(self.K_eff * U)  # created here: no H2D copies may happen during graph capture
        self.loss_buf = f32(8)       # [sup, sup_hits, unsup, -, dan, dan_hits, -, -]
        self.loss_hist = None
        self._tabs = None            # per-epoch sampler tables (device); see set_epoch_tables
        self._graph = None
        self._graph_key = None
        self.steps_done = 0
        self.kernels_per_step = None
        # Prologue of step t+1 (sampler hand-off, row gather + augmentation, MixUp plan, gene-block splits: ~18 us at the
        # front of the critical chain, none of it depends on the weights) assembled WHILE step t runs, into the second set of
        # batch buffers.  Needs everything the prologue produces to be independent of the teacher (no thresholding) and of
        # host-staged rows (sampler tables on the device), and the step to be a replayed graph.
        self.prefetch_ok = bool(mode == "mixmatch" and cfg.pseudolabel_min_confidence <= 0.0 and not self.teacher_own_spmm
                                and self.aug_code == 1 and self.w1_tc and self.tc_first and cfg.multi_stream
                                and os.environ.get("SCNB_PREFETCH", "1") != "0")
        if self.prefetch_ok:
            self._sets[1] = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in self._sets[0].items()}
        self._pf_graphs, self._next_p, self._prefetch_valid = {}, 0, False
        self._ev_adv, self._ev_fl, self._ev_pf = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self._ev_loss, self._ev_pub, self._ev_pplan, self._ev_pplan_done, self._ev_fetched = (torch.cuda.Event() for _ in range(5))

    # ------------------------------------------------------------------------------------
    def set_hyper(self, lr=None, weight_decay=None):
        c = self.cfg
        if lr is not None:
            c.lr = float(lr)
        if weight_decay is not None:
            c.weight_decay = float(weight_decay)
        eps = c.eps if c.eps is not None else (1e-6 if c.optimizer == "adadelta" else 1e-8)
        vals = [c.lr, c.weight_decay, c.betas[0], c.betas[1], eps, c.rho, c.momentum, 0.0]
        self.hp.copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=False)

    def set_weights(self, unsup_weight: float, dan_weight: float = 0.0):
        """ICLWeight(epoch) values (trainer.py:628,654); changing them re-captures the graph."""
        self.unsup_weight, self.dan_weight = float(unsup_weight), float(dan_weight)

    def P(self, name):
        return ptr(self.arena.w(name))

    def Gp(self, name):
        return ptr(self.arena.g(name))

    # ------------------------------------------------------------------------------------
    def _teacher_ptrs(self):
        """(weights, BN buffers) the eval-mode teacher reads."""
        if self.cfg.mean_teacher:
            off = self.arena.offsets
            tw = lambda name: self.teacher_flat[off[name]: off[name] + self.arena.sizes[name]]
            bn = self.teacher_bn
            return tw, (lambda a: (bn[a][0], bn[a][1]))
        return (lambda name: self.arena.w(name)), (lambda a: (self._blk[a]["bn"].running_mean, self._blk[a]["bn"].running_var))

    # ---- stream plumbing: the step is a DAG; independent kernels go to side streams so that a captured
    #      graph keeps only ~22 launches on its critical path.  Under data parallelism every collective is
    #      issued from the main stream, in the same order on every rank; the side branches carry compute only. ----
    def _streams(self):
        main = torch.cuda.current_stream()
        if not self.cfg.multi_stream:
            return main, main, main
        if self._side is None:
            self._side = tuple(torch.cuda.Stream(device=self.dev) for _ in range(6))
        return main, self._side[0], self._side[1]

    @staticmethod
    def _after(waiter, other):
        """`waiter` continues only after everything issued so far on `other`."""
        if waiter is not other:
            waiter.wait_stream(other)

    def _refresh_planes(self, main, sW):
        """bf16 hi/lo planes of the current W1T for the tensor-core first layer (on the side branch, overlapping
        the step prologue); recomputed every step because the optimizer has just changed W1T."""
        if not self.tc_first or self.w1_emits_planes:
            return
        self._after(sW, main)
        call("scnb_w1_planes", self.P(self._blk[0]["W"]), self.G, self.widths[0], ptr(self.w1_hi), ptr(self.w1_lo), sW.cuda_stream)
        if sW is not main:
            self._ev_planes.record(sW)

    def invalidate_planes(self):
        """Call after writing the first-layer weight from outside the engine's kernels through a path torch cannot see
        (raw `.data` edits, another library): the next step rebuilds the bf16 operand planes of W1T before it runs."""
        self._planes_version = None

    def _reserve_gemm_ws(self):
        """Workspaces for the operand planes of the large hidden-layer products (tc_gemm.cu, planes form), one per stream
        that issues them (main chain, weight-gradient branch, domain-adversary branch, teacher branch).  Only when the
        step has products above the tcgen05 take-over size (config 5: 16 384 x 1024 x 1024); allocated here, once, before
        any capture -- the library itself allocates nothing.  An engine without such products clears whatever a dead
        engine may have left registered for a recycled stream handle."""
        if self._gemm_ws is not None or not self.cfg.multi_stream or os.environ.get("SCNB_MAIN_PRIORITY", "1") == "0":
            return
        self._streams()
        if getattr(self, "_hi", None) is None:
            self._hi = torch.cuda.Stream(device=self.dev, priority=-1)
        handles = [s.cuda_stream for s in (self._hi, self._side[0], self._side[1], self._side[2])]
        mf = ctypes.c_longlong(0)
        call("scnb_gemm_tc_min_flops_get", ctypes.byref(mf))
        R, W = int(self.Rmax), int(max(self.widths))
        self._gemm_ws = []
        if not (mf.value > 0 and 2.0 * R * W * W >= mf.value and R >= 128 and W >= 64):
            for h in handles:
                call("scnb_gemm_tc_workspace", h, None, 0)
            return
        need = ctypes.c_longlong(0)
        n_bytes = 0
        for (m, n, k) in ((R, W, W), (W, W, R)):
            call("scnb_gemm_tc_workspace_bytes", m, n, k, ctypes.byref(need))
            n_bytes = max(n_bytes, int(need.value))
        for h in handles:
            buf = torch.empty(n_bytes + 1024, dtype=torch.uint8, device=self.dev)
            self._gemm_ws.append(buf)
            call("scnb_gemm_tc_workspace", h, buf.data_ptr() + (-buf.data_ptr()) % 1024, n_bytes)
        # Operand planes written by the producers (dense_ops.cu PlaneSink): the BatchNorm forward writes As[a] also as the planes
        # the next Linear's product reads, the BatchNorm backward writes dZs[a] also as the planes of the input-gradient
        # product; the weight-gradient product reads both sets as MN-major operands.  (Unfused hidden path, one GPU.)
        bound = []
        if not self.hs_fused and self.comm is None and os.environ.get("SCNB_GEMM_TC_BOUND", "1") != "0":
            Rp = (R + 255) // 256 * 256
            todo = [self.As[a] for a in range(self.n_app)] + [self.dZs[a] for a in range(1, self.n_app)]
            for t in todo:
                w = int(t.shape[1])
                if w % 32 or int(t.shape[0]) < R or not t.is_contiguous():
                    continue
                pair = []
                for _ in range(2):
                    buf = torch.zeros(Rp * w * 2 + 1024, dtype=torch.uint8, device=self.dev)
                    self._gemm_ws.append(buf)
                    pair.append(buf.data_ptr() + (-buf.data_ptr()) % 1024)
                call("scnb_gemm_tc_bind_planes", t.data_ptr(), R, w, pair[0], pair[1])
                bound.append((t.data_ptr(), R, w))
        import weakref
        weakref.finalize(self, _release_gemm_ws, handles, bound)

    def _sync_planes(self):
        """With `w1_emits_planes` the step itself keeps the operand planes current.  They are rebuilt here, eagerly, when
        something outside the engine's kernels may have written the first-layer weight since: `load_state_dict` (a
        post-hook on the model calls `invalidate_planes`), in-place torch ops on the Parameter or on the arena (a torch
        optimizer, `p.copy_()`: both version counters are tracked -- a Parameter re-pointed at an arena view keeps its
        OWN counter, the arena's does not move), every new epoch of sampler tables, and an explicit `invalidate_planes()`."""
        self._reserve_gemm_ws()
        if not self.w1_emits_planes:
            return
        w1 = self.arena.params[0]
        v = (self.arena.flat._version, w1._version)
        if v != self._planes_version:
            call("scnb_w1_planes", self.P(self._blk[0]["W"]), self.G, self.widths[0], ptr(self.w1_hi), ptr(self.w1_lo),
                 torch.cuda.current_stream().cuda_stream)
            self._planes_version = v

    def _first_layer(self, main, sW, n_rows, Z, prezeroed=False, splits_ready=False):
        st = main.cuda_stream
        blk0 = self._blk[0]
        if self.tc_first:
            if sW is not main and not self.w1_emits_planes:
                main.wait_event(self._ev_planes)
            if prezeroed and not getattr(self, "_zero_at_end", False):
                main.wait_event(self._ev_z1)
            if splits_ready and n_rows == self.nb and self.fwd_splits is not None and os.environ.get("SCNB_FWD_SPLITS", "1") != "0":
                call("scnb_csr_linear_fwd_tc_splits", ptr(self.b_indptr), ptr(self.b_idx), ptr(self.b_val), n_rows, self.G,
                     self.widths[0], ptr(self.w1_hi), ptr(self.w1_lo), self.P(blk0["b"]), ptr(Z), -1 if prezeroed else 0,
                     ptr(self.fwd_splits), st)
            else:
                call("scnb_csr_linear_fwd_tc", ptr(self.b_indptr), ptr(self.b_idx), ptr(self.b_val), n_rows, self.G, self.widths[0],
                     ptr(self.w1_hi), ptr(self.w1_lo), self.P(blk0["b"]), ptr(Z), -1 if prezeroed else 0, st)
        else:
            call("scnb_csr_linear_fwd", ptr(self.b_indptr), ptr(self.b_idx), ptr(self.b_val), n_rows, self.P(blk0["W"]),
                 self.P(blk0["b"]), self.widths[0], ptr(Z), st)

    def _prep(self, st, phases=3):
        t = self._tabs or {}
        mm = self.mode == "mixmatch"
        call("scnb_step_prep", ptr(t.get("l")), ptr(t.get("u")), ptr(t.get("k")), ptr(t.get("g")), int(t.get("n", 0)),
             ptr(self.state), self.L, self.U, ptr(self.l_rows), ptr(self.u_rows) if mm else None, ptr(self.perm_keys),
             ptr(self.gamma_raw), ptr(self.lab_y), ptr(self.lab_dom), ptr(self.unl_dom), int(self.use_dan), ptr(self.lab_ids),
             ptr(self.base_dom) if self.use_dan else None, ptr(self.lab.indptr), ptr(self.unl.indptr) if mm else None,
             ptr(self.b_indptr), ptr(self.loss_buf), 8, int(phases), st)

    def _gather_batch(self, st, rng, ctr_off=0, fetched=None):
        """Batch CSR [U_raw ; L_aug] (+ per-gene histogram for the SIMT dW1 path).  ctr_off = 1: the batch of the NEXT step."""
        c, L, U = self.cfg, self.L, self.U
        if self.host_mapped and self.aug_code == 1:
            call("scnb_csr_fetch_rows2", ptr(self.unl.indptr), ptr(self.unl.indices), ptr(self.unl.data), ptr(self.u_rows), U,
                 ptr(self.lab.indptr), ptr(self.lab.indices), ptr(self.lab.data), ptr(self.l_rows), L, ptr(self.b_indptr),
                 ptr(self.st_indptr), ptr(self.st_idx), ptr(self.st_val), int(os.environ.get("SCNB_FETCH_CTAS", "24")), st)
            if fetched is not None:     # (recorded on the stream that `st` belongs to: the prefetch branch)
                fetched.record(self._side[4])
            sp = self.st_indptr
            call("scnb_csr_gather_rows2_nocache", ptr(sp), ptr(self.st_idx), ptr(self.st_val), None, U, 0, 0,
                 ptr(sp[U:]), ptr(self.st_idx), ptr(self.st_val), None, L, 1, SID_IN_L,
                 ptr(self.b_indptr), ptr(self.b_idx), ptr(self.b_val), ptr(self.inj_in_keep_L), rng, c.p_in_drop,
                 self._gene_cnt_ptr(), int(ctr_off), st)
        elif self.host_mapped:
            raise NotImplementedError("host-mapped matrices are fed through the log1p_drop gather only")
        elif self.aug_code == 1:
            # (SCNB_PREFETCH_GATHER=nocache: the cache-free form, which could share an SM with the first layer's CTA; measured:
            # the block scheduler does not co-schedule it with the tcgen05 kernels either way -- no gain, off)
            light = ctr_off == 1 and os.environ.get("SCNB_PREFETCH_GATHER", "cache") == "nocache"
            call("scnb_csr_gather_rows2_nocache" if light else "scnb_csr_gather_rows2",
                 ptr(self.unl.indptr), ptr(self.unl.indices), ptr(self.unl.data), ptr(self.u_rows), U, 0, 0,
                 ptr(self.lab.indptr), ptr(self.lab.indices), ptr(self.lab.data), ptr(self.l_rows), L, 1, SID_IN_L,
                 ptr(self.b_indptr), ptr(self.b_idx), ptr(self.b_val), ptr(self.inj_in_keep_L), rng, c.p_in_drop,
                 self._gene_cnt_ptr(), int(ctr_off), st)
        else:   # the other schemes of dataprep.py:730-754: raw unlabeled rows, then the labeled rows through the scheme
            call("scnb_csr_gather_rows", ptr(self.unl.indptr), ptr(self.unl.indices), ptr(self.unl.data), ptr(self.u_rows), 0, U,
                 ptr(self.b_indptr), ptr(self.b_idx), ptr(self.b_val), 0, 0, 0, None, None, 0, 0.0, self._gene_cnt_ptr(), st)
            call("scnb_csr_gather_rows_aug", ptr(self.lab.indptr), ptr(self.lab.indices), ptr(self.lab.data), ptr(self.l_rows), 0, L,
                 ptr(self.b_indptr[U:]), ptr(self.b_idx), ptr(self.b_val), self.aug_code, self.G, c.p_in_drop, c.p_mask_apply,
                 ptr(self.inj_in_keep_L), ptr(self.inj_pois_L), ptr(self.inj_rowu_L), rng, SID_IN_L, self._gene_cnt_ptr(),
                 int(ctr_off), st)

    def _prologue_next(self, stream, rng, splits=True):
        """The whole prologue of the step that comes AFTER the counters' current value, into the current set of batch
        buffers, on one stream: sampler hand-off (table row state[0] % n), gather (RNG counter + 1), MixUp plan, splits
        (`splits=False`: the caller enqueues `_build_csc` itself, later in the step)."""
        st = stream.cuda_stream
        self._prep(st, phases=1)
        side = self._side[5] if (self.host_mapped and self._side is not None and stream is self._side[4]) else None
        if side is None:
            self._gather_batch(st, rng, ctr_off=1)
            self._plan(self._all_confident, st)
            if splits:
                self._build_csc(self.nb, st)
            return
        # Rows fetched from host memory take ~100 us.  What does not need them runs next to the fetch on a branch of its
        # own (the MixUp plan: only the hand-off), and what needs only their gene indices (the gene-block split points)
        # runs next to the augmenting gather that follows it -- not behind either.
        self._ev_pplan.record(stream)
        side.wait_event(self._ev_pplan)
        self._plan(self._all_confident, side.cuda_stream)
        self._gather_batch(st, rng, ctr_off=1, fetched=self._ev_fetched)
        side.wait_event(self._ev_fetched)
        self._build_csc(self.nb, side.cuda_stream, idx=self.st_idx)
        self._ev_pplan_done.record(side)
        stream.wait_event(self._ev_pplan_done)

    def _launch_mixmatch(self):
        c = self.cfg
        main, sW, sD = self._streams()   # main chain / weight-gradient + bookkeeping branch / domain-adversary branch
        if not self.use_dan:
            sD = main                    # never forked: a join would reference a stream outside the capture
        st = main.cuda_stream
        L, U, nb, R, C, H0 = self.L, self.U, self.nb, self.Rmax, self.C, self.widths[0]
        rng = ptr(self.state[1:3])
        blk = self._blk
        self._refresh_planes(main, sW)
        z1_pre = self.tc_first and sW is not main
        pre = bool(getattr(self, "_prefetched", False))   # this step's batch was assembled during the previous step
        # SCNB_ZERO_AT_END=1: replayed, prefetched steps clear Z1 and the gradient arena at the END of the previous step
        # (branch W, behind the optimizer and under the dW1 kernel) so that nothing is in front of the first layer.
        # Measured: 0.1604 vs 0.1613 ms -- not worth gradients that read as zero after a step; off by default
        self._zero_at_end = bool(pre and z1_pre and os.environ.get("SCNB_ZERO_AT_END", "0") == "1")
        if z1_pre and not self._zero_at_end:   # the split-K first layer accumulates into Z1: clear it on the side branch
            self._after(sW, main)
            self._zero_grads(sW.cuda_stream, also=self.Z1)     # one launch: Z1 and the gradient arena behind the first layer
            self._ev_z1.record(sW)
            self._ev_zero.record(sW)
        elif self._zero_at_end:
            self._after(sW, main)
        teacher_async = (c.pseudolabel_min_confidence <= 0.0) and sW is not main
        if pre:
            # only the counters are left of the prologue: advanced on branch W, needed from the hidden stack on
            call("scnb_step_prep", None, None, None, None, 0, ptr(self.state), L, U, ptr(self.l_rows), ptr(self.u_rows), None, None,
                 None, None, None, 0, ptr(self.lab_ids), None, ptr(self.lab.indptr), ptr(self.unl.indptr), ptr(self.b_indptr),
                 ptr(self.loss_buf), 8, 2, sW.cuda_stream)
            self._ev_adv.record(sW)
            self._ev_csc.record(sW)
        else:
            self._prep(st)
            # With pseudolabel_min_confidence <= 0 every unlabeled row is "confident": the MixUp plan depends only on the
            # step's random draws, so it runs on its own branch next to the gather and the first layer.
            if teacher_async:
                self._ev_prep.record(main)
            # -- batch CSR [U_raw ; L_aug] (+ per-gene histogram), one launch (issued before the plan: the graph releases
            #    the children of a node one after the other, and the gather is the one on the critical chain)
            self._gather_batch(st, rng)
            if teacher_async:
                sP = self._side[3]
                sP.wait_event(self._ev_prep)
                self._plan(self._all_confident, sP.cuda_stream)
                self._ev_plan.record(sP)
            # branch W: gene-major copy of the batch (needed only by the last kernel of the step), gradient zeroing
            self._after(sW, main)
            if not z1_pre:
                self._zero_grads(sW.cuda_stream)
                if sW is not main:
                    self._ev_zero.record(sW)
            self._build_csc(nb, sW.cuda_stream)
            if sW is not main:
                self._ev_csc.record(sW)
        # -- first layer, once
        self._first_layer(main, sW, nb, self.Z1, prezeroed=z1_pre, splits_ready=pre)
        if pre:
            # the prologue of the NEXT step, into the other set of batch buffers, on an idle low-priority branch (it reads
            # the sampler tables and the datasets only; it must see this step's counters: after the advance)
            main.wait_event(self._ev_adv)
            self._ev_fl.record(main)
            sF = self._side[4]
            sF.wait_event(self._ev_adv if os.environ.get("SCNB_PREFETCH_AT", "start") == "start" else self._ev_fl)
            cur = self.__dict__["_cur"]
            self.__dict__["_cur"] = 1 - cur
            # The split points of the next batch (one short CTA per row) are enqueued BEHIND the heads (SCNB_SPLITS_LATE=1,
            # default): launched next to them they delayed the cluster launch of the adversary head by 6 us (timeline:
            # hs_dan_head started when row_gene_splits drained).  0: with the rest of the prologue; 2: next to the dW1 kernel
            # (measured 0.1480 / 0.1466 / 0.1489 ms for 0 / 1 / 2)
            splits_late = os.environ.get("SCNB_SPLITS_LATE", "1") if not self.host_mapped else "0"
            splits_late = {"0": 0, "1": 1, "2": 2}.get(splits_late, 0)    # 1: behind the heads, 2: next to the dW1 kernel
            # SCNB_PREFETCH_AT=heads (experiment): the WHOLE prologue of the next step behind the heads
            pf_deferred = os.environ.get("SCNB_PREFETCH_AT", "start") == "heads" and not self.host_mapped
            if pf_deferred:
                splits_late = 3
                self.__dict__["_cur"] = cur
            else:
                try:
                    self._prologue_next(sF, rng, splits=splits_late == 0)
                finally:
                    self.__dict__["_cur"] = cur
            if splits_late == 0:
                self._ev_pf.record(sF)
        else:
            splits_late = 0
        # -- teacher (eval mode, losses.py:660-737).  With pseudolabel_min_confidence <= 0 every unlabeled row is
        #    "confident", so the MixUp plan and the student forward do not depend on the teacher: it runs on its own
        #    branch and is joined in front of the losses.
        if teacher_async:
            sT = self._side[2]
            self._after(sT, main)
            self._teacher_forward(sT.cuda_stream, rng)
            self._ev_teacher.record(sT)
        else:
            self._teacher_forward(st, rng)
        # -- MixUp plan + hidden-space MixUp (losses.py:811-875)
        if pre:
            pass                              # the MixUp plan of this step was part of the prefetched prologue
        elif teacher_async:
            main.wait_event(self._ev_plan)
        else:
            self._plan(self.confident, st)
        if self.hs_fused:    # hidden-space MixUp + tile statistics of its output, one launch
            ko = None if (self.inj_hidden_keep is not None or float(blk[0]["drop"].p) <= 0.0) else ptr(self.keepbuf[0])
            call("scnb_hs_mix", ptr(self.Z1), H0, ptr(self.srcA), ptr(self.srcB), ptr(self.gam), R, ptr(self.Zs[0]), ptr(self.pf[0]), ko, rng,
                 SID_HID, float(blk[0]["drop"].p), st)
        else:
            call("scnb_mix_rows", ptr(self.Z1), H0, ptr(self.srcA), ptr(self.srcB), ptr(self.gam), ptr(self.counts[3:]), R,
                 ptr(self.Zs[0]), st)
        # -- student super-batch forward
        self._student_forward(st, rng)
        # BatchNorm running statistics as soon as the forward statistics exist, off the critical path (the
        # stream of the MixUp plan is idle by now); the eval-mode teacher must have read the OLD running statistics first
        self._bn_running_done = False
        hf = self.heads_fused

        def running_update():
            sB = self._side[3]          # the plan's branch, long idle by now
            self._after(sB, main)
            if teacher_async:
                sB.wait_event(self._ev_teacher)
            self._bn_running_update(sB.cuda_stream)
            self._ev_bnrun.record(sB)
            self._bn_running_done = True
        if sW is not main and not hf:
            running_update()
        Hh = self.H
        dA_last = self._view(self.dA, R, Hh)
        if hf:
            last = self.n_app - 1
            pl = float(blk[last]["drop"].p)
            keep_last = self._hidden_keep(last) if self.inj_hidden_keep is not None else (ptr(self.keepbuf[last]) if pl > 0.0 else None)
            head_in = (ptr(self.Zs[last]), ptr(self.As[last]), None if self.comm is not None else ptr(self.pf[last]), ptr(self.stats[last]),
                       self.P(blk[last]["g"]), self.P(blk[last]["beta"]), keep_last, pl, Hh, R)
            seg = self._seg_host.ctypes.data
        # branch D: domain adversary head, its loss and the reversed gradient into the embedding
        if self.use_dan:
            d0, d1 = self._dan_names()
            Ae = self.As[-1][nb:]
            self._after(sD, main)
            sd = sD.cuda_stream
            scale = ptr(self.pconf) if c.dan_scale_loss_pseudoconf else None
            if hf:
                call("scnb_hs_dan_head", *head_in, nb, self.n_seg, seg, self.P(d0 + ".weight"), self.P(d0 + ".bias"),
                     self.P(d1 + ".weight"), self.P(d1 + ".bias"), self.D, ptr(self.counts[2:]), ptr(self.base_dom), ptr(self.srcA), nb,
                     scale, float(self.dan_weight), ptr(self.Hd), ptr(self.dHd), ptr(self.dlog), ptr(self.dlabel), ptr(self.ddlog),
                     ptr(self.loss_buf[4:]), ptr(self.dYs[last]), ptr(self.pb[last]), sd)
            else:
                call("scnb_sgemm", 0, 1, self.Rd, Hh, Hh, ptr(Ae), Hh, self.P(d0 + ".weight"), Hh, ptr(self.Hd), Hh, 1.0, 0.0,
                     self.P(d0 + ".bias"), 1, None, sd)
            if hf:
                pass
            elif self.fuse_dan:
                call("scnb_dan_tail_fused", ptr(self.Hd), self.P(d1 + ".weight"), self.P(d1 + ".bias"), Hh, self.D, self.Rd,
                     ptr(self.counts[2:]), ptr(self.base_dom), ptr(self.srcA), nb, scale, ptr(self.dlog), ptr(self.dlabel),
                     ptr(self.ddlog), ptr(self.dHd), ptr(self.loss_buf[4:]), sd)
            else:
                call("scnb_sgemm", 0, 1, self.Rd, self.D, Hh, ptr(self.Hd), Hh, self.P(d1 + ".weight"), Hh, ptr(self.dlog), self.D,
                     1.0, 0.0, self.P(d1 + ".bias"), 0, None, sd)
                call("scnb_onehot_rows", ptr(self.base_dom), ptr(self.srcA), nb, self.Rd, self.D, ptr(self.dlabel), sd)
                call("scnb_soft_ce_fwd_bwd", ptr(self.dlog), self.Rd, self.D, ptr(self.counts[2:]), ptr(self.dlabel), None, None,
                     None, 0, None, ptr(self.counts[2:]), 0, 1.0, scale, ptr(self.ddlog), ptr(self.loss_buf[4:]), 1, None, sd)
                call("scnb_sgemm", 0, 0, self.Rd, Hh, self.D, ptr(self.ddlog), self.D, self.P(d1 + ".weight"), Hh, ptr(self.dHd),
                     Hh, 1.0, 0.0, None, 0, ptr(self.Hd), sd)
            # gradient reversal: dEmbed = -w * dHd W  (model.py:338)
            if not hf:
                call("scnb_sgemm", 0, 0, self.Rd, Hh, Hh, ptr(self.dHd), Hh, self.P(d0 + ".weight"), Hh, ptr(dA_last[nb:]), Hh,
                     -self.dan_weight, 0.0, None, 0, None, sd)
            if sD is not main:
                self._ev_grl.record(sD)
            # the head's own weight gradients (needed only by the optimizer)
            if sD is not main:
                if not self._zero_at_end:
                    sD.wait_event(self._ev_zero)   # after the gradient zeroing
            call("scnb_sgemm", 1, 0, self.D, Hh, self.Rd, ptr(self.ddlog), self.D, ptr(self.Hd), Hh, self.Gp(d1 + ".weight"), Hh,
                 1.0, 1.0, None, 0, None, sd)
            call("scnb_colsum", ptr(self.ddlog), self.Rd, self.D, self.D, self.Gp(d1 + ".bias"), 1, sd)
            if self.hs_fused:
                call("scnb_hs_dw", ptr(self.dHd), Hh, ptr(Ae), Hh, self.Rd, self.Gp(d0 + ".weight"), self.Gp(d0 + ".bias"), sd)
            else:
                call("scnb_sgemm", 1, 0, Hh, Hh, self.Rd, ptr(self.dHd), Hh, ptr(Ae), Hh, self.Gp(d0 + ".weight"), Hh, 1.0, 1.0, None,
                     0, None, sd)
                call("scnb_colsum", ptr(self.dHd), self.Rd, Hh, Hh, self.Gp(d0 + ".bias"), 1, sd)
        # -- classifier, losses + dLogits (losses.py:896-923, 1224-1243; trainer.py:651-656) and dA of the classifier
        if teacher_async:
            main.wait_event(self._ev_teacher)
        if hf:
            call("scnb_hs_classif_head", *head_in, self.n_seg, seg, self.P("model.classif.0.weight"), self.P("model.classif.0.bias"), C, U, L,
                 ptr(self.counts), ptr(self.Ycat), ptr(self.srcA), ptr(self.srcB), ptr(self.gam), ptr(self.class_weight),
                 self.unsup_weight, ptr(self.logits), ptr(self.dlogits), ptr(self.loss_buf), ptr(self.dYs[last]), ptr(self.pb[last]), st)
        elif self.fuse_classif:
            call("scnb_classif_mixmatch_fused", ptr(self.As[-1]), self.P("model.classif.0.weight"), self.P("model.classif.0.bias"),
                 Hh, C, U, L, ptr(self.counts), ptr(self.Ycat), ptr(self.srcA), ptr(self.srcB), ptr(self.gam),
                 ptr(self.class_weight), self.unsup_weight, ptr(self.logits), ptr(self.dlogits), ptr(dA_last), ptr(self.loss_buf), st)
        else:
            call("scnb_sgemm", 0, 1, nb, C, Hh, ptr(self.As[-1]), Hh, self.P("model.classif.0.weight"), Hh,
                 ptr(self.logits), C, 1.0, 0.0, self.P("model.classif.0.bias"), 0, None, st)
            call("scnb_softmax_mse_fwd_bwd", ptr(self.logits), U, C, ptr(self.counts), ptr(self.Ycat), ptr(self.srcA),
                 ptr(self.srcB), ptr(self.gam), 0, U, None, self.unsup_weight, ptr(self.dlogits), ptr(self.loss_buf[2:]), st)
            call("scnb_soft_ce_fwd_bwd", ptr(self.logits[U:]), L, C, None, ptr(self.Ycat), ptr(self.srcA), ptr(self.srcB),
                 ptr(self.gam), U, ptr(self.class_weight), None, L, 1.0, None, ptr(self.dlogits[U:]), ptr(self.loss_buf), 1, None, st)
            # -- backward: the dA chain stays on the main stream, weight gradients go to branch W
            call("scnb_sgemm", 0, 0, nb, Hh, C, ptr(self.dlogits), C, self.P("model.classif.0.weight"), Hh, ptr(dA_last), Hh, 1.0,
                 0.0, None, 0, None, st)
        self._after(sW, main)
        sw = sW.cuda_stream
        call("scnb_sgemm", 1, 0, C, Hh, nb, ptr(self.dlogits), C, ptr(self.As[-1]), Hh, self.Gp("model.classif.0.weight"), Hh,
             1.0, 1.0, None, 0, None, sw)
        call("scnb_colsum", ptr(self.dlogits), nb, C, C, self.Gp("model.classif.0.bias"), 1, sw)
        if self.use_dan and sD is not main:
            main.wait_event(self._ev_grl)
        def late_splits():
            sF = self._side[4]
            self._ev_fl.record(main)
            sF.wait_event(self._ev_fl)
            cur = self.__dict__["_cur"]
            self.__dict__["_cur"] = 1 - cur
            try:
                self._build_csc(self.nb, sF.cuda_stream)
            finally:
                self.__dict__["_cur"] = cur
            self._ev_pf.record(sF)
        if splits_late == 1:
            late_splits()
        if splits_late == 3:
            sF = self._side[4]
            self._ev_fl.record(main)
            sF.wait_event(self._ev_fl)
            cur = self.__dict__["_cur"]
            self.__dict__["_cur"] = 1 - cur
            try:
                self._prologue_next(sF, rng, splits=True)
            finally:
                self.__dict__["_cur"] = cur
            self._ev_pf.record(sF)
        if sW is not main and hf:
            running_update()      # the statistics of the last application are written by the head kernels
        published = False
        if self.post_launch is not None and sW is not main:
            # the step's loss scalars are final once both heads have run: a feeding pipeline's read-back goes out now, on
            # the (idle) branch of the running statistics, instead of behind the dW1 kernel at the end of the step
            sQ = self._side[3]
            self._ev_loss.record(main)
            sQ.wait_event(self._ev_loss)
            with torch.cuda.stream(sQ):
                self.post_launch(self)
            self._ev_pub.record(sQ)
            published = True
        self._published = published
        dZ_first = self._student_backward(main, sW, rng, dA_last)
        self._dz_planes_ready = bool(self.w1_tc)
        if dZ_first is None:
            pass         # (scnb_hs_bwd_in_unmix_planes has written dZ1 and its planes)
        elif self.w1_tc:   # MixUp adjoint + bf16 operand planes of dZ1 for the tensor-core dW1 kernel, one launch
            call("scnb_unmix_rows_planes", ptr(dZ_first), H0, ptr(self.inv3), ptr(self.coef3), nb, ptr(self.dZ1), ptr(self.dz_hi),
                 ptr(self.dz_lo), st)
        else:
            call("scnb_unmix_rows", ptr(dZ_first), H0, ptr(self.inv3), ptr(self.coef3), nb, ptr(self.dZ1), st)
        if splits_late == 2:
            late_splits()
        self._finish_step(main, sW, sD, self.dZ1, nb)
        if published:
            main.wait_event(self._ev_pub)
        if pre:
            main.wait_event(self._ev_pf)   # the next step's batch is complete when this step is

    def _plan(self, conf_flags, st):
        c = self.cfg
        call("scnb_mixmatch_plan", ptr(conf_flags), ptr(self.perm_keys), ptr(self.gamma_raw), self.U, self.L, int(self.use_dan),
             int(c.dan_use_conf_pseudolabels), int(c.mixup_alpha > 0.0), ptr(self.counts), ptr(self.seg_off),
             ptr(self.plan_work), ptr(self.srcA), ptr(self.srcB), ptr(self.gam), self.Rmax, ptr(self.inv3), ptr(self.coef3),
             ptr(self.pconf), st)

    def _teacher_forward(self, st, rng):
        """Eval-mode teacher pass on stream `st`: pseudolabels q[0,U) (+ one-hot labels q[U,U+L)), confidences."""
        c = self.cfg
        L, U, C, H0 = self.L, self.U, self.C, self.widths[0]
        blk = self._blk
        tw, tbn = self._teacher_ptrs()
        if c.mean_teacher:
            call("scnb_ema_update", ptr(self.teacher_flat), ptr(self.arena.flat), self.n_model_params, c.decay_coef,
                 ptr(self.state), st)
        K = self.K_eff
        if self.teacher_own_spmm:
            for k in range(K):
                if c.augment_pseudolabels:
                    call("scnb_csr_row_offsets", ptr(self.unl.indptr), ptr(self.u_rows), 0, U, ptr(self.tb_indptr), None, st)
                    km = self.inj_in_keep_U[k] if self.inj_in_keep_U is not None else None
                    if self.aug_code == 1:
                        call("scnb_csr_gather_rows", ptr(self.unl.indptr), ptr(self.unl.indices), ptr(self.unl.data),
                             ptr(self.u_rows), 0, U, ptr(self.tb_indptr), ptr(self.tb_idx), ptr(self.tb_val), 1, 0, U, ptr(km), rng,
                             SID_IN_U + k, c.p_in_drop, None, st)
                    else:
                        pk = self.inj_pois_U[k] if self.inj_pois_U is not None else None
                        ru = self.inj_rowu_U[k] if self.inj_rowu_U is not None else None
                        call("scnb_csr_gather_rows_aug", ptr(self.unl.indptr), ptr(self.unl.indices), ptr(self.unl.data),
                             ptr(self.u_rows), 0, U, ptr(self.tb_indptr), ptr(self.tb_idx), ptr(self.tb_val), self.aug_code, self.G,
                             c.p_in_drop, c.p_mask_apply, ptr(km), ptr(pk), ptr(ru), rng, SID_IN_U + k, None, st)
                    ip, ix, vv = self.tb_indptr, self.tb_idx, self.tb_val
                else:
                    ip, ix, vv = self.b_indptr, self.b_idx, self.b_val
                call("scnb_csr_linear_fwd", ptr(ip), ptr(ix), ptr(vv), U, ptr(tw(blk[0]["W"])), ptr(tw(blk[0]["b"])), H0,
                     ptr(self.TZ1[k * U:]), st)
            tz = self.TZ1
        else:
            tz = self.Z1  # rows [0, U)
        KU = K * U
        seg_t = self._const_seg(KU)
        if self.hs_fused and KU % 32 == 0:
            # eval-mode BatchNorm -> ReLU folded into the next Linear's operand staging (hidden_ops.cu), one launch per layer
            seg1 = self._seg_teacher.ctypes.data
            for a in range(1, self.n_app):
                rm, rv = tbn(a - 1)
                k, n = self.widths[a - 1], self.widths[a]
                call("scnb_hs_fwd", ptr(tz), None, None, None, ptr(tw(blk[a - 1]["g"])), ptr(tw(blk[a - 1]["beta"])), None, 0.0,
                     ptr(tw(blk[a]["W"])), ptr(tw(blk[a]["b"])), ptr(self.TZ[a]), None, None, None, 0, 0.0, k, n, KU, 1, seg1, ptr(rm),
                     ptr(rv), st)
                tz = self.TZ[a]
            a0 = self.n_app - 1
        else:
            a0 = 0
        pl_args = (float(c.T if c.T is not None else 1.0), int(c.T is not None), float(c.pseudolabel_min_confidence))
        if (self.hs_fused and KU % 32 == 0 and self.fuse_classif and K == 1 and C <= 32 and self.H % 128 == 0
                and os.environ.get("SCNB_TEACHER_HEAD", "bn") == "bn"):
            # last eval-mode BatchNorm -> ReLU applied by the head kernel as it loads a row: one launch less on the teacher's
            # chain, which joins the student in front of the classifier head
            last = self.n_app - 1
            rm, rv = tbn(last)
            call("scnb_teacher_head_bn_fused", ptr(tz), ptr(tw(blk[last]["g"])), ptr(tw(blk[last]["beta"])), ptr(rm), ptr(rv),
                 ptr(tw("model.classif.0.weight")), ptr(tw("model.classif.0.bias")), self.H, C, K, U, pl_args[0], pl_args[1], pl_args[2],
                 ptr(self.t_logits), ptr(self.Ycat), ptr(self.conf), ptr(self.confident), ptr(self.pl_scratch), ptr(self.lab_ids), L, st)
            return
        for a in range(a0, self.n_app):
            rm, rv = tbn(a)
            call("scnb_bn_act_fwd", ptr(tz), ptr(self.TA[a]), None, self.widths[a], 1, ptr(seg_t), KU, 0, ptr(tw(blk[a]["g"])),
                 ptr(tw(blk[a]["beta"])), ptr(rm), ptr(rv), None, None, None, 0, 0.0, 1, None, None, st)
            if a + 1 < self.n_app:
                call("scnb_sgemm", 0, 1, KU, self.widths[a + 1], self.widths[a], ptr(self.TA[a]), self.widths[a],
                     ptr(tw(blk[a + 1]["W"])), self.widths[a], ptr(self.TZ[a + 1]), self.widths[a + 1], 1.0, 0.0,
                     ptr(tw(blk[a + 1]["b"])), 0, None, st)
                tz = self.TZ[a + 1]
        pl_args = (float(c.T if c.T is not None else 1.0), int(c.T is not None), float(c.pseudolabel_min_confidence))
        if self.fuse_classif:
            # teacher classifier + pseudolabels q[0,U) + one-hot labels q[U,U+L) in one launch
            call("scnb_teacher_head_fused", ptr(self.TA[-1]), ptr(tw("model.classif.0.weight")), ptr(tw("model.classif.0.bias")),
                 self.H, C, K, U, pl_args[0], pl_args[1], pl_args[2], ptr(self.t_logits), ptr(self.Ycat), ptr(self.conf),
                 ptr(self.confident), ptr(self.pl_scratch), ptr(self.lab_ids), L, st)
        else:
            call("scnb_sgemm", 0, 1, KU, C, self.H, ptr(self.TA[-1]), self.H, ptr(tw("model.classif.0.weight")), self.H,
                 ptr(self.t_logits), C, 1.0, 0.0, ptr(tw("model.classif.0.bias")), 0, None, st)
            call("scnb_pseudolabel", ptr(self.t_logits), K, U, C, pl_args[0], pl_args[1], pl_args[2], ptr(self.Ycat),
                 ptr(self.conf), ptr(self.confident), ptr(self.pl_scratch), ptr(self.lab_ids), L, st)

    def _view(self, buf, rows, width):
        """[rows, width] contiguous view at the start of a scratch buffer."""
        return buf.view(-1)[: rows * width].view(rows, width)

    def _const_seg(self, n):
        key = ("seg", n)
        if not hasattr(self, "_cseg"):
            self._cseg = {}
        if key not in self._cseg:
            self._cseg[key] = torch.tensor([0, n], dtype=torch.int32, device=self.dev)
        return self._cseg[key]

    def _dan_names(self):
        names = [n for n in self.arena.names if n.startswith("dan.") and n.endswith(".weight")]
        return names[0][: -len(".weight")], names[1][: -len(".weight")]

    def _hidden_keep(self, a):
        return ptr(self.inj_hidden_keep[a]) if self.inj_hidden_keep is not None else None

    def _xchg(self, kind, a, st):
        """Data parallel: make the statistics of application `a` global (scnb_hs_xchg over peer memory).  kind 0: forward tile
        statistics pf[a] -> stats[a] (+ global row counts for the running-statistics update); kind 1: backward tile sums pb[a]
        -> bmean[a], this rank's local sums into the BatchNorm parameter gradients."""
        px, blk = self.peer, self._blk
        w = self.widths[a]
        if kind == 0:
            call("scnb_hs_xchg", 0, ptr(self.pf[a]), w, self.Rmax, self.n_seg, self._seg_host.ctypes.data, ptr(px.ptrs), px.rank, px.world,
                 a, self.peer_slots, max(self.widths), ptr(self.state), self.peer_bn_bytes, ptr(self.stats[a]),
                 ptr(self.dp_fwd[a][self.n_seg * 2 * w:]), None, None, st)
        else:
            call("scnb_hs_xchg", 1, ptr(self.pb[a]), w, self.Rmax, self.n_seg, self._seg_host.ctypes.data, ptr(px.ptrs), px.rank, px.world,
                 self.n_app + a, self.peer_slots, max(self.widths), ptr(self.state), self.peer_bn_bytes, ptr(self.bmean[a]), None,
                 self.Gp(blk[a]["g"]), self.Gp(blk[a]["beta"]), st)

    def _student_forward(self, st, rng):
        blk, R = self._blk, self.Rmax
        if self.hs_fused:
            seg = self._seg_host.ctypes.data
            dp = self.comm is not None
            inj = self.inj_hidden_keep is not None
            pd = [float(b["drop"].p) for b in blk]
            hf = bool(getattr(self, "heads_fused", False))
            keep_out = lambda a: (ptr(self.keepbuf[a]) if ((a < self.n_app - 1 or hf) and not inj and pd[a] > 0.0) else None)
            keep_in = lambda a: (self._hidden_keep(a) if inj else (ptr(self.keepbuf[a]) if pd[a] > 0.0 else None))
            if self.mode != "mixmatch":   # no MixUp in front: tile statistics (and the dropout keep bytes) of the first layer's output
                call("scnb_hs_mix", ptr(self.Zs[0]), self.widths[0], None, None, None, R, ptr(self.Zs[0]), ptr(self.pf[0]), keep_out(0), rng,
                     SID_HID, pd[0], st)
            for a in range(1, self.n_app):
                k, n = self.widths[a - 1], self.widths[a]
                if dp:
                    self._xchg(0, a - 1, st)
                call("scnb_hs_fwd", ptr(self.Zs[a - 1]), ptr(self.As[a - 1]), None if dp else ptr(self.pf[a - 1]), ptr(self.stats[a - 1]),
                     self.P(blk[a - 1]["g"]), self.P(blk[a - 1]["beta"]), keep_in(a - 1), pd[a - 1], self.P(blk[a]["W"]), self.P(blk[a]["b"]),
                     ptr(self.Zs[a]), ptr(self.pf[a]), keep_out(a), rng, SID_HID + a, pd[a], k, n, R, self.n_seg, seg, None, None, st)
            a = self.n_app - 1
            if dp:
                self._xchg(0, a, st)
            if hf:
                return          # the head kernels apply the last BatchNorm -> Dropout -> ReLU themselves
            call("scnb_hs_act", ptr(self.Zs[a]), ptr(self.As[a]), self.widths[a], R, self.n_seg, seg, None if dp else ptr(self.pf[a]), self.P(blk[a]["g"]),
                 self.P(blk[a]["beta"]), ptr(self.stats[a]), self._hidden_keep(a), rng, SID_HID + a, float(blk[a]["drop"].p), st)
            return
        for a in range(self.n_app):
            w = self.widths[a]
            p = float(blk[a]["drop"].p)
            fwd = lambda part, sums: call(
                "scnb_bn_act_fwd", ptr(self.Zs[a]), ptr(self.As[a]), None, w, self.n_seg, ptr(self.seg_off), R, 1,
                self.P(blk[a]["g"]), self.P(blk[a]["beta"]), None, None, ptr(self.stats[a]), self._hidden_keep(a), rng,
                SID_HID + a, p, 1, part, sums, st)
            if self.comm is None:
                fwd(None, None)
            elif self.peer is not None:   # one launch: statistics all-reduced inside the kernel over peer memory
                px = self.peer
                call("scnb_bn_act_fwd_peer", ptr(self.Zs[a]), ptr(self.As[a]), None, w, self.n_seg, ptr(self.seg_off), R,
                     self.P(blk[a]["g"]), self.P(blk[a]["beta"]), ptr(self.stats[a]), self._hidden_keep(a), rng, SID_HID + a, p, 1,
                     ptr(px.ptrs), px.rank, px.world, a, self.peer_slots, ptr(self.state), self.peer_bn_bytes, ptr(self.dp_fwd[a]), st)
            else:  # cross-rank BatchNorm statistics: local sums -> all-reduce -> apply
                fwd(ptr(self.dp_fwd[a]), None)
                self.comm.all_reduce_sum(self.dp_fwd[a])
                fwd(None, ptr(self.dp_fwd[a]))
            if a + 1 < self.n_app:
                w2 = self.widths[a + 1]
                call("scnb_sgemm", 0, 1, R, w2, w, ptr(self.As[a]), w, self.P(blk[a + 1]["W"]), w, ptr(self.Zs[a + 1]), w2, 1.0,
                     0.0, self.P(blk[a + 1]["b"]), 0, None, st)

    def _student_backward(self, main, sW, rng, dA):
        """dA: gradient w.r.t. As[-1].  Returns dZ of the first application ([R, H0]).
        The dA chain runs on `main`; weight/bias gradients of the hidden Linear layers on `sW`."""
        blk, R = self._blk, self.Rmax
        st, sw = main.cuda_stream, sW.cuda_stream
        if self.hs_fused:
            return self._student_backward_fused(main, sW, dA)
        for a in range(self.n_app - 1, -1, -1):
            w = self.widths[a]
            p = float(blk[a]["drop"].p)
            dZ = self.dZs[a]
            bwd = lambda part, sums: call(
                "scnb_bn_act_bwd", ptr(dA), ptr(self.Zs[a]), ptr(self.As[a]), w, self.n_seg, ptr(self.seg_off), R,
                self.P(blk[a]["g"]), ptr(self.stats[a]), self._hidden_keep(a), rng, SID_HID + a, p, 1, ptr(dZ),
                self.Gp(blk[a]["g"]), self.Gp(blk[a]["beta"]), part, sums, st)
            if a == self.n_app - 1 and sW is not main:
                if not getattr(self, "_zero_at_end", False):
                    main.wait_event(self._ev_zero)   # BatchNorm gamma/beta gradients accumulate into zeroed buffers
            if self.comm is None:
                bwd(None, None)
            elif self.peer is not None:
                px = self.peer
                call("scnb_bn_act_bwd_peer", ptr(dA), ptr(self.Zs[a]), ptr(self.As[a]), w, self.n_seg, ptr(self.seg_off), R,
                     self.P(blk[a]["g"]), ptr(self.stats[a]), self._hidden_keep(a), rng, SID_HID + a, p, 1, ptr(dZ),
                     self.Gp(blk[a]["g"]), self.Gp(blk[a]["beta"]), ptr(px.ptrs), px.rank, px.world, self.n_app + a, self.peer_slots,
                     ptr(self.state), self.peer_bn_bytes, st)
            else:
                bwd(ptr(self.dp_bwd[a]), None)
                self.comm.all_reduce_sum(self.dp_bwd[a])
                bwd(None, ptr(self.dp_bwd[a]))
            if a == 0:
                return dZ
            wp = self.widths[a - 1]
            dA = self._view(self.dA, R, wp)
            # the weight / bias gradients need dZ only: fork branch W here, BEFORE the dA product joins the main chain, so
            # that the next BatchNorm backward (which waits for dA) is the only kernel released when that product ends
            # (the graph releases the children of a node one after the other, ~5 us apart: the dA product is issued first
            # and is the only child of its parent on the critical chain)
            ev = self._ev_dz[a]
            # Large products (tcgen05, one 197 KB CTA per SM): two of them cannot share an SM, so the weight-gradient product
            # is forked BEHIND the input-gradient product -- it then runs next to the HBM-bound BatchNorm backward of the
            # layer below instead of fighting the chain's own product for the SMs
            fork = os.environ.get("SCNB_WGRAD_FORK", "late") if self._gemm_ws else "early"
            late_fork = fork == "late"
            # the LAST weight-gradient product (a == 1) has nothing left on the chain to hide behind except the dW1 kernel,
            # which takes every SM: on the side branch it ran AFTER that kernel, a tail of ~120 us.  It goes on the chain in
            # front of the dW1 kernel (SCNB_WGRAD_LAST=side restores the branch; 2.69 -> 2.66 ms)
            if a == 1 and fork == "late" and os.environ.get("SCNB_WGRAD_LAST", "chain") == "chain":
                fork = "chain"
            if fork == "chain" and sW is not main:
                # (experiment) no fork at all: the weight-gradient product follows the input-gradient product on the chain
                call("scnb_sgemm", 0, 0, R, wp, w, ptr(dZ), w, self.P(blk[a]["W"]), wp, ptr(dA), wp, 1.0, 0.0, None, 0, None, st)
                call("scnb_sgemm", 1, 0, w, wp, R, ptr(dZ), w, ptr(self.As[a - 1]), wp, self.Gp(blk[a]["W"]), wp, 1.0, 1.0, None, 0,
                     None, st)
                call("scnb_colsum", ptr(dZ), R, w, w, self.Gp(blk[a]["b"]), 1, st)
                continue
            if sW is not main and not late_fork:
                ev.record(main)
            call("scnb_sgemm", 0, 0, R, wp, w, ptr(dZ), w, self.P(blk[a]["W"]), wp, ptr(dA), wp, 1.0, 0.0, None, 0, None, st)
            if sW is not main and late_fork:
                ev.record(main)
            if sW is not main:
                sW.wait_event(ev)
            call("scnb_sgemm", 1, 0, w, wp, R, ptr(dZ), w, ptr(self.As[a - 1]), wp, self.Gp(blk[a]["W"]), wp, 1.0, 1.0, None, 0,
                 None, sw)
            call("scnb_colsum", ptr(dZ), R, w, w, self.Gp(blk[a]["b"]), 1, sw)
        return None

    def _student_backward_fused(self, main, sW, dA):
        """Backward of the fused hidden stack (hidden_ops.cu).  The dA chain stays on `main`; the Linear weight / bias
        gradients (from the materialised dZ[a] and A[a-1]) go to branch W."""
        blk, R = self._blk, self.Rmax
        st, sw = main.cuda_stream, sW.cuda_stream
        seg = self._seg_host.ctypes.data
        last = self.n_app - 1
        if sW is not main:
            if not getattr(self, "_zero_at_end", False):
                main.wait_event(self._ev_zero)   # BatchNorm gamma / beta gradients accumulate into zeroed buffers
        dp = self.comm is not None
        if not getattr(self, "heads_fused", False):
            call("scnb_hs_bwd_act", ptr(dA), ptr(self.Zs[last]), ptr(self.As[last]), self.widths[last], R, self.n_seg, seg,
                 ptr(self.stats[last]), float(blk[last]["drop"].p), ptr(self.dYs[last]), ptr(self.pb[last]), st)
        for a in range(last, 0, -1):
            k, n = self.widths[a], self.widths[a - 1]
            if dp:
                self._xchg(1, a, st)
            call("scnb_hs_bwd", ptr(self.dYs[a]), ptr(self.Zs[a]), ptr(self.stats[a]), None if dp else ptr(self.pb[a]), self.P(blk[a]["g"]),
                 self.Gp(blk[a]["g"]), self.Gp(blk[a]["beta"]), ptr(self.dZs[a]), self.P(blk[a]["W"]), ptr(self.Zs[a - 1]),
                 ptr(self.As[a - 1]), ptr(self.stats[a - 1]), float(blk[a - 1]["drop"].p), ptr(self.dYs[a - 1]), ptr(self.pb[a - 1]),
                 k, n, R, self.n_seg, seg, ptr(self.bmean[a]) if dp else None, st)
            ev = self._ev_dz[a]
            if sW is not main:
                ev.record(main)
                sW.wait_event(ev)
            call("scnb_hs_dw", ptr(self.dZs[a]), k, ptr(self.As[a - 1]), n, R, self.Gp(blk[a]["W"]), self.Gp(blk[a]["b"]), sw)
        if dp:
            self._xchg(1, 0, st)
        if self.mode == "mixmatch" and self.w1_tc and os.environ.get("SCNB_BWD_IN", "split") == "fused":
            # BatchNorm backward of application 0 + MixUp adjoint + bf16 operand planes of dZ1 in one launch (head_ops.cu).
            # Measured slower than the two launches it replaces (16.4 us against 4.2 + 6.7: one column per thread is a chain of
            # ~8 dependent round trips), so it is an off-by-default switch (SCNB_BWD_IN=fused)
            call("scnb_hs_bwd_in_unmix_planes", ptr(self.dYs[0]), ptr(self.Zs[0]), self.widths[0], R, self.n_seg, seg, ptr(self.stats[0]),
                 None if dp else ptr(self.pb[0]), self.P(blk[0]["g"]), self.Gp(blk[0]["g"]), self.Gp(blk[0]["beta"]),
                 ptr(self.bmean[0]) if dp else None, self.Gp(blk[0]["b"]), ptr(self.inv3), ptr(self.coef3), self.nb, ptr(self.dZ1),
                 ptr(self.dz_hi), ptr(self.dz_lo), st)
            return None     # dZ1 and its planes are written
        call("scnb_hs_bwd_in", ptr(self.dYs[0]), ptr(self.Zs[0]), self.widths[0], R, self.n_seg, seg, ptr(self.stats[0]),
             None if dp else ptr(self.pb[0]), self.P(blk[0]["g"]), self.Gp(blk[0]["g"]), self.Gp(blk[0]["beta"]), ptr(self.dZs[0]),
             ptr(self.bmean[0]) if dp else None, self.Gp(blk[0]["b"]), st)
        return self.dZs[0]

    def _gene_cnt_ptr(self):
        return ptr(self.gene_cnt) if (self.fused_w1 and not self.w1_tc) else None

    def _build_csc(self, n_rows, st, idx=None):
        """Inputs of the first-layer backward that depend only on the batch: per-row gene-block split points
        (tensor-core kernel) or the gene-major copy of the batch (SIMT kernel)."""
        if self.w1_tc:
            # (idx: the staged rows of a host-mapped batch -- same entries, same positions as b_idx)
            if n_rows == self.nb and self.fwd_splits is not None:
                call("scnb_row_gene_splits2", ptr(self.b_indptr), ptr(self.b_idx if idx is None else idx), n_rows, self.G,
                     self.widths[0], ptr(self.w1_splits), ptr(self.fwd_splits), st)
            else:
                call("scnb_row_gene_splits", ptr(self.b_indptr), ptr(self.b_idx if idx is None else idx), n_rows, self.G,
                     ptr(self.w1_splits), st)
        elif self.fused_w1:
            call("scnb_csr_to_csc", ptr(self.b_indptr), ptr(self.b_idx), ptr(self.b_val), n_rows, self.G, ptr(self.gene_cnt),
                 ptr(self.c_ptr), ptr(self.c_cursor), ptr(self.c_ent), st)

    def _zero_grads(self, st, also=None):
        """optimizer.zero_grad() (trainer.py:608) for what the step accumulates into: the gradient arena (behind the first
        layer when the fused update never materialises dW1, or overwrites it) and, optionally, one more buffer."""
        g = self.arena.grad[self.P1:] if self.fused_w1 else self.arena.grad
        n2 = 0 if also is None else also.numel()
        call("scnb_zero2", ptr(g), g.numel(), ptr(also), n2, st)

    def _finish_step(self, main, sW, sD, dZ1, n_rows):
        """First-layer weight gradient, optimizer.step() and BatchNorm running statistics in the
        reference's update order.  dZ1: gradient w.r.t. the first Linear's output, rows aligned with
        the batch CSR.  main: fused dW1 + update of W1T; branch W: everything else."""
        c = self.cfg
        H0 = self.widths[0]
        blk0 = self._blk[0]
        W1 = blk0["W"]
        opt = OPTIMIZER_CODES[c.optimizer]
        hp, s4 = ptr(self.hp), ptr(self.state)
        ar = self.arena
        st, sw = main.cuda_stream, sW.cuda_stream
        self._after(sW, main)     # dZ1 ready
        self._after(sW, sD)       # domain-head gradients ready
        if not self.hs_fused:   # (fused hidden stack: db1 = column sums of dZ[0] over all passes, added by scnb_hs_bwd_in)
            call("scnb_colsum", ptr(dZ1), n_rows, H0, H0, self.Gp(blk0["b"]), 1, sw)
        if sW is not main:
            main.wait_event(self._ev_csc)   # batch CSC (built at the start of the step on branch W)
        if not self.fused_w1:
            self._after(main, sW)   # every other gradient is complete before the one optimizer pass
            call("scnb_csr_linear_bwd_dw", ptr(self.b_indptr), ptr(self.b_idx), ptr(self.b_val), n_rows, ptr(dZ1), H0, self.Gp(W1), st)
            if self.comm is not None:
                self.comm.all_reduce_avg(ar.grad)
            call("scnb_optim_step", opt, ptr(ar.flat), ptr(ar.grad), ptr(ar.state1), ptr(ar.state2), ar.total, hp, s4, st)
        elif self.comm is not None:
            # data parallel: gradient only (no atomics, every gene row written), exchange, then one update pass
            push = bool(self.w1_tc and self.peer is not None and 256 % H0 == 0 and os.environ.get("SCNB_DP_PUSH", "1") != "0")
            if self.w1_tc:
                if not self._dz_planes_ready:
                    call("scnb_w1_planes", ptr(dZ1), n_rows, H0, ptr(self.dz_hi), ptr(self.dz_lo), st)
                if push:   # gradient rows go straight to their owners' inboxes while the kernel runs
                    px = self.peer
                    call("scnb_csr_w1_bwd_push_tc", ptr(self.w1_splits), ptr(self.b_idx), ptr(self.b_val), n_rows, H0, self.G,
                         ptr(self.dz_hi), ptr(self.dz_lo), ptr(px.ptrs), px.rank, px.world, self.peer_off_inbox, ar.total,
                         self.Gp(W1) if c.keep_w1_grad else None, st)
                else:
                    call("scnb_csr_w1_bwd_optim_tc", -1, ptr(self.w1_splits), ptr(self.b_idx), ptr(self.b_val), n_rows, H0, self.G,
                         ptr(self.dz_hi), ptr(self.dz_lo), None, self.Gp(W1), None, None, None, None, None, None, st)
            else:
                call("scnb_csr_w1_bwd_optim", -1, ptr(self.c_ptr), ptr(self.c_ent), ptr(dZ1), n_rows, H0, self.G, None, self.Gp(W1),
                     None, None, None, None, st)
            self._after(main, sW)              # every other gradient (branch W, and branch D through it) is complete
            if self.peer is not None:
                # reduce-scatter (peer loads) + optimizer on the owned slice + all-gather (peer stores), one kernel
                px = self.peer
                call("scnb_dp_reduce_optim_bcast2", opt, ptr(px.ptrs), px.rank, px.world, self.peer_off_flags, self.peer_off_grad,
                     self.peer_off_flat, self.peer_off_inbox if push else -1, self.P1 if push else 0, ar.total, ptr(ar.state1),
                     ptr(ar.state2), hp, s4, s4, st)
            else:
                self.comm.all_reduce_avg(ar.grad)  # NCCL over NVLink
                call("scnb_optim_step", opt, ptr(ar.flat), ptr(ar.grad), ptr(ar.state1), ptr(ar.state2), ar.total, hp, s4, st)
        else:
            keep = self.Gp(W1) if c.keep_w1_grad else None
            if self.w1_tc:
                if not self._dz_planes_ready:
                    call("scnb_w1_planes", ptr(dZ1), n_rows, H0, ptr(self.dz_hi), ptr(self.dz_lo), st)
                call("scnb_csr_w1_bwd_optim_tc", opt, ptr(self.w1_splits), ptr(self.b_idx), ptr(self.b_val), n_rows, H0, self.G,
                     ptr(self.dz_hi), ptr(self.dz_lo), ptr(ar.flat), keep, ptr(ar.state1), ptr(ar.state2), hp, s4,
                     ptr(self.w1_hi) if self.w1_emits_planes else None, ptr(self.w1_lo) if self.w1_emits_planes else None, st)
            else:
                call("scnb_csr_w1_bwd_optim", opt, ptr(self.c_ptr), ptr(self.c_ent), ptr(dZ1), n_rows, H0, self.G, ptr(ar.flat), keep,
                     ptr(ar.state1), ptr(ar.state2), hp, s4, st)
            rest = ar.total - self.P1
            call("scnb_optim_step", opt, ptr(ar.flat[self.P1:]), ptr(ar.grad[self.P1:]), ptr(ar.state1[self.P1:]),
                 ptr(ar.state2[self.P1:]), rest, hp, s4, sw)
        if not self._bn_running_done:
            self._bn_running_update(sw)
        if self.loss_hist is not None:
            mm = self.mode == "mixmatch"
            call("scnb_record_step", ptr(self.loss_buf), ptr(self.loss_hist), int(self.loss_hist.shape[0]),
                 ptr(self.conf) if mm else None, self.U if mm else 0, ptr(self.conf_ring) if mm else None,
                 int(self.conf_ring.shape[0]) if mm else 0, ptr(self.state), sw)
        if self.mode == "mixmatch":
            call("scnb_mm_step_inc", ptr(self.state), sw)
        if getattr(self, "_zero_at_end", False):    # for the NEXT step (every consumer of this step's gradients is behind us on sW)
            self._zero_grads(sw, also=self.Z1)
        self._after(main, sW)
        self._after(main, sD)
        if self._bn_running_done:
            main.wait_event(self._ev_bnrun)

    def _bn_running_update(self, sw):
        """BatchNorm running statistics (momentum update over the step's forward passes, in the reference's order)."""
        done = set()
        for a in range(self.n_app):
            bn = self._blk[a]["bn"]
            if id(bn) in done:
                continue
            done.add(id(bn))
            apps = [b for b in range(self.n_app) if self._blk[b]["bn"] is bn]
            if len(apps) > 4:
                raise NotImplementedError("a BatchNorm module shared by more than 4 applications (n_layers > 5)")
            sp = [ptr(self.stats[b]) for b in apps] + [None] * (4 - len(apps))
            call("scnb_bn_running_update", ptr(bn.running_mean), ptr(bn.running_var), ptr(bn.num_batches_tracked), sp[0], sp[1],
                 sp[2], sp[3], ptr(self.seg_off), self.n_seg, self.widths[a], float(bn.momentum),
                 ptr(self.dp_fwd[a][self.n_seg * 2 * self.widths[a]:]) if self.comm is not None else None, sw)

    def enable_history(self, capacity: int, ring: int = 64):
        """Keep per-step loss scalars (and the last `ring` steps' pseudolabel confidences) on the
        device; `history()` reads them back once (epoch end) instead of the reference's per-step
        `.item()` / `.cpu()` synchronisations (trainer.py:666-681, losses.py:714-725)."""
        if self.loss_hist is None or self.loss_hist.shape[0] != capacity:
            self.loss_hist = torch.zeros((capacity, 8), dtype=torch.float32, device=self.dev)
            self.conf_ring = torch.zeros((ring, max(self.U, 1)), dtype=torch.float32, device=self.dev)
            self._graph = None
            self._pf_graphs = {}
        self._hist_t0 = int(self.steps_done)

    def history(self):
        """(losses [n,8] numpy in step order since enable_history, confidences [(<=ring), U])."""
        torch.cuda.synchronize()
        n = int(self.steps_done) - self._hist_t0
        cap = self.loss_hist.shape[0]
        t_first = int(self.state[0].item()) - n
        idx = [(t_first + i) % cap for i in range(n)]
        h = self.loss_hist.cpu().numpy()[idx]
        ring = self.conf_ring.shape[0]
        m = min(n, ring)
        t_end = int(self.state[0].item())
        ridx = [(t_end - m + i) % ring for i in range(m)]
        return h, self.conf_ring.cpu().numpy()[ridx]

    # ------------------------------------------------------------------------------------
    def _launch_supervised(self):
        """Trainer.train_epoch body (trainer.py:189-256): raw batch -> model -> CE -> backward -> step."""
        main, sW, _ = self._streams()
        sD = main
        st = main.cuda_stream
        L, C, H0 = self.L, self.C, self.widths[0]
        rng = ptr(self.state[1:3])
        blk = self._blk
        self._refresh_planes(main, sW)
        self._prep(st)
        call("scnb_csr_gather_rows", ptr(self.lab.indptr), ptr(self.lab.indices), ptr(self.lab.data), ptr(self.l_rows), 0, L,
             ptr(self.b_indptr), ptr(self.b_idx), ptr(self.b_val), 0, 0, 0, None, None, 0, 0.0, self._gene_cnt_ptr(), st)
        self._after(sW, main)
        self._zero_grads(sW.cuda_stream)
        if sW is not main:
            self._ev_zero.record(sW)
        self._build_csc(L, sW.cuda_stream)
        if sW is not main:
            self._ev_csc.record(sW)
        self._first_layer(main, sW, L, self.Zs[0])
        self._student_forward(st, rng)
        Hh = self.H
        dA_last = self._view(self.dA, L, Hh)
        call("scnb_onehot_rows", ptr(self.lab_ids), None, 0, L, C, ptr(self.Ycat), st)
        if self.fuse_classif:
            call("scnb_classif_supervised_fused", ptr(self.As[-1]), self.P("model.classif.0.weight"), self.P("model.classif.0.bias"),
                 Hh, C, L, ptr(self.Ycat), ptr(self.class_weight), ptr(self.logits), ptr(self.dlogits), ptr(dA_last),
                 ptr(self.loss_buf), st)
        else:
            call("scnb_sgemm", 0, 1, L, C, self.H, ptr(self.As[-1]), self.H, self.P("model.classif.0.weight"), self.H,
                 ptr(self.logits), C, 1.0, 0.0, self.P("model.classif.0.bias"), 0, None, st)
            call("scnb_soft_ce_fwd_bwd", ptr(self.logits), L, C, None, ptr(self.Ycat), None, None, None, 0, ptr(self.class_weight),
                 None, L, 1.0, None, ptr(self.dlogits), ptr(self.loss_buf), 1, None, st)
            call("scnb_sgemm", 0, 0, L, Hh, C, ptr(self.dlogits), C, self.P("model.classif.0.weight"), Hh, ptr(dA_last), Hh, 1.0, 0.0,
                 None, 0, None, st)
        self._after(sW, main)
        sw = sW.cuda_stream
        call("scnb_sgemm", 1, 0, C, Hh, L, ptr(self.dlogits), C, ptr(self.As[-1]), Hh, self.Gp("model.classif.0.weight"), Hh, 1.0,
             1.0, None, 0, None, sw)
        call("scnb_colsum", ptr(self.dlogits), L, C, C, self.Gp("model.classif.0.bias"), 1, sw)
        dZ1 = self._student_backward(main, sW, rng, dA_last)
        self._finish_step(main, sW, sD, dZ1, L)

    # ------------------------------------------------------------------------------------
    def _launch(self):
        if self.cfg.multi_stream and os.environ.get("SCNB_MAIN_PRIORITY", "1") != "0":
            # the critical chain runs on a high-priority stream: when its next kernel and a side branch's (weight gradients,
            # running statistics, teacher) both have CTAs pending, the chain's are placed first (the side streams keep the
            # default, lowest priority); captured graphs keep the priority as a kernel-node attribute
            if getattr(self, "_hi", None) is None:
                self._hi = torch.cuda.Stream(device=self.dev, priority=-1)
            cur = torch.cuda.current_stream()
            self._hi.wait_stream(cur)
            with torch.cuda.stream(self._hi):
                self._launch_inner()
            cur.wait_stream(self._hi)
        else:
            self._launch_inner()

    def _launch_inner(self):
        if self.mode == "mixmatch":
            self._launch_mixmatch()
        else:
            self._launch_supervised()
        if self.post_launch is not None and not getattr(self, "_published", False):
            self.post_launch(self)           # a feeding pipeline's read-back, inside the same graph (set before any capture)
        self._published = False

    def _fetch(self, st):
        if self._tabs is None:
            return
        t = self._tabs
        call("scnb_fetch_step", ptr(t["l"]), ptr(t.get("u")), ptr(t.get("k")), ptr(t.get("g")), self.L, self.U, t["n"],
             ptr(self.state), ptr(self.l_rows), ptr(self.u_rows), ptr(self.perm_keys), ptr(self.gamma_raw), st)

    def set_epoch_tables(self, l_tab, u_tab=None, key_tab=None, gam_tab=None):
        """Device-resident sampler: row ids (and MixUp randomness) for `n` consecutive steps.
        Step s reads row `(steps_taken % n)`; tables are copied into persistent buffers so a
        captured graph stays valid across epochs (same n) -- replaces DataLoader iteration
        (scnym/trainer.py:574-582) and the `repeater` around the unlabeled loader (main.py:48-68)."""
        n = int(l_tab.shape[0])
        mk = lambda t, dt: None if t is None else torch.as_tensor(t).to(device=self.dev, dtype=dt).contiguous()
        new = dict(n=n, l=mk(l_tab, torch.int64), u=mk(u_tab, torch.int64), k=mk(key_tab, torch.float32),
                   g=mk(gam_tab, torch.float32))
        if self.mode == "mixmatch" and (new["u"] is None or new["k"] is None or new["g"] is None):
            raise ValueError("mixmatch tables need unlabeled rows, permutation keys and gamma draws")
        old = self._tabs
        if old is not None and old["n"] == n and all((old[k] is None) == (new[k] is None) for k in "lukg"):
            for k in "lukg":
                if new[k] is not None:
                    old[k].copy_(new[k])
        else:
            self._tabs = new
            self._graph = None  # pointers changed: re-capture
            self._pf_graphs = {}
        self._prefetch_valid = False   # a batch assembled ahead of time came from the old tables
        self.invalidate_planes()   # once per epoch: one cheap plane rebuild guards against unseen writes to W1

    def sample_epoch_tables(self, n_steps: int, generator: Optional[torch.Generator] = None):
        """Draw one epoch of batches on the device: shuffled labeled rows without replacement
        (DataLoader(shuffle=True, drop_last=True), main.py:317-323), unlabeled rows likewise with
        wrap-around, uniform permutation keys and Beta(alpha, alpha) MixUp draws."""
        dev, L, U = self.dev, self.L, self.U
        def draw(n_rows, per):
            need = n_steps * per
            chunks, have = [], 0
            while have < need:
                chunks.append(torch.randperm(n_rows, device=dev, generator=generator))
                have += n_rows
            flat = torch.cat(chunks)
            # drop_last semantics inside each pass over the data
            usable = (n_rows // per) * per
            if usable >= need or usable == 0:
                return flat[:need].view(n_steps, per)
            out = torch.cat([c[:usable] for c in chunks])
            while out.numel() < need:
                out = torch.cat([out, torch.randperm(n_rows, device=dev, generator=generator)[:usable]])
            return out[:need].view(n_steps, per)
        l_tab = draw(self.lab.n_rows, L)
        if self.mode != "mixmatch":
            self.set_epoch_tables(l_tab)
            return
        u_tab = draw(self.unl.n_rows, U)
        keys = torch.rand((n_steps, L + U), device=dev, generator=generator)
        a = max(float(self.cfg.mixup_alpha), 1e-3)
        conc = torch.full((n_steps, L + U), a, device=dev)
        g1 = torch._standard_gamma(conc, generator=generator)
        g2 = torch._standard_gamma(conc, generator=generator)
        gam = g1 / (g1 + g2).clamp_min(1e-30)      # Beta(a, a) as a ratio of Gammas
        self.set_epoch_tables(l_tab, u_tab, keys, gam)

    def load_batch(self, l_rows, u_rows=None, perm_keys=None, gamma_raw=None):
        """Stage one step's sampled rows / MixUp randomness into the device input buffers."""
        self.l_rows.copy_(torch.as_tensor(l_rows, dtype=torch.int64), non_blocking=True)
        if self.mode == "mixmatch":
            self.u_rows.copy_(torch.as_tensor(u_rows, dtype=torch.int64), non_blocking=True)
            if perm_keys is not None:
                self.perm_keys.copy_(torch.as_tensor(perm_keys, dtype=torch.float32), non_blocking=True)
            if gamma_raw is not None:
                self.gamma_raw.copy_(torch.as_tensor(gamma_raw, dtype=torch.float32), non_blocking=True)

    def step(self):
        """Run one training step on whatever is staged in the input buffers."""
        if self.mode == "mixmatch" and self.cfg.mean_teacher and self.teacher_bn is None:
            self.teacher_bn = [(b["bn"].running_mean.clone(), b["bn"].running_var.clone()) for b in self._blk]
        injected = any(x is not None for x in (self.inj_in_keep_L, self.inj_in_keep_U, self.inj_hidden_keep, self.inj_pois_L,
                                               self.inj_rowu_L, self.inj_pois_U, self.inj_rowu_U))
        self._sync_planes()
        graph = self.cfg.use_cuda_graph and not injected
        if graph and self.prefetch_ok and self._tabs is not None and self.cfg.multi_stream:
            # replayed steps on device-resident sampler tables: the batch of step t+1 is assembled while step t runs
            p = self._next_p
            self.__dict__["_cur"] = p
            if not self._prefetch_valid:   # first step, or the tables changed: assemble this step's batch now
                self._prologue_next(torch.cuda.current_stream(), ptr(self.state[1:3]))
                if self.tc_first:          # (a replayed step leaves Z1 / the gradient arena cleared for its successor)
                    self._zero_grads(torch.cuda.current_stream().cuda_stream, also=self.Z1)
            key = (self.unsup_weight, self.dan_weight, p)
            if self._pf_graphs and next(iter(self._pf_graphs))[:2] != key[:2]:
                self._pf_graphs = {}      # the loss weights are launch arguments: a new epoch's weights retire the old graphs
            if key not in self._pf_graphs:
                self._prefetched = True
                try:
                    self._pf_graphs[key] = self._capture_graph()
                finally:
                    self._prefetched = False
            self._pf_graphs[key].replay()
            self._next_p, self._prefetch_valid = 1 - p, True
        elif graph:
            self._prefetch_valid = False
            key = (self.unsup_weight, self.dan_weight)
            if self._graph is None or self._graph_key != key:
                self._capture(key)
            self._graph.replay()
        else:
            self._prefetch_valid = False
            before = _lib.launch_count
            self._launch()
            self.kernels_per_step = _lib.launch_count - before
        self.steps_done += 1

    def prepare(self):
        """Capture the step graph for the current loss weights without running a step (so that the first
        `step()` of a pipelined caller does not capture while copies are in flight on other streams)."""
        if self.cfg.use_cuda_graph and (self._graph is None or self._graph_key != (self.unsup_weight, self.dan_weight)):
            if self.mode == "mixmatch" and self.cfg.mean_teacher and self.teacher_bn is None:
                self.teacher_bn = [(b["bn"].running_mean.clone(), b["bn"].running_var.clone()) for b in self._blk]
            self._reserve_gemm_ws()
            self._capture((self.unsup_weight, self.dan_weight))

    def begin_step(self):
        """What `step()` does on the host before the step's launches (for callers that replay a graph of their own made
        by `capture_step`)."""
        if self.mode == "mixmatch" and self.cfg.mean_teacher and self.teacher_bn is None:
            self.teacher_bn = [(b["bn"].running_mean.clone(), b["bn"].running_var.clone()) for b in self._blk]
        self._sync_planes()

    def end_step(self):
        self._prefetch_valid = False
        self.steps_done += 1

    def capture_step(self, pre=None, post=None):
        """One CUDA graph of [pre()] + the step + [post()]: lets a feeding pipeline put its staging copy and its loss
        read-back into the same launch as the step.  Replay it between `begin_step()` and `end_step()`."""
        self.begin_step()
        self.arena.check_aliasing()
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        before = _lib.launch_count
        with torch.cuda.graph(g):
            if pre is not None:
                pre()
            self._launch()
            if post is not None:
                post()
        self.kernels_per_step = _lib.launch_count - before
        return g

    def _capture_graph(self):
        # warm-up on a side stream is not needed: the library allocates nothing; but the first
        # eager pass must not double-apply a step, so capture directly.
        self.arena.check_aliasing()
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        before = _lib.launch_count
        with torch.cuda.graph(g):
            self._launch()
        self.kernels_per_step = _lib.launch_count - before
        return g

    def _capture(self, key):
        self._graph, self._graph_key = self._capture_graph(), key

    def losses(self) -> Dict[str, float]:
        """Read the last step's scalars (one D2H copy; the only host sync the engine exposes)."""
        v = self.loss_buf.cpu().tolist()
        out = {"sup_loss": v[0], "sup_acc": v[1] / max(self.L, 1)}
        if self.mode == "mixmatch":
            out["unsup_loss"] = v[2]
            if self.use_dan:
                n_dan = int(self.counts[2].item())
                out["dan_loss"] = v[4]
                out["dan_acc"] = v[5] / max(n_dan, 1)
            out["loss"] = v[0] + self.unsup_weight * v[2] + (v[4] if self.use_dan else 0.0)
        else:
            out["loss"] = v[0]
        return out


for _n in _BATCH_FIELDS:
    setattr(TrainEngine, _n, _batch_field(_n))
