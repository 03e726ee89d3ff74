"""Fused optimizers (`optimizer.step()` of scnym/trainer.py:234,671,1089).

`FusedOptimizer` is a `torch.optim.Optimizer` (param groups, `zero_grad`, `add_param_group`,
`state_dict` work as usual) whose `step()` runs the single elementwise kernel
`scnb_optim_step` per parameter tensor with torch.optim's update rules
(Adadelta / Adam / AdamW / SGD+momentum, L2-coupled weight decay except AdamW).  When a
`TrainEngine` owns the parameters it applies the same kernel to the whole flat arena in one
launch instead and keeps this object's hyper-parameters (`param_groups[0]`) as the source of
truth for lr / weight decay.
"""
import torch

from . import _lib
from ._lib import call, ptr, OPTIMIZER_CODES


class FusedOptimizer(torch.optim.Optimizer):
    def __init__(self, params, kind: str = "adadelta", lr: float = 1.0, weight_decay: float = 0.0, betas=(0.9, 0.999),
                 eps=None, rho: float = 0.9, momentum: float = 0.0):
        if kind not in OPTIMIZER_CODES:
            raise ValueError(f"unknown optimizer {kind}")
        if eps is None:
            eps = 1e-6 if kind == "adadelta" else 1e-8
        defaults = dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps, rho=rho, momentum=momentum)
        self.kind = kind
        super().__init__(params, defaults)
        self.engine = None  # set by a TrainEngine that took ownership of the update

    # ---- state owned by a TrainEngine: moments and step count live in the engine's arena ----
    def state_dict(self):
        """torch's layout, plus -- when a TrainEngine owns the update -- the engine's moment arenas and step counters under
        "engine" (otherwise saving / resuming would silently reset Adam / Adadelta state)."""
        sd = super().state_dict()
        eng = self.engine
        if eng is not None:
            sd["engine"] = {"state1": eng.arena.state1.detach().cpu().clone(), "state2": eng.arena.state2.detach().cpu().clone(),
                            "state4": eng.state.detach().cpu().clone(), "names": list(eng.arena.names),
                            "offsets": dict(eng.arena.offsets)}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        es = state_dict.pop("engine", None)
        super().load_state_dict(state_dict)
        if es is not None:
            eng = self.engine
            if eng is None:
                self._pending_engine_state = es      # adopted by the next TrainEngine that takes ownership
            else:
                self._apply_engine_state(eng, es)

    @staticmethod
    def _apply_engine_state(eng, es):
        if list(es["names"]) != list(eng.arena.names) or es["state1"].numel() != eng.arena.state1.numel():
            raise ValueError("optimizer state was saved for a different parameter arena")
        eng.arena.state1.copy_(es["state1"])
        eng.arena.state2.copy_(es["state2"])
        seed = eng.state[1].clone()
        eng.state.copy_(es["state4"])
        eng.state[1] = seed                          # the Philox seed stays the engine's own

    def attach_engine(self, eng):
        """Called by the trainer when a TrainEngine takes over `optimizer.step()` for these parameters."""
        self.engine = eng
        es = getattr(self, "_pending_engine_state", None)
        if es is not None:
            self._apply_engine_state(eng, es)
            self._pending_engine_state = None

    @torch.no_grad()
    def step(self, closure=None):
        if self.engine is not None:
            raise RuntimeError("FusedOptimizer.step(): a TrainEngine owns the update of these parameters (its step() applies the "
                               "optimizer with the moments kept in its arena); a separate step() here would start a second, "
                               "independent optimizer state -- detach the engine (optimizer.engine = None) first")
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        st = torch.cuda.current_stream().cuda_stream
        for group in self.param_groups:
            hp_vals = [group["lr"], group["weight_decay"], group["betas"][0], group["betas"][1], group["eps"], group["rho"],
                       group["momentum"], 0.0]
            for p in group["params"]:
                if p.grad is None:
                    continue
                _lib.require_cuda(p)
                state = self.state[p]
                # storage-order views (the gene-major first layer is contiguous as its transpose)
                pd = p.data if p.data.is_contiguous() else p.data.t()
                if not pd.is_contiguous():
                    raise RuntimeError("FusedOptimizer needs dense parameters")
                g = p.grad if p.data.is_contiguous() else p.grad.t()
                if not g.is_contiguous():
                    g = g.contiguous()
                if len(state) == 0:
                    state["s1"] = torch.zeros_like(pd)
                    state["s2"] = torch.zeros_like(pd)
                    state["t"] = torch.zeros(4, dtype=torch.int64, device=p.device)
                    state["hp"] = torch.empty(8, dtype=torch.float32, device=p.device)
                    state["hp_host"] = None
                if state["hp_host"] != hp_vals:
                    state["hp"].copy_(torch.tensor(hp_vals, dtype=torch.float32))
                    state["hp_host"] = list(hp_vals)
                call("scnb_step_begin", ptr(state["t"]), st)
                call("scnb_optim_step", OPTIMIZER_CODES[self.kind], ptr(pd), ptr(g), ptr(state["s1"]), ptr(state["s2"]), pd.numel(),
                     ptr(state["hp"]), ptr(state["t"]), st)
        return loss


def Adadelta(params, lr=1.0, weight_decay=0.0, rho=0.9, eps=1e-6):
    return FusedOptimizer(params, "adadelta", lr=lr, weight_decay=weight_decay, rho=rho, eps=eps)


def Adam(params, lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
    return FusedOptimizer(params, "adam", lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)


def AdamW(params, lr=1e-3, weight_decay=1e-2, betas=(0.9, 0.999), eps=1e-8):
    return FusedOptimizer(params, "adamw", lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)


def SGD(params, lr=1e-3, weight_decay=0.0, momentum=0.0):
    return FusedOptimizer(params, "sgd", lr=lr, weight_decay=weight_decay, momentum=momentum)
