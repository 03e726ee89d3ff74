"""Data-parallel plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink / NVSwitch).

The reference is single-process (no distributed code); the equivalence target of
data-parallel training is the reference step on the CONCATENATED global batch:
  * every rank draws its own labeled / unlabeled minibatch from its own shard of the cells;
  * BatchNorm batch statistics are global: per layer and per pass the ranks all-reduce
    (sum z, sum z^2, rows) in the forward and (sum dy, sum dy*zhat) in the backward
    (`scnb_bn_act_fwd/bwd` two-pass protocol);
  * parameter gradients are averaged with one all-reduce over the flat gradient arena, after
    which every rank applies the identical fused optimizer update;
  * MixUp permutations stay rank-local (block-diagonal in the global batch);
  * inference shards rows contiguously and needs no communication.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of `rank` (first `n_rows % world` ranks get one extra row)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class TorchDistComm(object):
    """Collectives used by `TrainEngine` on a `torch.distributed` process group."""

    def __init__(self, group=None) -> None:
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._avg = dist.get_backend(group) == "nccl"

    def all_reduce_sum(self, t: torch.Tensor) -> None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def all_reduce_avg(self, t: torch.Tensor) -> None:
        if self._avg:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:  # gloo has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.mul_(1.0 / self.world)

    def all_reduce_max(self, t: torch.Tensor) -> None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)

    def broadcast(self, t: torch.Tensor, src: int = 0) -> None:
        dist.broadcast(t, src=src, group=self.group)

    def barrier(self) -> None:
        dist.barrier(group=self.group)


def merge_bn_sums(sums: torch.Tensor, n_seg: int, width: int, eps: float = 1e-5):
    """Host restatement of what the kernels do with an all-reduced exchange buffer
    [n_seg][2][width] + [n_seg]: returns (mean, biased var, invstd, rows) per segment.  Used by the
    CPU (gloo) tests of the protocol; the product path does this inside `scnb_bn_act_fwd`."""
    s = sums[: n_seg * 2 * width].view(n_seg, 2, width)
    cnt = sums[n_seg * 2 * width: n_seg * 2 * width + n_seg].clamp_min(1.0)
    mean = s[:, 0] / cnt[:, None]
    var = (s[:, 1] / cnt[:, None] - mean * mean).clamp_min(0.0)
    return mean, var, 1.0 / torch.sqrt(var + eps), cnt


def init_from_env(device: Optional[torch.device] = None) -> Optional[TorchDistComm]:
    """Initialise NCCL from torchrun's environment (RANK / WORLD_SIZE / MASTER_*); None if single process."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    return TorchDistComm()
