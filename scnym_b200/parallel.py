"""Data-parallel plumbing: one process per GPU; `torch.distributed` (NCCL) is the RENDEZVOUS, peer memory the data plane.

The reference is single-process (no distributed code); the equivalence target of
data-parallel training is the reference step on the CONCATENATED global batch:
  * every rank draws its own labeled / unlabeled minibatch from its own shard of the cells;
  * BatchNorm batch statistics are global.  Product path (one node, NCCL process group): every rank owns a
    CUDA-IPC exchange buffer mapped by all ranks (`PeerExchange`); the BatchNorm kernels store their partial
    sums into every peer's buffer over NVLink / NVSwitch and add what the peers stored, inside the kernel
    (`scnb_bn_act_fwd_peer` / `_bwd_peer`) -- no collective launch.  Fallback (`SCNB_DP_PEER=0`, gloo tests):
    partial sums -> `all_reduce_sum` -> apply (`scnb_bn_act_fwd/bwd` two-pass protocol);
  * parameter gradients: `scnb_dp_reduce_optim_bcast` reduce-scatters them with peer loads, applies the
    optimizer to the owned slice and stores the new parameters into every rank's arena (fallback: one NCCL
    all-reduce over the flat gradient arena + the elementwise optimizer kernel);
  * MixUp permutations stay rank-local (block-diagonal in the global batch);
  * inference shards rows contiguously and needs no communication.
NCCL itself is used to exchange the IPC handles, to broadcast the initial replica and for barriers.
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

    def make_peer_exchange(self, nbytes: int, device) -> Optional["PeerExchange"]:
        """Peer-memory exchange buffers for the in-kernel BatchNorm all-reduce (NCCL process groups on one node only;
        SCNB_DP_PEER=0 keeps the partial -> NCCL all-reduce -> apply protocol)."""
        import os
        if not self._avg or os.environ.get("SCNB_DP_PEER", "1") == "0":
            return None
        return PeerExchange.from_dist(nbytes, device, self.group)


class PeerExchange(object):
    """Exchange buffers of all ranks as addressable from this process: `ptrs` is a device int64 array [world] (own
    buffer at [rank]) that the data-parallel BatchNorm kernels (`scnb_bn_act_fwd_peer` / `_bwd_peer`) store into and
    poll -- the cross-rank statistics all-reduce happens inside those kernels, over NVLink / NVSwitch peer memory."""

    def __init__(self, addrs, rank: int, world: int, nbytes: int, device, owned=None, opened=(), keep=None) -> None:
        self.rank, self.world, self.nbytes = int(rank), int(world), int(nbytes)
        self.ptrs = torch.tensor([int(a) for a in addrs], dtype=torch.int64, device=device)
        self._owned, self._opened, self._keep = owned, list(opened), keep

    @classmethod
    def from_dist(cls, nbytes: int, device, group=None) -> "PeerExchange":
        """One cudaMalloc buffer per rank, shared through CUDA IPC handles (one process per GPU, one node)."""
        import ctypes
        from . import _lib
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        with torch.cuda.device(device):
            own = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            _lib.call("scnb_peer_alloc", int(nbytes), ctypes.byref(own), handle)
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            addrs, opened = [], []
            for r in range(world):
                if r == rank:
                    addrs.append(own.value)
                else:
                    q = ctypes.c_void_p()
                    _lib.call("scnb_peer_open", ctypes.create_string_buffer(handles[r], 64), ctypes.byref(q))
                    addrs.append(q.value)
                    opened.append(q.value)
            torch.cuda.synchronize()
        dist.barrier(group=group)          # every buffer is zeroed and mapped everywhere before the first kernel polls it
        return cls(addrs, rank, world, nbytes, device, owned=own.value, opened=opened)

    @classmethod
    def local(cls, buffers, rank: int) -> "PeerExchange":
        """Emulated ranks inside one process (tests): `buffers` are the ranks' zeroed uint8 CUDA tensors."""
        return cls([b.data_ptr() for b in buffers], rank, len(buffers), buffers[0].numel(), buffers[0].device, keep=buffers)

    def tensor(self, offset_bytes: int, n_floats: int) -> torch.Tensor:
        """fp32 view of [offset_bytes, offset_bytes + 4 n_floats) of this rank's own buffer."""
        if offset_bytes % 16 or offset_bytes + 4 * n_floats > self.nbytes:
            raise ValueError("bad exchange-buffer region")
        if self._keep is not None:      # emulated ranks: the buffer is a torch tensor
            return self._keep[self.rank][offset_bytes: offset_bytes + 4 * n_floats].view(torch.float32)

        class _Raw(object):             # cudaMalloc memory owned by this object, exposed through the CUDA array interface
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(self._owned) + int(offset_bytes), False),
                                        "version": 2, "strides": None}
        raw._owner = self
        return torch.as_tensor(raw, device=self.ptrs.device)

    def close(self) -> None:
        from . import _lib
        for q in self._opened:
            _lib.call("scnb_peer_close", q)
        self._opened = []
        if self._owned:
            _lib.call("scnb_peer_free", self._owned)
            self._owned = None


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
