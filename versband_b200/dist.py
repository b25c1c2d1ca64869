"""Data-parallel sharding of clips over the GPUs of one box.

Clips are independent units: there is no collective on the data path. `torch.distributed`
(NCCL over NVLink on GPUs, gloo in the CPU tests) is used only for single-controller
orchestration — scatter the conditioning inputs from rank 0, gather the waveforms back — which is
what the reference's `mp.spawn` + `DistributedSampler` arrangement amounts to
(scripts/test_final.py:351-357,474-475: rendezvous only, no collective).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> tuple:
    """Contiguous block partition; the first (n % world) ranks get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class ClipExchange:
    """Scatter / gather of per-clip tensors between a controller rank and the workers, set up ONCE per
    (clip count, tensor signature): the shape/dtype handshake (the only pickled, host-synchronising step) happens
    on the first call, receive buffers are allocated once, and every later step is a single grouped
    collective per direction — `dist.scatter` / `dist.gather` when the shards are equal, one batch of
    point-to-point ops otherwise. Nothing on the per-step path touches the host."""

    def __init__(self, n_clips: int, *, src: int = 0, device=None, group=None):
        self.n, self.src, self.device, self.group = n_clips, src, device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.bounds = [shard_bounds(n_clips, r, self.world) for r in range(self.world)]
        self.equal = len({b - a for a, b in self.bounds}) == 1
        self.meta = None          # [(shape_tail, dtype)] of the scattered tensors
        self.recv: List[torch.Tensor] = []
        self.gather_out: Optional[torch.Tensor] = None

    def _handshake(self, tensors) -> None:
        meta = [[(tuple(t.shape[1:]), t.dtype) for t in tensors]] if self.rank == self.src else [None]
        dist.broadcast_object_list(meta, src=self.src, group=self.group)
        self.meta = meta[0]
        lo, hi = self.bounds[self.rank]
        self.recv = [torch.empty((hi - lo, *shape), dtype=dtype, device=self.device) for shape, dtype in self.meta]

    def scatter(self, tensors: Optional[Sequence[torch.Tensor]]) -> List[torch.Tensor]:
        """Rank `src` passes tensors shaped [n_clips, ...] (on `device`); every rank gets its shard."""
        if self.meta is None:
            self._handshake(tensors)
        src, group = self.src, self.group
        if self.equal:
            for i, buf in enumerate(self.recv):
                chunks = None
                if self.rank == src:
                    t = tensors[i].to(self.device)
                    chunks = [t[a:b].contiguous() for a, b in self.bounds]
                dist.scatter(buf, scatter_list=chunks, src=src, group=group)
            return self.recv
        ops_ = []
        for i, buf in enumerate(self.recv):
            if self.rank == src:
                t = tensors[i].to(self.device)
                for r, (a, b) in enumerate(self.bounds):
                    if r == src:
                        buf.copy_(t[a:b])
                    elif b > a:
                        ops_.append(dist.P2POp(dist.isend, t[a:b].contiguous(), r, group))
            elif buf.shape[0]:
                ops_.append(dist.P2POp(dist.irecv, buf, src, group))
        if ops_:
            for q in dist.batch_isend_irecv(ops_):
                q.wait()
        return self.recv

    def gather(self, local: torch.Tensor, dst: Optional[int] = None) -> Optional[torch.Tensor]:
        """Inverse direction for one tensor [local_clips, ...]: rank `dst` gets [n_clips, ...] (a buffer owned by
        the exchange and overwritten by the next gather)."""
        dst = self.src if dst is None else dst
        group = self.group
        if self.rank == dst and (self.gather_out is None or self.gather_out.shape[1:] != local.shape[1:]
                                 or self.gather_out.dtype != local.dtype):
            self.gather_out = torch.empty((self.n, *local.shape[1:]), dtype=local.dtype, device=local.device)
        if self.equal:
            outs = [self.gather_out[a:b] for a, b in self.bounds] if self.rank == dst else None
            dist.gather(local.contiguous(), gather_list=outs, dst=dst, group=group)
            return self.gather_out if self.rank == dst else None
        ops_ = []
        if self.rank == dst:
            for r, (a, b) in enumerate(self.bounds):
                if r == dst:
                    self.gather_out[a:b].copy_(local)
                elif b > a:
                    ops_.append(dist.P2POp(dist.irecv, self.gather_out[a:b], r, group))
        elif local.shape[0]:
            ops_.append(dist.P2POp(dist.isend, local.contiguous(), dst, group))
        if ops_:
            for q in dist.batch_isend_irecv(ops_):
                q.wait()
        return self.gather_out if self.rank == dst else None


_EXCHANGES: dict = {}


def _exchange(n_clips: int, device, group) -> ClipExchange:
    key = (n_clips, str(device), id(group))
    if key not in _EXCHANGES:
        _EXCHANGES[key] = ClipExchange(n_clips, device=device, group=group)
    return _EXCHANGES[key]


def scatter_clips(tensors: Optional[Sequence[torch.Tensor]], n_clips: int, *, src: int = 0, device=None,
                  group=None) -> List[torch.Tensor]:
    """One-shot form of ClipExchange.scatter (a fresh exchange: handshake included)."""
    return ClipExchange(n_clips, src=src, device=device, group=group).scatter(tensors)


def gather_clips(local: torch.Tensor, n_clips: int, *, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    return ClipExchange(n_clips, src=dst, device=local.device, group=group).gather(local)


def run_data_parallel(fn: Callable[[List[torch.Tensor]], torch.Tensor], inputs: Optional[Sequence[torch.Tensor]],
                      n_clips: int, *, device=None, group=None) -> Optional[torch.Tensor]:
    """scatter -> fn(local shard) -> gather. `fn` maps a list of [b, ...] tensors to a [b, ...] tensor. The
    exchange object is cached per (n_clips, device, group): only the first call pays the metadata handshake."""
    ex = _exchange(n_clips, device, group)
    local = ex.scatter(inputs)
    res = fn(local)
    return ex.gather(res)
