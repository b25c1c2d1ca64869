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


def scatter_clips(tensors: Optional[Sequence[torch.Tensor]], n_clips: int, *, src: int = 0, device=None,
                  group=None) -> List[torch.Tensor]:
    """Rank `src` holds tensors shaped [n_clips, ...]; every rank receives its shard_bounds slice.
    Non-src ranks pass `tensors=None` plus the metadata broadcast by src."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    meta = [[(tuple(t.shape[1:]), t.dtype) for t in tensors]] if rank == src else [None]
    dist.broadcast_object_list(meta, src=src, group=group)
    out = []
    lo, hi = shard_bounds(n_clips, rank, world)
    for i, (shape, dtype) in enumerate(meta[0]):
        recv = torch.empty((hi - lo, *shape), dtype=dtype, device=device)
        if rank == src:
            chunks = []
            for r in range(world):
                a, b = shard_bounds(n_clips, r, world)
                chunks.append(tensors[i][a:b].to(device).contiguous())
            # ragged shards: use point-to-point sends (scatter needs equal sizes)
            reqs = [dist.isend(chunks[r], dst=r, group=group) for r in range(world) if r != src]
            recv.copy_(chunks[src])
            for q in reqs:
                q.wait()
        else:
            dist.recv(recv, src=src, group=group)
        out.append(recv)
    return out


def gather_clips(local: torch.Tensor, n_clips: int, *, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Inverse of scatter_clips for one tensor [local_clips, ...]: rank `dst` gets [n_clips, ...]."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank != dst:
        if local.shape[0]:
            dist.send(local.contiguous(), dst=dst, group=group)
        return None
    out = torch.empty((n_clips, *local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        a, b = shard_bounds(n_clips, r, world)
        if r == dst:
            out[a:b].copy_(local)
        elif b > a:
            dist.recv(out[a:b], src=r, group=group)
    return out


def run_data_parallel(fn: Callable[[List[torch.Tensor]], torch.Tensor], inputs: Optional[Sequence[torch.Tensor]],
                      n_clips: int, *, device=None, group=None) -> Optional[torch.Tensor]:
    """scatter -> fn(local shard) -> gather. `fn` maps a list of [b, ...] tensors to a [b, ...] tensor."""
    local = scatter_clips(inputs, n_clips, device=device, group=group)
    res = fn(local)
    return gather_clips(res, n_clips, group=group)
