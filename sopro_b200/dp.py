"""Data-parallel plumbing: one process per GPU, utterances sharded by contiguous slices, no data-path collective.
The only collective of the whole path is the start-up weight broadcast: rank 0 loads/builds the
checkpoint, every other rank receives it over NCCL/NVLink (or gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """GPU `rank` of `world` owns utterances [lo, hi): contiguous, sizes differ by at most one."""
    lo = (n_items * rank) // world
    hi = (n_items * (rank + 1)) // world
    return lo, hi


def broadcast_state_dict(sd: Optional[Dict[str, torch.Tensor]], specs: Sequence[Tuple[str, Tuple[int, ...]]], *, src: int = 0,
                         device: Optional[torch.device] = None, group=None) -> Dict[str, torch.Tensor]:
    """Broadcast the fp32 tensors named in `specs` (name, shape) from rank `src` as ONE flat buffer.
    Returns CPU tensors on every rank (the engine copies them to its own device arena)."""
    rank = dist.get_rank(group)
    total = sum(int(torch.Size(s).numel()) for _, s in specs)
    dev = device or torch.device("cpu")
    flat = torch.empty(total, dtype=torch.float32, device=dev)
    if rank == src:
        assert sd is not None
        flat.copy_(torch.cat([sd[k].reshape(-1).to(torch.float32) for k, _ in specs]))
    dist.broadcast(flat, src=src, group=group)
    host = flat.cpu()
    out, off = {}, 0
    for k, s in specs:
        n = int(torch.Size(s).numel())
        out[k] = host[off: off + n].view(s).clone()
        off += n
    return out


def gather_token_lists(local: List[List[int]], group=None) -> List[List[int]]:
    """Optional: collect every rank's token lists on all ranks (tokens are tiny; waveforms stay rank-local)."""
    world = dist.get_world_size(group)
    bucket: List[Optional[List[List[int]]]] = [None] * world
    dist.all_gather_object(bucket, local, group=group)
    return [t for part in bucket for t in (part or [])]
