"""Data-parallel plumbing: one process per GPU, batch sharded by rank, one sum all-reduce per flat
gradient bucket (NCCL over NVLink on the GPUs, gloo in the CPU tests), 1/world folded into the Adam kernel.
The reference has no distributed code at all; this is new."""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_rows(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of the global batch that rank trains on (contiguous, equal shards)."""
    if global_batch % world:
        raise ValueError(f'global batch {global_batch} is not divisible by world size {world}')
    per = global_batch // world
    return rank * per, (rank + 1) * per


def allreduce_buckets(flat: torch.Tensor, spans: Sequence[Tuple[int, int]], group=None, async_op: bool = True) -> List:
    """Sum all-reduce of each [a,b) span of the flat gradient buffer, in the given order (the order in which backward
    finishes them).  Returns the work handles (already waited on when async_op=False)."""
    works = []
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return works
    for a, b in spans:
        w = dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def wait_all(works) -> None:
    for w in works:
        w.wait()
