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


class PeerBuckets:
    """The flat parameter and gradient buckets of a replica in SYMMETRIC memory (every replica's buffers mapped into every
    process over NVLink; torch.distributed._symmetric_memory does the handle exchange) plus the barrier flags, and the
    peer-pointer table ``kp_dp_adam_step`` takes.  Raises if the GPUs cannot map each other's memory."""

    def __init__(self, n: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import lib as L
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > L.DP_MAX_WORLD:
            raise RuntimeError(f'peer-memory data parallelism supports up to {L.DP_MAX_WORLD} replicas')
        self.p = symm.empty(n, dtype=torch.float32, device=device)
        self.g = symm.empty(n, dtype=torch.float32, device=device)
        self.flags = symm.empty(64, dtype=torch.int32, device=device)
        for t in (self.p, self.g, self.flags):
            t.zero_()
        self.handles = [symm.rendezvous(t, group=group) for t in (self.p, self.g, self.flags)]
        hp, hg, hf = self.handles
        self.peers = L.KpDpPeers()
        for j in range(self.world):
            self.peers.p[j], self.peers.g[j], self.peers.flag[j] = hp.buffer_ptrs[j], hg.buffer_ptrs[j], hf.buffer_ptrs[j]
        # NVSwitch multicast (multimem.ld_reduce / multimem.st): measured slower than plain peer loads/stores at 2 replicas
        # (0.283 vs 0.186 ms for the 98 MB bucket, gpurun_out/r2_dp_micro_n2.log); KP_DP_MULTICAST=0/1 overrides
        want = os.environ.get('KP_DP_MULTICAST')
        want = (self.world >= 4) if want is None else want != '0'
        self.multicast = bool(want and getattr(hp, 'has_multicast_support', False)
                              and getattr(hg, 'has_multicast_support', False))
        if self.multicast:
            self.peers.mc_p, self.peers.mc_g = hp.multicast_ptr, hg.multicast_ptr
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)                 # every replica's flags are zero before anyone signals

    def owned(self, n: int) -> Tuple[int, int]:
        """Element range of the flat bucket whose Adam moments this rank holds (kp_dp_adam_step's slice)."""
        n4 = n // 4
        chunk4 = (n4 + self.world - 1) // self.world
        return min(self.rank * chunk4, n4) * 4, min((self.rank + 1) * chunk4, n4) * 4
