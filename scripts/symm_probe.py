"""Probe: does torch symmetric memory (peer-mapped device buffers) work on this box?  torchrun --nproc-per-node 2 scripts/symm_probe.py"""
import os
import sys
import faulthandler
faulthandler.dump_traceback_later(60, exit=True)
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
print(rank, 'can_access_peer', [torch.cuda.can_device_access_peer(local, j) for j in range(world) if j != local], flush=True)
try:
    t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(float(rank + 1))
    h = symm.rendezvous(t, group=dist.group.WORLD)
    print(rank, 'rendezvous ok; ptrs', [hex(p) for p in h.buffer_ptrs], 'multicast', hex(h.multicast_ptr) if h.has_multicast_support else None,
          'signal pad size', h.signal_pad_size, flush=True)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, 'peer value', float(peer[0]), flush=True)
    h.barrier()
except Exception as e:
    print(rank, 'SYMM FAILED', type(e).__name__, e, flush=True)
dist.destroy_process_group()
