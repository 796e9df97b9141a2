"""Micro-benchmark of the data-parallel optimiser step alone (all ranks in lockstep):
fused peer-memory kernel (multicast / plain) vs NCCL all-reduce + replicated Adam.
torchrun --nproc-per-node N scripts/dp_microbench.py"""
import faulthandler
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(int(os.environ.get('KP_FAULT_S', '90')), exit=True)

import torch
import torch.distributed as dist
from keypoints_b200 import parallel
from keypoints_b200.models import keynet
from keypoints_b200.trainer import Trainer

rank, world, local = parallel.init_from_env('nccl')
dev = torch.device('cuda', local)
net = keynet.build('F', 3, 64, 10)
tr = Trainer(net, precision='bf16', use_graph=False, device=dev)
tr.flat_g.normal_()


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


res = {}
if tr.peer is not None:
    can_mc = bool(tr.peer.peers.mc_g) or tr.peer.multicast
    if not tr.peer.multicast and getattr(tr.peer.handles[0], 'has_multicast_support', False):
        tr.peer.peers.mc_p, tr.peer.peers.mc_g = tr.peer.handles[0].multicast_ptr, tr.peer.handles[1].multicast_ptr
        can_mc = True
    for mc in ([1, 0] if can_mc else [0]):
        tr.peer.multicast = bool(mc)
        res[f'fused p2p (multicast={mc})'] = timed(tr._dp_adam)
res['adam alone (replicated)'] = timed(tr._adam)
res['nccl all-reduce fp32 1 span'] = timed(lambda: dist.all_reduce(tr.flat_g))
spans = [u.span for u in tr.units.values()]
res['nccl all-reduce 3 spans + adam'] = timed(lambda: (parallel.wait_all(parallel.allreduce_buckets(tr.flat_g, spans, tr.pg)), tr._adam()))
if rank == 0:
    mb = tr.n_params * 4 / 1e6
    for k, v in res.items():
        print(f'{k:36s} {v:7.3f} ms   ({mb:.0f} MB bucket, world {world})', flush=True)
tr.close()
dist.destroy_process_group()
