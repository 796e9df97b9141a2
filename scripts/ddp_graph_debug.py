"""Debug aid: where does a data-parallel Trainer step stop?  torchrun --nproc-per-node 2 scripts/ddp_graph_debug.py [graph|eager]"""
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(int(os.environ.get('KP_FAULT_S', '60')), exit=True)

import torch
import torch.distributed as dist
from keypoints_b200 import parallel
from keypoints_b200.models import keynet
from keypoints_b200.trainer import Trainer


def mark(msg):
    print(f'[rank {os.environ.get("RANK")}] {time.time() % 1000:8.2f} {msg}', flush=True)


rank, world, local = parallel.init_from_env('nccl')
dev = torch.device('cuda', local)
mode = sys.argv[1] if len(sys.argv) > 1 else 'graph'
torch.manual_seed(rank)
net = keynet.build('F', 3, 64, 10)
tr = Trainer(net, precision='bf16', use_graph=mode == 'graph', device=dev, augment=dict(cntl_pts=4, variance=0.05, max_rotate=0.1))
mark(f'trainer built, dp_mode={tr.dp_mode} overlap={tr.overlap_allreduce} multicast={getattr(tr.peer, "multicast", None)}')
x = torch.rand(8, 3, 64, 64, device=dev)
orig_capture = tr._capture


def traced_capture(*a, **k):
    mark('capture: start')
    r = orig_capture(*a, **k)
    mark('capture: done')
    return r


tr._capture = traced_capture
for i in range(4):
    tr.step(x)
    mark(f'step {i} issued')
    torch.cuda.synchronize()
    mark(f'step {i} done, loss {tr.loss():.5f}')
p = tr.flat_p.clone()
gathered = [torch.zeros_like(p) for _ in range(world)]
dist.all_gather(gathered, p)
mark(f'replicas identical: {all(torch.equal(g, gathered[0]) for g in gathered)}')
tr.close()
dist.destroy_process_group()
