"""world_size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding, bucketed gradient all-reduce,
and the equivalence 'Adam on the averaged per-shard gradients == Adam on the full-batch gradient' that the
1/world scaling in kp_adam_step relies on.  The oracle is used as the checker for the Adam arithmetic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from keypoints_b200 import parallel
    r, w, _ = parallel.init_from_env('gloo')
    assert (r, w) == (rank, world)
    lo, hi = parallel.shard_rows(8, rank, world)
    assert hi - lo == 4 and lo == 4 * rank
    torch.manual_seed(100 + rank)
    flat = torch.randn(1000)
    mine = flat.clone()
    spans = [(0, 300), (300, 650), (650, 1000)]
    parallel.wait_all(parallel.allreduce_buckets(flat, spans))
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    assert torch.allclose(flat, sum(gathered), atol=1e-6)
    out[rank] = flat.clone()
    dist.destroy_process_group()


def test_bucket_allreduce_world2():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert torch.equal(out[0], out[1])


def test_mean_of_shard_gradients_equals_full_batch_gradient():
    """l2 loss is a mean over the batch, so with equal shards grad(full) == mean_r grad(shard_r) when BatchNorm is
    absent; with 1/world in Adam the replicas then take the same step as a single process would."""
    from oracle import keypoints_oracle as O
    torch.manual_seed(0)
    w = torch.randn(3, 3, 3, 3, requires_grad=True)
    x = torch.randn(8, 3, 10, 10)
    t = torch.randn(8, 3, 8, 8)

    def grad(xs, ts):
        loss = O.l2_reconstruction_loss(torch.nn.functional.conv2d(xs, w), ts)
        (g,) = torch.autograd.grad(loss, w)
        return g
    full = grad(x, t)
    shards = [grad(x[0:4], t[0:4]), grad(x[4:8], t[4:8])]
    summed = shards[0] + shards[1]
    assert torch.allclose(summed * 0.5, full, atol=1e-6)
    p1, p2 = w.detach().clone(), w.detach().clone()
    m1, v1, m2, v2 = (torch.zeros_like(p1) for _ in range(4))
    O.adam_step(p1, full, m1, v1, 1)
    O.adam_step(p2, summed * 0.5, m2, v2, 1)
    assert torch.allclose(p1, p2, atol=1e-7)


def test_shard_rows_rejects_ragged():
    from keypoints_b200 import parallel
    with pytest.raises(ValueError):
        parallel.shard_rows(10, 0, 4)
    assert parallel.shard_rows(64, 3, 8) == (24, 32)
