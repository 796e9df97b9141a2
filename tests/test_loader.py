"""Input pipeline (keypoints_b200/loader.py): pair sampling as AtariDataset.__getitem__ (datasets.py:173-183),
the pinned ring's recycling protocol (CPU), and on the GPU the double-buffered H2D path with the device-side
ToTensor + Normalize (datasets.py:287-295) feeding Trainer.step."""
import numpy as np
import pytest
import torch


def _trajs():
    rng = np.random.default_rng(3)
    return [rng.integers(0, 256, size=(t, 12, 10), dtype=np.uint8) for t in (40, 33)]


def test_frame_pairs_follow_the_reference_sampling():
    from keypoints_b200.loader import FramePairs
    tr = _trajs()
    ds = FramePairs(tr, min_frame_skip=5, max_frame_skip=20, seed=1)
    assert len(ds) == (40 - 20) + (33 - 20)                      # datasets.py:166-167: frames with max_frame_skip successors
    a, b = torch.empty(64, 12, 10, 1, dtype=torch.uint8), torch.empty(64, 12, 10, 1, dtype=torch.uint8)
    ds.fill(0, (a, b), 64)
    flat = [t.reshape(t.shape[0], -1) for t in tr]
    for j in range(64):
        hits = [(k, i) for k, t in enumerate(flat) for i in np.flatnonzero((t == a[j].reshape(-1).numpy()).all(1))]
        assert len(hits) == 1
        k, i = hits[0]
        later = np.flatnonzero((flat[k] == b[j].reshape(-1).numpy()).all(1))
        assert len(later) == 1 and 5 <= later[0] - i <= 20 and i < flat[k].shape[0] - 20
    a2, b2 = torch.empty_like(a), torch.empty_like(b)
    ds.fill(0, (a2, b2), 64)
    assert torch.equal(a, a2) and torch.equal(b, b2)             # seeded per batch index
    ds.fill(1, (a2, b2), 64)
    assert not torch.equal(a, a2)
    with pytest.raises(ValueError):
        FramePairs([tr[0][:10]], max_frame_skip=20)


def test_pinned_batcher_never_refills_a_slot_before_its_copy_is_reported():
    import threading
    import time
    from keypoints_b200.loader import PinnedBatcher
    n, depth = 40, 3
    lock = threading.Lock()
    filled = []

    def fn(i, out):
        with lock:
            filled.append(i)
        out[0].fill_(i)

    pb = PinnedBatcher(fn, [((4,), torch.int64)], n, workers=3, depth=depth, pin=False)
    seen = []
    for i, batch in enumerate(pb):
        time.sleep(0.002)
        with lock:                                               # nobody may have started batch >= i + depth yet
            assert max(filled) < i + depth, (i, max(filled))
        seen.append(int(batch[0][0]))
        batch.on_copied()                                        # report the (here instantaneous) copy
    assert seen == list(range(n))


@pytest.mark.gpu
def test_prefetcher_feeds_the_trainer_from_uint8_frames():
    from keypoints_b200.loader import DevicePrefetcher, FramePairs, PinnedBatcher, u8_pairs_to_float
    from keypoints_b200.models import transporter
    from keypoints_b200.trainer import Trainer
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(0)
    trajs = [rng.integers(0, 256, size=(60, 32, 32), dtype=np.uint8) for _ in range(3)]
    ds = FramePairs(trajs, seed=2)
    B, nb = 8, 12
    pb = PinnedBatcher(lambda i, out: ds.fill(i, out, B), [((B, 32, 32, 1), torch.uint8)] * 2, nb, workers=2, depth=4)
    pf = DevicePrefetcher(pb, dev, convert=u8_pairs_to_float(0.5, 0.5))
    torch.manual_seed(1)
    tr = Trainer(transporter.make('VGG_PONG', 1, 8, 3), precision='fp32', use_graph=True, device=dev)
    losses, count = [], 0
    for xa, xb in pf:
        assert xa.shape == (B, 1, 32, 32) and xa.dtype == torch.float32
        if count == 0:                                           # ToTensor + Normalize((0.5,), (0.5,)) of the first batch
            ha, hb = torch.empty(B, 32, 32, 1, dtype=torch.uint8), torch.empty(B, 32, 32, 1, dtype=torch.uint8)
            ds.fill(0, (ha, hb), B)
            ref = (ha.permute(0, 3, 1, 2).float() / 255 - 0.5) / 0.5
            assert float((xa.cpu() - ref).abs().max()) < 1e-6
        tr.step(xa, xb)
        losses.append(tr.loss())
        count += 1
    assert count == nb and all(np.isfinite(losses))
