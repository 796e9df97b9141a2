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


def _write_jpegs(root, n, seed=0, subsampling=0):
    """A small ImageFolder of CelebA-sized (178x218) JPEGs: smooth gradients + rectangles, written with Pillow."""
    import os
    from PIL import Image
    rng = np.random.default_rng(seed)
    paths = []
    for cls in ('a', 'b'):
        os.makedirs(os.path.join(root, cls), exist_ok=True)
    for i in range(n):
        h, w = 218, 178
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        img = np.stack([(np.sin(xx / rng.uniform(8, 30) + rng.uniform(0, 6)) * 0.5 + 0.5) * 255,
                        (np.cos(yy / rng.uniform(8, 30)) * 0.5 + 0.5) * 255,
                        ((xx + yy) / (h + w)) * 255], axis=-1)
        for _ in range(3):
            y0, x0 = rng.integers(0, h - 40), rng.integers(0, w - 40)
            img[y0:y0 + rng.integers(8, 40), x0:x0 + rng.integers(8, 40)] = rng.integers(0, 256, size=3)
        p = os.path.join(root, 'a' if i % 2 == 0 else 'b', f'img_{i:03d}.jpg')
        Image.fromarray(img.astype(np.uint8)).save(p, quality=92, subsampling=subsampling)
        paths.append(p)
    return paths


def test_image_folder_listing_matches_torchvision(tmp_path):
    import torchvision as tv
    from keypoints_b200.loader import image_folder_files
    _write_jpegs(str(tmp_path), 7)
    ours = image_folder_files(str(tmp_path))
    ref = tv.datasets.ImageFolder(str(tmp_path)).samples
    assert ours == [(p, c) for p, c in ref]


@pytest.mark.gpu
def test_jpeg_decode_and_resize_match_the_reference_transform(tmp_path):
    """nvJPEG decode + the resize kernels against the reference's celeba_transform (datasets.py:297-300: Pillow decode,
    transforms.Resize((128,128)), ToTensor).  The resize is checked tightly on identical decoded pixels (<= 1 grey level,
    Pillow works in fixed point); decode + resize end to end within the IDCT tolerance between libjpeg and nvJPEG."""
    import ctypes
    import torchvision.transforms as T
    from PIL import Image
    from keypoints_b200 import lib as L
    from keypoints_b200.loader import JpegBatchDecoder, jpeg_folder_batches
    dev = torch.device('cuda:0')
    paths = _write_jpegs(str(tmp_path), 6, seed=1, subsampling=0)
    celeba_transform = T.Compose([T.Resize((128, 128)), T.ToTensor()])
    # (a) resize kernels on Pillow-decoded pixels
    for p in paths[:3]:
        im = Image.open(p).convert('RGB')
        arr = torch.from_numpy(np.asarray(im).copy()).to(dev)                    # [H][W][3] uint8
        H, W = arr.shape[:2]
        tmp = torch.empty(H * 128 * 3, dtype=torch.uint8, device=dev)
        out = torch.empty(3, 128, 128, device=dev)
        L.call('kp_resize_to_f32', L.stream(), L.ptr(arr), H, W, 3, L.ptr(tmp), L.ptr(out), 128, 128)
        ref = celeba_transform(im)
        d = (out.cpu() - ref).abs() * 255
        assert float(d.max()) <= 1.01, float(d.max())
        assert float((d > 0.5).float().mean()) < 0.02                             # almost every pixel identical
    # (b) the whole pipeline
    dec = JpegBatchDecoder(dev, (128, 128), threads=3)
    blobs = [open(p, 'rb').read() for p in paths]
    out = dec.decode(blobs)
    dec.close()
    for i, p in enumerate(paths):
        ref = celeba_transform(Image.open(p).convert('RGB'))
        d = (out[i].cpu() - ref).abs() * 255
        assert float(d.mean()) < 0.6 and float(d.max()) <= 6.0, (p, float(d.mean()), float(d.max()))
    # (c) batches of the folder in ImageFolder order feed the trainer
    got = list(jpeg_folder_batches(str(tmp_path), 2, dev, shuffle_seed=None))
    assert len(got) == 3 and got[0].shape == (2, 3, 128, 128)
    from keypoints_b200.loader import image_folder_files
    order = [p for p, _ in image_folder_files(str(tmp_path))]
    ref0 = celeba_transform(Image.open(order[0]).convert('RGB'))
    assert float((got[0][0].cpu() - ref0).abs().max()) * 255 <= 6.0
    from keypoints_b200.models import keynet
    from keypoints_b200.trainer import Trainer
    torch.manual_seed(0)
    tr = Trainer(keynet.build('VGG_PONG', 3, 8, 4), precision='bf16', use_graph=False, device=dev,
                 augment=dict(cntl_pts=4, variance=0.05, max_rotate=0.1))
    for x in got:
        tr.step(x)
        assert np.isfinite(tr.loss())


@pytest.mark.gpu
def test_train_example_end_to_end(tmp_path):
    """scripts/train_example.py: the reference's outer loop assembled from loader + Trainer + runlog + checkpoints — on an
    ImageFolder of JPEGs (KeyNet, TPS+rotate) and on synthetic frame pairs (Transporter), with a save / resume."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'scripts'))
    import train_example
    _write_jpegs(str(tmp_path / 'img'), 8, seed=3)
    run = str(tmp_path / 'run')
    h = train_example.main(['--data', str(tmp_path / 'img'), '--model', 'keynet', '--model-type', 'VGG_PONG', '--z', '8',
                            '--keypoints', '4', '--batch', '4', '--steps', '6', '--run-dir', run, '--checkpoint-freq', '3'])
    assert [s for s, _ in h] == list(range(6)) and all(np.isfinite(v) for _, v in h)
    assert sorted(os.listdir(run)) == ['decoder', 'encoder', 'keypoint', 'trainer.pt']
    h2 = train_example.main(['--data', str(tmp_path / 'img'), '--model', 'keynet', '--model-type', 'VGG_PONG', '--z', '8',
                             '--keypoints', '4', '--batch', '4', '--steps', '2', '--run-dir', run, '--resume'])
    assert [s for s, _ in h2] == [6, 7]                         # the step counter came back with the optimiser state
    h3 = train_example.main(['--synthetic-frames', '--model', 'transporter', '--z', '16', '--keypoints', '4', '--batch', '4',
                             '--steps', '5'])
    assert len(h3) == 5 and all(np.isfinite(v) for _, v in h3)
