#!/usr/bin/env python
"""End-to-end training loop on the B200 path — what `keypoints.py` / `transporter.py` / `autoencode.py` of the reference do
(keypoints.py:30-110), assembled from this package's pieces:

  JPEG ImageFolder  -> nvJPEG decode + Pillow-exact Resize((128,128)) + ToTensor on the GPU   (loader.jpeg_folder_batches)
  or uint8 frame pairs -> pinned ring -> double-buffered H2D -> device-side Normalize         (loader.FramePairs ...)
  -> fused Trainer.step (TPS+rotate, forward, loss, backward, [gradient exchange], Adam; one CUDA graph)
  -> non-blocking loss log + ReduceLROnPlateau semantics                                       (runlog.RunLog / PlateauLR)
  -> checkpoints in the reference's .mdl layout + optimiser state                              (Trainer.save / load)

  python scripts/train_example.py --data /path/to/imagefolder --model keynet --steps 200
  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/train_example.py --data ... (each rank reads its own shard)
  python scripts/train_example.py --synthetic-frames --model transporter      (no dataset: random Atari-like trajectories)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--data', default=None, help='ImageFolder root of JPEGs (CelebA layout, datasets.py:237-254)')
    ap.add_argument('--synthetic-frames', action='store_true', help='random uint8 trajectories instead of a dataset')
    ap.add_argument('--model', default='keynet', choices=['keynet', 'transporter', 'autoencoder'])
    ap.add_argument('--model-type', default=None)
    ap.add_argument('--keypoints', type=int, default=10)
    ap.add_argument('--z', type=int, default=64)
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--lr', type=float, default=1e-4)
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--run-dir', default=None)
    ap.add_argument('--resume', action='store_true')
    ap.add_argument('--checkpoint-freq', type=int, default=100)
    args = ap.parse_args(argv)

    from keypoints_b200 import loader, parallel, runlog
    from keypoints_b200.models import autoencoder, keynet, transporter
    from keypoints_b200.trainer import Trainer

    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    frames = args.synthetic_frames or args.data is None
    cin = 1 if frames else 3
    mt = args.model_type or ('VGG_PONG_LAYERNECK' if frames else 'F')
    torch.manual_seed(0)
    if args.model == 'keynet':
        net = keynet.build(mt, cin, args.z, args.keypoints)
    elif args.model == 'transporter':
        net = transporter.make(mt, cin, args.z, args.keypoints)
    else:
        net = autoencoder.make(mt, cin, args.z)
    aug = dict(cntl_pts=4, variance=0.05, max_rotate=0.1) if (not frames and args.model != 'autoencoder') else None
    tr = Trainer(net, precision=args.precision, lr=args.lr, augment=aug, device=dev, seed=0)
    if args.resume and args.run_dir and os.path.isdir(args.run_dir):
        tr.load(args.run_dir)
    sched = runlog.PlateauLR(tr, factor=0.5, patience=20)
    log = runlog.RunLog(tr, slots=256, every=16, scheduler=sched,
                        on_loss=(lambda s, v: print(f'step {s:6d}  loss {v:.5f}  lr {tr.lr:g}', flush=True)) if rank == 0 else None)

    if frames:
        rng = np.random.default_rng(100 + rank)
        trajs = [rng.integers(0, 256, size=(80, 84, 84), dtype=np.uint8) for _ in range(4)]
        ds = loader.FramePairs(trajs, seed=rank)
        pb = loader.PinnedBatcher(lambda i, out: ds.fill(i, out, args.batch), [((args.batch, 84, 84, 1), torch.uint8)] * 2,
                                  args.steps, workers=4, depth=6)
        batches = loader.DevicePrefetcher(pb, dev, convert=loader.u8_pairs_to_float(0.5, 0.5))   # grey_transform
    else:
        batches = ((x,) for x in loader.jpeg_folder_batches(args.data, args.batch, dev, (128, 128), shuffle_seed=rank, epochs=10 ** 6))

    for i, xs in enumerate(batches):
        if i >= args.steps:
            break
        if args.model == 'autoencoder' or aug is not None:
            tr.step(xs[0])
        else:
            tr.step(xs[0], xs[-1])
        log.after_step()                          # never blocks on the step in flight
        if args.run_dir and (i + 1) % args.checkpoint_freq == 0:
            tr.save(args.run_dir)                 # every rank calls it (the sharded Adam moments are gathered); rank 0 writes
    log.flush()
    if args.run_dir:
        tr.save(args.run_dir)
    tr.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return log.history


if __name__ == '__main__':
    main()
