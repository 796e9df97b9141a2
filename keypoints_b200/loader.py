"""Input pipeline for the fused trainer: the reference feeds its loop from
``DataLoader(num_workers=0)`` and a synchronous ``.to(device)`` (transporter.py:33,78-79), which cannot keep a
6 k pairs/s/GPU step busy.  Here batches are assembled by worker threads into PINNED host buffers and copied to
the device on a copy stream while the previous step computes (two device slots, events order the streams).

  FramePairs        — Atari-style pair sampling (datasets.py:109-183: frame t and frame t+skip of one trajectory) over
                      uint8 trajectories held in host memory; batches travel as uint8 (4x fewer PCIe bytes than fp32)
                      and are normalised on the device (grey_transform / color_transform, datasets.py:287-295).
  PinnedBatcher     — background threads filling a ring of pinned batches from any ``batch_fn(index) -> tensors``.
  DevicePrefetcher  — double-buffered H2D in front of ``Trainer.step``.
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, Iterator, Optional, Sequence, Tuple

import numpy as np
import torch


class DevicePrefetcher:
    """Double-buffered host->device copies.  ``next()`` returns the device tensors of the current batch (the compute
    stream waits for their copy) and starts the copy of the following batch on the copy stream; call ``release()`` once
    the consumer has taken what it needs from the slot (``Trainer.step`` copies its inputs into its static graph buffers
    first thing), so the slot can be overwritten."""

    def __init__(self, batches: Iterator[Sequence[torch.Tensor]], device, convert: Optional[Callable] = None):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.convert = convert               # optional device-side post-processing of a slot (e.g. uint8 -> float)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.filled = [False, False]
        self.i = 0
        main = torch.cuda.current_stream(self.device)
        for ev in self.consumed:
            ev.record(main)
        self._prefetch(0)

    def _prefetch(self, slot):
        try:
            host = next(self.it)
        except StopIteration:
            self.filled[slot] = False
            return
        if self.slots[slot] is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(self.slots[slot], host)):
            self.slots[slot] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])       # the step that last read this slot has taken its copy
            for d, h in zip(self.slots[slot], host):
                d.copy_(h, non_blocking=True)
            self.ready[slot].record(self.copy_stream)
            cb = getattr(host, 'on_copied', None)
            if cb is not None:                                      # lets a PinnedBatcher recycle the host buffer safely
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                cb(ev)
        self.filled[slot] = True

    def next(self) -> Tuple[torch.Tensor, ...]:
        cur = self.i & 1
        if not self.filled[cur]:
            raise StopIteration
        self.i += 1
        self._prefetch(cur ^ 1)                                     # next batch's H2D overlaps this step
        torch.cuda.current_stream(self.device).wait_event(self.ready[cur])
        self._cur = cur
        out = self.slots[cur]
        return self.convert(*out) if self.convert is not None else out

    def release(self):
        self.consumed[self._cur].record(torch.cuda.current_stream(self.device))

    def __iter__(self):
        while True:
            try:
                yield self.next()
            except StopIteration:
                return
            self.release()


class PinnedBatch(tuple):
    """A tuple of pinned host tensors that wants to know when its host->device copy has been issued."""
    on_copied: Optional[Callable] = None


class PinnedBatcher:
    """Worker threads assemble batches into a ring of `depth` pinned host buffers.  ``batch_fn(i, out)`` fills the tuple of
    pinned tensors ``out`` for batch index i (numpy / torch CPU work releases the GIL in its inner loops).  Ring slot
    i % depth is refilled for batch i + depth only after the device copy that read batch i has COMPLETED: the consumer
    (DevicePrefetcher) reports the CUDA event of that copy through ``PinnedBatch.on_copied``."""

    def __init__(self, batch_fn: Callable, shapes: Sequence[Tuple[Tuple[int, ...], torch.dtype]], n_batches: int,
                 workers: int = 4, depth: int = 6, pin: bool = True):
        self.batch_fn, self.n, self.depth = batch_fn, n_batches, depth
        mk = (lambda s, d: torch.empty(s, dtype=d).pin_memory()) if pin else (lambda s, d: torch.empty(s, dtype=d))
        self.ring = [tuple(mk(s, d) for s, d in shapes) for _ in range(depth)]
        self.done = [threading.Event() for _ in range(n_batches)]
        self.copied = [threading.Event() for _ in range(n_batches)]
        self.copy_event = [None] * n_batches
        self.todo: "queue.Queue[int]" = queue.Queue()
        for i in range(n_batches):
            self.todo.put(i)
        self.errors = []
        self.threads = [threading.Thread(target=self._work, daemon=True) for _ in range(workers)]
        for t in self.threads:
            t.start()

    def _work(self):
        while True:
            try:
                i = self.todo.get_nowait()
            except queue.Empty:
                return
            try:
                if i >= self.depth:                  # the previous tenant of this ring slot must be on the device
                    self.copied[i - self.depth].wait()
                    ev = self.copy_event[i - self.depth]
                    if ev is not None:
                        ev.synchronize()
                self.batch_fn(i, self.ring[i % self.depth])
            except Exception as e:                   # surface in the consumer
                self.errors.append(e)
            self.done[i].set()

    def _mark(self, i, event=None):
        self.copy_event[i] = event
        self.copied[i].set()

    def __iter__(self):
        for i in range(self.n):
            self.done[i].wait()
            if self.errors:
                raise self.errors[0]
            b = PinnedBatch(self.ring[i % self.depth])
            b.on_copied = lambda ev=None, i=i: self._mark(i, ev)
            yield b
            if not self.copied[i].is_set() and i >= 1:
                # a consumer that does not report copies (plain iteration): batch i-1 is no longer in use once i is asked for
                self._mark(i - 1)


class FramePairs:
    """Pairs (frame t, frame t + skip) from uint8 trajectories [T][H][W][C] (AtariDataset.__getitem__, datasets.py:173-183:
    skip ~ U{min_frame_skip..max_frame_skip}, index = every frame with max_frame_skip successors)."""

    def __init__(self, trajectories: Sequence[np.ndarray], min_frame_skip: int = 5, max_frame_skip: int = 20, seed: int = 0):
        self.traj = [np.ascontiguousarray(t if t.ndim == 4 else t[..., None]) for t in trajectories]
        self.min_skip, self.max_skip = min_frame_skip, max_frame_skip
        self.index = [(k, i) for k, t in enumerate(self.traj) for i in range(t.shape[0] - max_frame_skip)]
        if not self.index:
            raise ValueError('trajectories shorter than max_frame_skip')
        self.seed = seed
        self.shape = self.traj[0].shape[1:]          # (H, W, C)

    def __len__(self):
        return len(self.index)

    def fill(self, batch_index: int, out, batch: int):
        """Write batch `batch_index` (a seeded random draw of `batch` pairs) into the pinned uint8 tensors out = (a, b),
        each [batch][H][W][C]."""
        rng = np.random.default_rng((self.seed, batch_index))
        a, b = out[0].numpy(), out[1].numpy()
        picks = rng.integers(0, len(self.index), size=batch)
        skips = rng.integers(self.min_skip, self.max_skip + 1, size=batch)
        for j, (p, s) in enumerate(zip(picks, skips)):
            k, i = self.index[p]
            a[j] = self.traj[k][i]
            b[j] = self.traj[k][i + s]


def u8_pairs_to_float(mean: float = 0.5, std: float = 0.5):
    """Device-side ``ToTensor() + Normalize(mean, std)`` (grey_transform / color_transform, datasets.py:287-295) for uint8
    NHWC slots: returns a convert() for DevicePrefetcher producing fp32 NCHW tensors."""
    from . import lib as L

    def convert(*slots):
        outs = []
        for s in slots:
            n, h, w, c = s.shape
            o = torch.empty((n, c, h, w), dtype=torch.float32, device=s.device)
            L.call('kp_u8_to_f32', L.stream(), L.ptr(s), L.ptr(o), n, h, w, c, 1.0 / (255.0 * std), -mean / std)
            outs.append(o)
        return tuple(outs)
    return convert


# ---- real images: ImageFolder of JPEGs -> nvJPEG decode -> Pillow-exact resize -> fp32 NCHW batch on the device ---------
IMG_EXTENSIONS = ('.jpg', '.jpeg', '.png', '.ppm', '.bmp', '.pgm', '.tif', '.tiff', '.webp')


def image_folder_files(root: str):
    """The sample list of ``torchvision.datasets.ImageFolder(root)`` (datasets.py:248): classes = sorted sub-directories,
    files of a class in sorted walk order; returns [(path, class_index)]."""
    import os
    classes = sorted(e.name for e in os.scandir(root) if e.is_dir())
    if not classes:
        raise FileNotFoundError(f"Couldn't find any class folder in {root}.")
    out = []
    for ci, cname in enumerate(classes):
        for d, _, fnames in sorted(os.walk(os.path.join(root, cname), followlinks=True)):
            for f in sorted(fnames):
                if f.lower().endswith(IMG_EXTENSIONS):
                    out.append((os.path.join(d, f), ci))
    return out


class JpegBatchDecoder:
    """Decodes a batch of JPEG byte strings on the GPU and applies the reference's ``celeba_transform`` (datasets.py:297-300:
    Resize((128,128)) + ToTensor) into an fp32 NCHW batch.  `threads` host threads each own an nvJPEG context, a CUDA stream
    and scratch buffers (nvJPEG's Huffman stage runs on the host; ctypes releases the GIL during the calls)."""

    def __init__(self, device, size=(128, 128), threads: int = 4):
        import ctypes
        from concurrent.futures import ThreadPoolExecutor
        from . import lib as L
        self.L, self.device, self.size = L, torch.device(device), tuple(size)
        L.load()
        self.workers = []
        with torch.cuda.device(self.device):
            for _ in range(threads):
                ctx = ctypes.c_void_p()
                L.call('kp_jpeg_create', ctypes.byref(ctx), count=False)
                self.workers.append({'ctx': ctx, 'stream': torch.cuda.Stream(device=self.device), 'rgb': None, 'tmp': None,
                                     'event': torch.cuda.Event()})
        self.pool = ThreadPoolExecutor(max_workers=threads)
        self.free: "queue.Queue[dict]" = queue.Queue()
        for w in self.workers:
            self.free.put(w)

    def _one(self, data: bytes, out_slot: torch.Tensor):
        import ctypes
        L = self.L
        w = self.free.get()
        try:
            with torch.cuda.device(self.device):
                buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data)
                wd, ht, nc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
                L.call('kp_jpeg_info', w['ctx'], buf, len(data), ctypes.byref(wd), ctypes.byref(ht), ctypes.byref(nc), count=False)
                W, H = wd.value, ht.value
                oh, ow = self.size
                if w['rgb'] is None or w['rgb'].numel() < H * W * 3:
                    w['rgb'] = torch.empty(H * W * 3, dtype=torch.uint8, device=self.device)
                if w['tmp'] is None or w['tmp'].numel() < H * ow * 3:
                    w['tmp'] = torch.empty(H * ow * 3, dtype=torch.uint8, device=self.device)
                st = ctypes.c_void_p(w['stream'].cuda_stream)
                L.call('kp_jpeg_decode', w['ctx'], st, buf, len(data), L.ptr(w['rgb']), W, H)
                L.call('kp_resize_to_f32', st, L.ptr(w['rgb']), H, W, 3, L.ptr(w['tmp']), L.ptr(out_slot), oh, ow)
                w['event'].record(w['stream'])
                w['event'].synchronize()          # scratch buffers and the host byte buffer are reused by the next image
        finally:
            self.free.put(w)

    def decode(self, blobs: Sequence[bytes], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """blobs: JPEG byte strings -> out [len(blobs)][3][oh][ow] fp32 in [0,1] on the device (complete on return)."""
        oh, ow = self.size
        if out is None:
            out = torch.empty((len(blobs), 3, oh, ow), dtype=torch.float32, device=self.device)
        futs = [self.pool.submit(self._one, b, out[i]) for i, b in enumerate(blobs)]
        for f in futs:
            f.result()
        return out

    def close(self):
        self.pool.shutdown(wait=True)
        for w in self.workers:
            self.L.call('kp_jpeg_destroy', w['ctx'], count=False)
        self.workers = []


def jpeg_folder_batches(root: str, batch: int, device, size=(128, 128), shuffle_seed: Optional[int] = 0, epochs: int = 1,
                        decode_threads: int = 4, read_threads: int = 4, drop_last: bool = True):
    """Batches of an ImageFolder of JPEGs as fp32 NCHW device tensors (the CelebA datapack of the reference,
    datasets.py:237-254,297-300 + DataLoader(shuffle=True), keypoints.py:36): files are read by a thread pool one batch
    ahead of the decoder, decoded and resized on the GPU.  Feed the result to ``Trainer.step(x)`` with TpsAndRotate on."""
    from concurrent.futures import ThreadPoolExecutor
    files = [p for p, _ in image_folder_files(root)]
    dec = JpegBatchDecoder(device, size, decode_threads)
    readers = ThreadPoolExecutor(max_workers=read_threads)

    def read(path):
        with open(path, 'rb') as fh:
            return fh.read()

    try:
        for ep in range(epochs):
            order = np.arange(len(files))
            if shuffle_seed is not None:
                np.random.default_rng((shuffle_seed, ep)).shuffle(order)
            starts = list(range(0, len(order) - (batch - 1 if drop_last else 0), batch))
            pending = None
            for s in starts + [None]:
                nxt = None if s is None else [readers.submit(read, files[i]) for i in order[s:s + batch]]
                if pending is not None:
                    yield dec.decode([f.result() for f in pending])
                pending = nxt
    finally:
        readers.shutdown(wait=False)
        dec.close()
