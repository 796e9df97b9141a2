"""Run services of the training loop that must not stall the GPU.

The reference logs every step with ``loss.item()`` (utils.py:122-130), a device->host sync that drains the queue once
per step, and drives ``ReduceLROnPlateau`` from the same host value (autoencode.py:73,128).  Here the step writes its
loss into a device ring (``kp_loss_ring_push``, part of the captured graph); the host copies the ring back every
``every`` steps on a side stream into pinned memory and consumes the (step, loss) pairs that have ARRIVED — it never
waits for the step in flight.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch

from . import lib as L


class LossRing:
    """Device ring of (step, mean loss) pairs + asynchronous read-back."""

    def __init__(self, device, slots: int = 256, every: int = 32):
        if every > slots:
            raise ValueError('read-back period must not exceed the ring size')
        self.slots, self.every = int(slots), int(every)
        self.ring = torch.full((slots, 2), -1.0, dtype=torch.float64, device=device)
        self.host = [torch.full((slots, 2), -1.0, dtype=torch.float64).pin_memory() for _ in range(2)]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.pending = [False, False]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.seen = -1                    # last step index already handed out
        self.flip = 0
        self.scale = 1.0

    # called by the trainer inside the step (graph-capturable: one tiny kernel, no host interaction)
    def push(self, loss_sum: torch.Tensor, step_dev: torch.Tensor):
        L.call('kp_loss_ring_push', L.stream(), L.ptr(loss_sum), float(self.scale), L.ptr(step_dev), L.ptr(self.ring),
               self.slots)

    def poll(self, steps_done: int, force: bool = False) -> List[Tuple[int, float]]:
        """Call once per step on the host.  Every `every` steps starts an asynchronous copy of the ring; returns the
        (step, loss) pairs of copies that have completed since the last call (possibly empty), oldest first."""
        out = []
        for b in (0, 1):
            if self.pending[b] and (force or self.events[b].query()):
                if force:
                    self.events[b].synchronize()
                out += self._harvest(self.host[b])
                self.pending[b] = False
        if (force or steps_done % self.every == 0) and not self.pending[self.flip]:
            b = self.flip
            self.copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.copy_stream):
                self.host[b].copy_(self.ring, non_blocking=True)
                self.events[b].record(self.copy_stream)
            self.pending[b] = True
            self.flip ^= 1
            if force:
                self.events[b].synchronize()
                out += self._harvest(self.host[b])
                self.pending[b] = False
        out.sort()
        return out

    def _harvest(self, host) -> List[Tuple[int, float]]:
        rows = [(int(s), float(v)) for s, v in host.tolist() if s > self.seen]
        rows.sort()
        if rows:
            self.seen = rows[-1][0]
        return rows


class PlateauLR:
    """``torch.optim.lr_scheduler.ReduceLROnPlateau`` semantics (mode='min', rel threshold; autoencode.py:73,128) fed by the
    ring: ``step(metric)`` may be called late and in batches — the decision only needs the values, not their timing."""

    def __init__(self, trainer, factor: float = 0.1, patience: int = 10, threshold: float = 1e-4, cooldown: int = 0,
                 min_lr: float = 0.0, eps: float = 1e-8):
        self.tr, self.factor, self.patience, self.threshold = trainer, factor, patience, threshold
        self.cooldown, self.min_lr, self.eps = cooldown, min_lr, eps
        self.best, self.bad, self.cool = float('inf'), 0, 0

    def step(self, metric: float):
        if metric < self.best * (1.0 - self.threshold):
            self.best, self.bad = metric, 0
        else:
            self.bad += 1
        if self.cool > 0:
            self.cool -= 1
            self.bad = 0
        if self.bad > self.patience:
            new = max(self.tr.lr * self.factor, self.min_lr)
            if self.tr.lr - new > self.eps:
                self.tr.lr = new              # drops the captured graph; re-captured on the next step
            self.cool, self.bad = self.cooldown, 0


class RunLog:
    """Drop-in for the part of ``utils.ResultsLogger`` the training loops use per step (`log(...)`, utils.py:122-130):
    keeps running train-loss statistics and forwards to an optional callback (TensorBoard writer, print, ...), fed from
    the ring instead of `loss.item()`."""

    def __init__(self, trainer, slots: int = 256, every: int = 32, on_loss: Optional[Callable[[int, float], None]] = None,
                 scheduler: Optional[PlateauLR] = None):
        self.tr = trainer
        self.ring = LossRing(trainer.device, slots, every)
        trainer.attach_loss_ring(self.ring)
        self.on_loss, self.scheduler = on_loss, scheduler
        self.history: List[Tuple[int, float]] = []
        self.best = float('inf')

    def after_step(self, force: bool = False):
        rows = self.ring.poll(self.tr.steps_done, force)
        for step, loss in rows:
            self.history.append((step, loss))
            self.best = min(self.best, loss)
            if self.on_loss is not None:
                self.on_loss(step, loss)
            if self.scheduler is not None:
                self.scheduler.step(loss)
        return rows

    def flush(self):
        return self.after_step(force=True)
