#!/usr/bin/env python
"""bench.py — training images(pairs)/s of the keypoint hot path on N B200s (driver contract).

Workload (BASELINE.json north_star target / configs[2]): KeyNet `F`, synthetic 128x128x3, K=10, bf16 tensor-core
path, per-GPU batch 64 (weak scaling), TPS+rotate augmentation of the pair on the device.  One step =
augment + forward + masked L2 loss + backward + gradient all-reduce + Adam.  1 image = 1 training pair.

  python bench.py --gpus N --steps K --warmup W          our arm (torchrun for N > 1)
  python bench.py --impl reference ...                   the reference algorithm (oracle port, CPU, all host threads)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, model_type, cin, z, K, H, W, augment)
    'keynet_F_128_K10': ('keynet', 'F', 3, 64, 10, 128, 128, True),
    'transporter_F_128_K30': ('transporter', 'F', 3, 64, 30, 128, 128, True),
    'transporter_pong_84_K4': ('transporter', 'VGG_PONG_LAYERNECK', 1, 16, 4, 84, 84, False),
    'keynet_F_256_K64': ('keynet', 'F', 3, 64, 64, 256, 256, False),
}
# algorithmic conv GFLOP per training pair (fwd + wgrad + dgrad of every conv call)
GFLOP_PER_PAIR = {'keynet_F_128_K10': 196.517, 'transporter_F_128_K30': 235.254, 'transporter_pong_84_K4': 1.447,
                  'keynet_F_256_K64': 787.766}
# minimal conv bytes per training pair of the HBM-bound workload (bf16, each operand once per pass)
HBM_BYTES_PER_PAIR = {'transporter_pong_84_K4': 20.45e6}
# per-GPU batch of each workload (cfg 2 batch 64, cfg 3 sweep point 64, cfg 4 / 5 batch 32)
DEFAULT_BATCH = {'keynet_F_128_K10': 64, 'transporter_F_128_K30': 32, 'transporter_pong_84_K4': 64, 'keynet_F_256_K64': 32}
AUG = dict(cntl_pts=4, variance=0.05, max_rotate=0.1)          # configs/keypoints_celeba.yaml:13-16


def synth_batch(n, c, h, w, seed, lo=0.0, hi=1.0):
    """Smooth noise + bright rectangles (SURVEY.md 8d synthetic inputs)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(n, c, max(h // 8, 2), max(w // 8, 2), generator=g)
    x = torch.nn.functional.interpolate(base, size=(h, w), mode='bilinear', align_corners=False)
    for i in range(n):
        for _ in range(3):
            y0, x0 = int(torch.randint(0, h - 4, (1,), generator=g)), int(torch.randint(0, w - 4, (1,), generator=g))
            hh, ww = int(torch.randint(2, max(h // 4, 3), (1,), generator=g)), int(torch.randint(2, max(w // 4, 3), (1,), generator=g))
            x[i, :, y0:y0 + hh, x0:x0 + ww] = torch.rand(c, 1, 1, generator=g) * 0.5 + 0.5
    return (x * (hi - lo) + lo).contiguous()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sustained=p['bf16_tflops_sustained'], source='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.path = tempfile.mktemp(suffix='.csv')
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(index),
                                          '-lms', '20'], stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            f = [t.strip() for t in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        os.unlink(self.path)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_reference_rate(wl, batch, steps, warmup, threads=None, check=None):
    """The reference algorithm (oracle port of the reference's PyTorch graph, fp32) on the host cores: full train
    step (forward, loss, autograd backward, Adam) on a bounded sample of the workload.  Returns (pairs/s, cores, s/step).
    `check` (optional callable) receives the first step's inputs and the oracle's outputs, so the baseline leg doubles as
    the checker of the CUDA path on the benchmark shapes (the second half of BASELINE's metric: keypoint max-abs-err)."""
    import torch
    from oracle import keypoints_oracle as O
    kind, mt, cin, z, K, H, W, aug = WORKLOADS[wl]
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    ops = O.transporter_ops(mt, cin, z, K) if kind == 'transporter' else O.keynet_ops(mt, cin, z, K)
    tr = O.OracleTrainer(kind, mt, cin, z, K, O.init_state_dict(ops, 0))
    x = synth_batch(batch, cin, H, W, 1234)
    gen = torch.Generator().manual_seed(5)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if aug:
            p1 = O.sample_perturb_params(batch, AUG['cntl_pts'], AUG['variance'], AUG['max_rotate'], gen)
            p2 = O.sample_perturb_params(batch, AUG['cntl_pts'], AUG['variance'], AUG['max_rotate'], gen)
            a, b, mask = O.tps_and_rotate(x, p1, p2)
        else:
            a, b, mask = x, x.flip(0), None
        loss, out = tr.step(a, b, mask)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
        if i == 0 and check is not None:
            check(a, b, mask, float(loss), out)
    per = sum(times) / len(times)
    return batch / per, torch.get_num_threads(), per


def parity_for(wl, dev, batch=2):
    """fp32 (parity-mode) CUDA path against the CPU oracle port on the workload's own shapes: identical weights
    (oracle.init_state_dict seed 0), identical (augmented) inputs, one train step.  This is the second half of BASELINE's
    metric (keypoint (x,y) max-abs-err vs ref); bar 1e-3."""
    import torch
    from oracle import keypoints_oracle as O
    from keypoints_b200.models import keynet, transporter
    from keypoints_b200.trainer import Trainer
    kind, mt, cin, z, K, H, W, aug = WORKLOADS[wl]
    torch.set_num_threads(os.cpu_count() or 1)
    ops = O.transporter_ops(mt, cin, z, K) if kind == 'transporter' else O.keynet_ops(mt, cin, z, K)
    lo, hi = (-1.0, 1.0) if cin == 1 else (0.0, 1.0)
    x = synth_batch(batch, cin, H, W, 4321, lo, hi)
    if aug:
        gen = torch.Generator().manual_seed(5)
        p1 = O.sample_perturb_params(batch, AUG['cntl_pts'], AUG['variance'], AUG['max_rotate'], gen)
        p2 = O.sample_perturb_params(batch, AUG['cntl_pts'], AUG['variance'], AUG['max_rotate'], gen)
        a, b, mask = O.tps_and_rotate(x, p1, p2)
    else:
        a, b, mask = x, x.flip(0).contiguous(), None
    ref = O.OracleTrainer(kind, mt, cin, z, K, O.init_state_dict(ops, 0))
    ref_loss, ref_out = ref.step(a, b, mask)
    net = transporter.make(mt, cin, z, K) if kind == 'transporter' else keynet.build(mt, cin, z, K)
    net.load_state_dict(O.init_state_dict(ops, 0), strict=True)
    tr = Trainer(net, precision='fp32', use_graph=False, augment=None, device=dev, process_group=False)
    tr.step(a.to(dev), b.to(dev), None if mask is None else mask.to(dev))
    k_t, xhat = tr.outputs()
    ref_x, ref_k = ref_out[0].detach(), ref_out[2].detach()
    out = {'keypoint_max_abs_err': float((k_t.cpu() - ref_k).abs().max()),
           'recon_max_rel_err': float((xhat.cpu() - ref_x).abs().max() / ref_x.abs().max()),
           'loss_rel_err': abs(tr.loss() - float(ref_loss)) / abs(float(ref_loss)),
           'what': f'fp32 parity-mode CUDA path vs the CPU oracle port, one train step of {wl} at batch {batch} '
                   f'(same weights and augmented inputs); bar 1e-3'}
    del tr, net
    torch.cuda.empty_cache()
    return out


def gpu_eager_rate(wl, batch, steps, warmup, dev, mode):
    """The "kernel to beat" (BASELINE.md 4(4)): the reference's graph — the oracle's functional restatement of it,
    the same ATen calls — in torch eager on this GPU: cuDNN convolutions, ATen BatchNorm / pooling / upsampling /
    grid_sample, autograd, torch.optim.Adam.  mode 'fp32': PyTorch defaults (TF32 convolutions allowed, as the reference
    scripts would run); mode 'bf16': autocast(bf16) + channels_last, both with cudnn.benchmark=True.  Generous to the
    baseline: augmentation on the GPU (the reference builds its TPS grids on the CPU), no per-step loss.item()."""
    import torch
    from oracle import keypoints_oracle as O
    kind, mt, cin, z, K, H, W, aug = WORKLOADS[wl]
    torch.backends.cudnn.benchmark = True
    ops = O.transporter_ops(mt, cin, z, K) if kind == 'transporter' else O.keynet_ops(mt, cin, z, K)
    bf16 = mode == 'bf16'
    sd = {}
    for k, v in O.init_state_dict(ops, 0).items():
        v = v.to(dev)
        if bf16 and v.dim() == 4:
            v = v.contiguous(memory_format=torch.channels_last)
        sd[k] = v
    keys = O.trainable_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
    optim = torch.optim.Adam([sd[k] for k in keys], lr=1e-4)
    lo, hi = (-1.0, 1.0) if cin == 1 else (0.0, 1.0)
    x = synth_batch(batch, cin, H, W, 1234, lo, hi).to(dev)
    x2 = x.roll(1, 0).contiguous()
    if bf16:
        x, x2 = x.contiguous(memory_format=torch.channels_last), x2.contiguous(memory_format=torch.channels_last)
    gen = torch.Generator().manual_seed(5)

    def step():
        if aug:
            p1 = tuple(t.to(dev) for t in O.sample_perturb_params(batch, AUG['cntl_pts'], AUG['variance'], AUG['max_rotate'], gen))
            p2 = tuple(t.to(dev) for t in O.sample_perturb_params(batch, AUG['cntl_pts'], AUG['variance'], AUG['max_rotate'], gen))
            a, b, mask = O.tps_and_rotate(x, p1, p2)
        else:
            a, b, mask = x, x2, None
        optim.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=bf16):
            out = O.transporter_forward(a, b, sd, ops) if kind == 'transporter' else O.keynet_forward(a, b, sd, ops)
            loss = O.l2_reconstruction_loss(out[0].float(), b, mask)
        loss.backward()
        optim.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lv = float(loss)
    del sd, optim
    torch.cuda.empty_cache()
    return {'value': batch / (ms / 1e3), 'unit': 'pairs/s', 'ms_per_step': ms, 'loss': lv}


def gpu_eager_baseline(wl, batch, dev, ours):
    out = {'what': 'reference graph in torch eager on this GPU (cuDNN / ATen / autograd / torch.optim.Adam), batch '
                   f'{batch}, cudnn.benchmark=True; fp32 = PyTorch defaults (TF32 convolutions), bf16 = autocast + channels_last'}
    for mode in ('fp32', 'bf16'):
        try:
            out[mode] = gpu_eager_rate(wl, batch, 5, 3, dev, mode)
        except Exception as e:                       # e.g. out of memory at a large shape: report, do not hide
            out[mode] = {'error': f'{type(e).__name__}: {str(e)[:200]}'}
    best = max((out[m].get('value', 0.0) for m in ('fp32', 'bf16')), default=0.0)
    out['ours_over_best_eager'] = ours / best if best else None
    return out


def module_api_rate(wl, batch, steps, warmup, dev, precision):
    """The path the reference's scripts drive UNCHANGED: the reference's inner loop
    (transporter.py:75-89 / keypoints.py:70-84) written against its own import paths, resolved by keypoints_b200.dropin —
    nn.Module forward, autograd backward, torch.optim.Adam, TpsAndRotate from data_augments."""
    import torch
    import keypoints_b200
    from keypoints_b200 import dropin
    kind, mt, cin, z, K, H, W, aug = WORKLOADS[wl]
    dropin.install(precision=precision)
    from keypoints.models import transporter as ref_transporter       # noqa: E402  (resolved by dropin)
    from keypoints.models import keynet as ref_keynet
    from data_augments import TpsAndRotate
    torch.manual_seed(0)
    net = (ref_transporter.make(mt, cin, z, K) if kind == 'transporter' else ref_keynet.build(mt, cin, z, K)).to(dev)
    optim = torch.optim.Adam(net.parameters(), lr=1e-4)
    augment = TpsAndRotate(AUG['cntl_pts'], AUG['variance'], AUG['max_rotate']) if aug else None
    lo, hi = (-1.0, 1.0) if cin == 1 else (0.0, 1.0)
    x = synth_batch(batch, cin, H, W, 1234, lo, hi).to(dev)
    x2 = x.roll(1, 0).contiguous()

    def step():
        if augment is not None:
            a, b, mask = augment(x, x)
        else:
            a, b, mask = x, x2, None
        optim.zero_grad()
        out = net(a, b)
        loss = ((out[0] - b) ** 2 * mask).mean() if mask is not None else ((out[0] - b) ** 2).mean()
        loss.backward()
        optim.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lv = float(loss)
    keypoints_b200.set_precision('fp32')
    del net, optim
    torch.cuda.empty_cache()
    return {'value': batch / (ms / 1e3), 'unit': 'pairs/s', 'ms_per_step': ms, 'loss': lv,
            'what': 'reference training loop through the module API (dropin import paths, autograd, torch.optim.Adam), '
                    f'{precision} mode, batch {batch}'}


def workload_config(args, wl, batch, world, graph=None):
    """The `config` object of the JSON line (identical for both arms)."""
    kind, mt, cin, z, K, H, W, aug = WORKLOADS[wl]
    cfg = {'workload': wl, 'per_gpu_batch': batch, 'global_batch': batch * world, 'image': [cin, H, W],
           'keypoints': K, 'augment': 'TpsAndRotate(4,0.05,0.1) on device' if aug else 'none', 'parallelism': f'dp{world}'}
    if graph is not None:
        cfg['cuda_graph'] = graph
    return cfg


def step_roofline(wl, pairs_per_s_per_gpu, pk):
    """Whole-step roofline of a workload: the F nets are bound by the tensor pipe (algorithmic conv GFLOP per
    pair against the measured sustained bf16 peak), the Pong net by HBM (minimal conv bytes per pair against the measured
    copy bandwidth)."""
    if wl in HBM_BYTES_PER_PAIR:
        ach = pairs_per_s_per_gpu * HBM_BYTES_PER_PAIR[wl] / 1e9
        return {'bound': 'hbm', 'scope': 'whole step', 'achieved': ach, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': ach / pk['hbm'],
                'peak_source': pk['source'] + ' hbm_gbs', 'bytes_per_pair': HBM_BYTES_PER_PAIR[wl]}
    ach = pairs_per_s_per_gpu * GFLOP_PER_PAIR[wl] / 1e3
    return {'bound': 'tensor', 'scope': 'whole step', 'achieved': ach, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
            'frac': ach / pk['tf_sustained'], 'peak_source': pk['source'] + ' bf16_tflops_sustained',
            'gflop_per_pair': GFLOP_PER_PAIR[wl]}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = args.workload
    batch = args.cpu_batch
    steps = max(1, min(args.steps, 3))
    rate, cores, per = cpu_reference_rate(wl, batch, steps, 1)
    line = {'metric': 'training images/sec', 'value': rate, 'unit': 'pairs/s', 'impl': 'reference', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': 1, 'ms_per_step': per * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload_config(args, wl, args.batch, max(args.gpus, 1)), sample_batch=batch,
                           augment='TpsAndRotate(4,0.05,0.1) on the host' if WORKLOADS[wl][7] else 'none'),
            'cpu_baseline': {'value': rate, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{steps} full train steps of {wl} at batch {batch} (fp32, torch CPU ops)'},
            'e2e': {'value': rate, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


class Run:
    """One workload on this rank's GPU: builds the net and the fused trainer, holds the synthetic batch."""

    def __init__(self, args, wl, batch, dev, rank, world):
        import torch
        from keypoints_b200.models import keynet, transporter
        from keypoints_b200.trainer import Trainer
        self.wl, self.B, self.dev, self.rank, self.world = wl, batch, dev, rank, world
        kind, mt, cin, z, K, H, W, aug = WORKLOADS[wl]
        self.aug = aug
        torch.manual_seed(0)                                   # (the trainer broadcasts rank 0's weights anyway)
        net = transporter.make(mt, cin, z, K) if kind == 'transporter' else keynet.build(mt, cin, z, K)
        self.tr = Trainer(net, precision=args.precision, use_graph=not args.no_graph, augment=AUG if aug else None, device=dev)
        lo, hi = (-1.0, 1.0) if cin == 1 else (0.0, 1.0)
        self.x_host = synth_batch(batch, cin, H, W, 1234 + rank, lo, hi).pin_memory()
        self.x2_host = self.x_host.roll(1, 0).contiguous().pin_memory()
        self.x_dev, self.x2_dev = self.x_host.to(dev), self.x2_host.to(dev)

    def one_step(self):
        if self.aug:
            self.tr.step(self.x_dev)
        else:
            self.tr.step(self.x_dev, self.x2_dev)

    def sync(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def timed(self, steps, warmup):
        """W untimed + K timed steps with the inputs resident in HBM; CUDA events; returns ms (this rank)."""
        import torch
        for _ in range(warmup):
            self.one_step()
        self.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.one_step()
        e1.record()
        self.sync()
        return e0.elapsed_time(e1)

    def timed_e2e(self, steps):
        """End to end through the public API with HOST buffers: H2D of the step's inputs + D2H of the loss inside the
        timed region.  Double-buffered input pipeline, as a data loader with pinned memory drives the public API: the H2D
        copy of batch i+1 runs on a copy stream while step i computes; every step still copies its own inputs from pinned
        host memory and reads its own loss back."""
        import torch
        from keypoints_b200.loader import DevicePrefetcher
        tr, aug = self.tr, self.aug
        batches = ((self.x_host,) if aug else (self.x_host, self.x2_host))
        pf = DevicePrefetcher(iter(lambda: batches, None), self.dev)

        # the loss of every step is read back on the host inside the timed region, one step late: the 8-byte D2H copy of
        # step i is queued behind it and consumed after step i+1 has been launched, so the host never drains the queue
        # (what runlog.LossRing does for a training loop; the reference's per-step loss.item() stalls here)
        pinned = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)]
        evs = [torch.cuda.Event(), torch.cuda.Event()]
        state = {'i': 0, 'loss': 0.0}

        def e2e_step():
            cur = pf.next()
            tr.step(*cur)
            pf.release()                                       # step() has copied its inputs into the static graph buffers
            k = state['i'] & 1
            pinned[k].copy_(tr.loss_sum, non_blocking=True)    # device -> host read of this step's loss
            evs[k].record()
            if state['i'] > 0:
                evs[k ^ 1].synchronize()
                state['loss'] = float(pinned[k ^ 1]) / tr.numel
            state['i'] += 1

        def drain():
            k = (state['i'] - 1) & 1
            evs[k].synchronize()
            state['loss'] = float(pinned[k]) / tr.numel

        for _ in range(2):                                     # untimed: first use of the staging tensors / pinned copies
            e2e_step()
        drain()
        self.sync()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(steps):
            e2e_step()
        drain()
        f1.record()
        self.sync()
        loss_val = state['loss']
        h2d = self.x_host.numel() * 4 * (1 if aug else 2)
        return f0.elapsed_time(f1), h2d, loss_val

    def kernel_timing(self):
        """Per-launch CUDA events, eager, one stream, rank 0 only (no collective in this leg)."""
        import torch
        from keypoints_b200 import lib as L
        tr = self.tr
        # one stream: with the side streams on (encoder beside keypoint branch, weight gradients beside the dgrad chain) a
        # kernel's event pair also spans the time it waits for SMs held by the kernel running next to it
        was = tr.use_graph, tr.two_streams, tr.wgrad_streams, tr.world
        tr.use_graph, tr.two_streams, tr.wgrad_streams, tr.world = False, False, False, 1
        self.one_step()
        torch.cuda.synchronize()
        runs = []
        for _ in range(3):                       # three timed steps: the per-call durations are averaged call by call
            L.timing = []
            self.one_step()
            torch.cuda.synchronize()
            runs.append(L.timing)
        L.timing = None
        tr.use_graph, tr.two_streams, tr.wgrad_streams, tr.world = was
        assert all(len(r) == len(runs[0]) for r in runs)
        rec = []
        for calls in zip(*runs):
            name, fl, _, _, tg = calls[0]
            ms = sum(a.elapsed_time(b) for _, _, a, b, _ in calls) / len(calls)
            rec.append((name, fl, ms, tg))
        return rec

    def close(self):
        import torch
        self.tr.close()                        # the captured NCCL collectives must be gone before the process group is
        del self.tr
        torch.cuda.empty_cache()


def max_over_ranks(vals, dev, world):
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return [float(v) for v in t]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='keynet_F_128_K10', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=None, help='per-GPU batch (weak scaling); default: the workload\'s own')
    ap.add_argument('--cpu-batch', type=int, default=16, help='batch of the bounded CPU sample')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--api', default='trainer', choices=['trainer', 'module'],
                    help="'module': time the reference loop through the module API (dropin) instead of the fused trainer")
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-timing', action='store_true')
    ap.add_argument('--no-eager-baseline', action='store_true')
    ap.add_argument('--no-other-workloads', action='store_true')
    ap.add_argument('--no-batch-sweep', action='store_true')
    ap.add_argument('--no-module-api', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.batch is None:
        args.batch = DEFAULT_BATCH[args.workload]
    if os.environ.get('KP_FAULT_S'):          # debugging aid: dump every thread's stack and exit if the run hangs
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['KP_FAULT_S']), exit=True)
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from keypoints_b200 import lib as L

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    pk = peaks()
    wl, B = args.workload, args.batch

    if args.api == 'module':                   # rank-local, no data parallelism on this path (the scripts have none)
        res = module_api_rate(wl, B, args.steps, args.warmup, dev, args.precision)
        if rank == 0:
            print(json.dumps({'metric': 'training images/sec', 'value': res['value'], 'unit': 'pairs/s', 'n_gpus': 1,
                              'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': res['ms_per_step'],
                              'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.precision,
                              'data': 'synthetic', 'api': 'module', 'config': workload_config(args, wl, B, 1, False),
                              'what': res['what'], 'loss': res['loss']}))
        if world > 1:
            dist.destroy_process_group()
        return

    run = Run(args, wl, B, dev, rank, world)
    tr = run.tr
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = L.launches
    ms = run.timed(args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    calls_per_step = getattr(tr, 'calls_per_step', None)
    launches = (L.launches - l0) * args.steps // (args.steps + args.warmup) if args.no_graph else None
    ms_e2e, h2d, loss_val = run.timed_e2e(args.steps)
    ms, ms_e2e = max_over_ranks([ms, ms_e2e], dev, world)
    value = args.steps * B * world / (ms / 1e3)
    e2e = args.steps * B * world / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel family (tcgen05 convs), per-launch CUDA events, eager (no graph) ----
    roof = None
    if rank == 0 and not args.no_kernel_timing and args.precision == 'bf16':
        rec = run.kernel_timing()
        agg = {}
        calls = []
        for name, fl, t_ms, tg in rec:
            d = agg.setdefault(name, [0, 0.0, 0.0])
            d[0] += 1; d[1] += t_ms; d[2] += fl
            calls.append((t_ms, name, tg, fl))
        if os.environ.get('KP_BENCH_CALLS'):
            with open(os.environ['KP_BENCH_CALLS'], 'w') as fh:
                for t_ms, name, tg, fl in sorted(calls, key=lambda c: -c[0]):
                    fh.write(f'{t_ms:8.4f} ms  {name:24s} {tg:40s} {fl / t_ms / 1e9 if fl else 0:8.1f} TFLOP/s\n')
        total_ms = sum(d[1] for d in agg.values())
        fam = ('kp_conv_tc', 'kp_conv_wgrad_tc', 'kp_conv_wgrad_tc_img')
        # the deferred fold of the staging gradients belongs to the weight-gradient kernels' time (no flops of its own)
        tc_ms = sum(agg[k][1] for k in fam + ('kp_wgrad_finalize_multi',) if k in agg)
        tc_fl = sum(agg[k][2] for k in fam if k in agg)
        tc_n = sum(agg[k][0] for k in fam if k in agg)
        step_roof = step_roofline(wl, value / world, pk)
        if tc_n and step_roof['bound'] == 'tensor':
            ach = tc_fl / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
            traffic, traffic_note = None, None
            tpath = next((q for q in (os.path.join(ROOT, 'profiles', f'{t}_tc_traffic.json') for t in ('r2b', 'r2', 'r1'))
                          if os.path.exists(q)), '')
            if wl == 'keynet_F_128_K10' and B == 64 and os.path.exists(tpath):
                tj = json.load(open(tpath))
                traffic = tj['bytes_per_launch']
                traffic_note = (f"dram read+write bytes per launch, mean over the {tj['launches']} tcgen05 launches of one step "
                                f"(ncu, {os.path.relpath(tpath, ROOT).replace('.json', '.md')}); algorithmic minimum 219.3 MB/pair x 64 / "
                                f"{tj['launches']} = {219.3e6 * 64 / tj['launches'] / 1e6:.0f} MB per launch (SURVEY 8d)")
            roof = {'bound': 'tensor', 'kernel': 'conv_tc_k + wgrad_tc_k (tcgen05 implicit-GEMM convs)', 'achieved': ach,
                    'peak': pk['tf_sustained'], 'peak_source': pk['source'] + ' bf16_tflops_sustained', 'unit': 'TFLOP/s',
                    'frac': ach / pk['tf_sustained'], 'traffic': traffic, 'traffic_note': traffic_note, 'launches_per_step': tc_n,
                    'share_of_step': tc_ms / total_ms if total_ms else None,
                    'step_frac_of_roofline': step_roof['frac']}
        else:
            # HBM-bound workload (the Pong net): the whole step against the minimal conv traffic; the dominant kernels are
            # the 16/32-channel mma.sync convolutions and the BatchNorm passes, all L2-resident at this size
            roof = dict(step_roof, traffic=None, kernel='whole step (small_mma convs + BatchNorm passes)')
        roof['per_call_ms'] = {k: round(v[1], 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}
        calls_per_step = len([r for r in rec if r[0] != 'kp_zero'])
    activation_bytes, n_params = tr.activation_bytes(), tr.n_params
    run.close()

    # ---- the other BASELINE configs, short runs on every rank (same weak-scaling contract) ----
    others = []
    if not args.no_other_workloads and args.precision == 'bf16':
        for owl in ('transporter_pong_84_K4', 'transporter_F_128_K30', 'keynet_F_256_K64'):
            if owl == wl:
                continue
            ob = DEFAULT_BATCH[owl]
            r2 = Run(args, owl, ob, dev, rank, world)
            osteps = 10
            oms = r2.timed(osteps, 3)
            (oms,) = max_over_ranks([oms], dev, world)
            oval = osteps * ob * world / (oms / 1e3)
            entry = {'workload': owl, 'config': workload_config(args, owl, ob, world, not args.no_graph), 'value': oval,
                     'unit': 'pairs/s', 'steps': osteps, 'warmup': 3, 'ms_per_step': oms / osteps,
                     'roofline': step_roofline(owl, oval / world, pk), 'loss': r2.tr.loss(),
                     'gpu_launches_per_step': getattr(r2.tr, 'calls_per_step', None)}
            r2.close()
            others.append(entry)

    # ---- batch sweep of the headline workload (SURVEY 8 cfg 3: 16 = the reference's default batch, 64, 256), rank 0 only ----
    sweep = None
    if not (args.no_batch_sweep or args.no_other_workloads) and args.precision == 'bf16' and world == 1 and wl == 'keynet_F_128_K10':
        sweep = []
        for sb in (16, 128, 256):
            if sb == B:
                continue
            r3 = Run(args, wl, sb, dev, rank, world)
            ssteps = 10
            sms = r3.timed(ssteps, 3)
            sval = ssteps * sb / (sms / 1e3)
            sweep.append({'per_gpu_batch': sb, 'value': sval, 'unit': 'pairs/s', 'ms_per_step': sms / ssteps, 'steps': ssteps,
                          'warmup': 3, 'step_frac_of_roofline': step_roofline(wl, sval, pk)['frac']})
            r3.close()
        sweep.append({'per_gpu_batch': B, 'value': value, 'unit': 'pairs/s', 'ms_per_step': ms / args.steps, 'steps': args.steps,
                      'warmup': args.warmup, 'step_frac_of_roofline': step_roofline(wl, value, pk)['frac']})
        sweep.sort(key=lambda e: e['per_gpu_batch'])

    cpu = None
    parity = None
    eager = None
    module_api = None
    if rank == 0 and world == 1:
        if not args.no_module_api and args.precision == 'bf16':
            module_api = module_api_rate(wl, B, 5, 3, dev, 'bf16')
            module_api['trainer_over_module'] = value / module_api['value']
        if not args.no_eager_baseline:
            eager = gpu_eager_baseline(wl, B, dev, value)
            for entry in others:
                entry['gpu_eager_baseline'] = gpu_eager_baseline(entry['workload'], entry['config']['per_gpu_batch'], dev,
                                                                 entry['value'])
        if not args.no_cpu_baseline:
            rate, cores, per = cpu_reference_rate(wl, args.cpu_batch, 2, 1)
            cpu = {'value': rate, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                   'sample': f'2 full train steps of {wl} at batch {args.cpu_batch} ({per:.2f} s/step, fp32 torch CPU ops)'}
            parity = parity_for(wl, dev, batch=4)
            for entry in others:
                entry['parity'] = parity_for(entry['workload'], dev, batch=2)

    if rank == 0:
        line = {'metric': 'training images/sec', 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
                'config': dict(workload_config(args, wl, B, world, not args.no_graph),
                               l2='per-step working set (activations, several GB) far exceeds the 126 MB L2',
                               frames_per_s=2 * value),
                'e2e': {'value': e2e, 'unit': 'pairs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 8,
                        'ms_per_step': ms_e2e / args.steps},
                'gpu_launches': (launches if launches is not None else (calls_per_step or 0) * args.steps),
                'gpu_launches_note': 'C-ABI kernel-launching calls inside the timed region (each >= 1 kernel; replayed from a CUDA graph)',
                'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu, 'parity': parity, 'gpu_eager_baseline': eager,
                'module_api': module_api, 'other_workloads': others or None, 'batch_sweep': sweep, 'loss': loss_val,
                'activation_bytes': activation_bytes, 'params': n_params}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
