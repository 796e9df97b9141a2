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


def parity_check(args, dev, result):
    """Returns the `check` callback: the fp32 (parity-mode) CUDA path on the same weights (oracle.init_state_dict seed 0 is
    what cpu_reference_rate trains) and the same first-step inputs, compared with the oracle's outputs."""
    def check(a, b, mask, ref_loss, ref_out):
        import torch
        from oracle import keypoints_oracle as O
        from keypoints_b200.models import keynet, transporter
        from keypoints_b200.trainer import Trainer
        kind, mt, cin, z, K, H, W, aug = WORKLOADS[args.workload]
        ops = O.transporter_ops(mt, cin, z, K) if kind == 'transporter' else O.keynet_ops(mt, cin, z, K)
        net = transporter.make(mt, cin, z, K) if kind == 'transporter' else keynet.build(mt, cin, z, K)
        net.load_state_dict(O.init_state_dict(ops, 0), strict=True)
        tr = Trainer(net, precision='fp32', use_graph=False, augment=None, device=dev, process_group=False)
        tr.step(a.to(dev), b.to(dev), None if mask is None else mask.to(dev))
        k_t, xhat = tr.outputs()
        ref_x, ref_k = ref_out[0].detach(), ref_out[2].detach()
        result.update({
            'keypoint_max_abs_err': float((k_t.cpu() - ref_k).abs().max()),
            'recon_max_rel_err': float((xhat.cpu() - ref_x).abs().max() / ref_x.abs().max()),
            'loss_rel_err': abs(tr.loss() - ref_loss) / abs(ref_loss),
            'what': f'fp32 parity-mode CUDA path vs the CPU oracle port, first train step of the cpu_baseline sample '
                    f'({args.workload}, batch {a.shape[0]}, same weights and augmented inputs); bar: 1e-3'})
        del tr, net
        torch.cuda.empty_cache()
    return check


def workload_config(args, world, graph=None):
    """The `config` object of the JSON line (identical for both arms)."""
    kind, mt, cin, z, K, H, W, aug = WORKLOADS[args.workload]
    cfg = {'workload': args.workload, 'per_gpu_batch': args.batch, 'global_batch': args.batch * world, 'image': [cin, H, W],
           'keypoints': K, 'augment': 'TpsAndRotate(4,0.05,0.1) on device' if aug else 'none', 'parallelism': f'dp{world}'}
    if graph is not None:
        cfg['cuda_graph'] = graph
    return cfg


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
            'config': dict(workload_config(args, max(args.gpus, 1)), sample_batch=batch,
                           augment='TpsAndRotate(4,0.05,0.1) on the host' if WORKLOADS[wl][7] else 'none'),
            'cpu_baseline': {'value': rate, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{steps} full train steps of {wl} at batch {batch} (fp32, torch CPU ops)'},
            'e2e': {'value': rate, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='keynet_F_128_K10', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=64, help='per-GPU batch (weak scaling)')
    ap.add_argument('--cpu-batch', type=int, default=16, help='batch of the bounded CPU sample')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-timing', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if os.environ.get('KP_FAULT_S'):          # debugging aid: dump every thread's stack and exit if the run hangs
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['KP_FAULT_S']), exit=True)
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from keypoints_b200 import lib as L
    from keypoints_b200.models import keynet, transporter
    from keypoints_b200.trainer import Trainer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    kind, mt, cin, z, K, H, W, aug = WORKLOADS[args.workload]
    torch.manual_seed(0)                                   # identical initial weights on every rank
    net = transporter.make(mt, cin, z, K) if kind == 'transporter' else keynet.build(mt, cin, z, K)
    tr = Trainer(net, precision=args.precision, use_graph=not args.no_graph, augment=AUG if aug else None, device=dev)
    B = args.batch
    lo, hi = (-1.0, 1.0) if cin == 1 else (0.0, 1.0)
    x_host = synth_batch(B, cin, H, W, 1234 + rank, lo, hi).pin_memory()
    x2_host = x_host.roll(1, 0).contiguous().pin_memory()
    x_dev, x2_dev = x_host.to(dev), x2_host.to(dev)

    def one_step():
        if aug:
            tr.step(x_dev)
        else:
            tr.step(x_dev, x2_dev)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    sync()
    l0 = L.launches
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    calls_per_step = getattr(tr, 'calls_per_step', None)
    launches = (L.launches - l0) if args.no_graph else None

    # ---- end to end through the public API with HOST buffers: H2D of the step's inputs + D2H of the loss ----
    # double-buffered input pipeline, as a data loader with pinned memory would drive the public API: the H2D copy of
    # batch i+1 runs on a copy stream while step i computes; every step still copies its own inputs from pinned host
    # memory and reads its own loss back, all inside the timed region
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [(torch.empty_like(x_dev), torch.empty_like(x_dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream()

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])             # the step that last read this slot has taken its copy
            slots[slot][0].copy_(x_host, non_blocking=True)
            if not aug:
                slots[slot][1].copy_(x2_host, non_blocking=True)
            ready[slot].record(copy_stream)

    state = {'i': 0}
    for ev in consumed:
        ev.record(main)
    prefetch(0)

    def e2e_step():
        cur = state['i'] & 1
        state['i'] += 1
        prefetch(cur ^ 1)                                      # next batch's H2D overlaps this step
        main.wait_event(ready[cur])
        if aug:
            tr.step(slots[cur][0])
        else:
            tr.step(slots[cur][0], slots[cur][1])
        consumed[cur].record(main)                             # step() copies its inputs into the static graph buffers first
        return tr.loss()                                       # device -> host read of the step's loss

    for _ in range(2):                                         # untimed: first use of the staging tensors / pinned copies
        e2e_step()
    sync()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    loss_val = 0.0
    for _ in range(args.steps):
        loss_val = e2e_step()
    f1.record()
    sync()
    ms_e2e = f0.elapsed_time(f1)
    h2d = x_host.numel() * 4 * (1 if aug else 2)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    value = args.steps * B * world / (ms / 1e3)
    e2e = args.steps * B * world / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel family (tcgen05 convs), per-launch CUDA events, eager (no graph) ----
    roof = None
    pk = peaks()
    if rank == 0 and not args.no_kernel_timing and args.precision == 'bf16':
        tr2 = tr
        was, was2, was_world = tr2.use_graph, tr2.two_streams, tr2.world
        tr2.use_graph = False
        tr2.two_streams = False                    # time every kernel alone on one stream
        tr2.world = 1                              # rank 0 only: no collective in this leg (the other ranks are not in it)
        one_step()
        torch.cuda.synchronize()
        L.timing = []
        one_step()
        torch.cuda.synchronize()
        rec, L.timing = L.timing, None
        tr2.use_graph, tr2.two_streams, tr2.world = was, was2, was_world
        agg = {}
        calls = []
        for name, fl, a, b, tg in rec:
            d = agg.setdefault(name, [0, 0.0, 0.0])
            t_ms = a.elapsed_time(b)
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
        ach = tc_fl / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, 'profiles', 'r1_tc_traffic.json')
        if args.workload == 'keynet_F_128_K10' and B == 64 and os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj['bytes_per_launch']
            traffic_note = (f"dram read+write bytes per launch, mean over the {tj['launches']} tcgen05 launches of one step "
                            f"(ncu, profiles/r1_tc_traffic.md); algorithmic minimum 219.3 MB/pair x 64 / {tj['launches']} = "
                            f"{219.3e6 * 64 / tj['launches'] / 1e6:.0f} MB per launch (SURVEY 8d)")
        roof = {'bound': 'tensor', 'kernel': 'conv_tc_k + wgrad_tc_k (tcgen05 implicit-GEMM convs)', 'achieved': ach,
                'peak': pk['tf_sustained'], 'peak_source': pk['source'] + ' bf16_tflops_sustained', 'unit': 'TFLOP/s',
                'frac': ach / pk['tf_sustained'], 'traffic': traffic, 'traffic_note': traffic_note, 'launches_per_step': tc_n,
                'share_of_step': tc_ms / total_ms if total_ms else None,
                'step_frac_of_roofline': (value / world) * GFLOP_PER_PAIR[args.workload] / 1e3 / pk['tf_sustained'],
                'per_call_ms': {k: round(v[1], 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}
        calls_per_step = len(rec)

    cpu = None
    parity = {}
    if rank == 0 and not args.no_cpu_baseline:
        rate, cores, per = cpu_reference_rate(args.workload, args.cpu_batch, 2, 1, check=parity_check(args, dev, parity))
        cpu = {'value': rate, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
               'sample': f'2 full train steps of {args.workload} at batch {args.cpu_batch} ({per:.2f} s/step, fp32 torch CPU ops)'}

    if rank == 0:
        line = {'metric': 'training images/sec', 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
                'config': dict(workload_config(args, world, not args.no_graph),
                               l2='per-step working set (activations, several GB) far exceeds the 126 MB L2',
                               frames_per_s=2 * value),
                'e2e': {'value': e2e, 'unit': 'pairs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 8,
                        'ms_per_step': ms_e2e / args.steps},
                'gpu_launches': (launches if launches is not None else (calls_per_step or 0) * args.steps),
                'gpu_launches_note': 'C-ABI kernel-launching calls inside the timed region (each >= 1 kernel; replayed from a CUDA graph)',
                'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu, 'parity': parity or None, 'loss': loss_val,
                'activation_bytes': tr.activation_bytes(), 'params': tr.n_params}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
