"""Turn the ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (run in the build container)."""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r1'

METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
           'launch__shared_mem_per_block_dynamic', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
           'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'smsp__issue_active.avg.pct_of_peak_sustained_active',
           # shared-memory side (round 2: the 64/128-channel tcgen05 variants are bounded by it)
           'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed',
           'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max']


def launches():
    path = os.path.join(ROOT, 'gpurun_out', 'launches_graph.csv')
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 1]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    names = [r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '') for r in data]
    vals = []
    for r in data:
        v = float(r[vi].replace(',', ''))
        vals.append(v / 1e6 if r[ui] == 'ns' else v / 1e3 if r[ui] == 'us' else v)
    idx = [i for i, n in enumerate(names) if n.startswith('adam_k')]
    a, b = idx[-2] + 1, idx[-1] + 1
    agg = collections.OrderedDict()
    for n, v in zip(names[a:b], vals[a:b]):
        d = agg.setdefault(n[:70], [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(vals[a:b])
    with open(os.path.join(OUT, f'{TAG}_launches_one_step.md'), 'w') as fh:
        fh.write(f'# One graph-replayed training step, per-kernel device time ({TAG})\n\n')
        fh.write('Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 2 --warmup 3 '
                 '--no-cpu-baseline --no-kernel-timing [--no-eager-baseline --no-module-api --no-other-workloads]` (KeyNet F 128x128 K=10, batch 64, bf16).  Launches between the last two '
                 '`adam_k` launches = one step.  ncu serialises kernels and runs them cold: compare SHARES, not absolutes.\n\n')
        fh.write(f'{b - a} kernel launches, sum {tot:.3f} ms\n\n| ms | share | launches | kernel |\n|---:|---:|---:|---|\n')
        for k, d in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f'| {d[1]:.3f} | {100 * d[1] / tot:.1f}% | {d[0]} | `{k}` |\n')
        tc = sum(d[1] for k, d in agg.items() if 'tc_' in k)
        bn = sum(d[1] for k, d in agg.items() if k.startswith('bn_'))
        fh.write(f'\ntcgen05 conv family share of the step: {100 * tc / tot:.1f}%\n')
        fh.write(f'\nBatchNorm pass family (bn_*) share of the step: {100 * bn / tot:.1f}%\n')
    with open(os.path.join(OUT, f'{TAG}_launches_one_step.csv'), 'w') as fh:
        fh.write('kernel,ms\n')
        for n, v in zip(names[a:b], vals[a:b]):
            fh.write(f'"{n}",{v:.5f}\n')


def reports():
    lines = [f'# ncu --set full summaries ({TAG})\n',
             'Captured by `scripts/profile_round.sh` (`ncu --set full --clock-control none -k regex:<kernel> ... python bench.py '
             '--steps 1 --warmup 3 --no-graph ...`) on one B200; values per launch.  The raw pages are exported on the GPU box '
             '(`ncu -i <rep> --page raw --csv`) because gpurun returns at most 64 MiB.\n']
    gdir = os.path.join(ROOT, 'gpurun_out')
    for f in sorted(os.listdir(gdir)):
        if not (f.startswith('prof_') and f.endswith('.raw.csv')):
            continue
        rows = list(csv.reader(open(os.path.join(gdir, f))))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        lines.append(f'\n## {f[:-8]}\n')
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')][:110]
            lines.append(f'\n`{name}`\n\n| metric | value |\n|---|---|\n')
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    lines.append(f'| {m} | {r[i]} {units[i]} |\n')
    with open(os.path.join(OUT, f'{TAG}_ncu_full_summary.md'), 'w') as fh:
        fh.writelines(lines)


def traffic():
    """DRAM bytes of every tcgen05 conv launch of one step -> profiles/<tag>_tc_traffic.json (bench.py roofline.traffic)."""
    import json
    path = os.path.join(ROOT, 'gpurun_out', 'tc_traffic.csv')
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[hi]
    ki, mi, vi, ui, ii = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        d = per.setdefault(r[ii], {'kernel': r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')})
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(u, 1)
        d[r[mi]] = v * scale
    launches = list(per.values())
    tot_r = sum(l.get('dram__bytes_read.sum', 0) for l in launches)
    tot_w = sum(l.get('dram__bytes_write.sum', 0) for l in launches)
    tot_t = sum(l.get('gpu__time_duration.sum', 0) for l in launches)
    out = {'launches': len(launches), 'dram_bytes_read': tot_r, 'dram_bytes_write': tot_w, 'time_ms_under_ncu': tot_t,
           'bytes_per_launch': (tot_r + tot_w) / max(len(launches), 1),
           'how': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:conv_tc|wgrad_tc over the 69 tcgen05 launches '
                  'of one eager step (scripts/profile_round.sh), KeyNet F 128x128 K=10 batch 64'}
    json.dump(out, open(os.path.join(OUT, f'{TAG}_tc_traffic.json'), 'w'), indent=1)
    with open(os.path.join(OUT, f'{TAG}_tc_traffic.md'), 'w') as fh:
        fh.write(f'# DRAM traffic of the tcgen05 conv launches of one step ({TAG})\n\n{out["how"]}\n\n')
        fh.write(f'{len(launches)} launches: read {tot_r / 1e9:.3f} GB, written {tot_w / 1e9:.3f} GB, {tot_t:.3f} ms under ncu\n\n')
        fh.write('| kernel | ms | read MB | write MB |\n|---|---:|---:|---:|\n')
        for l in launches:
            fh.write(f"| `{l['kernel'][:60]}` | {l.get('gpu__time_duration.sum', 0):.4f} | {l.get('dram__bytes_read.sum', 0) / 1e6:.1f} | "
                     f"{l.get('dram__bytes_write.sum', 0) / 1e6:.1f} |\n")


def bottleneck():
    """Per-kernel time and DRAM bytes of the bottleneck / warp / loss / optimiser / thin-conv kernels of one eager step
    (scripts/profile_bottleneck.sh) -> profiles/<tag>_bottleneck_kernels.md, achieved GB/s against the measured copy bandwidth."""
    import json
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6453.1)
    out = [f'# HBM-bound kernels of the path: achieved DRAM GB/s ({TAG})\n\n',
           '`scripts/profile_bottleneck.sh`: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... '
           '--clock-control none -k regex:<bottleneck|warp|loss|adam|thin-conv kernels> python bench.py --workload <w> --steps 1 '
           '--warmup 3 --no-graph`; rows are the launches of the LAST step (between the last two `adam_k`), averaged per kernel.  '
           f'GB/s = (DRAM read + written bytes) / kernel time; peak = measured copy bandwidth {peak:.0f} GB/s (MEASURED_PEAKS.json).  '
           'Kernels whose working set fits the 126 MB L2 move few DRAM bytes by design (their inputs were just written by the '
           'producer; at 84x84 batch 64 a whole activation is 15-30 MB): for those the L2 GB/s column (lts__t_bytes / time) and the '
           'time itself are the figures of merit.\n']
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3}
    for wl, title in (('pong', 'Transporter Pong-grey 84x84 K=4, batch 64, bf16 (bottleneck kernels at full 84x84 resolution)'),
                      ('keynet', 'KeyNet F 128x128 K=10, batch 64, bf16 (TPS + rotate warps at 128x128; bottleneck at 16x16)')):
        path = os.path.join(ROOT, 'gpurun_out', f'bneck_{wl}.csv')
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
        hdr = rows[hi]
        ki, mi, vi, ui, ii = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
        per = collections.OrderedDict()
        for r in rows[hi + 1:]:
            if len(r) <= vi:
                continue
            d = per.setdefault(r[ii], {'kernel': r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')})
            d[r[mi]] = float(r[vi].replace(',', '')) * scale.get(r[ui], 1)
        ls = list(per.values())
        idx = [i for i, l in enumerate(ls) if l['kernel'].startswith('adam_k')]
        ls = ls[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else ls
        agg = collections.OrderedDict()
        for l in ls:
            a = agg.setdefault(l['kernel'][:64], [0, 0.0, 0.0, 0.0, 0.0, 0, 0.0])
            a[0] += 1
            a[1] += l.get('gpu__time_duration.sum', 0)
            a[2] += l.get('dram__bytes_read.sum', 0)
            a[3] += l.get('dram__bytes_write.sum', 0)
            a[4] += l.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0)
            a[5] = int(l.get('launch__grid_size', 0))
            a[6] += l.get('lts__t_bytes.sum', 0)
        out.append(f'\n## {title}\n\n| kernel | launches | us / launch | DRAM read MB | written MB | DRAM GB/s | frac of {peak:.0f} | L2 GB/s | ncu dram % | grid |\n'
                   '|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n')
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            n = a[0]
            gbs = (a[2] + a[3]) / max(a[1], 1e-9) / 1e3
            out.append(f'| `{k}` | {n} | {a[1] / n:.1f} | {a[2] / n / 1e6:.2f} | {a[3] / n / 1e6:.2f} | {gbs:.0f} | {gbs / peak:.2f} | {a[6] / max(a[1], 1e-9) / 1e3:.0f} | '
                       f'{a[4] / n:.0f} | {a[5]} |\n')
    if len(out) > 2:
        with open(os.path.join(OUT, f'{TAG}_bottleneck_kernels.md'), 'w') as fh:
            fh.writelines(out)


def sass():
    so = os.path.join(ROOT, 'keypoints_b200', 'lib', 'libkeypoints_b200.so')
    txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    c = collections.Counter()
    import re
    for m in re.finditer(r'\b(LDGMC[.A-Z0-9_]*|UTCHMMA[.A-Z0-9]*|UTMALDG[.A-Z0-9_]*|UTMAPF[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*|LDTM[.A-Za-z0-9_]*|USETMAXREG[.A-Z0-9_]*|HMMA[.A-Z0-9_]*|REDG[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*|UBLKCP[.A-Z0-9_]*|FFMA2[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*)', txt):
        c[m.group(1)] += 1
    with open(os.path.join(OUT, f'{TAG}_sass_evidence.md'), 'w') as fh:
        fh.write(f'# SASS mnemonics in libkeypoints_b200.so ({TAG})\n\n`cuobjdump -sass keypoints_b200/lib/libkeypoints_b200.so`\n\n| mnemonic | count |\n|---|---:|\n')
        for k, v in sorted(c.items()):
            fh.write(f'| {k} | {v} |\n')
        fh.write('\nUTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), UTMALDG = cp.async.bulk.tensor (TMA), LDTM = tcgen05.ld, '
                 '(LDTM.x16 / .x32 = 32x32b row-per-thread loads, LDTM.16dp256bit = 16x256b fragment loads of the dgrad epilogue), USETMAXREG = setmaxnreg (register hand-over between the warpgroups of the register-statistics conv tiles), UTCBAR = tcgen05.commit, UTMASTG = TMA store (conv epilogue), UBLKCP = cp.async.bulk (BatchNorm streaming kernels), '
                 'FFMA2 = packed fp32x2 math; LDGMC / STG...MC = multimem.ld_reduce / multimem.st (kp_dp.cu); HMMA (mma.sync) in the kernels '
                 'for shapes a 64-channel tcgen05 k-block cannot fill: first-layer Cin<=3 and narrow 1x1 heads (kp_conv_thin_mma.cu: '
                 'thin_mma_*, head_mma_*) and the 16/32-channel VGG_PONG layers (kp_conv_small_mma.cu: small_mma_*).\n')


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    launches()
    reports()
    traffic()
    bottleneck()
    sass()
    print(os.listdir(OUT))
