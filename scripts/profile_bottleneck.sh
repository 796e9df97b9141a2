#!/bin/bash
# Bottleneck / warp / loss / optimiser kernels (the HBM-bound side of the path): per-launch time and DRAM bytes of one eager
# step, at the Pong-grey 84x84 K=4 batch-64 workload (bottleneck at full resolution) and at KeyNet F 128x128 K=10 batch 64
# (TPS + rotate warps).  Run under gpurun, ONE GPU.  Summarised by scripts/summarize_profiles.py -> profiles/<tag>_bottleneck_kernels.md
set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size
K='regex:ssm_|gaussian_|transport_|concat|l2_loss_k|tps_warp_k|rotate_warp_k|adam_k|c1_|small_mma_|head1x1_|thin_|bn_'
E="--steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-timing"
ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/bneck_pong.csv python bench.py --workload transporter_pong_84_K4 $E > gpurun_out/ncu_bneck_pong.log 2>&1
ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/bneck_keynet.csv python bench.py $E > gpurun_out/ncu_bneck_keynet.log 2>&1
ls -la gpurun_out/bneck_*
