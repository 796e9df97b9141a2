#!/bin/bash
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-timing"
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_persist_k<128" -s 8 -c 2 -o gpurun_out/prof_tc128 -f $B > gpurun_out/ncu_tc128.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"thin_wgrad_k" -s 4 -c 1 -o gpurun_out/prof_thinw -f $B > gpurun_out/ncu_thinw.log 2>&1
ls -la gpurun_out | tail -5
