#!/bin/bash
# ncu captures for profiles/ (run under gpurun, one GPU).  Numbers printed by bench.py under ncu are not bench values.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-timing"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 460 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_k -s 40 -c 3 -o gpurun_out/prof_conv_tc -f $B > gpurun_out/ncu_bench2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_k -s 20 -c 3 -o gpurun_out/prof_wgrad_tc -f $B > gpurun_out/ncu_bench3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bn_act_bwd_reduce_k -s 30 -c 3 -o gpurun_out/prof_bn_bwd -f $B > gpurun_out/ncu_bench4.log 2>&1
ls -la gpurun_out
