#!/bin/bash
# Round profile set (run under gpurun, ONE GPU).  Numbers printed by bench.py under ncu are never bench values.
#   1. launch list of one graph-replayed step (per-kernel device time, cold-cache and serialised: compare SHARES)
#   2. ncu --set full of the dominant kernels (one launch each)
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-timing"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph.csv $B > gpurun_out/ncu_launches.log 2>&1
E="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-timing"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_pair_k -s 40 -c 2 -o gpurun_out/prof_conv_pair -f $E > gpurun_out/ncu_p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_pair_k -s 14 -c 2 -o gpurun_out/prof_wgrad_pair -f $E > gpurun_out/ncu_p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bn_act_bwd_rows_k -s 24 -c 2 -o gpurun_out/prof_bn_bwd -f $E > gpurun_out/ncu_p3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bn_bwd_apply_rows_k -s 24 -c 1 -o gpurun_out/prof_bn_apply -f $E > gpurun_out/ncu_p4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_resident_k -s 6 -c 1 -o gpurun_out/prof_conv_resident -f $E > gpurun_out/ncu_p5.log 2>&1
ls -la gpurun_out/*.ncu-rep
