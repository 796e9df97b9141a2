#!/bin/bash
# Round profile set (run under gpurun, ONE GPU).  Numbers printed by bench.py under ncu are never bench values.
#   1. launch list of one graph-replayed step (per-kernel device time, cold-cache and serialised: compare SHARES)
#   2. DRAM traffic of every tcgen05 conv launch of one (eager) step -> roofline.traffic in bench.py
#   3. ncu --set full of the dominant kernels (one or two launches each)
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-timing"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph.csv $B > gpurun_out/ncu_launches.log 2>&1
E="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-timing"
# warm-up = 3 eager steps + 1 timed step; the family launches 69 + 23 finalize kernels per step: skip the first 3 steps
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_tc_pair|wgrad_tc_pair_k|conv_tc_persist_k|wgrad_tc_persist_k" -s 207 -c 69 --csv --log-file gpurun_out/tc_traffic.csv $E > gpurun_out/ncu_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_pair3_k -s 40 -c 2 -o gpurun_out/prof_conv_pair3 -f $E > gpurun_out/ncu_p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_pair_k -s 14 -c 2 -o gpurun_out/prof_wgrad_pair -f $E > gpurun_out/ncu_p2.log 2>&1
ncu --set full --clock-control none -k regex:bn_bwd_none_pipe_k -s 24 -c 2 -o gpurun_out/prof_bn_bwd_none_pipe -f $E > gpurun_out/ncu_p3.log 2>&1
ncu --set full --clock-control none -k regex:bn_fwd_none_pipe_k -s 24 -c 1 -o gpurun_out/prof_bn_fwd_none_pipe -f $E > gpurun_out/ncu_p4.log 2>&1
ncu --set full --clock-control none -k regex:bn_bwd_up_pipe_k -s 6 -c 1 -o gpurun_out/prof_bn_bwd_up_pipe -f $E > gpurun_out/ncu_p5.log 2>&1
ncu --set full --clock-control none -k regex:bn_bwd_pool_pipe_k -s 6 -c 1 -o gpurun_out/prof_bn_bwd_pool_pipe -f $E > gpurun_out/ncu_p6.log 2>&1
ls -la gpurun_out/*.ncu-rep
# gpurun merges at most 64 MiB back: keep raw / details CSV pages of every report, drop reports above 8 MB
for r in gpurun_out/prof_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null
  ncu -i $r --page details --csv > ${r%.ncu-rep}.details.csv 2>/dev/null
  if [ $(stat -c %s $r) -gt 8000000 ]; then rm -f $r; fi
done
ls -la gpurun_out
