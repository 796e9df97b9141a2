#!/bin/bash
# Round-2 profile set (run under gpurun, ONE GPU).  Numbers printed by bench.py under ncu are never bench values.
#   1. launch list of one graph-replayed step (per-kernel device time, cold-cache and serialised: compare SHARES)
#   2. DRAM traffic of every tcgen05 conv launch of one (eager) step -> roofline.traffic in bench.py
#   3. ncu --set full (+ shared-memory pipe counters) of the tcgen05 variants VERDICT r1 asked about: the 64/128-channel
#      128x128 convs, the old one-tap 64-channel wgrad (KP_WGRAD_ROWS=0) and the row-stacked wgrad that replaced it
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-timing --no-eager-baseline --no-module-api --no-other-workloads"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph.csv $B > gpurun_out/ncu_launches.log 2>&1
E="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-timing --no-eager-baseline --no-module-api --no-other-workloads"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_tc_pair|wgrad_tc_pair_k|conv_tc_persist_k|wgrad_tc_persist_k|wgrad_tc_rows_k" -s 207 -c 69 --csv --log-file gpurun_out/tc_traffic.csv $E > gpurun_out/ncu_traffic.log 2>&1
X="--set full --clock-control none --import-source on --kernel-name-base demangled --metrics l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_uniform.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max"
ncu $X -k regex:'conv_tc_pair3_k<\(int\)128' -s 8 -c 2 -o gpurun_out/prof_conv_pair3_128 -f $E > gpurun_out/ncu_q1.log 2>&1
ncu $X -k regex:'conv_tc_pair3_k<\(int\)64' -s 2 -c 1 -o gpurun_out/prof_conv_pair3_64 -f $E > gpurun_out/ncu_q2.log 2>&1
ncu $X -k regex:'conv_tc_pair3_k<\(int\)256' -s 20 -c 1 -o gpurun_out/prof_conv_pair3_256 -f $E > gpurun_out/ncu_q3.log 2>&1
ncu $X -k regex:wgrad_tc_rows_k -s 3 -c 2 -o gpurun_out/prof_wgrad_rows -f $E > gpurun_out/ncu_q4.log 2>&1
KP_WGRAD_ROWS=0 ncu $X -k regex:'wgrad_tc_persist_k<\(int\)64' -s 3 -c 2 -o gpurun_out/prof_wgrad_persist64_old -f $E > gpurun_out/ncu_q5.log 2>&1
ncu $X -k regex:wgrad_tc_pair_k -s 14 -c 1 -o gpurun_out/prof_wgrad_pair -f $E > gpurun_out/ncu_q6.log 2>&1
ncu --set full --clock-control none -k regex:bn_bwd_none_pipe_k -s 24 -c 2 -o gpurun_out/prof_bn_bwd_none_pipe -f $E > gpurun_out/ncu_q7.log 2>&1
ncu --set full --clock-control none -k regex:bn_fwd_up_pipe_k -s 6 -c 1 -o gpurun_out/prof_bn_fwd_up_pipe -f $E > gpurun_out/ncu_q8.log 2>&1
ls -la gpurun_out/*.ncu-rep
for r in gpurun_out/prof_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null
  ncu -i $r --page details --csv > ${r%.ncu-rep}.details.csv 2>/dev/null
  if [ $(stat -c %s $r) -gt 6000000 ]; then rm -f $r; fi
done
ls -la gpurun_out | tail -30
