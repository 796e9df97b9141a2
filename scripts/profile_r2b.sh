#!/bin/bash
# Round-2 (second session) profile set, ONE GPU.  Numbers printed by bench.py under ncu are never bench values.
#   1. launch list of one graph-replayed step
#   2. DRAM traffic of every tcgen05 conv launch of one eager step (roofline.traffic)
#   3. ncu --set full of the kernels touched or targeted this session: the register-statistics / resident-weight conv
#      variants, the 256-wide conv (fprop with statistics), the first-layer and head kernels, weight pack and gradient fold
set -x
mkdir -p gpurun_out
if [ -x ./scripts/_bin/exp_tmem_ld_shapes ]; then ./scripts/_bin/exp_tmem_ld_shapes > gpurun_out/r2b_tmem_ld_shapes.txt 2>&1; fi
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-timing --no-eager-baseline --no-module-api --no-other-workloads"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph.csv $B > gpurun_out/ncu_launches.log 2>&1
E="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-timing --no-eager-baseline --no-module-api --no-other-workloads"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_tc_pair|wgrad_tc_pair_k|conv_tc_persist_k|wgrad_tc_persist_k|wgrad_tc_rows_k" -s 207 -c 69 --csv --log-file gpurun_out/tc_traffic.csv $E > gpurun_out/ncu_traffic.log 2>&1
X="--set full --clock-control none --import-source on --kernel-name-base demangled --metrics l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max"
ncu $X -k regex:'conv_tc_pair3_k<\(int\)128, \(int\)5' -s 4 -c 2 -o gpurun_out/prof_conv_rs128 -f $E > gpurun_out/ncu_q1.log 2>&1
ncu $X -k regex:'conv_tc_pair3_k<\(int\)64' -s 2 -c 2 -o gpurun_out/prof_conv_rs64 -f $E > gpurun_out/ncu_q2.log 2>&1
ncu $X -k regex:'conv_tc_pair3_k<\(int\)256' -s 20 -c 1 -o gpurun_out/prof_conv_256_fprop -f $E > gpurun_out/ncu_q3.log 2>&1
ncu $X -k regex:'conv_tc_pair3_k<\(int\)256' -s 52 -c 1 -o gpurun_out/prof_conv_256_dgrad -f $E > gpurun_out/ncu_q4.log 2>&1
if [ -z "$CONV_ONLY" ]; then   # the first-layer / head / optimiser kernels (unchanged since the first capture of the session)
ncu $X -k regex:thin_mma_fprop_k -s 6 -c 1 -o gpurun_out/prof_thin_fprop -f $E > gpurun_out/ncu_q5.log 2>&1
ncu $X -k regex:thin_mma_wgrad_k -s 6 -c 1 -o gpurun_out/prof_thin_wgrad -f $E > gpurun_out/ncu_q6.log 2>&1
ncu $X -k regex:'head1x1_wgrad_k|head1x1_dgrad_k|head_mma_fprop_k' -s 9 -c 3 -o gpurun_out/prof_head -f $E > gpurun_out/ncu_q7.log 2>&1
ncu $X -k regex:'pack_weights_multi_k|wgrad_finalize_multi_k|adam_k' -s 9 -c 3 -o gpurun_out/prof_opt -f $E > gpurun_out/ncu_q8.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep
for r in gpurun_out/prof_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null
  ncu -i $r --page details --csv > ${r%.ncu-rep}.details.csv 2>/dev/null
done
for r in thin_fprop thin_wgrad head; do
  if [ -f gpurun_out/prof_$r.ncu-rep ]; then ncu -i gpurun_out/prof_$r.ncu-rep --page source --csv > gpurun_out/prof_$r.source.csv 2>/dev/null; fi
done
for r in gpurun_out/prof_*.ncu-rep; do
  if [ $(stat -c %s $r) -gt 5000000 ]; then rm -f $r; fi
done
ls -la gpurun_out | tail -40
