#!/bin/bash
# A/B of the conv_tc_pair3_k epilogue / resident-weight variants (round 2, second session): parity tests first, then
# per-call timings (single stream, eager) and the graph-replayed step of the product library and of the experiments build
# with the resident-B path switched off.
# needs the experiments build: make -C keypoints_b200/csrc EXPERIMENTS=1 OUT=../lib/libkeypoints_b200_exp.so BUILD=../_build_exp
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py -x -q -k "tcgen05" > gpurun_out/r3_test_tc.log 2>&1
tail -3 gpurun_out/r3_test_tc.log
python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core or fused_trainer or bf16" > gpurun_out/r3_test_par.log 2>&1
tail -3 gpurun_out/r3_test_par.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-module-api --no-other-workloads"
KP_BENCH_CALLS=gpurun_out/r3_calls_new.txt KP_WGRAD_STREAM=0 KP_TWO_STREAMS=0 $B > gpurun_out/r3_b_new_1s.log 2>&1
KP_LIB=$PWD/keypoints_b200/lib/libkeypoints_b200_exp.so KP_TC_RESB=0 KP_BENCH_CALLS=gpurun_out/r3_calls_resb0.txt KP_WGRAD_STREAM=0 KP_TWO_STREAMS=0 $B > gpurun_out/r3_b_resb0_1s.log 2>&1
for i in 1 2; do
  $B --no-kernel-timing > gpurun_out/r3_b_new_$i.log 2>&1
  KP_LIB=$PWD/keypoints_b200/lib/libkeypoints_b200_exp.so KP_TC_RESB=0 $B --no-kernel-timing > gpurun_out/r3_b_resb0_$i.log 2>&1
done
grep -h -E "kp_conv_tc .*(64->128|128->64|256->128)@128" gpurun_out/r3_calls_new.txt
echo ---
grep -h -E "kp_conv_tc .*(64->128|128->64|256->128)@128" gpurun_out/r3_calls_resb0.txt
for f in gpurun_out/r3_b_*.log; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"frac": [0-9.]*' $f | head -1)"; done
