#!/bin/bash
# Fragment epilogue: KP_TC_FE = 3 (default: everywhere but the 64 -> 128 fprop) vs 1 (dgrad only), with and without the
# co-resident BatchNorm hint (the 256-wide fragment tiles use 162 registers: a 288-thread BatchNorm CTA no longer fits beside them).
# needs the experiments build: make -C keypoints_b200/csrc EXPERIMENTS=1 OUT=../lib/libkeypoints_b200_exp.so BUILD=../_build_exp
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -k "tcgen05 or tensor_core or fused_trainer or normalised" > gpurun_out/r3_test_fe.log 2>&1; tail -3 gpurun_out/r3_test_fe.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-module-api --no-other-workloads --no-kernel-timing"
X="KP_LIB=$PWD/keypoints_b200/lib/libkeypoints_b200_exp.so"
for i in 1 2 3 4; do
  env $X KP_TC_FE=3 KP_BN_CORESIDENT=1 $B > gpurun_out/r3_b_f3c1_$i.log 2>&1
  env $X KP_TC_FE=3 KP_BN_CORESIDENT=0 $B > gpurun_out/r3_b_f3c0_$i.log 2>&1
  env $X KP_TC_FE=1 KP_BN_CORESIDENT=1 $B > gpurun_out/r3_b_f1c1_$i.log 2>&1
  env $X KP_TC_FE=0 KP_BN_CORESIDENT=1 $B > gpurun_out/r3_b_f0c1_$i.log 2>&1
done
for f in gpurun_out/r3_b_f?c?_*.log; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1)"; done
