#!/bin/bash
# Fragment epilogue (tcgen05.ld.16x256b) for the dgrad launches (KP_TC_FE=1, the default) vs never (0) vs always (2).
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -k "tcgen05 or tensor_core or fused_trainer or normalised" > gpurun_out/r3_test_fe.log 2>&1; tail -3 gpurun_out/r3_test_fe.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-module-api --no-other-workloads"
X="KP_LIB=$PWD/keypoints_b200/lib/libkeypoints_b200_exp.so"
for i in 1 2 3 4; do
  env $X KP_TC_FE=1 $B --no-kernel-timing > gpurun_out/r3_b_fe1_$i.log 2>&1
  env $X KP_TC_FE=0 $B --no-kernel-timing > gpurun_out/r3_b_fe0_$i.log 2>&1
done
$B > gpurun_out/r3_b_fe1_1s.log 2>&1
for f in gpurun_out/r3_b_fe*.log; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"frac": [0-9.]*' $f | head -1)"; done
