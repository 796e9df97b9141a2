#!/bin/bash
# A/B of this session's kernel work on one box: library built from the session-start commit (e7a806f, KP_LIB) against the
# current one, same host code, 3 interleaved graph-replayed runs per workload.
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-module-api --no-other-workloads --no-kernel-timing"
for w in keynet_F_128_K10 transporter_F_128_K30 keynet_F_256_K64 transporter_pong_84_K4; do
for i in 1 2 3; do
  $B --workload $w > gpurun_out/r3_s_new_${w}_$i.log 2>&1
  KP_LIB=$PWD/keypoints_b200/lib/libkeypoints_b200_old.so $B --workload $w > gpurun_out/r3_s_old_${w}_$i.log 2>&1
done
done
for f in gpurun_out/r3_s_*.log; do echo "$f $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"ms_per_step": [0-9.]*' $f | head -1)"; done
