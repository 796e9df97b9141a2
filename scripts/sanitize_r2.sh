#!/bin/bash
# compute-sanitizer (memcheck) over the kernels added in round 2, through the small-shape tests that launch them:
# row-stacked wgrad (8x8 / 12x20 layers), aug draw + constant-one warp, loss ring, u8 conversion + prefetcher, nvJPEG resize,
# fused trainer for the auto-encoder and the loop / sum combine modes, kp_ctx map cache.  One GPU.
set -x
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_tensor_core_conv_vs_fp32 tests/test_gpu_widening.py::test_aug_draw_distribution_and_replay tests/test_gpu_widening.py::test_loss_ring_and_plateau_scheduler tests/test_gpu_widening.py::test_fused_trainer_autoencoder_vs_reference_golden tests/test_gpu_widening.py::test_fused_trainer_combine_modes_vs_reference_golden tests/test_gpu_widening.py::test_ctx_tensor_map_cache_and_sm_limit tests/test_loader.py::test_prefetcher_feeds_the_trainer_from_uint8_frames tests/test_loader.py::test_jpeg_decode_and_resize_match_the_reference_transform"
compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2_sanitizer_raw.log python -m pytest $T -m gpu -q -x --timeout 1200 > gpurun_out/r2_sanitizer_pytest.log 2>&1
echo "exit $?" >> gpurun_out/r2_sanitizer_pytest.log
tail -5 gpurun_out/r2_sanitizer_pytest.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2_sanitizer_raw.log
tail -5 gpurun_out/r2_sanitizer_raw.log
