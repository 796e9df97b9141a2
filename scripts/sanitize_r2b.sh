#!/bin/bash
# compute-sanitizer (memcheck) over the conv_tc_pair3_k variants of the second round-2 session (register-resident BatchNorm
# statistics with setmaxnreg, resident weight tiles) through the small-shape tensor-core tests.  One GPU.
set -x
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2b_sanitizer_raw.log python -m pytest tests/test_gpu_parity.py::test_tensor_core_conv_vs_fp32 -m gpu -q -x --timeout 1200 > gpurun_out/r2b_sanitizer_pytest.log 2>&1
echo "exit $?" >> gpurun_out/r2b_sanitizer_pytest.log
tail -5 gpurun_out/r2b_sanitizer_pytest.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2b_sanitizer_raw.log
tail -5 gpurun_out/r2b_sanitizer_raw.log
