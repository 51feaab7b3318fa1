#!/bin/bash
# compute-sanitizer over what changed late in round 2: the residual-by-TMA epilogue and two-taps-per-k-block gather of the
# tcgen05 conv (memcheck), the chunked NMS path with named barriers (memcheck + racecheck), the laned pipeline (memcheck)
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $SAN --tool memcheck --print-limit 20 --launch-timeout 0 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu -x --timeout 1100 > gpurun_out/sanitizer2_memcheck_conv_tc.log 2>&1
echo "memcheck conv_tc exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer2_memcheck_conv_tc.log | tail -2
for tool in memcheck racecheck; do
  timeout 900 $SAN --tool $tool --print-limit 20 --launch-timeout 0 python -m pytest tests/test_gpu_detect.py -q -m gpu -x --timeout 800 > gpurun_out/sanitizer2_${tool}_detect.log 2>&1
  echo "$tool detect exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer2_${tool}_detect.log | tail -2
done
timeout 900 $SAN --tool memcheck --print-limit 20 --launch-timeout 0 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -x --timeout 800 -k "laned or pipelined" > gpurun_out/sanitizer2_memcheck_lanes.log 2>&1
echo "memcheck lanes exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer2_memcheck_lanes.log | tail -2
