#!/bin/bash
# compute-sanitizer over the tcgen05 kernels (warp-specialised mbarrier / TMEM pipelines, cluster multicast, fused dw->pw) and
# the decode/NMS kernels.  Summaries -> gpurun_out/sanitizer_*.log (copied to profiles/r02_sanitizer.md by hand).
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 1200 $SAN --tool $tool --print-limit 20 --launch-timeout 0 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu -x --timeout 1100 \
      > gpurun_out/sanitizer_${tool}_conv_tc.log 2>&1
  echo "$tool conv_tc exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_conv_tc.log | tail -3
  K2Y_DWPW=1 timeout 1200 $SAN --tool $tool --print-limit 20 --launch-timeout 0 python -m pytest tests/test_gpu_network.py -q -m gpu -x --timeout 1100 \
      -k "fused_depthwise" > gpurun_out/sanitizer_${tool}_dwpw.log 2>&1
  echo "$tool dwpw exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_dwpw.log | tail -3
  timeout 900 $SAN --tool $tool --print-limit 20 --launch-timeout 0 python -m pytest tests/test_gpu_detect.py -q -m gpu -x --timeout 800 \
      -k "seeded or ties or strided" > gpurun_out/sanitizer_${tool}_detect.log 2>&1
  echo "$tool detect exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_detect.log | tail -3
done
