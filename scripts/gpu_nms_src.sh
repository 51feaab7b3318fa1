#!/bin/bash
# source-level ncu capture of the NMS kernel (one launch) + the end-to-end host-time probe
set -u
mkdir -p gpurun_out
timeout 300 python scripts/e2e_probe.py 2> gpurun_out/e2e_probe.err | tee gpurun_out/e2e_probe.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:detect_ --profile-from-start off -o gpurun_out/nms_src -f python bench.py --profile-step --no-cpu > gpurun_out/nms_src.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/nms_src.ncu-rep
