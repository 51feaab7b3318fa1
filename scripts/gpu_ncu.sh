#!/bin/bash
# ncu --set full captures of selected kernels of one bench step.  Usage: gpu_ncu.sh <tag> <kernel-regex> <skip> <count> [bench args]
set -u
mkdir -p gpurun_out
TAG=$1; REGEX=$2; SKIP=$3; COUNT=$4; shift 4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c $COUNT -f -o gpurun_out/$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu "$@" > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu $TAG exit $?"; tail -2 gpurun_out/ncu_$TAG.log
