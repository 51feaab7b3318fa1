#!/bin/bash
# ncu --set full over exactly one bench step (all kernels), per config.  Usage: gpu_ncu_full.sh <cfg id> [bench args]
set -u
CFG=${1:-2}; shift || true
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_full_cfg$CFG \
    python bench.py --config $CFG --warmup 3 --profile-step --no-cpu "$@" > gpurun_out/profile_step_cfg$CFG.json 2> gpurun_out/ncu_full_cfg$CFG.log
echo "ncu cfg$CFG exit $?"; tail -2 gpurun_out/ncu_full_cfg$CFG.log; ls -la gpurun_out/r02_full_cfg$CFG.ncu-rep
