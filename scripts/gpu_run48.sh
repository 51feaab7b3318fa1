SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $SAN --tool racecheck --print-limit 20 --launch-timeout 0 python -m pytest tests/test_gpu_detect.py -q -m gpu -x --timeout 800 > gpurun_out/sanitizer2_racecheck_detect.log 2>&1
echo "racecheck detect exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer2_racecheck_detect.log | tail -2
timeout 600 python -m pytest tests/test_gpu_detect.py tests/test_gpu_fullsize.py -q -m gpu --timeout 500 -x 2>&1 | tail -2
K2Y_PROBE_CFG=3 K2Y_PROBE_HOSTS=6 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -8
