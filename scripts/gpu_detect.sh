#!/bin/bash
# decode/NMS iteration loop: parity tests, a short bench line, per-launch times of the two detect kernels under ncu
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_detect.py tests/test_gpu_fullsize.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_detect.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_detect.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_detect.json 2> gpurun_out/bench_detect.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_detect.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "detect_ms", round(d["detect_ms"], 4), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_checked"))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:detect_ --profile-from-start off --csv --log-file gpurun_out/detect_launches.csv python bench.py --profile-step --no-cpu > /dev/null 2>&1
grep -E "detect_" gpurun_out/detect_launches.csv | awk -F'","' '{print $5, $NF}' | head
