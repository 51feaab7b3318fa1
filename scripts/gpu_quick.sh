#!/bin/bash
# Quick perf iteration: network parity tests + bench without the CPU leg + per-layer trace.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_network.py -q -x --tb=short --timeout 300 > gpurun_out/pytest_net.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_net.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "det_ms", round(d["detect_ms"], 4))
print(d["top_launches"])
PY
