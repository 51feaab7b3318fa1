#!/bin/bash
# end-of-round record: the driver's own sequence (pytest -m gpu, smoke, bench with default flags, reference arm) + cfg 3/4/5 lines
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"
bash scripts/gpu_cfgs.sh
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("cfg2 value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d["parity_checked"], "flushed", d["ms_per_step_flushed"], "clocks", d["clocks"])
r = json.load(open("gpurun_out/bench_reference.json"))
print("reference arm", round(r["value"], 2), r["unit"], r["cpu_baseline"]["cores"], "cores")
PY
