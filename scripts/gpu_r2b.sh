#!/bin/bash
# detect re-test + fused-kernel evaluation: default vs K2Y_DWPW=1, with per-layer launch tables
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_detect.py tests/test_gpu_fullsize.py tests/test_gpu_pipeline.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_detect.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_detect.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; echo "plain exit $?"
K2Y_DWPW=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_dwpw.json 2> gpurun_out/bench_dwpw.err; echo "dwpw exit $?"
python - <<'PY'
import json
for tag in ("plain", "dwpw"):
    try:
        d = json.load(open(f"gpurun_out/bench_{tag}.json"))
    except Exception as e:
        print(tag, "no json", e); continue
    print(tag, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "detect_ms", round(d["detect_ms"], 4), "parity", d.get("parity_checked"))
    print("   " + "  ".join(f"{a['name']}={a['us']}" for a in d["launch_table"]))
PY
