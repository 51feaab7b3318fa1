#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_network.py -q -m gpu --timeout 300 -x -k "fused or real_weights" > gpurun_out/pytest_fused.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_fused.log
timeout 200 python scripts/dwpw_trace.py 1 3 5 7 2> gpurun_out/dwpw_trace.log; grep "dwpw-trace\|====" gpurun_out/dwpw_trace.log | grep -v "cta 14\|cta 9" | head -80
for mask in 0x1 0x5 0x15 0x7d5; do
  K2Y_DWPW=1 K2Y_DWPW_MASK=$mask timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_mask_$mask.json 2> gpurun_out/bench_mask_$mask.err; echo "mask $mask exit $?"
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; echo "plain exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_mask_*.json")) + ["gpurun_out/bench_plain.json"]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "no json", e); continue
    print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "detect_ms", round(d["detect_ms"], 4))
    print("   " + "  ".join(f"{a['name'].replace('conv_','')}={a['us']}" for a in d["launch_table"] if a['us'] > 3))
PY
