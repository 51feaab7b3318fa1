#!/bin/bash
# bench lines of the other BASELINE configs (cfg 3, 4, 5) on one GPU -> gpurun_out/bench_cfgN.json
set -u
mkdir -p gpurun_out
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err
  echo "cfg $c exit $?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_cfg$c.json"))
    print("cfg $c", d["config"]["workload"][:70], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1),
          "parity", d.get("parity_checked"), "cpu", round(d["cpu_baseline"]["value"], 2), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
except Exception as e:
    print("cfg $c: no json", e)
PY
done
