#!/bin/bash
# bench --no-cpu under a list of environment settings.  CASES='"A=1 B=2" "C=3"' (quoted groups), LANES (default 2)
i=0
for c in "$@"; do
  i=$((i+1))
  env $c timeout 300 python bench.py --lanes ${LANES:-2} --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_env_$i.json 2> gpurun_out/bench_env_$i.err
  python - "$c" gpurun_out/bench_env_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(f"[{sys.argv[1]}] value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "flushed", round(d["ms_per_step_flushed"], 4), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(f"[{sys.argv[1]}] failed", e)
PY
done
