#!/bin/bash
# N-GPU bench repeated (value stability)
N=${1:-2}
for r in 1 2 3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$r bench.py --gpus $N --steps ${STEPS:-50} --warmup 5 --no-cpu ${EXTRA:-} \
      > gpurun_out/bench_n${N}_r$r.json 2> gpurun_out/bench_n${N}_r$r.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_n${N}_r$r.json"))
print("run $r N=$N value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["e2e"]["windows_ms"], "3xtf32", round(d["precision_matched"]["value"]))
PY
done
