#!/bin/bash
# N-GPU bench under environment / flag variants: args are "ENV1=a ENV2=b -- flags" strings
N=${1:-2}; shift
i=0
for spec in "$@"; do
  i=$((i+1))
  envs="${spec%%--*}"; flags="${spec#*--}"
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$i bench.py --gpus $N --steps 50 --warmup 5 --no-cpu $flags \
      > gpurun_out/bench_n${N}_e$i.json 2> gpurun_out/bench_n${N}_e$i.err
  python - "$spec" gpurun_out/bench_n${N}_e$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(f"[{sys.argv[1]}] value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(f"[{sys.argv[1]}] failed", e)
PY
done
