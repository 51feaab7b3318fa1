#!/bin/bash
# one N-GPU bench line -> gpurun_out/bench_nN.json
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$N.json"))
print("N=$N value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["e2e"]["windows_ms"])
PY
tail -2 gpurun_out/bench_n$N.err
