#!/bin/bash
# N-GPU checks (gpurun --gpus N): gathered records == single-GPU records, then the scaling bench lines.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 580 > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi exit $?"; tail -6 gpurun_out/pytest_multi.log
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N exit $?"
python - <<PY
import json
for n in (1, $N):
    try:
        d = json.load(open(f"gpurun_out/bench_n{n}.json"))
        print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]))
    except Exception as e:
        print(n, "no json", e)
PY
tail -5 gpurun_out/bench_n$N.err
