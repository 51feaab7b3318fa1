"""N > 1 on real GPUs: the records gathered by ONE ncclAllGather (k2y_allgather_detections) equal the single-GPU records
of the same images, in global image order.  Needs >= 2 visible GPUs (`gpurun --gpus 2`); skipped on a 1-GPU box."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import ROOT

WORKER = r'''
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["K2Y_ROOT"])
import bench_workloads as wl
from k210_yolo_framework_b200.pipeline import DetectionPipeline
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg, B = wl.CONFIGS[2], 4
def make(world_, rank_, batch):
    p = DetectionPipeline(cfg["model"], cfg["in_hw"], wl.anchors(cfg), cfg["classes"], cfg["alpha"], batch, wl.OBJ_THRESH,
                          wl.IOU_THRESH, wl.MAX_PER_CLASS, device=local, world=world_, rank=rank_)
    p.engine.set_weights(wl.bench_weights(cfg, p.engine.expected_variables()))
    return p
x_all = wl.synthetic_batch(cfg, 77, world * B)
pipe = make(world, rank, B)
outs = []
for rep in range(3):                                  # three steps: both gather slots and the slot-reuse wait are exercised
    xs = np.roll(x_all, rep, axis=0)
    d, c = pipe.detect_host(torch.from_numpy(xs[rank * B:(rank + 1) * B].copy()).pin_memory())
    outs.append(DetectionPipeline.records(d.clone(), c.clone()))
ok, why = True, ""
if rank == 0:
    # reference: ONE GPU, the same per-rank batch size (same tile plans, same accumulation order -> same bits), shard by shard
    single = make(1, 0, B)
    for rep in range(3):
        xs = np.roll(x_all, rep, axis=0)
        ref = []
        for r in range(world):
            d, c = single.detect_host(torch.from_numpy(xs[r * B:(r + 1) * B].copy()).pin_memory())
            ref += DetectionPipeline.records(d.clone(), c.clone())
        if ref != outs[rep]:
            bad = [i for i, (a, b) in enumerate(zip(ref, outs[rep])) if a != b]
            why += f" rep {rep}: images {bad} differ ({[len(ref[i]) for i in bad]} vs {[len(outs[rep][i]) for i in bad]} records);"
            ok = False
        ok &= sum(len(i) for i in ref) > 50
    print("K2Y_MULTI " + json.dumps({"ok": bool(ok), "world": world, "nccl": pipe.gather.comm.nccl_version(), "why": why}), flush=True)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.timeout(600)
def test_gathered_records_equal_single_gpu(tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, K2Y_ROOT=ROOT, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", str(script)], capture_output=True, text=True, timeout=540, env=env)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    line = [l for l in r.stdout.splitlines() if l.startswith("K2Y_MULTI ")]
    assert line, r.stdout[-2000:]
    res = json.loads(line[0][len("K2Y_MULTI "):])
    assert res["ok"] and res["world"] == world, res
