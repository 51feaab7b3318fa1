"""Where does an end-to-end step spend its host time?  Times submit() and collect() separately over the streaming loop
of bench.py (cfg 2, uint8 pinned host batches) and prints per-step averages beside the device-resident step time."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as wl
from k210_yolo_framework_b200.pipeline import DetectionPipeline

cfg = wl.CONFIGS[int(os.environ.get('K2Y_PROBE_CFG', '2'))]
B = cfg["batch"]
torch.cuda.set_device(0)
pipe = DetectionPipeline(cfg["model"], cfg["in_hw"], wl.anchors(cfg), cfg["classes"], cfg["alpha"], B, wl.OBJ_THRESH, wl.IOU_THRESH,
                         wl.MAX_PER_CLASS, device=0)
pipe.engine.set_weights(wl.bench_weights(cfg, pipe.engine.expected_variables()))
hosts = [torch.from_numpy(wl.synthetic_batch_u8(cfg, 2000 + j)).pin_memory() for j in range(int(os.environ.get('K2Y_PROBE_HOSTS', '24')))]
for j in range(6):
    pipe.collect(pipe.submit(hosts[j % len(hosts)]))
torch.cuda.synchronize()
for trial in range(3):
    steps = 60
    ts, tc = 0.0, 0.0
    t0 = time.perf_counter()
    prev = None
    for i in range(steps):
        a = time.perf_counter()
        tk = pipe.submit(hosts[i % len(hosts)])
        b = time.perf_counter()
        if prev is not None:
            pipe.collect(prev)
        c = time.perf_counter()
        ts += b - a
        tc += c - b
        prev = tk
    pipe.collect(prev)
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print(f"trial {trial}: step {1e3 * tot / steps:.4f} ms  submit {1e3 * ts / steps:.4f} ms  collect-wait {1e3 * tc / steps:.4f} ms  "
          f"-> {B * steps / tot:.0f} img/s", flush=True)
# device-resident, for comparison
xs = [h.cuda() for h in hosts[:min(20, len(hosts))]]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for x in xs:
    pipe.engine.bind_input(x); pipe.step_device()
torch.cuda.synchronize()
e0.record()
for i in range(60):
    pipe.engine.bind_input(xs[i % len(xs)]); pipe.step_device()
e1.record(); torch.cuda.synchronize()
print(f"device-resident u8 step {e0.elapsed_time(e1) / 60:.4f} ms")

# the same step from float32 device inputs, and the first conv alone in both forms (eager profile)
xf = [torch.from_numpy(wl.synthetic_batch(cfg, 3000 + j)).cuda() for j in range(min(4, len(hosts)))]
for x in xf:
    pipe.engine.bind_input(x); pipe.step_device()
torch.cuda.synchronize()
e0.record()
for i in range(40):
    pipe.engine.bind_input(xf[i % len(xf)]); pipe.step_device()
e1.record(); torch.cuda.synchronize()
print(f"device-resident f32 step {e0.elapsed_time(e1) / 40:.4f} ms")
for name, x in (("f32", xf[0]), ("u8", xs[0])):
    pipe.engine.bind_input(x)
    prof = pipe.engine.profile(B)
    print(name, "first launches:", [(a["name"], round(1e3 * a["ms"], 1)) for a in prof[:3]])
