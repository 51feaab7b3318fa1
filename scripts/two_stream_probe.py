"""Experiment: does running two half-batches (16 + 16 images) concurrently on two streams beat one batch of 32?
Two DetectionPipelines (own engines, own streams), steps issued alternately; device time over K rounds."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workloads as wl
from k210_yolo_framework_b200.pipeline import DetectionPipeline

cfg = dict(wl.CONFIGS[2])
torch.cuda.set_device(0)


def make(batch):
    p = DetectionPipeline(cfg["model"], cfg["in_hw"], wl.anchors(cfg), cfg["classes"], cfg["alpha"], batch, wl.OBJ_THRESH, wl.IOU_THRESH,
                          wl.MAX_PER_CLASS, device=0)
    p.engine.set_weights(wl.bench_weights(cfg, p.engine.expected_variables()))
    return p


def run(pipes, streams, batch, rounds=40):
    xs = [[torch.from_numpy(wl.synthetic_batch(cfg, 1000 + 16 * k + j, batch)).cuda() for j in range(8)] for k in range(len(pipes))]
    def issue(i):
        for k, (p, st) in enumerate(zip(pipes, streams)):
            with torch.cuda.stream(st):
                p.engine.bind_input(xs[k][i % 8])
                p.step_device(pipelined=True)
    for i in range(18):
        issue(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in streams:
        st.wait_event(e0)
    for i in range(rounds):
        issue(i)
    main = torch.cuda.current_stream()
    for p, st in zip(pipes, streams):
        with torch.cuda.stream(st):
            p.wait_gathered()
        ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / rounds
    return ms, len(pipes) * batch / ms * 1e3


one = make(32)
ms, ips = run([one], [torch.cuda.Stream()], 32)
print(f"1 x 32: {ms:.4f} ms/round  {ips:.0f} img/s", flush=True)
del one
a, b = make(16), make(16)
ms, ips = run([a, b], [torch.cuda.Stream(), torch.cuda.Stream()], 16)
print(f"2 x 16 concurrent: {ms:.4f} ms/round  {ips:.0f} img/s", flush=True)
ms, ips = run([a], [torch.cuda.Stream()], 16)
print(f"1 x 16: {ms:.4f} ms/round  {ips:.0f} img/s", flush=True)
del a, b
a, b = make(32), make(32)
ms, ips = run([a, b], [torch.cuda.Stream(), torch.cuda.Stream()], 32)
print(f"2 x 32 concurrent: {ms:.4f} ms/round  {ips:.0f} img/s", flush=True)
