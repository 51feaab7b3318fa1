"""Full-size parity on the BASELINE.json workloads (the bench's own weights and inputs, through DetectionPipeline):

  (1) identical post-NMS (class, box index) lists for EVERY image against the oracle decode + NMS run on the GPU's own head
      tensors — the bit-identical-survivor-set requirement, independent of the network's arithmetic;
  (2) decoded confidences and normalised box coordinates of ALL boxes within 1e-3 of the oracle's own forward pass.

cfg 2, 3 and 4 run their full per-GPU batch; cfg 5 (Darknet-53, 608x608, 80 classes, 22 743 boxes per image) runs two
images so that the CPU oracle stays within a minute."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import bench_workloads as wl
from k210_yolo_framework_b200.pipeline import DetectionPipeline
from oracle import decode_ref, keras_ref

CASES = [(2, None, None), (3, None, 16), (4, None, None), (5, 2, 1)]   # (config, batch override, images with a CPU forward)


@pytest.mark.parametrize("cfg_id,batch,n_forward", CASES)
def test_bench_workload_parity(cfg_id, batch, n_forward):
    cfg = dict(wl.CONFIGS[cfg_id])
    B = batch or cfg["batch"]
    hw = cfg["in_hw"]
    pipe = DetectionPipeline(cfg["model"], hw, wl.anchors(cfg), cfg["classes"], cfg["alpha"], B, wl.OBJ_THRESH, wl.IOU_THRESH,
                             wl.MAX_PER_CLASS)
    weights = wl.bench_weights(cfg, pipe.engine.expected_variables())
    pipe.engine.set_weights(weights)
    x = wl.synthetic_batch(cfg, 1000, B)                      # rank 0's first bench batch
    dets, counts = pipe.detect_host(torch.from_numpy(x).pin_memory())
    got = DetectionPipeline.records(dets.clone(), counts.clone())
    heads_gpu = [t[:B].cpu().numpy() for t in pipe.engine.head_buffers]
    A, C = 3, cfg["classes"]
    helper = decode_ref.HelperRef(wl.anchors(cfg), list(hw), wl.out_hw(cfg), C)
    # (1) survivor sets on identical head tensors, every image
    ref = decode_ref.detect_batch_fast(heads_gpu, helper, list(hw), [hw] * B, wl.OBJ_THRESH, wl.IOU_THRESH, wl.MAX_PER_CLASS)
    n_det = 0
    for b in range(B):
        assert [(d[0], d[1]) for d in got[b]] == [(d[0], d[1]) for d in ref[b]], f"cfg{cfg_id} image {b}: survivor sets differ"
        if ref[b]:
            np.testing.assert_array_equal(np.array([d[2] for d in got[b]], np.float32), np.array([d[2] for d in ref[b]], np.float32))
            np.testing.assert_allclose(np.array([d[3:] for d in got[b]]), np.array([[float(v) for v in d[3:]] for d in ref[b]]),
                                       rtol=1e-6, atol=1e-4)
        n_det += len(ref[b])
    assert n_det > 10 * B, "the workload is meant to be detection-rich"
    # (2) decoded quantities against the oracle's own forward pass (fp32 CPU restatement of the reference graph)
    nf = min(B, n_forward or B)
    heads_ref = keras_ref.forward(cfg["model"], weights, x[:nf], alpha=cfg["alpha"])
    norm = np.array([hw[0], hw[1], hw[0], hw[1]], np.float64)
    for b in range(nf):
        gb, gs = decode_ref.decode_layers([t[b].reshape(t.shape[1], t.shape[2], A, 5 + C) for t in heads_gpu], helper, list(hw), hw)
        rb, rs = decode_ref.decode_layers([t[b].reshape(t.shape[1], t.shape[2], A, 5 + C) for t in heads_ref], helper, list(hw), hw)
        assert float(np.abs(gs.astype(np.float64) - rs).max()) < 1e-3, f"cfg{cfg_id} image {b}: confidence error"
        err = np.abs(gb / norm - rb / norm) / np.maximum(1.0, np.abs(rb / norm))
        assert float(err.max()) < 1e-3, f"cfg{cfg_id} image {b}: normalised box error {float(err.max()):.2e}"


def test_pipelined_steps_equal_serial_steps():
    """step_device(pipelined=True): the decode/NMS of step i (own stream, head set i % 2) overlaps the convolutions of step
    i+1; its records are read only AFTER step i+1 has been issued, and must equal those of the same batch run alone."""
    cfg = dict(wl.CONFIGS[2])
    B, hw = cfg["batch"], cfg["in_hw"]
    pipe = DetectionPipeline(cfg["model"], hw, wl.anchors(cfg), cfg["classes"], cfg["alpha"], B, wl.OBJ_THRESH, wl.IOU_THRESH,
                             wl.MAX_PER_CLASS)
    pipe.engine.set_weights(wl.bench_weights(cfg, pipe.engine.expected_variables()))
    xs = [torch.from_numpy(wl.synthetic_batch(cfg, 4000 + j, B)).cuda() for j in range(5)]
    serial = []
    for x in xs:
        pipe.engine.bind_input(x)
        d, c = pipe.step_device()           # current stream waits for the decode
        serial.append((d.clone(), c.clone()))
    torch.cuda.synchronize()
    for rep in range(2):                    # second pass: every (input, head set) graph already captured, launches back to back
        got, prev = [], None
        for x in xs + [xs[0]]:
            pipe.engine.bind_input(x)
            views = pipe.step_device(pipelined=True)
            if prev is not None:
                pipe.wait_gathered()        # decode stream is in order: step i is complete once step i+1's decode is
                got.append((prev[0].clone(), prev[1].clone()))
            prev = views
        torch.cuda.synchronize()
        for j, ((d, c), (sd, sc)) in enumerate(zip(got, serial)):
            assert torch.equal(c, sc), f"pass {rep} step {j}: counts differ"
            rec, ref = DetectionPipeline.records(d, c), DetectionPipeline.records(sd, sc)
            assert rec == ref, f"pass {rep} step {j}: records differ"


def test_laned_pipeline_equals_serial_steps():
    """LanedPipeline: two batches in flight on two engines / stream sets; device-resident and host-streaming forms both
    return, batch for batch, the records of the same batch run alone."""
    from collections import deque
    from k210_yolo_framework_b200.pipeline import LanedPipeline
    cfg = dict(wl.CONFIGS[2])
    B, hw = cfg["batch"], cfg["in_hw"]
    lp = LanedPipeline(2, cfg["model"], hw, wl.anchors(cfg), cfg["classes"], cfg["alpha"], B, wl.OBJ_THRESH, wl.IOU_THRESH,
                       wl.MAX_PER_CLASS)
    lp.set_weights(wl.bench_weights(cfg, lp.lanes[0].engine.expected_variables()))
    xs_host = [torch.from_numpy(wl.synthetic_batch(cfg, 5000 + j, B)).pin_memory() for j in range(6)]
    xs = [x.cuda() for x in xs_host]
    serial = []
    for x in xs:
        lp.lanes[0].engine.bind_input(x)
        d, c = lp.lanes[0].step_device()
        serial.append(DetectionPipeline.records(d.clone(), c.clone()))
    torch.cuda.synchronize()
    # device-resident: a lane's views stay valid until it has run two more batches -> read batch i after issuing batch i+1
    for rep in range(2):
        got, pending = [], deque()
        for x in xs:
            lp.bind_input(x)
            pending.append(lp.step_device())
            if len(pending) == 2:
                lp.wait_all()
                d, c = pending.popleft()
                got.append(DetectionPipeline.records(d.clone(), c.clone()))
        lp.wait_all()
        while pending:
            d, c = pending.popleft()
            got.append(DetectionPipeline.records(d.clone(), c.clone()))
        torch.cuda.synchronize()
        assert got == serial, f"device-resident pass {rep}"
    # host streaming
    for rep in range(2):
        got, q = [], deque()
        for x in xs_host:
            q.append(lp.submit(x))
            if len(q) >= lp.in_flight_limit():
                d, c = lp.collect(q.popleft())
                got.append(DetectionPipeline.records(d.clone(), c.clone()))
        while q:
            d, c = lp.collect(q.popleft())
            got.append(DetectionPipeline.records(d.clone(), c.clone()))
        assert got == serial, f"host streaming pass {rep}"
