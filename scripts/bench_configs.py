"""BASELINE.json configs 3-5 (and 2 for reference) on one GPU: full-size parity spot-check against the oracle on the
first image(s) + device timing of forward + decode/NMS.  Prints one JSON object; used to fill DESIGN.md §5."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from k210_yolo_framework_b200 import _lib  # noqa: E402
from k210_yolo_framework_b200.pipeline import DetectionPipeline  # noqa: E402
from k210_yolo_framework_b200.weights import random_weights  # noqa: E402
from oracle import decode_ref, keras_ref  # noqa: E402

ANCH2 = np.load(os.path.join(ROOT, "tests", "golden", "voc_anchor.npy"))
ANCH3 = np.concatenate([ANCH2, ANCH2[1:] * 0.5], 0)
CONFIGS = [
    ("cfg2", "yolo_mobilev1", 0.75, (224, 320), 32, 20, ANCH2, 1.465),
    ("cfg3", "tiny_yolo", 1.0, (416, 416), 64, 20, ANCH2, 5.472),
    ("cfg4", "yolo_mobilev2", 1.0, (224, 320), 32, 20, ANCH2, 1.494),
    ("cfg5", "yolo", 1.0, (608, 608), 16, 80, ANCH3, 140.69),
]


def main():
    which = sys.argv[1:] or [c[0] for c in CONFIGS]
    out = {}
    for name, model_def, alpha, hw, batch, classes, anchors, gflop in CONFIGS:
        if name not in which:
            continue
        pipe = DetectionPipeline(model_def, hw, anchors, classes, alpha, batch, obj_thresh=0.7, iou_thresh=0.5)
        w = random_weights(pipe.engine.expected_variables(), seed=0, detection_rich=True, head_bias=-0.2, head_bias_std=1.5)
        pipe.engine.set_weights(w)
        rng = np.random.default_rng(1)
        x = rng.random((batch, hw[0], hw[1], 3), dtype=np.float32)
        # low-frequency structure so that activations vary over the image
        x[:, ::2, ::3] *= 0.5
        pipe.engine.input_buffer.copy_(torch.from_numpy(x))
        for _ in range(3):
            pipe.step_device()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            dets, counts = pipe.step_device()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        got = DetectionPipeline.records(dets.clone(), counts.clone())
        heads_gpu = [t[:1].cpu().numpy() for t in pipe.engine.head_buffers]
        t0 = time.time()
        heads = keras_ref.forward(model_def, w, x[:1], alpha=alpha)
        t_cpu = time.time() - t0
        err = [float(np.abs(a - b).max()) for a, b in zip(heads_gpu, heads)]
        scale = [float(np.abs(b).max()) for b in heads]
        out_hw = [(h.shape[1], h.shape[2]) for h in heads]
        helper = decode_ref.HelperRef(anchors, list(hw), out_hw, classes)
        ref = decode_ref.detect_batch_fast(heads_gpu, helper, list(hw), [hw], 0.7, 0.5)   # oracle NMS on the GPU's heads
        same = [(d[0], d[1]) for d in got[0]] == [(d[0], d[1]) for d in ref[0]]
        out[name] = {"model": model_def, "alpha": alpha, "in_hw": hw, "batch": batch, "classes": classes,
                     "ms_per_step": ms, "images_per_sec": batch / (ms * 1e-3), "tflops_algorithmic": batch * gflop / ms,
                     "launches": pipe.launches_per_step(), "math": _lib.MATH_NAMES[pipe.engine.get_math()],
                     "head_max_abs_err_vs_fp32_oracle": err, "head_max_abs": scale, "oracle_cpu_s_per_image": t_cpu,
                     "detections_image0": len(got[0]), "nms_sets_identical_on_gpu_heads": bool(same)}
        print(name, json.dumps(out[name]), file=sys.stderr, flush=True)
        del pipe
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
