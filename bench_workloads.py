"""The BASELINE.json workloads (configs 2-5) as data + seeded generators, shared by bench.py's two arms and the
full-size parity tests.  Imports neither the CUDA package nor the oracle: the reference arm of bench.py must run
without loading this repo's native library.

A workload = model builder name, depth multiplier, network input size, per-GPU batch, class count, anchors,
algorithmic FLOP per image and the seed / head-bias settings of the "detection-rich" random-init
weights (Keras default initialisers restated in k210_yolo_framework_b200/weights.py; that file is pure numpy and
is loaded here BY PATH, so the package's __init__ — which loads libk210yolo_b200.so — never runs).
"""
from __future__ import annotations

import importlib.util
import json
import os
from typing import Dict

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _anchors2():
    return np.load(os.path.join(GOLDEN, "voc_anchor.npy"))


def _anchors3():
    a = _anchors2()
    return np.concatenate([a, a[1:] * 0.5], 0)


# name -> workload.  batch = images per GPU (weak scaling).  `gpus` = the GPU count BASELINE.json quotes the config on.
CONFIGS: Dict[int, dict] = {
    2: dict(name="cfg2", model="yolo_mobilev1", alpha=0.75, in_hw=(224, 320), batch=32, classes=20, layers=2, gflop=1.465,
            gpus=1, metric="images/sec @320x224 yolo_mobilev1-0.75",
            desc="yolo_mobilev1 a=0.75, 320x224 (HxW 224x320), batch 32/GPU, VOC-20 anchors"),
    3: dict(name="cfg3", model="tiny_yolo", alpha=1.0, in_hw=(416, 416), batch=64, classes=20, layers=2, gflop=5.472,
            gpus=1, metric="images/sec @416x416 tiny_yolo",
            desc="tiny_yolo, 416x416, batch 64/GPU, 2 output scales x 3 anchors, VOC-20 anchors"),
    4: dict(name="cfg4", model="yolo_mobilev2", alpha=1.0, in_hw=(224, 320), batch=32, classes=20, layers=2, gflop=1.494,
            gpus=8, metric="images/sec @320x224 yolo_mobilev2-1.0",
            desc="yolo_mobilev2 a=1.0, 320x224, batch 32/GPU (256 over 8 GPUs), image-shard + all-gather of detections"),
    5: dict(name="cfg5", model="yolo", alpha=1.0, in_hw=(608, 608), batch=16, classes=80, layers=3, gflop=140.69,
            gpus=8, metric="images/sec @608x608 yolo (Darknet-53), 80 classes",
            desc="full yolo (Darknet-53), 608x608, batch 16/GPU (128 over 8 GPUs), 3 output scales, 80 classes, NMS stress"),
}
OBJ_THRESH, IOU_THRESH, MAX_PER_CLASS = 0.7, 0.5, 30
HEAD_BIAS, HEAD_BIAS_STD, WEIGHT_SEED = -0.2, 1.5, 0


def anchors(cfg: dict) -> np.ndarray:
    return _anchors3() if cfg["layers"] == 3 else _anchors2()


def out_hw(cfg: dict):
    h, w = cfg["in_hw"]
    return [(h // 32 * 2 ** l, w // 32 * 2 ** l) for l in range(cfg["layers"])]


def workload_string(cfg: dict) -> str:
    return f"{cfg['name']}: {cfg['desc']}, random-init detection-rich weights (seed {WEIGHT_SEED})"


def _weights_module():
    spec = importlib.util.spec_from_file_location("_k2y_weights_standalone",
                                                  os.path.join(ROOT, "k210_yolo_framework_b200", "weights.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def expected_variables(cfg: dict) -> Dict[str, Dict[str, tuple]]:
    """{keras_layer: {var: shape}} of the workload's network, from the committed fixture (written by
    scripts/make_expected_vars.py from the native graph builder; tests/test_graph_builder.py keeps the two equal)."""
    with open(os.path.join(GOLDEN, f"vars_{cfg['name']}.json")) as fh:
        raw = json.load(fh)
    return {layer: {var: tuple(shape) for var, shape in vars_.items()} for layer, vars_ in raw.items()}


def bench_weights(cfg: dict, expected=None):
    exp = expected if expected is not None else expected_variables(cfg)
    return _weights_module().random_weights(exp, seed=WEIGHT_SEED, detection_rich=True, head_bias=HEAD_BIAS,
                                            head_bias_std=HEAD_BIAS_STD)


def synthetic_batch(cfg: dict, seed: int, n: int = None) -> np.ndarray:
    """float32 NHWC in [0, 1): what `img / np.max(img)` hands to predict."""
    n = cfg["batch"] if n is None else n
    h, w = cfg["in_hw"]
    return np.random.default_rng(seed).random((n, h, w, 3), dtype=np.float32)


def synthetic_batch_u8(cfg: dict, seed: int, n: int = None) -> np.ndarray:
    """uint8 NHWC letterboxed RGB: what `_read_img/_process_img` hold before the normalisation."""
    n = cfg["batch"] if n is None else n
    h, w = cfg["in_hw"]
    return np.random.default_rng(seed).integers(0, 256, (n, h, w, 3), dtype=np.uint8)
