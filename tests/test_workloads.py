"""bench_workloads: the committed variable fixtures equal the native graph builder, and both arms of bench.py build
bit-identical seeded weights without the reference arm ever importing the CUDA package."""
import subprocess
import sys

import numpy as np
import pytest

import bench_workloads as wl
from conftest import ROOT


@pytest.mark.parametrize("cfg_id", sorted(wl.CONFIGS))
def test_fixture_equals_graph_builder(cfg_id):
    from k210_yolo_framework_b200 import yolonet
    cfg = wl.CONFIGS[cfg_id]
    h, w = cfg["in_hw"]
    m, _ = getattr(yolonet, cfg["model"])([h, w, 3], 3, cfg["classes"], alpha=cfg["alpha"], max_batch=1)
    exp = m.engine.expected_variables()
    fix = wl.expected_variables(cfg)
    assert list(exp) == list(fix)                      # same layers, same (creation) order -> same RNG stream
    assert exp == fix
    assert [(hh, ww) for hh, ww, _ in m.engine.out_shapes] == wl.out_hw(cfg)
    assert wl.anchors(cfg).shape == (cfg["layers"], 3, 2)


def test_weights_identical_through_both_paths():
    from k210_yolo_framework_b200 import yolonet
    from k210_yolo_framework_b200.weights import random_weights
    cfg = wl.CONFIGS[2]
    m, _ = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75, max_batch=1)
    a = random_weights(m.engine.expected_variables(), seed=wl.WEIGHT_SEED, detection_rich=True, head_bias=wl.HEAD_BIAS,
                       head_bias_std=wl.HEAD_BIAS_STD)
    b = wl.bench_weights(cfg)
    assert list(a) == list(b)
    for layer in a:
        for var in a[layer]:
            np.testing.assert_array_equal(a[layer][var], b[layer][var])


def test_reference_arm_never_loads_the_cuda_package():
    code = ("import sys, bench, bench_workloads as wl\n"
            "cp = bench.CpuPath(wl.CONFIGS[2])\n"
            "bad = [m for m in sys.modules if m.startswith('k210_yolo_framework_b200')]\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert not bad, bad\n"
            "assert 'libk210yolo_b200' not in maps\n"
            "print('clean')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0 and "clean" in r.stdout, r.stderr[-2000:]
