"""keras_inference.py drop-in: the reference's cfg-1 (dog.jpg, trained yolo_mobilev1-0.75) end to end on the GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from conftest import GOLDEN, ROOT


@pytest.fixture()
def workdir(tmp_path, dog_u8, voc_anchors):
    from PIL import Image
    (tmp_path / "data").mkdir()
    np.save(tmp_path / "data" / "voc_anchor.npy", voc_anchors)
    Image.fromarray(dog_u8).save(tmp_path / "dog.png")     # lossless: same pixels the oracle saw
    return tmp_path


def test_main_prints_reference_format(workdir, dog_golden):
    ckpt = os.path.join(GOLDEN, "yolo_mobilev1_075_voc_weights.npz")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "keras_inference.py"), ckpt, "dog.png", "--model_def",
                        "yolo_mobilev1", "--depth_multiplier", "0.75", "--image_size", "224", "320", "--output_size",
                        "7", "10", "14", "20", "--obj_thresh", "0.7", "--iou_thresh", "0.5"],
                       cwd=workdir, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("[")]
    assert lines[1] == "[top\tleft\tbottom\tright\tscore\tclass]"
    gold = dog_golden["keras"]["detections"]
    assert len(lines) == 2 + len(gold)
    for line, g in zip(lines[2:], gold):
        vals = line.strip("[]").split("\t")
        np.testing.assert_allclose([float(v) for v in vals[:4]], g[3:], atol=0.06)   # printed with 1 decimal
        assert abs(float(vals[4]) - g[2]) < 0.006 and int(vals[5]) == g[0]


def test_detect_function_matches_golden(workdir, dog_golden, monkeypatch):
    monkeypatch.chdir(workdir)
    sys.path.insert(0, ROOT)
    import keras_inference
    ckpt = os.path.join(GOLDEN, "yolo_mobilev1_075_voc_weights.npz")
    _, _, found = keras_inference.detect(ckpt, [224, 320], [7, 10, 14, 20], "yolo_mobilev1", 20, 0.75, 0.7, 0.5, "voc", "dog.png")
    gold = dog_golden["keras"]["detections"]
    assert [(d[0], d[1]) for d in found] == [(g[0], g[1]) for g in gold]            # identical post-NMS box indices
    np.testing.assert_allclose([d[2] for d in found], [g[2] for g in gold], atol=1e-3)   # confidences within 1e-3
    np.testing.assert_allclose(np.array([d[3:] for d in found]) / [224, 320, 224, 320],
                               np.array([g[3:] for g in gold]) / [224, 320, 224, 320], atol=1e-3)  # normalised coords
    # no detections -> the NOTE line
    _, _, none = keras_inference.detect(ckpt, [224, 320], [7, 10, 14, 20], "yolo_mobilev1", 20, 0.75, 0.9999, 0.5, "voc", "dog.png")
    assert none == []
