"""GPU letterbox (k2y_letterbox_u8) vs the oracle restatement — bit-exact uint8 — and the people.jpg path end to end."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from k210_yolo_framework_b200 import Helper, KerasDetector, preprocess, yolonet
from oracle import decode_ref, keras_ref, preprocess_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gpu_letterbox(img, in_hw):
    return preprocess.letterbox_device(torch.from_numpy(img).cuda(), in_hw).cpu().numpy()


@pytest.mark.parametrize("src_hw,in_hw,lo", [((375, 500), (224, 320), 0), ((100, 37), (96, 96), 0), ((224, 320), (224, 320), 0),
                                             ((480, 640), (416, 416), 0), ((33, 61), (224, 320), 25), ((1, 1), (32, 32), 7),
                                             ((720, 1280), (608, 608), 1), ((500, 333), (224, 320), 0)])
def test_letterbox_bit_exact(src_hw, in_hw, lo):
    """down-/up-scaling, width- and height-limited, identity, min > 0 (the clip branch that preserves the zero fill)."""
    rng = np.random.default_rng(src_hw[0] * 1000 + src_hw[1])
    img = rng.integers(lo, 256, (src_hw[0], src_hw[1], 3), dtype=np.uint8)
    ref = preprocess_ref.letterbox(img, in_hw)
    got = _gpu_letterbox(img, in_hw)
    assert got.shape == ref.shape and got.dtype == np.uint8
    np.testing.assert_array_equal(got, ref)


def test_people_letterbox_matches_golden(people_u8):
    with open(os.path.join(GOLDEN, "people_golden.json")) as fh:
        gold = json.load(fh)
    got = _gpu_letterbox(people_u8, (224, 320))
    assert hashlib.sha256(got.tobytes()).hexdigest() == gold["letterbox_sha256"]
    for r, c, *rgb in gold["letterbox_samples_r_c_rgb"]:
        assert got[r, c].tolist() == rgb


def test_helper_process_img_equals_oracle(people_u8, dog_u8, voc_anchors):
    h = Helper(None, 20, voc_anchors, [[224, 320]] * 2, [[7, 10], [14, 20]])
    for img in (people_u8, dog_u8):
        x, _ = h._process_img(img.copy(), None, False, True)
        ref = preprocess_ref.process_img(img, (224, 320))
        assert x.dtype == np.float64
        np.testing.assert_array_equal(x, ref)
    with pytest.raises(ValueError):
        preprocess.letterbox_device(torch.zeros((4, 4, 3), dtype=torch.uint8), (8, 8))   # host tensor: no CPU path


def test_people_end_to_end(golden_weights, people_u8, voc_anchors):
    """decoded uint8 image -> GPU letterbox -> uint8 front end -> network -> decode + NMS, against the oracle run on the
    oracle-letterboxed input and against the five class-14 boxes of asset/people_res.jpg (people_golden.json)."""
    with open(os.path.join(GOLDEN, "people_golden.json")) as fh:
        gold = json.load(fh)
    m, w = yolonet.yolo_mobilev1([224, 320, 3], 3, 20, alpha=0.75, max_batch=1)
    m.set_weights_dict(golden_weights)
    h = Helper(None, 20, voc_anchors, [[224, 320]] * 2, [[7, 10], [14, 20]])
    y = w.predict_device_u8(h.letterbox_device(people_u8)[None])
    det = KerasDetector(h.anchors, [224, 320], h.out_hw, 20, 0.7, 0.5, max_per_class=30, max_batch=1)
    dets, counts = det.run([t.contiguous() for t in y], people_u8.shape[:2])
    found = KerasDetector.to_host(dets, counts)[0]
    assert [(d[0], d[1]) for d in found] == [(d[0], d[1]) for d in gold["detections"]]
    for g, r in zip(found, gold["detections"]):
        assert abs(g[2] - r[2]) < 1e-3
        assert np.abs(np.array(g[3:]) - np.array(r[3:])).max() < 0.05
    assert [d[1] for d in found] == [637, 778, 745, 676, 590] and all(d[0] == 14 for d in found)
