"""Pins the oracle to the only golden material the reference ships:
asset/yolo_model.h5 + data/dog.jpg must reproduce the boxes drawn in asset/dog_res.jpg."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_ref, keras_ref, region_c

# Known answers derived during the survey from the reference's shipped weights/image.
KNOWN_KERAS = [  # (class, flat index, score, [top, left, bottom, right])
    (6, 53, 0.815598, [25.901, 188.631, 74.671, 309.192]),
    (11, 765, 0.996330, [90.045, 40.739, 212.476, 142.740]),
]
KNOWN_REGION = [[(188, 25, 309, 74, 6, 0.8213)], [(40, 90, 142, 212, 11, 0.9965)]]


def _heads(golden_weights, dog_u8, dtype=torch.float32):
    x = (dog_u8 / np.max(dog_u8)).astype(np.float64 if dtype == torch.float64 else np.float32)[None]
    return keras_ref.forward("yolo_mobilev1", golden_weights, x, alpha=0.75, dtype=dtype)


def test_network_oracle_reproduces_committed_heads(golden_weights, dog_u8, dog_heads):
    heads = _heads(golden_weights, dog_u8)
    assert heads[0].shape == (1, 7, 10, 75) and heads[1].shape == (1, 14, 20, 75)
    np.testing.assert_allclose(heads[0], dog_heads["l0_f32"], atol=2e-4)
    np.testing.assert_allclose(heads[1], dog_heads["l1_f32"], atol=2e-4)
    # fp32 stand-in vs fp64 ground truth: the tolerance budget the GPU path is judged against
    assert np.abs(heads[0] - dog_heads["l0_f64"]).max() < 1e-3
    assert abs(np.abs(dog_heads["l0_f64"]).max() - 17.30) < 0.01 and abs(np.abs(dog_heads["l1_f64"]).max() - 30.81) < 0.01


def test_keras_dialect_known_answers(dog_heads, voc_anchors, dog_u8):
    h = decode_ref.HelperRef(voc_anchors, [224, 320], [7, 10, 14, 20], 20)
    yp = [dog_heads["l0_f32"][0].reshape(7, 10, 3, 25), dog_heads["l1_f32"][0].reshape(14, 20, 3, 25)]
    boxes, scores = decode_ref.decode_layers(yp, h, [224, 320], dog_u8.shape[:2])
    assert boxes.shape == (1050, 4) and scores.shape == (1050, 20)
    cand = np.argwhere(scores >= 0.7)
    assert sorted(map(tuple, cand.tolist())) == [(53, 6), (705, 11), (765, 11), (768, 11)]
    assert abs(scores[705, 11] - 0.9768) < 1e-3 and abs(scores[768, 11] - 0.7570) < 1e-3
    assert abs(scores[708, 11] - 0.6712) < 1e-3
    det = decode_ref.detect_image(yp, h, [224, 320], dog_u8.shape[:2], 0.7, 0.5)
    assert len(det) == len(KNOWN_KERAS)
    for got, (c, idx, score, box) in zip(det, KNOWN_KERAS):
        assert got[0] == c and got[1] == idx
        assert abs(got[2] - score) < 1e-3
        np.testing.assert_allclose(got[3:], box, atol=0.05)


def test_keras_dialect_matches_committed_json(dog_heads, voc_anchors, dog_u8, dog_golden):
    h = decode_ref.HelperRef(voc_anchors, [224, 320], [7, 10, 14, 20], 20)
    yp = [dog_heads["l0_f32"][0].reshape(7, 10, 3, 25), dog_heads["l1_f32"][0].reshape(14, 20, 3, 25)]
    det = decode_ref.detect_image(yp, h, [224, 320], dog_u8.shape[:2], 0.7, 0.5)
    gold = dog_golden["keras"]["detections"]
    assert [(d[0], d[1]) for d in det] == [(g[0], g[1]) for g in gold]
    np.testing.assert_allclose([list(d[2:]) for d in det], [g[2:] for g in gold], rtol=1e-6, atol=1e-5)


def test_region_dialect_known_answers_numpy(dog_heads, voc_anchors):
    for l, (W, H) in enumerate([(10, 7), (20, 14)]):
        chw = region_c.nhwc_to_chw(dog_heads[f"l{l}_f32"][0], 3)
        out, _, _ = region_c.region_layer_np(chw, W, H, voc_anchors[l].reshape(-1), 0.6, 0.3, 320, 224)
        assert [t[:5] for t in out] == [t[:5] for t in KNOWN_REGION[l]]
        assert abs(out[0][5] - KNOWN_REGION[l][0][5]) < 1e-3


@pytest.mark.skipif(not region_c.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_region_dialect_known_answers_compiled_reference(dog_heads, voc_anchors, dog_golden):
    for l, (W, H) in enumerate([(10, 7), (20, 14)]):
        chw = region_c.nhwc_to_chw(dog_heads[f"l{l}_f32"][0], 3)
        r = region_c.RegionLayerRef(W, H, 75, 320, 224, voc_anchors[l].reshape(-1), 0.6, 0.3)
        out = r.run(chw)
        assert [list(t[:5]) for t in out] == [g[:5] for g in dog_golden["region_c"]["layers"][l]]
        assert [t[:5] for t in out] == [t[:5] for t in KNOWN_REGION[l]]


def test_nms_semantics_small_cases():
    # ties -> ascending index; suppression strictly '>' threshold; cap at max_output_size
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [0, 5, 10, 15], [20, 20, 30, 30]], np.float32)
    scores = np.array([0.9, 0.9, 0.8, 0.7], np.float32)
    assert decode_ref.nms_tf(boxes, scores, 30, 0.5).tolist() == [0, 2, 3]   # box1 (IoU 1) dropped, box2 IoU 1/3 kept
    assert decode_ref.nms_tf(boxes, scores, 2, 0.5).tolist() == [0, 2]
    iou = decode_ref.iou_yxyx(boxes[0], boxes[2])
    assert decode_ref.nms_tf(boxes, scores, 30, float(iou)).tolist() == [0, 2, 3]       # equal to threshold: kept
    assert decode_ref.nms_tf(boxes, scores, 30, float(iou) - 1e-6).tolist() == [0, 3]
    assert decode_ref.nms_tf(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 30, 0.5).tolist() == []
    # degenerate (zero-area) boxes never overlap anything
    z = np.array([[5, 5, 5, 9], [5, 5, 5, 9]], np.float32)
    assert decode_ref.nms_tf(z, np.array([0.9, 0.8], np.float32), 30, 0.1).tolist() == [0, 1]


def test_correct_box_letterbox_math():
    # 374x499 image into 224x320: new_shape = round(img * min(in/img)) = (224, 299); offset/scale per keras_inference.py:55-57
    xy = np.array([[0.5, 0.5]], np.float32)
    wh = np.array([[0.2, 0.4]], np.float32)
    b = decode_ref.correct_box(xy, wh, [224, 320], [374, 499])
    new_w = round(499 * min(224 / 374, 320 / 499))
    sx = 320 / new_w
    cx = (0.5 - (320 - new_w) / 2 / 320) * sx
    np.testing.assert_allclose(b[0], [(0.5 - 0.2) * 374, (cx - 0.1 * sx) * 499, (0.5 + 0.2) * 374, (cx + 0.1 * sx) * 499], rtol=1e-5)


def test_fast_nms_equals_reference_nms():
    rng = np.random.default_rng(0)
    for trial in range(20):
        n = int(rng.integers(1, 120))
        yx = rng.random((n, 2)) * 100
        hw = rng.random((n, 2)) * 40
        boxes = np.concatenate([yx, yx + hw], 1).astype(np.float32)
        if trial % 4 == 0:
            boxes[:, [0, 2]] = boxes[:, [2, 0]]          # un-normalised corners
        scores = np.round(rng.random(n), 2).astype(np.float32)   # many ties
        a = decode_ref.nms_tf(boxes, scores, 30, 0.4)
        b = decode_ref.nms_tf_fast(boxes, scores, 30, 0.4)
        assert a.tolist() == b.tolist()
