"""Operator-level entry points (tf_xywh_to_all / correct_box / non_max_suppression) vs the oracle's restatements."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from k210_yolo_framework_b200 import Helper, ops
from oracle import decode_ref


def test_decode_operators_match_oracle(voc_anchors, dog_heads):
    h = Helper(None, 20, voc_anchors, [[224, 320]] * 2, [[7, 10], [14, 20]])
    href = decode_ref.HelperRef(voc_anchors, [224, 320], [7, 10, 14, 20], 20)
    rng = np.random.default_rng(2)
    for layer, (hh, ww) in enumerate([(7, 10), (14, 20)]):
        y = dog_heads[f"l{layer}_f32"].reshape(1, hh, ww, 3, 25)
        y = np.concatenate([y, rng.normal(0, 2, y.shape).astype(np.float32)], 0)            # batch of 2
        t = torch.from_numpy(y).cuda()
        xy, wh = ops.tf_xywh_to_all(t[..., 0:2], t[..., 2:4], layer, h)
        rxy, rwh = decode_ref.xywh_to_all(y[..., 0:2], y[..., 2:4], layer, href)
        np.testing.assert_allclose(xy.cpu().numpy(), rxy, rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(wh.cpu().numpy(), rwh, rtol=2e-6, atol=1e-7)
        for image_shape in ((224, 320), (374, 499), (500, 333)):
            boxes = ops.correct_box(xy, wh, (224, 320), image_shape)
            ref = decode_ref.correct_box(rxy, rwh, [224, 320], image_shape)
            assert boxes.shape == ref.shape
            np.testing.assert_allclose(boxes.cpu().numpy(), ref, rtol=1e-5, atol=2e-3)


@pytest.mark.parametrize("n,cap,thr", [(1, 30, 0.5), (7, 30, 0.5), (33, 30, 0.3), (300, 30, 0.5), (1050, 30, 0.5), (1050, 200, 0.45), (64, 3, 0.0)])
def test_nms_matches_oracle(n, cap, thr):
    rng = np.random.default_rng(n * 7 + cap)
    yx = rng.uniform(0, 200, (n, 2)).astype(np.float32)
    hw = rng.uniform(5, 90, (n, 2)).astype(np.float32)
    boxes = np.concatenate([yx, yx + hw], 1).astype(np.float32)
    flip = rng.random(n) < 0.2                                  # TF normalises flipped corners
    boxes[flip] = boxes[flip][:, [2, 3, 0, 1]]
    scores = rng.random(n).astype(np.float32)
    scores[rng.integers(0, n, max(n // 8, 1))] = scores[0]     # ties -> lower index first
    got = ops.non_max_suppression(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), cap, thr).cpu().numpy()
    ref = decode_ref.nms_tf_fast(boxes, scores, cap, thr)   # == nms_tf (tests/test_oracle_golden.py::test_fast_nms_equals_reference_nms)
    assert got.dtype == np.int32
    np.testing.assert_array_equal(got, ref)


def test_nms_edge_cases():
    z = torch.zeros((0, 4), device="cuda")
    assert ops.non_max_suppression(z, torch.zeros((0,), device="cuda"), 30, 0.5).numel() == 0
    b = torch.tensor([[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 20, 30], [0, 0, 10, 10.5]], dtype=torch.float32, device="cuda")
    s = torch.tensor([0.5, 0.9, -1.0, 0.9], device="cuda")     # duplicates, a zero-area box, a negative score
    assert ops.non_max_suppression(b, s, 30, 0.5).cpu().tolist() == [1, 2]
    assert ops.non_max_suppression(b, s, 0, 0.5).numel() == 0
    assert ops.non_max_suppression(b, s, 1, 0.5).cpu().tolist() == [1]
    with pytest.raises(ValueError):
        ops.non_max_suppression(b.cpu(), s.cpu(), 30, 0.5)      # host tensors: no CPU path
