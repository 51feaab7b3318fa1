"""Oracle restatement of the KERAS-dialect decode + per-class NMS (TEST INFRASTRUCTURE).

Follows /root/reference/keras_inference.py:32-72 (``correct_box``), :94-135 (decode loop,
threshold mask, per-class ``tf.image.non_max_suppression``), and
/root/reference/tools/utils.py:53-82, 233-271 (``Helper`` anchors / xy_offset),
:524-547 (``tf_xywh_to_all``).  All arithmetic in numpy float32, in the reference's
operation order.

``tf.image.non_max_suppression`` lives in TensorFlow 1.14 (third-party, not vendored):
greedy, candidates visited in descending score, a candidate is dropped iff
IoU(candidate, some already-selected box) > iou_threshold, stop at max_output_size;
IoU on (min,max)-normalised corners, 0 for non-positive areas.  TF's heap leaves the
order of equal scores unspecified; this oracle (and the CUDA path) break ties by
ascending box index.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

f32 = np.float32


class HelperRef:
    """The inference-relevant part of ``Helper.__init__`` (tools/utils.py:53-82)."""

    def __init__(self, anchors: np.ndarray, in_hw: Sequence[int], out_hw: Sequence[Sequence[int]], class_num: int):
        self.in_hw = np.reshape(np.array(in_hw), (-1, 2))
        self.out_hw = np.reshape(np.array(out_hw), (-1, 2))
        self.grid_wh = (1 / self.out_hw)[:, [1, 0]]
        self.class_num = class_num
        self.anchors = np.asarray(anchors)  # [L, A, 2] (w, h) as fraction of the net input, f64
        self.anchor_number = len(self.anchors[0])
        self.output_number = len(self.anchors)
        self.xy_offset = self._coordinate_offset(self.anchors, self.out_hw)

    @staticmethod
    def _coordinate_offset(anchors, out_hw):
        """tools/utils.py:233-253 — per layer [h, w, 1, 2] grid of (col, row)."""
        grid = []
        for l in range(len(anchors)):
            gy = np.tile(np.reshape(np.arange(0, out_hw[l][0]), [-1, 1, 1, 1]), [1, out_hw[l][1], 1, 1])
            gx = np.tile(np.reshape(np.arange(0, out_hw[l][1]), [1, -1, 1, 1]), [out_hw[l][0], 1, 1, 1])
            grid.append(np.concatenate([gx, gy], axis=-1))
        return grid


def exp32(x: np.ndarray) -> np.ndarray:
    """float32 exp, DEFINED as the correctly rounded result (double-precision exp rounded once to float32).

    The reference evaluates exp inside TensorFlow 1.14 (Eigen's vectorised pexp, third-party, not vendored); its last-bit
    behaviour is not reproducible offline, and numpy's own float32 SIMD exp differs from the correctly rounded value
    in ~40 % of the arguments.  Pinning exp to the correctly rounded value makes the score ORDER (and with it the NMS
    survivor sets) a property of the head tensors alone; the CUDA path (csrc/detect.cu exp_cr) uses the same definition."""
    return np.exp(np.asarray(x, f32).astype(np.float64)).astype(f32)


def sigmoid32(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x).astype(f32)
    return (f32(1) / (f32(1) + exp32(-x))).astype(f32)


def xywh_to_all(pred_xy: np.ndarray, pred_wh: np.ndarray, layer: int, h: HelperRef):
    """tools/utils.py:544-547."""
    xy = (sigmoid32(pred_xy) + h.xy_offset[layer].astype(f32)) / h.out_hw[layer][::-1].astype(f32)
    wh = exp32(pred_wh) * h.anchors[layer].astype(f32)
    return xy.astype(f32), wh.astype(f32)


def correct_box(box_xy, box_wh, input_shape, image_shape) -> np.ndarray:
    """keras_inference.py:32-72 — returns [..., 4] = (ymin, xmin, ymax, xmax) in image pixels."""
    box_yx = box_xy[..., ::-1]
    box_hw = box_wh[..., ::-1]
    input_shape = np.asarray(input_shape, f32)
    image_shape = np.asarray(image_shape, f32)
    new_shape = np.round(image_shape * np.min(input_shape / image_shape)).astype(f32)  # tf.round: half-to-even
    offset = ((input_shape - new_shape) / f32(2.0) / input_shape).astype(f32)
    scale = (input_shape / new_shape).astype(f32)
    box_yx = ((box_yx - offset) * scale).astype(f32)
    box_hw = (box_hw * scale).astype(f32)
    mins = (box_yx - (box_hw / f32(2.0))).astype(f32)
    maxes = (box_yx + (box_hw / f32(2.0))).astype(f32)
    boxes = np.concatenate([mins[..., 0:1], mins[..., 1:2], maxes[..., 0:1], maxes[..., 1:2]], axis=-1)
    boxes = (boxes * np.concatenate([image_shape, image_shape], axis=-1)).astype(f32)
    return boxes


def decode_layers(y_pred: List[np.ndarray], h: HelperRef, image_size, image_shape):
    """keras_inference.py:94-116 for ONE image: y_pred[l] is [h_l, w_l, A, 5+C].

    Returns (boxes [Nbox,4] f32, scores [Nbox,C] f32) with flat box index
    ``off_l + (row*W + col)*A + a`` (layer 0 first).
    """
    bl, sl = [], []
    for l, p in enumerate(y_pred):
        p = p.astype(f32)
        scores = (sigmoid32(p[..., 5:]) * sigmoid32(p[..., 4:5])).astype(f32)
        xy, wh = xywh_to_all(p[..., 0:2], p[..., 2:4], l, h)
        boxes = correct_box(xy, wh, image_size, image_shape)
        bl.append(boxes.reshape(-1, 4))
        sl.append(scores.reshape(-1, h.class_num))
    return np.concatenate(bl, 0), np.concatenate(sl, 0)


def iou_yxyx(a: np.ndarray, b: np.ndarray) -> np.float32:
    """TF NonMaxSuppression IoU on corner boxes, float32 op order."""
    ymin_a, ymax_a = min(a[0], a[2]), max(a[0], a[2])
    xmin_a, xmax_a = min(a[1], a[3]), max(a[1], a[3])
    ymin_b, ymax_b = min(b[0], b[2]), max(b[0], b[2])
    xmin_b, xmax_b = min(b[1], b[3]), max(b[1], b[3])
    area_a = f32(f32(ymax_a - ymin_a) * f32(xmax_a - xmin_a))
    area_b = f32(f32(ymax_b - ymin_b) * f32(xmax_b - xmin_b))
    if area_a <= 0 or area_b <= 0:
        return f32(0)
    iy = f32(max(f32(min(ymax_a, ymax_b) - max(ymin_a, ymin_b)), f32(0)))
    ix = f32(max(f32(min(xmax_a, xmax_b) - max(xmin_a, xmin_b)), f32(0)))
    inter = f32(iy * ix)
    return f32(inter / f32(f32(area_a + area_b) - inter))


def nms_tf(boxes: np.ndarray, scores: np.ndarray, max_output_size: int, iou_threshold: float) -> np.ndarray:
    """``tf.image.non_max_suppression`` (V3, score_threshold=-inf); returns selected indices."""
    n = len(scores)
    order = sorted(range(n), key=lambda i: (-float(scores[i]), i))
    thr = f32(iou_threshold)
    sel: List[int] = []
    for i in order:
        if len(sel) >= max_output_size:
            break
        keep = True
        for j in reversed(sel):
            if iou_yxyx(boxes[i], boxes[j]) > thr:
                keep = False
                break
        if keep:
            sel.append(i)
    return np.asarray(sel, np.int32)


def detect_image(y_pred: List[np.ndarray], h: HelperRef, image_size, image_shape,
                 obj_thresh: float, iou_thresh: float, max_per_class: int = 30):
    """keras_inference.py:94-135 for one image.

    Returns a list of detections ``(class, flat_index, score, ymin, xmin, ymax, xmax)`` in the
    reference's output order (class ascending, score descending within class).
    """
    boxes, scores = decode_layers(y_pred, h, image_size, image_shape)
    mask = scores >= f32(obj_thresh)
    out = []
    for c in range(h.class_num):
        idx = np.nonzero(mask[:, c])[0]
        if len(idx) == 0:
            continue
        cb = boxes[idx]
        cs = scores[idx, c]
        sel = nms_tf(cb, cs, max_per_class, iou_thresh)
        for s in sel:
            b = cb[s]
            out.append((c, int(idx[s]), f32(cs[s]), f32(b[0]), f32(b[1]), f32(b[2]), f32(b[3])))
    return out


def nms_tf_fast(boxes: np.ndarray, scores: np.ndarray, max_output_size: int, iou_threshold: float) -> np.ndarray:
    """Same result as ``nms_tf`` with the inner loop vectorised (used for the CPU-baseline timing, where
    TF's NMS kernel is compiled code; validated against ``nms_tf`` in tests/test_oracle_golden.py)."""
    n = len(scores)
    if n == 0:
        return np.zeros((0,), np.int32)
    order = np.lexsort((np.arange(n), -scores.astype(np.float64)))
    b = boxes.astype(f32)
    ymin, ymax = np.minimum(b[:, 0], b[:, 2]), np.maximum(b[:, 0], b[:, 2])
    xmin, xmax = np.minimum(b[:, 1], b[:, 3]), np.maximum(b[:, 1], b[:, 3])
    area = ((ymax - ymin).astype(f32) * (xmax - xmin).astype(f32)).astype(f32)
    thr = f32(iou_threshold)
    sel: List[int] = []
    for i in order:
        if len(sel) >= max_output_size:
            break
        if sel:
            s = np.asarray(sel)
            iy = np.maximum((np.minimum(ymax[i], ymax[s]) - np.maximum(ymin[i], ymin[s])).astype(f32), f32(0))
            ix = np.maximum((np.minimum(xmax[i], xmax[s]) - np.maximum(xmin[i], xmin[s])).astype(f32), f32(0))
            inter = (iy * ix).astype(f32)
            with np.errstate(divide="ignore", invalid="ignore"):
                iou = (inter / ((area[i] + area[s]).astype(f32) - inter).astype(f32)).astype(f32)
            iou = np.where((area[i] <= 0) | (area[s] <= 0), f32(0), iou)
            if np.any(iou > thr):
                continue
        sel.append(int(i))
    return np.asarray(sel, np.int32)


def detect_batch_fast(heads: List[np.ndarray], h: HelperRef, image_size, image_shapes, obj_thresh, iou_thresh,
                      max_per_class: int = 30):
    """keras_inference.py:94-135 over a batch (heads[l]: [N,h,w,A*(5+C)]), vectorised NMS inner loop."""
    out = []
    A = h.anchor_number
    for b in range(heads[0].shape[0]):
        yp = [hd[b].reshape(hd.shape[1], hd.shape[2], A, 5 + h.class_num) for hd in heads]
        boxes, scores = decode_layers(yp, h, image_size, image_shapes[b])
        mask = scores >= f32(obj_thresh)
        img = []
        for c in range(h.class_num):
            idx = np.nonzero(mask[:, c])[0]
            if len(idx) == 0:
                continue
            sel = nms_tf_fast(boxes[idx], scores[idx, c], max_per_class, iou_thresh)
            for s in sel:
                bb = boxes[idx[s]]
                img.append((c, int(idx[s]), f32(scores[idx[s], c]), f32(bb[0]), f32(bb[1]), f32(bb[2]), f32(bb[3])))
        out.append(img)
    return out
