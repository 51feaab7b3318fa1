"""Operator-level mirror of the reference's decode loop (keras_inference.py:94-135), on CUDA tensors.

For callers that keep the reference's Python loop and swap one operator at a time:

    tf_xywh_to_all(pred_xy, pred_wh, layer, helper)          tools/utils.py:524-547
    correct_box(box_xy, box_wh, input_shape, image_shape)    keras_inference.py:32-72
    non_max_suppression(boxes, scores, max_output_size, iou_threshold)   tf.image.non_max_suppression (:125-126)

Each call is one hand-written kernel behind the C-ABI (k2y_xywh_to_all / k2y_correct_box / k2y_nms_boxes) with the
arithmetic of the fused ``k2y_detect_keras`` path; the fused path (``KerasDetector``) is the fast one.  No CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._lib import check, lib


def _stream(t: torch.Tensor):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _f32_cuda(t, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32:
        raise ValueError(f"{name}: expected a CUDA float32 tensor")
    return t.contiguous()


def tf_xywh_to_all(pred_xy: torch.Tensor, pred_wh: torch.Tensor, layer: int, h):
    """pred_xy, pred_wh: [..., h_l, w_l, A, 2] raw head slices -> (xy in [0, 1], wh as a fraction of the network input)."""
    pred_xy, pred_wh = _f32_cuda(pred_xy, "pred_xy"), _f32_cuda(pred_wh, "pred_wh")
    hh, ww = (int(v) for v in h.out_hw[layer])
    anchors = np.ascontiguousarray(np.asarray(h.anchors[layer], np.float32).reshape(-1))
    a = anchors.size // 2
    if pred_xy.shape != pred_wh.shape or pred_xy.dim() < 4 or tuple(pred_xy.shape[-4:]) != (hh, ww, a, 2):
        raise ValueError(f"expected [..., {hh}, {ww}, {a}, 2] for layer {layer}, got {tuple(pred_xy.shape)} / {tuple(pred_wh.shape)}")
    xy, wh = torch.empty_like(pred_xy), torch.empty_like(pred_wh)
    with torch.cuda.device(pred_xy.device):
        check(lib.k2y_xywh_to_all(pred_xy.data_ptr(), pred_wh.data_ptr(), pred_xy.numel() // 2, hh, ww, a,
                                  anchors.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), xy.data_ptr(), wh.data_ptr(), _stream(pred_xy)))
    return xy, wh


def correct_box(box_xy: torch.Tensor, box_wh: torch.Tensor, input_shape, image_shape) -> torch.Tensor:
    """[..., 2] xy / wh (relative to the letterboxed input) -> [..., 4] (ymin, xmin, ymax, xmax) in original-image pixels."""
    box_xy, box_wh = _f32_cuda(box_xy, "box_xy"), _f32_cuda(box_wh, "box_wh")
    if box_xy.shape != box_wh.shape or box_xy.shape[-1] != 2:
        raise ValueError("box_xy and box_wh must both be [..., 2]")
    boxes = torch.empty(tuple(box_xy.shape[:-1]) + (4,), dtype=torch.float32, device=box_xy.device)
    if boxes.numel():
        with torch.cuda.device(box_xy.device):
            check(lib.k2y_correct_box(box_xy.data_ptr(), box_wh.data_ptr(), box_xy.numel() // 2, float(input_shape[0]), float(input_shape[1]),
                                      float(image_shape[0]), float(image_shape[1]), boxes.data_ptr(), _stream(box_xy)))
    return boxes


def non_max_suppression(boxes: torch.Tensor, scores: torch.Tensor, max_output_size: int, iou_threshold: float = 0.5) -> torch.Tensor:
    """boxes [n, 4] (ymin, xmin, ymax, xmax), scores [n] -> int32 indices of the selected boxes, score-descending."""
    boxes, scores = _f32_cuda(boxes, "boxes"), _f32_cuda(scores, "scores")
    if boxes.dim() != 2 or boxes.shape[1] != 4 or scores.dim() != 1 or scores.shape[0] != boxes.shape[0]:
        raise ValueError("expected boxes [n, 4] and scores [n]")
    n, k = int(boxes.shape[0]), max(int(max_output_size), 0)
    need = ctypes.c_size_t()
    check(lib.k2y_nms_workspace_bytes(n, k, ctypes.byref(need)))
    dev = boxes.device
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    idx = torch.empty(max(min(n, k), 1), dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    if n and k:
        with torch.cuda.device(dev):
            check(lib.k2y_nms_boxes(boxes.data_ptr(), scores.data_ptr(), n, k, float(iou_threshold), idx.data_ptr(), cnt.data_ptr(),
                                    ws.data_ptr(), ws.numel(), _stream(boxes)))
    return idx[:int(cnt.item())]
