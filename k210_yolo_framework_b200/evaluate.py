"""Precision / recall of the confidence channel, as the reference's Keras metrics define them
(/root/reference/tools/custom.py:13-75 ``Yolo_Precision`` / ``Yolo_Recall``), counted on the GPU (k2y_pr_counts)."""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import torch

from ._lib import check, lib, require_cuda


class PrecisionRecall:
    """Accumulates tp / fp / fn over batches of (label, prediction) head tensors ``[..., A, 5+C]`` or ``[..., A*(5+C)]``."""

    def __init__(self, class_num: int, threshold: float = 0.5, apply_sigmoid: bool = False, device=None):
        require_cuda()
        self.entry = 5 + int(class_num)
        self.threshold, self.apply_sigmoid = float(threshold), bool(apply_sigmoid)
        self.counts = torch.zeros(3, dtype=torch.int64, device=device or "cuda")

    def update(self, y_true: Sequence[torch.Tensor], y_pred: Sequence[torch.Tensor]) -> None:
        for t, p in zip(y_true, y_pred):
            if not (t.is_cuda and p.is_cuda and t.dtype == p.dtype == torch.float32 and t.is_contiguous() and p.is_contiguous()
                    and t.numel() == p.numel() and t.numel() % self.entry == 0):
                raise ValueError("expected contiguous CUDA float32 tensors of equal size, a multiple of 5 + class_num floats")
            st = torch.cuda.current_stream(t.device.index)
            check(lib.k2y_pr_counts(t.data_ptr(), p.data_ptr(), t.numel() // self.entry, self.entry, self.threshold,
                                    int(self.apply_sigmoid), self.counts.data_ptr(), ctypes.c_void_p(st.cuda_stream)))

    def result(self) -> Tuple[float, float]:
        tp, fp, fn = (int(v) for v in self.counts.cpu())
        return (tp / (tp + fp) if tp + fp else 0.0), (tp / (tp + fn) if tp + fn else 0.0)

    def reset(self) -> None:
        self.counts.zero_()
