"""Host wrappers over the fused decode + NMS kernels (C-ABI ``k2y_detect_keras`` / ``k2y_region_run``).

``KerasDetector`` is the GPU form of keras_inference.py:94-135: it takes the head tensors ``predict``
returned (device or host) and yields, per image, the per-class survivors in the reference's output order
(class ascending, score descending).  ``RegionDetector`` is the batched device form of region_layer.c's
``region_layer_run`` for one output layer.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

DET_DTYPE = np.dtype([("ymin", "<f4"), ("xmin", "<f4"), ("ymax", "<f4"), ("xmax", "<f4"), ("score", "<f4"),
                      ("index", "<i4")])
DET_WORDS = 6  # 24-byte records == 6 x 32-bit words (k2y_det)


class KerasDetector:
    def __init__(self, anchors: np.ndarray, in_hw: Sequence[int], out_hw: Sequence[Sequence[int]], class_num: int,
                 obj_thresh: float = 0.7, iou_thresh: float = 0.3, max_per_class: int = 30, max_batch: int = 32,
                 device: Optional[int] = None):
        _lib.require_cuda()
        anchors = np.asarray(anchors, np.float64)
        out_hw = np.reshape(np.asarray(out_hw), (-1, 2))
        if anchors.ndim != 3 or anchors.shape[0] != len(out_hw) or anchors.shape[2] != 2:
            raise ValueError(f"anchors must be [L,A,2] with L == len(out_hw); got {anchors.shape} vs {len(out_hw)}")
        if not 1 <= len(out_hw) <= 3 or anchors.shape[1] > 8:
            raise ValueError("1..3 output layers and <= 8 anchors per layer are supported")
        self.cfg = _lib.DetectCfg()
        self.cfg.n_layers = len(out_hw)
        for l, (h, w) in enumerate(out_hw):
            self.cfg.layer_h[l], self.cfg.layer_w[l] = int(h), int(w)
        self.cfg.anchor_num, self.cfg.class_num = anchors.shape[1], int(class_num)
        flat = anchors.astype(np.float32).reshape(-1)
        for i, v in enumerate(flat):
            self.cfg.anchors[i] = float(v)
        self.cfg.in_h, self.cfg.in_w = int(in_hw[0]), int(in_hw[1])
        self.cfg.obj_thresh, self.cfg.iou_thresh = float(obj_thresh), float(iou_thresh)
        self.cfg.max_per_class = int(max_per_class)
        self.A, self.C, self.K = anchors.shape[1], int(class_num), int(max_per_class)
        self.out_hw = [tuple(int(v) for v in hw) for hw in out_hw]
        self.max_batch = int(max_batch)
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        dev = torch.device("cuda", self.device_index)
        nbytes = ctypes.c_size_t()
        check(lib.k2y_detect_workspace_bytes(ctypes.byref(self.cfg), self.max_batch, ctypes.byref(nbytes)))
        self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        # records as int32 words so that torch can hold / all-gather them; viewed as k2y_det on the host
        self.dets = torch.zeros((self.max_batch, self.C, self.K, DET_WORDS), dtype=torch.int32, device=dev)
        self.counts = torch.zeros((self.max_batch, self.C), dtype=torch.int32, device=dev)
        self._img_hw = torch.empty((self.max_batch, 2), dtype=torch.float32, device=dev)

    def set_thresholds(self, obj_thresh: float, iou_thresh: float) -> None:
        self.cfg.obj_thresh, self.cfg.iou_thresh = float(obj_thresh), float(iou_thresh)

    def run(self, heads: List[torch.Tensor], image_hw, dets_out: Optional[torch.Tensor] = None,
            counts_out: Optional[torch.Tensor] = None, stream: Optional[torch.cuda.Stream] = None):
        """heads[l]: CUDA f32 [N,h_l,w_l,A*(5+C)] (or [N,h,w,A,5+C]); image_hw: [N,2] (orig_h, orig_w) tensor/array
        or a single (h, w).  Returns (dets int32 [N,C,K,6] device, counts int32 [N,C] device) — asynchronous."""
        n = heads[0].shape[0]
        if n > self.max_batch:
            raise ValueError(f"batch {n} > max_batch {self.max_batch}")
        for l, t in enumerate(heads):
            h, w = self.out_hw[l]
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n * h * w * self.A * (5 + self.C):
                raise ValueError(f"head {l}: expected contiguous CUDA float32 with {n * h * w * self.A * (5 + self.C)} elements")
        if isinstance(image_hw, torch.Tensor) and image_hw.is_cuda:
            img = image_hw.to(torch.float32).reshape(n, 2).contiguous()
        else:
            arr = np.asarray(image_hw, np.float32)
            arr = np.broadcast_to(arr.reshape(-1, 2), (n, 2))
            self._img_hw[:n].copy_(torch.from_numpy(np.array(arr, dtype=np.float32, order="C", copy=True)), non_blocking=True)
            img = self._img_hw
        dets = self.dets if dets_out is None else dets_out
        counts = self.counts if counts_out is None else counts_out
        # outputs may be strided per image (records and counts of one image adjacent inside a gather block, dist.py)
        if (dets.dtype != torch.int32 or counts.dtype != torch.int32 or tuple(dets.shape[1:]) != (self.C, self.K, DET_WORDS) or
                tuple(dets.stride()[1:]) != (self.K * DET_WORDS, DET_WORDS, 1) or tuple(counts.shape[1:]) != (self.C,) or
                counts.stride(1) != 1 or dets.shape[0] < n or counts.shape[0] < n):
            raise ValueError("dets_out must be int32 [>=N,C,K,6] dense per image, counts_out int32 [>=N,C]")
        st = stream if stream is not None else torch.cuda.current_stream(self.device_index)
        ptrs = (ctypes.c_void_p * len(heads))(*[t.data_ptr() for t in heads])
        check(lib.k2y_detect_keras_strided(ctypes.byref(self.cfg), ptrs, n, img.data_ptr(), dets.data_ptr(), counts.data_ptr(),
                                           dets.stride(0), counts.stride(0), self._ws.data_ptr(), self._ws.numel(),
                                           ctypes.c_void_p(st.cuda_stream)))
        return dets[:n], counts[:n]

    @staticmethod
    def to_host(dets: torch.Tensor, counts: torch.Tensor) -> List[List[Tuple]]:
        """Device records -> per image list of (class, index, score, ymin, xmin, ymax, xmax) in reference order."""
        d = dets.cpu().numpy()
        c = counts.cpu().numpy()
        rec = np.ascontiguousarray(d).view(DET_DTYPE)[..., 0]  # [N, C, K]
        out = []
        for b in range(rec.shape[0]):
            img = []
            for cls in range(rec.shape[1]):
                for k in range(int(c[b, cls])):
                    r = rec[b, cls, k]
                    img.append((cls, int(r["index"]), float(r["score"]), float(r["ymin"]), float(r["xmin"]),
                                float(r["ymax"]), float(r["xmax"])))
            out.append(img)
        return out


class RegionDetector:
    """Batched REGION_C decode+NMS of one output layer (region_layer.c:378-383 semantics)."""

    def __init__(self, width, height, anchors, classes, net_w, net_h, threshold, nms_value, image_w=320, image_h=224,
                 max_batch: int = 32, device: Optional[int] = None):
        _lib.require_cuda()
        anchors = np.asarray(anchors, np.float32).reshape(-1)
        self.cfg = _lib.RegionCfg()
        self.cfg.layer_w, self.cfg.layer_h = int(width), int(height)
        self.cfg.anchor_num, self.cfg.classes = len(anchors) // 2, int(classes)
        self.cfg.net_w, self.cfg.net_h, self.cfg.image_w, self.cfg.image_h = int(net_w), int(net_h), int(image_w), int(image_h)
        for i, v in enumerate(anchors):
            self.cfg.anchors[i] = float(v)
        self.cfg.threshold, self.cfg.nms_value = float(threshold), float(nms_value)
        self.N = int(width) * int(height) * (len(anchors) // 2)
        self.C = int(classes)
        self.max_batch = int(max_batch)
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        dev = torch.device("cuda", self.device_index)
        nbytes = ctypes.c_size_t()
        check(lib.k2y_region_workspace_bytes(ctypes.byref(self.cfg), self.max_batch, ctypes.byref(nbytes)))
        self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        self.probs = torch.zeros((self.max_batch, self.N, self.C + 1), dtype=torch.float32, device=dev)
        self.boxes = torch.zeros((self.max_batch, self.N, 4), dtype=torch.float32, device=dev)

    def run(self, chw: torch.Tensor):
        """chw: CUDA f32 [B, A, 5+C, H, W].  Returns (probs [B,N,C+1], boxes [B,N,4]) device tensors."""
        n = chw.shape[0]
        if not chw.is_cuda or chw.dtype != torch.float32 or not chw.is_contiguous() or chw.numel() != n * self.N * (5 + self.C):
            raise ValueError("expected contiguous CUDA float32 [B, A, 5+C, H, W]")
        st = torch.cuda.current_stream(self.device_index)
        check(lib.k2y_region_run(ctypes.byref(self.cfg), chw.data_ptr(), n, None, self.probs.data_ptr(),
                                 self.boxes.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                 ctypes.c_void_p(st.cuda_stream)))
        return self.probs[:n], self.boxes[:n]
