"""Batched detection pipeline: the whole hot path behind one call.

``DetectionPipeline.detect_host(x)`` is the user-facing end-to-end call for a batch of pre-processed images
held in (pinned) host memory: H2D copy -> network (one CUDA graph) -> fused decode+NMS -> [all-gather when the
job spans several GPUs] -> D2H of the fixed-size record buffers.  ``step_device()`` is the same work on inputs
already resident in HBM (what ``bench.py``'s ``value`` times).  It is the batched form of
keras_inference.py:87-135, which the reference runs for one image per process.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import yolonet
from .detect import DET_WORDS, KerasDetector
from .dist import DetectionGather


class DetectionPipeline:
    def __init__(self, model_def: str, image_size: Sequence[int], anchors: np.ndarray, class_num: int, alpha: float,
                 batch: int, obj_thresh: float = 0.7, iou_thresh: float = 0.3, max_per_class: int = 30,
                 device: Optional[int] = None, world: int = 1, rank: int = 0):
        self.batch, self.world, self.rank = int(batch), int(world), int(rank)
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        builder = getattr(yolonet, model_def)
        self.model, self.wrapper = builder([image_size[0], image_size[1], 3], len(anchors[0]), class_num, alpha=alpha,
                                           max_batch=batch, device=self.device_index)
        self.engine = self.model.engine
        out_hw = [(h, w) for h, w, _ in self.engine.out_shapes]
        self.detector = KerasDetector(anchors, image_size, out_hw, class_num, obj_thresh, iou_thresh, max_per_class,
                                      max_batch=batch, device=self.device_index)
        dev = torch.device("cuda", self.device_index)
        self.gather = DetectionGather(batch, class_num, max_per_class, dev, world=self.world, rank=self.rank, words=DET_WORDS)
        self._img_hw = torch.tensor([[image_size[0], image_size[1]]] * batch, dtype=torch.float32, device=dev)
        self._host_dets = torch.empty(self.gather.dets.shape, dtype=torch.int32).pin_memory()
        self._host_counts = torch.empty(self.gather.counts.shape, dtype=torch.int32).pin_memory()

    def set_image_shapes(self, image_hw) -> None:
        """Original (pre-letterbox) image sizes, [batch, 2] (h, w); defaults to the network input size."""
        arr = np.broadcast_to(np.asarray(image_hw, np.float32).reshape(-1, 2), (self.batch, 2))
        self._img_hw.copy_(torch.from_numpy(np.ascontiguousarray(arr)))

    def step_device(self, n: Optional[int] = None):
        """Network + decode/NMS (+ all-gather) on the bound device input; asynchronous.  Returns the gathered
        (dets [world*batch, C, K, 6] int32, counts [world*batch, C]) device tensors."""
        n = self.batch if n is None else n
        heads = self.engine.run(n)
        self.detector.run(heads, self._img_hw[:n], dets_out=self.gather.local_dets, counts_out=self.gather.local_counts)
        return self.gather.gather()

    def detect_host(self, x_host: torch.Tensor):
        """x_host: CPU float32 [batch,H,W,3] tensor (pinned for an asynchronous copy).  Returns host tensors
        (dets, counts) for all images of the job, after a stream synchronise."""
        n = x_host.shape[0]
        self.engine.input_buffer[:n].copy_(x_host, non_blocking=True)
        dets, counts = self.step_device(n)
        self._host_dets.copy_(dets, non_blocking=True)
        self._host_counts.copy_(counts, non_blocking=True)
        torch.cuda.current_stream(self.device_index).synchronize()
        return self._host_dets, self._host_counts

    def launches_per_step(self) -> int:
        return self.engine.launches_per_run() + 1  # + the fused decode/NMS kernel

    @staticmethod
    def records(dets: torch.Tensor, counts: torch.Tensor) -> List[list]:
        return KerasDetector.to_host(dets, counts)
