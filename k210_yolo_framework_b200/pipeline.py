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
        # streaming mode (submit / collect): two staging slots so that the H2D copy of batch i+1 overlaps batch i
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._stage = [torch.empty_like(self.engine.input_buffer) for _ in range(2)]
        self._slot_dets = [torch.empty(self.gather.dets.shape, dtype=torch.int32).pin_memory() for _ in range(2)]
        self._slot_counts = [torch.empty(self.gather.counts.shape, dtype=torch.int32).pin_memory() for _ in range(2)]
        self._h2d_done = [torch.cuda.Event() for _ in range(2)]
        self._stage_free = [torch.cuda.Event() for _ in range(2)]
        self._d2h_done = [torch.cuda.Event() for _ in range(2)]
        self._submitted = 0
        self._stage_u8 = None

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
        """x_host: CPU float32 (normalised) or uint8 (raw letterboxed RGB) [batch,H,W,3] tensor, pinned for an
        asynchronous copy.  Returns host tensors (dets, counts) for all images of the job, after a stream synchronise."""
        n = x_host.shape[0]
        if x_host.dtype == torch.uint8:
            self.engine.enable_u8_input()[:n].copy_(x_host, non_blocking=True)
        else:
            self.engine.disable_u8_input()
            self.engine.input_buffer[:n].copy_(x_host, non_blocking=True)
        dets, counts = self.step_device(n)
        self._host_dets.copy_(dets, non_blocking=True)
        self._host_counts.copy_(counts, non_blocking=True)
        torch.cuda.current_stream(self.device_index).synchronize()
        return self._host_dets, self._host_counts

    def submit(self, x_host: torch.Tensor) -> int:
        """Streaming form of ``detect_host``: enqueue one batch (pinned host float32 [n,H,W,3]) and return a ticket.
        The H2D copy runs on a side stream into a staging slot, so it overlaps the previous batch's kernels; at
        most two batches are in flight — ``collect`` the ticket of batch i-1 before submitting batch i+1."""
        n = x_host.shape[0]
        slot = self._submitted % 2
        self._submitted += 1
        compute = torch.cuda.current_stream(self.device_index)
        u8 = x_host.dtype == torch.uint8
        if u8:  # raw letterboxed RGB: 4x fewer PCIe bytes, normalisation (img / max(img)) fused into the first conv
            if self._stage_u8 is None:
                self._stage_u8 = [torch.empty(self.engine.input_buffer.shape, dtype=torch.uint8, device=self.engine.input_buffer.device)
                                  for _ in range(2)]
            stage, dst = self._stage_u8[slot], self.engine.enable_u8_input()
        else:
            self.engine.disable_u8_input()
            stage, dst = self._stage[slot], self.engine.input_buffer
        self._copy_stream.wait_event(self._stage_free[slot])
        with torch.cuda.stream(self._copy_stream):
            stage[:n].copy_(x_host, non_blocking=True)
            self._h2d_done[slot].record(self._copy_stream)
        compute.wait_event(self._h2d_done[slot])
        dst[:n].copy_(stage[:n], non_blocking=True)
        self._stage_free[slot].record(compute)
        dets, counts = self.step_device(n)
        self._slot_dets[slot].copy_(dets, non_blocking=True)
        self._slot_counts[slot].copy_(counts, non_blocking=True)
        self._d2h_done[slot].record(compute)
        return slot

    def collect(self, ticket: int):
        """Blocks until the batch behind `ticket` is on the host; returns (dets, counts) pinned host tensors (valid
        until the ticket's slot is reused two submits later)."""
        self._d2h_done[ticket].synchronize()
        return self._slot_dets[ticket], self._slot_counts[ticket]

    def launches_per_step(self) -> int:
        return self.engine.launches_per_run() + 1  # + the fused decode/NMS kernel

    @staticmethod
    def records(dets: torch.Tensor, counts: torch.Tensor) -> List[list]:
        return KerasDetector.to_host(dets, counts)
