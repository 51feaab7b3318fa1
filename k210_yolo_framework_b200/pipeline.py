"""Batched detection pipeline: the whole hot path behind one call.

``DetectionPipeline.detect_host(x)`` is the user-facing end-to-end call for a batch of pre-processed images
held in (pinned) host memory: H2D copy -> network (one CUDA graph) -> decode + NMS -> [all-gather when the
job spans several GPUs] -> D2H of the fixed-size record blocks.  ``step_device()`` is the same work on inputs
already resident in HBM (what ``bench.py``'s ``value`` times).  ``submit``/``collect`` is the streaming form.
It is the batched form of keras_inference.py:87-135, which the reference runs for one image per process.

Streams: the network (one CUDA graph) runs on the caller's current stream; decode + NMS run on a second, high-priority
stream and read one of TWO head-buffer sets the network alternates between, so the decode of batch i overlaps the first
convolutions of batch i+1 (the small-grid layers leave SMs idle; the decode kernels are latency-bound).  With more than one
rank the all-gather (and the D2H copy of the gathered blocks) runs on a third stream into one of two gather buffers; H2D
copies use a fourth stream and land directly in one of two device input buffers the network is re-pointed at (one CUDA graph
per (input buffer, head set)), so there is no staging copy.  ``step_device(pipelined=False)`` (the default) makes the
caller's stream wait for the decode before returning — plain stream semantics; the streaming calls (``submit``/``collect``,
``step_device(pipelined=True)`` + ``wait_gathered()``) keep the overlap.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import yolonet
from .detect import DET_WORDS, KerasDetector
from .dist import DetectionGather

SLOTS = 2


class DetectionPipeline:
    def __init__(self, model_def: str, image_size: Sequence[int], anchors: np.ndarray, class_num: int, alpha: float,
                 batch: int, obj_thresh: float = 0.7, iou_thresh: float = 0.3, max_per_class: int = 30,
                 device: Optional[int] = None, world: int = 1, rank: int = 0, comm=None):
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
        self.gather = DetectionGather(batch, class_num, max_per_class, dev, world=self.world, rank=self.rank, words=DET_WORDS,
                                      slots=SLOTS, comm=comm)
        self._img_hw = torch.tensor([[image_size[0], image_size[1]]] * batch, dtype=torch.float32, device=dev)
        self._side = torch.cuda.Stream(device=dev)          # all-gather + D2H of gathered blocks
        self._det_stream = torch.cuda.Stream(device=dev, priority=-1)   # decode + NMS (its CTAs go first when SM slots free up)
        self._head_sets = None                              # two sets of head buffers, created on first use
        self._heads_ready = [torch.cuda.Event() for _ in range(SLOTS)]   # network finished writing head set
        self._copy_stream = torch.cuda.Stream(device=dev)   # H2D
        self._host = [torch.empty(self.gather.bufs[0].shape, dtype=torch.int32).pin_memory() for _ in range(SLOTS)]
        self._in_f32: Optional[List[torch.Tensor]] = None   # device input buffers of the streaming API, created on first use
        self._in_u8: Optional[List[torch.Tensor]] = None
        self._h2d_done = [torch.cuda.Event() for _ in range(SLOTS)]
        self._in_free = [torch.cuda.Event() for _ in range(SLOTS)]       # network finished reading input slot
        self._det_done = [torch.cuda.Event() for _ in range(SLOTS)]      # local block of gather slot written
        self._gather_done = [torch.cuda.Event() for _ in range(SLOTS)]   # gather slot complete on this rank
        self._slot_free = [torch.cuda.Event() for _ in range(SLOTS)]     # last reader of gather slot is done
        self._d2h_done = [torch.cuda.Event() for _ in range(SLOTS)]
        self._steps = 0
        self._submitted = 0

    def set_image_shapes(self, image_hw) -> None:
        """Original (pre-letterbox) image sizes, [batch, 2] (h, w); defaults to the network input size."""
        arr = np.broadcast_to(np.asarray(image_hw, np.float32).reshape(-1, 2), (self.batch, 2))
        self._img_hw.copy_(torch.from_numpy(np.array(arr, dtype=np.float32, order="C", copy=True)))

    # -- device-resident step ---------------------------------------------------------------------------------------
    def _step(self, n: int) -> int:
        """Network on the current stream into head set s, decode/NMS on the decode stream into gather slot s, then
        (world > 1) the all-gather on the side stream.  Returns s."""
        s = self._steps % SLOTS
        self._steps += 1
        compute = torch.cuda.current_stream(self.device_index)
        if self._head_sets is None:
            self._head_sets = [list(self.engine.head_buffers), self.engine.new_head_set()]
        # head set s was last read by the decode of two steps ago; that decode wrote gather slot s, too
        compute.wait_event(self._det_done[s])
        self.engine.bind_heads(self._head_sets[s])
        heads = self.engine.run(n)
        self._heads_ready[s].record(compute)
        det = self._det_stream
        det.wait_event(self._heads_ready[s])
        det.wait_event(self._slot_free[s])          # whoever still reads this gather slot (all-gather / D2H two steps ago)
        dets, counts = self.gather.local(s)
        self.detector.run(heads, self._img_hw[:n], dets_out=dets, counts_out=counts, stream=det)
        self._det_done[s].record(det)
        if self.world > 1:
            self._side.wait_event(self._det_done[s])
            self.gather.gather(s, stream=self._side)
            self._gather_done[s].record(self._side)
            self._slot_free[s].record(self._side)
        return s

    def step_device(self, n: Optional[int] = None, pipelined: bool = False):
        """Network + decode/NMS (+ all-gather) on the bound device input; asynchronous.  Returns the gathered
        (dets [world*batch, C, K, 6] int32, counts [world*batch, C]) device views.  ``pipelined=False``: the current stream
        waits for them (use them like any other result of the stream).  ``pipelined=True``: the current stream does NOT wait,
        so the next step's convolutions overlap this step's decode; the views are complete once the stream has passed
        ``wait_gathered()`` (or after a device synchronise), and stay valid for one more ``step_device``."""
        s = self._step(self.batch if n is None else n)
        self._last_slot = s
        if not pipelined:
            self.wait_gathered()
        return self.gather.views(s)

    def wait_gathered(self) -> None:
        """Makes the current stream wait for the decode (and all-gather) of the latest ``step_device``."""
        ev = self._gather_done[self._last_slot] if self.world > 1 else self._det_done[self._last_slot]
        torch.cuda.current_stream(self.device_index).wait_event(ev)

    # -- host API ---------------------------------------------------------------------------------------------------
    def _input_slots(self, u8: bool) -> List[torch.Tensor]:
        if u8:
            if self._in_u8 is None:
                self._in_u8 = [torch.empty(self.engine.input_buffer.shape, dtype=torch.uint8, device=self.engine.input_buffer.device)
                               for _ in range(SLOTS)]
            return self._in_u8
        if self._in_f32 is None:
            self._in_f32 = [torch.empty_like(self.engine.input_buffer) for _ in range(SLOTS)]
        return self._in_f32

    def detect_host(self, x_host: torch.Tensor):
        """x_host: CPU float32 (normalised) or uint8 (raw letterboxed RGB) [batch,H,W,3] tensor, pinned for an
        asynchronous copy.  Returns host tensors (dets, counts) for all images of the job, after a stream synchronise."""
        tk = self.submit(x_host)
        return self.collect(tk)

    def submit(self, x_host: torch.Tensor) -> int:
        """Streaming form of ``detect_host``: enqueue one batch (pinned host float32 or uint8 [n,H,W,3]) and return a
        ticket.  The H2D copy runs on a side stream straight into one of two device input buffers (the network is
        re-pointed at it), so it overlaps the previous batch's kernels; at most two batches are in flight —
        ``collect`` the ticket of batch i-1 before submitting batch i+1."""
        n = x_host.shape[0]
        slot = self._submitted % SLOTS
        self._submitted += 1
        compute = torch.cuda.current_stream(self.device_index)
        # raw letterboxed RGB (uint8): 4x fewer PCIe bytes, the normalisation (img / max(img)) is fused into the first conv
        buf = self._input_slots(x_host.dtype == torch.uint8)[slot]
        self._copy_stream.wait_event(self._in_free[slot])
        with torch.cuda.stream(self._copy_stream):
            buf[:n].copy_(x_host, non_blocking=True)
            self._h2d_done[slot].record(self._copy_stream)
        compute.wait_event(self._h2d_done[slot])
        self.engine.bind_input(buf)
        s = self._step(n)
        self._in_free[slot].record(compute)
        # D2H of the gathered blocks: behind the all-gather on the side stream (world > 1), else behind the decode
        st = self._side if self.world > 1 else self._det_stream
        with torch.cuda.stream(st):
            self._host[s].copy_(self.gather.bufs[s], non_blocking=True)
            self._d2h_done[s].record(st)
            self._slot_free[s].record(st)
        return s

    def collect(self, ticket: int):
        """Blocks until the batch behind `ticket` is on the host; returns (dets, counts) pinned host views (valid
        until the ticket's slot is reused two submits later)."""
        self._d2h_done[ticket].synchronize()
        return DetectionGather.split(self._host[ticket], self.gather.C, self.gather.K, DET_WORDS)

    def launches_per_step(self) -> int:
        return self.engine.launches_per_run() + 2  # + the decode scan and the per-class NMS kernel

    @staticmethod
    def records(dets: torch.Tensor, counts: torch.Tensor) -> List[list]:
        return KerasDetector.to_host(dets, counts)


class LanedPipeline:
    """Several batches in flight: ``lanes`` DetectionPipelines (own engine + activation arena, own streams, one shared NCCL
    communicator) fed round-robin, each on its own compute stream.  Consecutive batches are independent, so batch i+1's
    bandwidth-bound early layers run beside batch i's small-grid late layers — the multi-stream serving form of the same
    per-batch work; every batch still goes through the same kernels with the same batch size.  With ``sm_split`` (default) every
    lane sizes its persistent tensor-core grids for 1/lanes of the SMs, so the lanes' kernels share the GPU side by side.

    ``bind_input(x)`` / ``step_device()`` / ``wait_all()`` for device-resident inputs, ``submit`` / ``collect`` for pinned
    host batches (tickets carry the lane)."""

    def __init__(self, lanes: int, *args, sm_split: bool = True, sm_limit: Optional[int] = None, max_queued: Optional[int] = None, **kw):
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        first = DetectionPipeline(*args, **kw)
        self.lanes: List[DetectionPipeline] = [first] + [DetectionPipeline(*args, comm=first.gather.comm, **kw) for _ in range(lanes - 1)]
        self.batch, self.world, self.rank, self.device_index = first.batch, first.world, first.rank, first.device_index
        dev = torch.device("cuda", self.device_index)
        self._streams = [torch.cuda.Stream(device=dev) for _ in self.lanes]
        self._fork = torch.cuda.Event()
        self._next = 0          # next lane for step_device
        self._next_host = 0     # next lane for submit
        self._ran = [False] * lanes
        # host-side flow control of step_device: at most `max_queued` batches issued and not yet through their network — a bounded
        # queue, as a server has.  Default 2 per lane (= the gather / head-set slots of a lane, and what submit/collect allows):
        # the GPU always has a full step queued behind the running one, and the host never runs further ahead.  Deeper queues
        # measured WORSE with more than one rank (8 or unbounded: 119 k images/s on 2 GPUs, 4: 152 k — profiles/r02_lanes.md).
        self.max_queued = max(lanes, int(max_queued)) if max_queued is not None else 2 * lanes
        self._flow_events = [torch.cuda.Event() for _ in range(self.max_queued)]
        self._flow_pending = 0
        self._issued = 0
        # each lane plans its persistent tensor-core kernels for 1/lanes of the SMs: two lanes' one-CTA-per-SM grids then run
        # side by side (measured on cfg 2, two lanes: 69.3 k -> 77 k images/s; see profiles/r02_lanes.md)
        self.sm_limit = 0
        if sm_limit is not None:
            self.sm_limit = int(sm_limit)
        elif sm_split and lanes > 1:
            self.sm_limit = torch.cuda.get_device_properties(dev).multi_processor_count // lanes
        if self.sm_limit:
            for ln in self.lanes:
                ln.engine.set_sm_limit(self.sm_limit)

    def set_weights(self, weights) -> None:
        for ln in self.lanes:
            ln.engine.set_weights(weights)

    def set_math(self, mode) -> None:
        for ln in self.lanes:
            ln.engine.set_math(mode)

    def set_image_shapes(self, image_hw) -> None:
        for ln in self.lanes:
            ln.set_image_shapes(image_hw)

    def bind_input(self, x: torch.Tensor) -> None:
        """Device input of the NEXT ``step_device`` (float32 or uint8 [batch,H,W,3])."""
        self.lanes[self._next % len(self.lanes)].engine.bind_input(x)

    def step_device(self, n: Optional[int] = None):
        """One batch on the next lane (asynchronous, on the lane's stream, ordered after the caller's current stream).  The
        returned views are complete after ``wait_all()`` and stay valid until that lane has run two more batches."""
        k = self._next % len(self.lanes)
        self._next += 1
        cur = torch.cuda.current_stream(self.device_index)
        self._fork.record(cur)
        st = self._streams[k]
        st.wait_event(self._fork)
        slot = self._issued % self.max_queued
        if self._issued >= self.max_queued:
            self._flow_events[slot].synchronize()     # the batch issued max_queued steps ago has left its network
        with torch.cuda.stream(st):
            views = self.lanes[k].step_device(n, pipelined=True)
            self._flow_events[slot].record(st)
        self._issued += 1
        self._ran[k] = True
        return views

    def wait_all(self) -> None:
        """Makes the current stream wait for everything issued with ``step_device`` so far (decode / all-gather of every lane)."""
        for k, ln in enumerate(self.lanes):
            if self._ran[k]:
                ln.wait_gathered()

    def submit(self, x_host: torch.Tensor):
        k = self._next_host % len(self.lanes)
        self._next_host += 1
        with torch.cuda.stream(self._streams[k]):
            return (k, self.lanes[k].submit(x_host))

    def collect(self, ticket):
        k, t = ticket
        return self.lanes[k].collect(t)

    def in_flight_limit(self) -> int:
        """Batches that may be submitted before the oldest must be collected."""
        return SLOTS * len(self.lanes)

    def launches_per_step(self) -> int:
        return self.lanes[0].launches_per_step()
