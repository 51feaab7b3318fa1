"""Multi-GPU collation: images shard across ranks, ONE all-gather of fixed-size detection blocks per batch.

The reference is single-process (keras_inference.py:12-17); this is the one collective the B200 build adds
(SURVEY.md §8e).  Rank r owns images [r*B/G, (r+1)*B/G).  A gather buffer holds, per image, its ``C*K`` 24-byte
records followed by its ``C`` counts (``C*(6K+1)`` int32 words), so that one rank's block is contiguous: the NMS
kernel writes records and counts straight into this rank's block (``k2y_detect_keras_strided``) and a single in-place
``ncclAllGather`` — issued through the C-ABI (``k2y_allgather_detections``, the library's own communicator) —
completes the buffer on every rank, in global image order.  Buffers come in ``slots`` (two by default) so the gather of
batch i can run on a side stream while batch i+1 computes.  On CPU tensors (the gloo tests) the same layout is gathered
with ``torch.distributed``.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous image shard of `rank`; the remainder goes to the first ranks."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class Communicator:
    """The library's NCCL communicator (k2y_comm).  The 128-byte unique id is created on rank 0 and distributed with
    whatever process group launched the ranks (torch.distributed here)."""

    def __init__(self, device_index: int, world: Optional[int] = None, rank: Optional[int] = None):
        from ._lib import check, lib
        self.world = dist.get_world_size() if world is None else int(world)
        self.rank = dist.get_rank() if rank is None else int(rank)
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            check(lib.k2y_comm_unique_id(buf))
            uid = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
        if self.world > 1:
            on_gpu = dist.get_backend() == "nccl"
            t = uid.cuda(device_index) if on_gpu else uid
            dist.broadcast(t, src=0)
            uid = t.cpu()
        raw = (ctypes.c_ubyte * 128)(*uid.tolist())
        h = ctypes.c_void_p()
        check(lib.k2y_comm_create(raw, self.world, self.rank, int(device_index), ctypes.byref(h)))
        self._h = h

    def all_gather(self, buf: torch.Tensor, bytes_per_rank: int, stream: torch.cuda.Stream) -> None:
        from ._lib import check, lib
        check(lib.k2y_allgather_detections(self._h, buf.data_ptr(), int(bytes_per_rank), ctypes.c_void_p(stream.cuda_stream)))

    def nccl_version(self) -> int:
        from ._lib import check, lib
        v = ctypes.c_int()
        check(lib.k2y_comm_info(self._h, None, None, ctypes.byref(v)))
        return v.value

    def close(self) -> None:
        if getattr(self, "_h", None):
            from ._lib import lib
            lib.k2y_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DetectionGather:
    def __init__(self, per_rank_batch: int, class_num: int, max_per_class: int, device, world: int = None,
                 rank: int = None, words: int = 6, slots: int = 1, comm: Optional[Communicator] = None):
        self.world = dist.get_world_size() if world is None else world
        self.rank = dist.get_rank() if rank is None else rank
        self.n, self.C, self.K, self.words = int(per_rank_batch), int(class_num), int(max_per_class), int(words)
        self.img_words = self.C * (self.K * self.words + 1)   # per image: C*K records, then C counts
        self.bufs = [torch.zeros((self.world * self.n, self.img_words), dtype=torch.int32, device=device) for _ in range(slots)]
        self.comm = comm
        if self.world > 1 and self.bufs[0].is_cuda and self.comm is None:
            self.comm = Communicator(self.bufs[0].device.index)

    # -- layout ---------------------------------------------------------------
    @property
    def bytes_per_rank(self) -> int:
        return self.n * self.img_words * 4

    @staticmethod
    def split(buf: torch.Tensor, C: int, K: int, words: int = 6):
        """(dets [N,C,K,words], counts [N,C]) views of a gather buffer [N, C*(K*words+1)] (device or host)."""
        n, iw = buf.shape
        dets = buf.as_strided((n, C, K, words), (iw, K * words, words, 1))
        counts = buf.as_strided((n, C), (iw, 1), C * K * words)
        return dets, counts

    def views(self, slot: int = 0):
        return self.split(self.bufs[slot], self.C, self.K, self.words)

    def local(self, slot: int = 0):
        d, c = self.views(slot)
        return d[self.rank * self.n:(self.rank + 1) * self.n], c[self.rank * self.n:(self.rank + 1) * self.n]

    # views of slot 0 (the single-buffer form the gloo tests use)
    @property
    def dets(self) -> torch.Tensor:
        return self.views(0)[0]

    @property
    def counts(self) -> torch.Tensor:
        return self.views(0)[1]

    @property
    def local_dets(self) -> torch.Tensor:
        return self.local(0)[0]

    @property
    def local_counts(self) -> torch.Tensor:
        return self.local(0)[1]

    # -- the collective ---------------------------------------------------------
    def gather(self, slot: int = 0, stream: Optional[torch.cuda.Stream] = None):
        """In-place all-gather of buffer `slot` (no-op for world == 1); asynchronous on `stream` for CUDA buffers."""
        buf = self.bufs[slot]
        if self.world > 1:
            if buf.is_cuda:
                st = stream if stream is not None else torch.cuda.current_stream(buf.device.index)
                self.comm.all_gather(buf, self.bytes_per_rank, st)
            else:
                dist.all_gather_into_tensor(buf, buf[self.rank * self.n:(self.rank + 1) * self.n])
        return self.views(slot)
