"""Multi-GPU collation: images shard across ranks, one NCCL all-gather of fixed-size detection records.

The reference is single-process (keras_inference.py:12-17); this is the one collective the B200 build adds
(SURVEY.md §8e).  Rank r owns images [r*B/G, (r+1)*B/G); the NMS kernel writes its records straight into
this rank's slice of the gather buffer, and ``all_gather_into_tensor`` (in place) completes the buffer on
every rank, so gathered results are in global image order.  Works with the ``nccl`` backend on GPUs and
with ``gloo`` on CPU tensors (tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous image shard of `rank`; the remainder goes to the first ranks."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class DetectionGather:
    def __init__(self, per_rank_batch: int, class_num: int, max_per_class: int, device, world: int = None,
                 rank: int = None, words: int = 6):
        self.world = dist.get_world_size() if world is None else world
        self.rank = dist.get_rank() if rank is None else rank
        self.n = per_rank_batch
        self.dets = torch.zeros((self.world * self.n, class_num, max_per_class, words), dtype=torch.int32, device=device)
        self.counts = torch.zeros((self.world * self.n, class_num), dtype=torch.int32, device=device)

    @property
    def local_dets(self) -> torch.Tensor:
        return self.dets[self.rank * self.n:(self.rank + 1) * self.n]

    @property
    def local_counts(self) -> torch.Tensor:
        return self.counts[self.rank * self.n:(self.rank + 1) * self.n]

    def gather(self):
        """In-place all-gather of both buffers (no-op for world == 1)."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.dets, self.local_dets)
            dist.all_gather_into_tensor(self.counts, self.local_counts)
        return self.dets, self.counts
