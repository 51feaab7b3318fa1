"""N>1 host logic on CPU: image sharding + in-place all-gather of detection records (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from k210_yolo_framework_b200.dist import DetectionGather, shard_range
    total, C, K = 6, 4, 5
    lo, hi = shard_range(total, rank, world)
    g = DetectionGather(hi - lo, C, K, torch.device("cpu"))
    # every rank fills its slice with values that encode the GLOBAL image index
    for i in range(lo, hi):
        g.local_dets[i - lo] = i * 1000 + torch.arange(C * K * 6, dtype=torch.int32).reshape(C, K, 6)
        g.local_counts[i - lo] = i
    dets, counts = g.gather()
    ok = True
    for i in range(total):
        ok &= bool((dets[i] == i * 1000 + torch.arange(C * K * 6, dtype=torch.int32).reshape(C, K, 6)).all())
        ok &= bool((counts[i] == i).all())
    q.put((rank, ok, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    from k210_yolo_framework_b200.dist import shard_range
    for total in (1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_all_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in procs)
    for p in procs:
        p.join(30)
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [(0, 3), (3, 6)]
