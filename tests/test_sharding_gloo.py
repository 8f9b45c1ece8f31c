"""
Multi-GPU host logic on CPU: world_size-2 gloo processes exercise the machine partition broadcast and the
summary gather that bench.py / the fleet API use over NCCL (the data path itself has no collective).
"""
import os
import socket

import numpy as np
import pytest

from gordo_components_b200 import fleet


def test_partition_is_contiguous_balanced_and_complete():
    for n, w in ((1000, 8), (1000, 3), (5, 8), (0, 2), (256, 8)):
        parts = fleet.partition(n, w)
        assert len(parts) == w
        flat = [m for p in parts for m in p]
        assert flat == list(range(n))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
    assert [len(p) for p in fleet.partition(1000, 8)] == [125] * 8


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = fleet.assign_machines(10, world, rank, dist)
        # stand-in for the per-machine score summary a rank would compute on its GPU
        local = torch.tensor([float(m) * 1.5 for m in mine], dtype=torch.float32)
        if len(mine) < 5:  # gather needs equal sizes: callers pad to the largest block
            local = torch.cat([local, torch.full((5 - len(mine),), float("nan"))])
        allv = fleet.gather_summaries(local, world, dist)
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max-over-ranks timing reduction used by bench.py
        q.put((rank, mine.tolist(), allv.tolist(), float(t.item())))
    finally:
        dist.destroy_process_group()


def test_assign_and_gather_world2_gloo():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, g0, t0), (r1, m1, g1, t1) = res
    assert m0 == [0, 1, 2, 3, 4] and m1 == [5, 6, 7, 8, 9]
    assert g0 == g1 == [m * 1.5 for m in range(10)]
    assert t0 == t1 == 2.0
