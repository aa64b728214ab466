"""world_size-2 gloo test of the data-parallel gradient exchange (CPU tensors): an all-reduce
SUM over R ranks of per-rank accumulators equals the reference's NUM_MINIBATCHES=R accumulation
(lib/ssnet.py:76-79), bucket by bucket."""
import os
import socket
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uresnet_b200 import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _table():
    t, off = OrderedDict(), 0
    for l, (wn, bn) in enumerate(((144, 16), (4608, 32), (9216, 32), (27, 3))):
        t["UResNet/l%d/weights" % l] = ((wn,), off, wn)
        off = (off + wn + 63) // 64 * 64
        t["UResNet/l%d/BatchNorm/beta" % l] = ((bn,), off, bn)
        off = (off + bn + 63) // 64 * 64
    return t, off


def _worker(rank, world, port, average, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table, total = _table()
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(total, generator=g)
    mine = flat.clone()
    ex = dp.GradientExchange(table, 4, flat, bucket_mb=0.01, average=average)
    assert len(ex.buckets) >= 2 and ex.events_array() is None      # CPU: synchronous path
    ex.zeroed()
    ex.backward_enqueued()
    ex.finish()
    others = [torch.randn(total, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    expect = sum(others) / (world if average else 1)
    ok = torch.allclose(flat, expect, atol=1e-6)
    s = ex.allreduce_scalars([1.0 + rank, 10.0])
    ok = ok and s == [sum(1.0 + r for r in range(world)), 10.0 * world]
    q.put((rank, bool(ok), float((flat - mine).abs().max())))
    dist.destroy_process_group()


@pytest.mark.parametrize("average", [False, True])
def test_gradient_exchange_world2(average):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, average, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(d > 0 for _, _, d in res)      # the exchange really changed the buffers
