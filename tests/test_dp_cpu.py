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


# --------------------------------------------------------------------------------------------------
# The semantics anchor with the REFERENCE'S OWN CODE on every rank: R ranks, each running
# /root/reference/lib/ssnet.py's accumulate step on its own minibatch (over oracle/tf1_shim.py), their
# accumulators all-reduced by the product's GradientExchange over gloo, one Adam apply per rank
#   ==  one process running the reference with NUM_MINIBATCHES = R on the same R minibatches.
# --------------------------------------------------------------------------------------------------
_HAVE_REF = os.path.exists("/root/reference/lib/ssnet.py")
_DIMS, _F, _S, _B = [16, 16, 1], 2, 2, 2


def _ref_net():
    from oracle import tf1_shim as shim
    from oracle import uresnet_oracle as O
    shim.reset_default_graph(torch.float64)
    uresnet = shim.load_reference()
    import tensorflow as tf
    net = uresnet(dims=_DIMS, num_class=3, num_strides=_S, base_num_outputs=_F)
    net.construct(trainable=True, use_weight=True, learning_rate=1e-3)
    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    P = O.init_params(2, base=_F, num_strides=_S, seed=21)
    tvs = tf.trainable_variables()
    for v in tvs:
        v.load(P[v.op.name], sess)
    return net, sess, tvs


def _minibatch(rank):
    from uresnet_b200 import synthetic_io as S
    d, l, w = S.make_batch(_B, _DIMS, (555, rank, 0))          # what synthetic_threadio draws on that rank
    return d, l, w / w.sum(axis=1, keepdims=True)


def _ref_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net, sess, tvs = _ref_net()
    net.zero_gradients(sess)
    net.accum_gradients(sess, *_minibatch(rank))
    # the product's layout: one flat accumulator, tensors at 64-element aligned offsets, two per layer
    table, off = OrderedDict(), 0
    for v in tvs:
        cnt = int(np.prod(v.shape.as_list()))
        table[v.op.name] = (tuple(v.shape.as_list()), off, cnt)
        off = (off + cnt + 63) // 64 * 64
    flat = torch.zeros(off, dtype=torch.float64)
    for a, (s, o, c) in zip(net._accum_vars, table.values()):
        flat[o:o + c] = a.value.reshape(-1)
    ex = dp.GradientExchange(table, len(tvs) // 2, flat, bucket_mb=0.0005)
    assert len(ex.buckets) >= 3
    ex.zeroed()
    ex.backward_enqueued()
    ex.finish()
    for a, (s, o, c) in zip(net._accum_vars, table.values()):
        a.load(flat[o:o + c].reshape(s).numpy(), sess)
    net.apply_gradients(sess)
    q.put((rank, {v.op.name: v.value.numpy().copy() for v in tvs}))
    dist.destroy_process_group()


@pytest.mark.skipif(not _HAVE_REF, reason="reference tree not on this box")
def test_two_ranks_of_reference_code_equal_its_two_minibatch_accumulation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ref_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    # one process, NUM_MINIBATCHES = 2 (ssnet_trainval.py:164-191)
    net, sess, tvs = _ref_net()
    before = {v.op.name: v.value.numpy().copy() for v in tvs}
    net.zero_gradients(sess)
    for r in range(2):
        net.accum_gradients(sess, *_minibatch(r))
    net.apply_gradients(sess)
    from oracle import tf1_shim as shim
    for v in tvs:
        want = v.value.numpy()
        assert not np.array_equal(want, before[v.op.name])                 # Adam moved it
        assert np.array_equal(res[0][v.op.name], want), v.op.name          # a + b is order-free: bit-identical
        assert np.array_equal(res[1][v.op.name], want), v.op.name          # replicas stay identical
    shim.uninstall()
