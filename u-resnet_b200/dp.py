"""Data-parallel gradient exchange: bucketed all-reduce overlapped with backward.

The reference has no multi-GPU path; its only batch-scaling mechanism is
``NUM_MINIBATCHES`` gradient accumulation -- a SUM of per-minibatch gradients, each
minibatch normalised with its own batch-norm statistics, followed by one Adam apply
(/root/reference/lib/ssnet.py:76-79, lib/ssnet_trainval.py:164-191).  R-rank data
parallelism with an all-reduce SUM of the accumulators and per-replica batch-norm
statistics is exactly ``MINIBATCH_SIZE=B/R, NUM_MINIBATCHES=R`` of the reference
(SURVEY.md 8e); ``average=True`` divides by R instead.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  The flat
gradient buffer is cut into contiguous buckets of whole layers in the order backward
finishes them (last layer first).  ``urn_backward`` records one CUDA event per layer when
that layer's filter+beta gradients are final; the communication stream waits on the
events of a bucket and launches its all-reduce while the rest of backward is still
running on the compute stream.  Adam waits for the last bucket.

On CPU tensors (gloo; tests/test_dp.py) the same bucket plan runs synchronously.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist


def plan_buckets(layer_ranges, bucket_elems):
    """layer_ranges: [(lo, hi)] element ranges per layer in creation order (contiguous,
    increasing).  Returns buckets [(lo, hi, [layer indices])] walking from the LAST layer
    to the first, each at least ``bucket_elems`` long (except the final one)."""
    buckets, cur = [], []
    for l in range(len(layer_ranges) - 1, -1, -1):
        cur.append(l)
        lo, hi = layer_ranges[cur[-1]][0], layer_ranges[cur[0]][1]
        if hi - lo >= bucket_elems:
            buckets.append((lo, hi, list(cur)))
            cur = []
    if cur:
        buckets.append((layer_ranges[cur[-1]][0], layer_ranges[cur[0]][1], list(cur)))
    return buckets


class GradientExchange(object):

    def __init__(self, param_table, num_layers, flat, group=None, bucket_mb=16.0, average=False,
                 force=False):
        """param_table: OrderedDict name -> (shape, offset, count) in creation order (two
        entries per layer: weights, beta); flat: the flat gradient accumulator tensor.
        ``force``: run the event / side-stream protocol even in a group of one rank (tests of the
        overlap machinery on a single GPU; the all-reduce of a one-rank group is the identity)."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.active = self.world > 1 or (bool(force) and dist.is_initialized())
        self.average = bool(average)
        self.time_finish = False          # record CUDA events around finish(): exposed all-reduce time
        self._finish_events = []
        self.flat = flat
        items = list(param_table.values())
        assert len(items) == 2 * num_layers
        self.layer_ranges = []
        for l in range(num_layers):
            (_, wo, wc), (_, bo, bc) = items[2 * l], items[2 * l + 1]
            hi = items[2 * l + 2][1] if l + 1 < num_layers else flat.numel()
            self.layer_ranges.append((wo, hi))          # includes the alignment padding
        bucket_mb = float(os.environ.get("URESNET_B200_DP_BUCKET_MB", bucket_mb))
        self.buckets = plan_buckets(self.layer_ranges, int(bucket_mb * (1 << 20) / 4))
        self.cuda = flat.is_cuda
        # URESNET_B200_DP_OVERLAP=0: no per-layer events; ONE all-reduce of the whole buffer before Adam
        self.overlap = os.environ.get("URESNET_B200_DP_OVERLAP", "1") != "0"
        self.expected_minibatches = 1
        self._seen = 0
        self._works = []
        self._launched = False
        if self.cuda:
            self.comm_stream = torch.cuda.Stream(device=flat.device)
            self.events = [torch.cuda.Event() for _ in range(num_layers)]
            for e in self.events:                       # force creation of the CUDA handles
                e.record()
            self._ev_array = (C.c_void_p * num_layers)(*[C.c_void_p(e.cuda_event) for e in self.events])
        else:
            self.events, self._ev_array = None, None

    # -- hooks called by ssnet_base ------------------------------------------------
    def zeroed(self):
        self._seen = 0
        self._launched = False

    def events_array(self):
        """Event array for urn_backward, or None when this backward is not the last
        minibatch of the step (gradients are still being accumulated locally)."""
        if not self.active or not self.cuda or not self.overlap:
            return None
        return self._ev_array if self._seen + 1 >= self.expected_minibatches else None

    def backward_enqueued(self):
        """Called right after urn_backward returned (work is only enqueued): launch every
        bucket's all-reduce on the communication stream behind that bucket's events."""
        self._seen += 1
        if not self.active or not self.cuda or not self.overlap or self._seen < self.expected_minibatches:
            return
        for lo, hi, layers in self.buckets:
            for l in layers:
                self.comm_stream.wait_event(self.events[l])
            with torch.cuda.stream(self.comm_stream):
                self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM,
                                                   group=self.group, async_op=True))
        self._launched = True

    def finish(self, flat=None):
        """Before the optimizer: make the compute stream wait for the exchange (or run it
        now if it was not overlapped)."""
        if not self.active:
            return
        timed = self.time_finish and self.cuda
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()                   # compute stream: backward has been enqueued up to here
        if not self._launched:
            if self.cuda and not self.overlap:
                self._works.append(dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            else:
                for lo, hi, _ in self.buckets:
                    self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM,
                                                       group=self.group, async_op=True))
        for w in self._works:
            w.wait()
        self._works = []
        self._launched = False
        if self.average:
            self.flat.mul_(1.0 / self.world)
        if timed:
            e1.record()                   # compute stream resumes here: e1 - e0 = exposed exchange time
            self._finish_events.append((e0, e1))

    def exposed_ms(self):
        """Mean milliseconds the compute stream waited for the exchange per step since the last
        call (needs ``time_finish``; synchronises)."""
        if not self._finish_events:
            return None
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in self._finish_events]
        self._finish_events = []
        return sum(ms) / len(ms)

    def allreduce_scalars(self, values):
        """Sum a small list of python floats over ranks (loss / accuracy counters)."""
        if self.world == 1:
            return list(values)
        t = torch.tensor(list(values), dtype=torch.float64, device=self.flat.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return [float(v) for v in t.cpu()]


_compute_stream = None


def use_high_priority_compute_stream(device):
    """Make a HIGH-priority stream the calling thread's current stream (once per process).  NCCL's
    kernels run on lowest-priority streams; with the compute kernels at default (= lowest) priority
    too, the CTAs of an in-flight bucket all-reduce displaced CTAs of the 148-CTA persistent conv /
    batch-norm kernels (measured at 2 GPUs: 2D 28.01 ms single-GPU step -> 28.79 ms; with the compute
    stream at high priority 28.50 ms, a single un-overlapped all-reduce 28.39 ms).
    ``URESNET_B200_DP_HIPRI=0`` keeps the default stream."""
    global _compute_stream
    if os.environ.get("URESNET_B200_DP_HIPRI", "1") == "0" or not torch.cuda.is_available():
        return None
    if _compute_stream is None:
        torch.cuda.synchronize(device)
        _compute_stream = torch.cuda.Stream(device=device, priority=-1)
    torch.cuda.set_stream(_compute_stream)
    return _compute_stream


def attach(net, group=None, bucket_mb=16.0, average=False, force=False):
    """Give a constructed, trainable ``uresnet`` a gradient exchange over ``group``."""
    if net._accum.is_cuda and (force or (dist.is_initialized() and dist.get_world_size(group) > 1)):
        use_high_priority_compute_stream(net._accum.device)
    ex = GradientExchange(net._param_table, net._num_layers, net._accum, group=group,
                          bucket_mb=bucket_mb, average=average, force=force)
    net._exchange = ex
    return ex


def broadcast_variables(net, src=0, group=None):
    """Make every replica start from rank ``src``'s variables."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(net._params, src=src, group=group)
