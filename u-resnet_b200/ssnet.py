"""``ssnet_base``: host-side mirror of /root/reference/lib/ssnet.py:8-153.

Same constructor, same method names / arity / return structure; ``sess`` arguments are
accepted and ignored (there is no TensorFlow session -- the work is done by the CUDA
kernels behind ``include/uresnet_b200.h``).  PyTorch is used only for device buffers,
pinned staging memory and streams.

Differences that are deliberate and visible:
  * ``construct`` takes extra keyword-only knobs (``dtype``, ``impl``, ``device``, ``seed``); the default
    ``dtype`` is ``'f32'`` like the reference, ``'mix16'`` (fp16 activations, bf16 gradients, tcgen05 tensor
    cores: ~13x faster, gradients within the graph's own sensitivity to 2^-11 operand rounding) is opt-in.
  * variables are exposed through ``get_variables`` / ``set_variables`` keyed by the TF
    variable names (``UResNet/<scope>/weights``, ``.../BatchNorm/beta``), TF layouts.
  * out-of-range labels raise ``ValueError`` (TF-CPU raises, TF-GPU yields NaN).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

from . import _capi


class _optimizer_state(object):
    """Stands in for ``tf.train.AdamOptimizer`` (ssnet.py:72-75); ``_lr`` is read by the
    report line of the driver (ssnet_trainval.py:209)."""

    def __init__(self, lr):
        self._lr = lr
        self._beta1, self._beta2, self._epsilon = 0.9, 0.999, 1e-8
        self._t = 0


# bytes per host->device chunk of the pinned staging pipeline
_H2D_CHUNK_BYTES = int(float(os.environ.get("URESNET_B200_H2D_CHUNK_MB", "4")) * (1 << 20))


class ssnet_base(object):

    def __init__(self, dims, num_class):
        self._dims = np.array(dims, np.int32)
        if not len(self._dims) in [3, 4]:
            print('Error: len(dims) =', len(self._dims), 'but only 3 (H,W,C) or 4 (H,W,D,C) supported!')
            raise NotImplementedError
        self._num_class = int(num_class)
        self._constructed = False

    def _build(self, input_tensor=None):
        raise NotImplementedError

    # ------------------------------------------------------------------ construct
    def construct(self, trainable=True, use_weight=True, learning_rate=None, *, dtype=None,
                  impl=None, device=None, seed=1234):
        import os
        self._trainable = bool(trainable)
        self._use_weight = bool(use_weight)
        self._learning_rate = learning_rate
        # default numeric mode = the reference's: float32 storage and arithmetic (ssnet.py:30-32).  The 16-bit
        # tensor-core mode is opt-in: dtype='mix16' / DTYPE 'mix16' in the .cfg / URESNET_B200_DTYPE=mix16
        self._dtype = _capi.DTYPES[(dtype or os.environ.get("URESNET_B200_DTYPE", "f32")).lower()]
        self._impl = _capi.IMPLS[(impl or os.environ.get("URESNET_B200_IMPL", "auto")).lower()]
        self._lib = _capi.load()          # raises if the CUDA library is missing: no fallback
        if not torch.cuda.is_available():
            raise RuntimeError("uresnet_b200 needs a CUDA device (B200, sm_100a); no CPU fallback")
        self._device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._net_cfg = self._build(None)  # dict(nd, spatial, in_channels, base_filters, num_strides)
        self._handles = {}                 # batch -> (handle, workspace tensor)
        self._data_size = int(np.prod(self._dims))
        self._label_size = int(np.prod(self._dims[:-1]))

        # parameter table from a batch-1 handle (layout does not depend on the batch)
        h = self._create_handle(1, training=False)
        self._param_table = OrderedDict()
        n = self._lib.urn_param_count(h)
        for i in range(n):
            name, rank = C.c_char_p(), C.c_int()
            shape = (C.c_int64 * 6)()
            off, cnt = C.c_int64(), C.c_int64()
            _capi.check(self._lib.urn_param_info(h, i, C.byref(name), C.byref(rank), shape,
                                                 C.byref(off), C.byref(cnt)))
            self._param_table[name.value.decode()] = (tuple(shape[k] for k in range(rank.value)),
                                                      off.value, cnt.value)
        self._param_total = int(self._lib.urn_param_total(h))
        self._num_layers = int(self._lib.urn_layer_count(h))
        self._lib.urn_destroy(h)

        dev = self._device
        self._params = torch.zeros(self._param_total, dtype=torch.float32, device=dev)
        self._init_variables(seed)
        self._softmax = None
        if self._trainable:
            # accum_vars / Adam slots (ssnet.py:53-55,72-79)
            self._accum = torch.zeros_like(self._params)
            self._zero_pending = False
            self._adam_m = torch.zeros_like(self._params)
            self._adam_v = torch.zeros_like(self._params)
            lr = 0.001 if (learning_rate is None or learning_rate <= 0) else float(learning_rate)
            self._opt = _optimizer_state(lr)
        self._result = torch.zeros(4, dtype=torch.float64, device=dev)
        self._result_host = torch.zeros(4, dtype=torch.float64).pin_memory()
        self._staging = {}
        self._pinned_arrays = []
        self._graphs = {}                 # CUDA graphs of the forward / loss+backward launch sequences
        self._graph_calls = {}
        self._use_graphs = os.environ.get("URESNET_B200_GRAPHS", "1") != "0"
        self._profiling = False
        self._copy_stream = None          # side stream for label / weight H2D copies
        self._side_pending = []
        self._step_start = None
        self._exchange = None              # set by dp.attach(): gradient allreduce hook
        self._last_handle = None           # (handle, workspace) of the last loss call
        self._constructed = True

    # --------------------------------------------------------------- variables
    def _init_variables(self, seed):
        """slim default xavier_initializer() for filters, zeros for beta."""
        rng = np.random.Generator(np.random.PCG64(int(seed)))
        host = np.zeros(self._param_total, np.float32)
        for name, (shape, off, cnt) in self._param_table.items():
            if name.endswith("/weights"):
                rf = int(np.prod(shape[:-2]))
                lim = math.sqrt(6.0 / (rf * shape[-2] + rf * shape[-1]))
                host[off:off + cnt] = rng.uniform(-lim, lim, cnt).astype(np.float32)
        self._params.copy_(torch.from_numpy(host))

    def trainable_variables(self):
        """[(name, shape)] in tf.trainable_variables() order."""
        return [(k, v[0]) for k, v in self._param_table.items()]

    def get_variables(self, what="params"):
        if what == "grads" and getattr(self, "_zero_pending", False):
            self._materialise_zero()
        src = {"params": self._params, "grads": getattr(self, "_accum", None),
               "adam_m": getattr(self, "_adam_m", None), "adam_v": getattr(self, "_adam_v", None)}[what]
        host = src.detach().cpu().numpy()
        return OrderedDict((k, host[o:o + c].reshape(s).copy()) for k, (s, o, c) in self._param_table.items())

    def set_variables(self, values, strict=True):
        host = self._params.detach().cpu().numpy().copy()
        for k, v in values.items():
            if k not in self._param_table:
                if strict:
                    raise KeyError(k)
                continue
            s, o, c = self._param_table[k]
            v = np.asarray(v, np.float32)
            if tuple(v.shape) != tuple(s):
                raise ValueError("shape mismatch for %s: %s vs %s" % (k, v.shape, s))
            host[o:o + c] = v.reshape(-1)
        self._params.copy_(torch.from_numpy(host))

    # ----------------------------------------------------------------- handles
    def _create_handle(self, batch, training):
        cfg = _capi.urn_net_cfg()
        nc = self._net_cfg
        cfg.nd = nc["nd"]
        for i in range(3):
            cfg.spatial[i] = nc["spatial"][i] if i < len(nc["spatial"]) else 1
        cfg.in_channels = nc["in_channels"]
        cfg.batch = int(batch)
        cfg.base_filters = nc["base_filters"]
        cfg.num_strides = nc["num_strides"]
        cfg.num_class = self._num_class
        cfg.dtype, cfg.impl, cfg.training = self._dtype, self._impl, int(bool(training))
        h = C.c_void_p()
        _capi.check(self._lib.urn_create(C.byref(cfg), C.byref(h)))
        return h

    def _handle(self, batch):
        key = int(batch)
        if key not in self._handles:
            for k in list(self._handles):           # one live workspace at a time per mode
                hh, _ = self._handles.pop(k)
                self._lib.urn_destroy(hh)
            h = self._create_handle(batch, self._trainable)
            nbytes = int(self._lib.urn_workspace_bytes(h))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
            self._handles[key] = (h, ws)
        return self._handles[key]

    def workspace_bytes(self, batch):
        h = self._create_handle(batch, self._trainable)
        n = int(self._lib.urn_workspace_bytes(h))
        self._lib.urn_destroy(h)
        return n

    def __del__(self):
        try:
            for h, _ in getattr(self, "_handles", {}).values():
                self._lib.urn_destroy(h)
        except Exception:
            pass

    # ------------------------------------------------------------------- feeds
    def feed_dict(self, input_data, input_label=None, input_weight=None):
        """Validates like ssnet.py:141-153 and stages the host arrays on the device."""
        if input_weight is None and self._use_weight:
            sys.stderr.write('Network configured to use loss pixel-weighting. Cannot run w/ input_weight=None...\n')
            raise TypeError
        feed = {"data": self._to_device("data", input_data, self._data_size)}
        batch = feed["data"].shape[0]
        if input_label is not None:
            feed["label"] = self._to_device("label", input_label, self._label_size, batch)
        if input_weight is not None:
            feed["weight"] = self._to_device("weight", input_weight, self._label_size, batch)
        return feed

    # Staging: every stream (data / label / weight) owns a ring of two slots, each a pinned host
    # buffer + a device buffer, used alternately so that the transfer of minibatch k+1 can run while
    # the kernels of minibatch k still read the other slot.
    def _slot(self, key, shape, index=None):
        ring = self._staging.get(key)
        if ring is None or ring["shape"] != tuple(shape):
            ring = self._staging[key] = {"shape": tuple(shape), "slots": [None, None], "next": 0}
        i = ring["next"] if index is None else index
        ring["next"] = 1 - i
        if ring["slots"][i] is None:
            ring["slots"][i] = {"pin": None, "dev": torch.empty(shape, dtype=torch.float32, device=self._device),
                                "event": None, "sig": None}
        return ring["slots"][i]

    @staticmethod
    def _signature(a):
        return (a.ctypes.data, a.shape, a.strides)

    def _copy_in(self, slot, a, stream):
        """host array -> slot's device buffer on `stream` (async).  A pinned source is DMA'd directly;
        a pageable one goes through the slot's pinned buffer in chunks, so that the DMA of chunk k runs
        while the host copies chunk k+1 into pinned memory."""
        if slot["event"] is not None:
            slot["event"].synchronize()             # previous transfer out of this slot's pinned buffer is done
        src = torch.from_numpy(a).view(-1)
        dst = slot["dev"].view(-1)
        with torch.cuda.stream(stream):
            if src.is_pinned():
                dst.copy_(src, non_blocking=True)
            else:
                if slot["pin"] is None:
                    slot["pin"] = torch.empty(a.shape, dtype=torch.float32).pin_memory()
                pin = slot["pin"].view(-1)
                step = max(1, _H2D_CHUNK_BYTES // 4)
                for i in range(0, max(src.numel(), 1), step):
                    pin[i:i + step].copy_(src[i:i + step])                  # multi-threaded host copy
                    dst[i:i + step].copy_(pin[i:i + step], non_blocking=True)
            slot["event"] = torch.cuda.Event()
            slot["event"].record(stream)

    def _copy_stream_ready(self):
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self._device)
        if self._step_start is not None:
            # every kernel enqueued before the current step started has finished: the slot that is
            # about to be overwritten (last read one step earlier) is free
            self._copy_stream.wait_event(self._step_start)
        return self._copy_stream

    def _to_device(self, key, arr, row, batch=None, side=False):
        """Host array -> device (async).  If the array was announced with ``stage_next`` its transfer is
        already in flight (or done) and only an event wait is enqueued.  With ``side=True`` the copy
        is enqueued on the copy stream so that it (and the host-side memcpy into pinned memory that
        precedes it) overlaps the forward pass already running on the compute stream; the compute
        stream is made to wait for it by ``_join_side_copies``."""
        if isinstance(arr, torch.Tensor) and arr.is_cuda:
            t = arr.to(torch.float32).reshape(-1, row)
        else:
            a = np.ascontiguousarray(np.asarray(arr, dtype=np.float32).reshape(-1, row))
            ring = self._staging.get(key)
            slot = None
            if ring is not None and ring["shape"] == a.shape:
                for i, sl in enumerate(ring["slots"]):
                    if sl is not None and sl["sig"] is not None and sl["sig"] == self._signature(a):
                        slot, sl["sig"] = sl, None                # prefetched by stage_next: consume once
                        ring["next"] = 1 - i
            if slot is not None:
                self._side_pending.append(slot["event"])
                if not side:
                    self._join_side_copies()
            else:
                if ring is not None:
                    for sl in ring["slots"]:                      # a staged minibatch that was skipped expires
                        if sl is not None:
                            sl["sig"] = None
                slot = self._slot(key, a.shape)
                stream = self._copy_stream_ready() if side else torch.cuda.current_stream(self._device)
                self._copy_in(slot, a, stream)
                if side:
                    self._side_pending.append(slot["event"])
            t = slot["dev"]
        if batch is not None and t.shape[0] != batch:
            raise ValueError("batch mismatch between data and %s" % key)
        return t

    def stage_next(self, input_data, input_label=None, input_weight=None):
        """Announce the NEXT minibatch: its host->device transfer starts now on the copy stream, under
        the kernels of the minibatch that was just enqueued (``accum_gradients(..., fetch=False)``).
        The ``accum_gradients`` / ``run_test`` / ``inference`` call that later receives the same arrays
        (recognised by address and shape) only waits for the transfer.  The arrays must not be
        modified in between.  Pinned arrays (``pinned_empty``) are DMA'd without a host-side copy."""
        stream = self._copy_stream_ready()
        for key, arr, row in (("data", input_data, self._data_size), ("label", input_label, self._label_size),
                              ("weight", input_weight, self._label_size)):
            if arr is None or (isinstance(arr, torch.Tensor) and arr.is_cuda):
                continue
            a = np.asarray(arr, dtype=np.float32).reshape(-1, row)
            if not a.flags["C_CONTIGUOUS"]:
                continue                                   # would be copied again anyway
            slot = self._slot(key, a.shape)
            self._copy_in(slot, a, stream)
            slot["sig"] = self._signature(a)

    def pinned_empty(self, shape):
        """float32 numpy array in page-locked host memory (kept alive by this object): inputs built in
        such arrays reach the device without the pageable -> pinned staging copy."""
        t = torch.empty(tuple(int(v) for v in shape), dtype=torch.float32).pin_memory()
        self._pinned_arrays.append(t)
        return t.numpy()

    def _join_side_copies(self):
        cur = torch.cuda.current_stream(self._device)
        for ev in self._side_pending:
            cur.wait_event(ev)
        self._side_pending = []

    def _feed_overlapped(self, input_data, input_label, input_weight):
        """feed_dict for the hot entry points: data first, forward enqueued, THEN label / weight
        are staged while the GPU is busy (they are only needed by the loss kernel)."""
        if input_weight is None and self._use_weight:
            sys.stderr.write('Network configured to use loss pixel-weighting. Cannot run w/ input_weight=None...\n')
            raise TypeError
        feed = {"data": self._to_device("data", input_data, self._data_size)}
        batch = feed["data"].shape[0]
        self._step_start = torch.cuda.Event()
        self._step_start.record(torch.cuda.current_stream(self._device))
        h, ws, batch = self._forward(feed)
        if input_label is not None:
            feed["label"] = self._to_device("label", input_label, self._label_size, batch, side=True)
        if input_weight is not None:
            feed["weight"] = self._to_device("weight", input_weight, self._label_size, batch, side=True)
        self._join_side_copies()
        return feed, h, ws, batch

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def _graphed(self, key, launch, dp_safe=False):
        """Run `launch()` (a sequence of library launches on the current stream with fixed
        pointers) -- eagerly the first two times for a key, then captured once into a CUDA graph
        and replayed: ~300 kernel launches become one cudaGraphLaunch.  In data-parallel runs the
        per-layer ready events of the backward sequence are captured as external event-record
        nodes (urn_backward), so the communication stream keeps waiting on them after each replay;
        nothing is captured while the per-layer profiler records."""
        if not self._use_graphs or (self._exchange is not None and not dp_safe) or self._profiling:
            return launch()
        g = self._graphs.get(key)
        if g is not None:
            g.replay()
            return
        n = self._graph_calls.get(key, 0) + 1
        self._graph_calls[key] = n
        if n <= 2:
            return launch()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                launch()
        except Exception as e:                      # capture not possible: stay eager, say so once
            sys.stderr.write("uresnet_b200: CUDA graph capture failed (%r); running eagerly\n" % (e,))
            self._use_graphs = False
            torch.cuda.synchronize(self._device)
            return launch()
        self._graphs[key] = g
        g.replay()

    def _forward(self, feed):
        batch = feed["data"].shape[0]
        h, ws = self._handle(batch)
        dptr, pptr, wptr = feed["data"].data_ptr(), self._params.data_ptr(), ws.data_ptr()
        self._graphed(("fwd", batch, dptr, wptr),
                      lambda: _capi.check(self._lib.urn_forward(h, pptr, dptr, wptr, self._stream())), dp_safe=True)
        return h, ws, batch

    def _loss_metrics(self, h, ws, feed, want_grad, want_softmax):
        sm = None
        if want_softmax:
            batch = feed["data"].shape[0]
            sm = torch.empty((batch,) + tuple(int(d) for d in self._dims[:-1]) + (self._num_class,),
                             dtype=torch.float32, device=self._device)
        lab = feed.get("label")
        wgt = feed.get("weight") if self._use_weight else None
        self._last_handle = (h, ws)
        _capi.check(self._lib.urn_loss_metrics(
            h, lab.data_ptr() if lab is not None else None,
            wgt.data_ptr() if wgt is not None else None, feed["data"].data_ptr(),
            sm.data_ptr() if sm is not None else None, self._result.data_ptr(), int(want_grad),
            ws.data_ptr(), self._stream()))
        return sm

    def _fetch_metrics(self):
        """Device->host read of (loss, n_correct_all, n_correct_nz, n_nz); synchronises."""
        self._result_host.copy_(self._result, non_blocking=True)
        torch.cuda.current_stream(self._device).synchronize()
        loss, ca, cnz, nnz = (float(v) for v in self._result_host)
        if math.isnan(loss):
            # two causes: a label outside [0, num_class) (the loss kernel raises a device flag and
            # poisons the loss; TF-CPU raises, TF-GPU yields NaN) or a genuinely diverged run
            # (non-finite logits / weights), which TF reports as a NaN loss and so does this
            if self._last_handle is not None:
                h, ws = self._last_handle
                if self._lib.urn_check_labels(h, ws.data_ptr(), self._stream()) != 0:
                    raise ValueError("label out of range [0, %d)" % self._num_class)
            # diverged run: like TF, hand the NaN back to the caller
        return loss, ca, cnz, nnz

    def _metrics(self, batch):
        loss, ca, cnz, nnz = self._fetch_metrics()
        acc_all = ca / float(batch * self._label_size)
        acc_nz = (cnz / nnz) if nnz > 0 else float("nan")   # tf.reduce_mean of empty -> NaN
        return np.float32(loss), np.float32(acc_all), np.float32(acc_nz)

    def fetch_metrics(self, batch):
        """Metrics of the last ``accum_gradients(..., fetch=False)`` call, in its return format
        (synchronises the compute stream)."""
        loss, acc_all, acc_nz = self._metrics(batch)
        return [None, loss, acc_all, acc_nz], ['', 'loss', 'acc. all', 'acc. nonzero']

    # ------------------------------------------------------- reference methods
    def make_summary(self, sess, input_data, input_label, input_weight=None):
        """ssnet.py:91-97 -- another full forward; returns the scalar summaries as a dict."""
        (loss, acc_all, acc_nz), _ = self.run_test(sess, input_data, input_label, input_weight)
        return {"loss": float(loss), "accuracy_all": float(acc_all), "accuracy_nonzero": float(acc_nz)}

    def zero_gradients(self, sess=None):
        """ssnet.py:99-101 (assign zeros to the accumulators).  No memset is launched: the next
        ``accum_gradients`` runs ``urn_backward`` with accumulate = 0, i.e. it STORES the gradients
        instead of adding them to zeros (the accumulators are neither cleared nor re-read: one launch
        and 2 x 67 / 199 MB of traffic less per step); whoever looks at the accumulators before that
        (``get_variables('grads')``, ``apply_gradients``) gets real zeros."""
        self._zero_pending = True
        if self._exchange is not None:
            self._exchange.zeroed()
        return [None]

    def _materialise_zero(self):
        if self._zero_pending:
            self._accum.zero_()
            self._zero_pending = False

    def accum_gradients(self, sess, input_data, input_label, input_weight=None, fetch=True):
        """ssnet.py:103-115: one fused forward+backward, gradients ADDED to the accumulators.
        Inputs may be host numpy arrays (staged through pinned memory) or CUDA tensors.
        ``fetch=False`` skips the device->host read of the metrics (nothing synchronises)."""
        if not self._trainable:
            raise RuntimeError("network constructed with trainable=False")
        feed, h, ws, batch = self._feed_overlapped(input_data, input_label, input_weight)
        events = self._exchange.events_array() if self._exchange is not None else None
        accumulate = 0 if self._zero_pending else 1      # first minibatch after zero_gradients: store
        self._zero_pending = False

        def loss_and_backward():
            self._loss_metrics(h, ws, feed, want_grad=True, want_softmax=False)
            _capi.check(self._lib.urn_backward(h, self._params.data_ptr(), self._accum.data_ptr(), accumulate,
                                               ws.data_ptr(), events, self._stream()))
        # `events is not None` is part of the key: the capture of the last minibatch of a
        # data-parallel step holds the external per-layer event-record nodes, the capture of the
        # other minibatches (NUM_MINIBATCHES > 1) must not be replayed in its place
        key = ("bwd", batch, feed["data"].data_ptr(), feed["label"].data_ptr(),
               feed["weight"].data_ptr() if (self._use_weight and "weight" in feed) else 0, ws.data_ptr(),
               events is not None, accumulate)
        self._graphed(key, loss_and_backward, dp_safe=os.environ.get("URESNET_B200_DP_GRAPH", "1") != "0")
        if self._exchange is not None:
            self._exchange.backward_enqueued()
        if not fetch:
            return [None, None, None, None], ['', 'loss', 'acc. all', 'acc. nonzero']
        loss, acc_all, acc_nz = self._metrics(batch)
        return [None, loss, acc_all, acc_nz], ['', 'loss', 'acc. all', 'acc. nonzero']

    def apply_gradients(self, sess=None):
        """ssnet.py:117-119: Adam on (accumulated gradient, variable) pairs."""
        self._materialise_zero()                  # zero_gradients() with no minibatch in between
        if self._exchange is not None:
            self._exchange.finish(self._accum)
        o = self._opt
        o._t += 1
        _capi.check(self._lib.urn_adam_step(
            self._params.data_ptr(), self._accum.data_ptr(), self._adam_m.data_ptr(),
            self._adam_v.data_ptr(), self._param_total, o._lr, o._beta1, o._beta2, o._epsilon, o._t,
            self._stream()))
        return [None]

    def run_test(self, sess, input_data, input_label, input_weight=None):
        feed, h, ws, batch = self._feed_overlapped(input_data, input_label, input_weight)
        self._loss_metrics(h, ws, feed, want_grad=False, want_softmax=False)
        loss, acc_all, acc_nz = self._metrics(batch)
        return [loss, acc_all, acc_nz], ['loss', 'acc. all', 'acc. nonzero']

    def inference(self, sess, input_data, input_label=None, with_decision=False, threshold=1.0):
        """ssnet.py:130-139: [softmax] or [softmax, acc_all, acc_nonzero]; softmax is a host
        numpy array [B,H,W,(D,)num_class].  ``with_decision`` appends the shower/track map
        of ssnet_trainval.py:269-287 ([B,H,W,(D)] f32: 0 where data <= threshold, 1 shower,
        2 track), computed on the device from the same forward pass."""
        # like the reference, this raises TypeError when constructed with use_weight=True
        # (ssnet.py:132 feeds no weights; ana configs set USE_WEIGHTS False)
        feed = self.feed_dict(input_data=input_data, input_label=input_label)
        h, ws, batch = self._forward(feed)
        sm = self._loss_metrics(h, ws, feed, want_grad=False, want_softmax=True)
        dec = self._decision(h, ws, feed, threshold) if with_decision else None
        out = [sm.cpu().numpy()]
        if input_label is not None:
            _, acc_all, acc_nz = self._metrics(batch)
            out += [acc_all, acc_nz]
        if dec is not None:
            out.append(dec.cpu().numpy())
        return out

    def _decision(self, h, ws, feed, threshold=1.0):
        batch = feed["data"].shape[0]
        dec = torch.empty((batch,) + tuple(int(d) for d in self._dims[:-1]), dtype=torch.float32,
                          device=self._device)
        _capi.check(self._lib.urn_decision(h, feed["data"].data_ptr(), float(threshold), dec.data_ptr(),
                                           ws.data_ptr(), self._stream()))
        return dec

    def decision(self, sess, input_data, input_label=None, threshold=1.0):
        """Analysis fast path: [decision] or [decision, acc_all, acc_nonzero] without the
        device->host read of the softmax (1 instead of num_class floats per pixel come back)."""
        feed = self.feed_dict(input_data=input_data, input_label=input_label)
        h, ws, batch = self._forward(feed)
        if input_label is not None:
            self._loss_metrics(h, ws, feed, want_grad=False, want_softmax=False)
        out = [self._decision(h, ws, feed, threshold).cpu().numpy()]
        if input_label is not None:
            _, acc_all, acc_nz = self._metrics(batch)
            out += [acc_all, acc_nz]
        return out

    # ------------------------------------------------------------ measurement
    PHASES = ("conv_fwd", "bn_stats", "bn_apply", "bn_bwd", "wgrad", "dgrad", "misc", "-")

    def profile_enable(self, batch, on=True):
        self._profiling = bool(on)
        h, _ = self._handle(batch)
        _capi.check(self._lib.urn_profile_enable(h, int(on)))

    def profile_collect(self, batch):
        """Returns (ms[layers,8], counts[layers,8]) accumulated since the last collect."""
        h, _ = self._handle(batch)
        n = self._num_layers * 8
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        _capi.check(self._lib.urn_profile_collect(h, ms, cnt))
        return (np.array(ms[:]).reshape(self._num_layers, 8), np.array(cnt[:]).reshape(self._num_layers, 8))

    def layer_table(self, batch):
        """[(scope, desc, is_transpose)] of the plan (for FLOP / byte accounting)."""
        h, _ = self._handle(batch)
        out = []
        for i in range(self._num_layers):
            scope, d, tr = C.c_char_p(), _capi.urn_conv_desc(), C.c_int()
            _capi.check(self._lib.urn_layer_info(h, i, C.byref(scope), C.byref(d), C.byref(tr), None, None, None, None))
            out.append((scope.value.decode(), d, bool(tr.value)))
        return out

    # ------------------------------------------------------------ conveniences
    def logits(self, batch):
        """f32 logits [B, ..., num_class] left in the workspace by the last forward (a view)."""
        h, ws = self._handle(batch)
        off = int(self._lib.urn_logits_offset_bytes(h))
        n = batch * self._label_size * self._num_class
        return ws[off:off + 4 * n].view(torch.float32).view(
            (batch,) + tuple(int(d) for d in self._dims[:-1]) + (self._num_class,))
