"""A small emulation of the TensorFlow-1.x API surface that the reference's model code uses, so that
``/root/reference/lib/{ssnet,uresnet,resnet_module}.py`` can be IMPORTED AND EXECUTED UNMODIFIED in a
container that has neither TensorFlow 1.x nor Python 2.  TEST INFRASTRUCTURE ONLY (see
``oracle/__init__.py``): nothing under ``u-resnet_b200/`` imports this.

Why: the reference keeps its arithmetic in TensorFlow (un-vendored, un-installable here), but the graph
COMPOSITION -- which layer feeds which, the residual / skip / concat wiring, scopes and variable creation
order, the loss reduction, the accuracy expressions, the accumulate / zero / Adam-apply protocol -- is the
reference's own Python (lib/uresnet.py:22-123, lib/resnet_module.py:10-87, lib/ssnet.py:20-153).  With this
module installed as ``tensorflow`` that code builds a lazy graph of the nodes below and ``Session.run``
evaluates it with torch-CPU (float64 by default) and autograd.  ``oracle/uresnet_oracle.py`` is thereby
checked against *the reference's code running*, not only against a reading of it
(tests/test_reference_run.py; golden vectors tests/golden/gref*.npz, made by
tests/golden/make_reference_run_golden.py).

What is still a restatement: the semantics of the ~25 TensorFlow / slim ops themselves, written here from
the TF-1.x definitions, independently of the oracle's formulation:

* ``slim.conv{2,3}d``: NHWC, ``padding='same'`` = TF SAME (out = ceil(n/s), total pad
  max((out-1)s + k - n, 0), the smaller half first), Xavier-uniform ``weights`` [k..,Cin,Cout]; with a
  ``normalizer_fn`` there are no biases (slim's rule).
* ``slim.conv{2,3}d_transpose``: TF lowers it to ``Conv{2,3}DBackpropInput`` with output size n*s (the
  reference's GraphDef dump shows exactly that, tests/test_reference_graph.py): evaluated here LITERALLY as
  the gradient, with respect to its input, of the SAME stride-s convolution -- not as a cropped
  ``conv_transpose`` as the oracle states it.
* ``slim.batch_norm`` defaults: decay 0.999, center, no scale, epsilon 0.001, ``is_training=True``: batch
  moments over all axes but the last, biased variance, ``x * rsqrt(var + eps) + (beta - mean * rsqrt(var +
  eps))`` (the expansion in the dump); the moving-average updates go to UPDATE_OPS, which the reference
  never runs (ssnet.py:72-79 has no control dependency on them), so the moving statistics stay at 0 / 1.
* ``tf.train.AdamOptimizer``: ``lr_t = lr sqrt(1 - b2^t) / (1 - b1^t)``; ``m += (1-b1)(g-m)``;
  ``v += (1-b2)(g^2-v)``; ``var -= lr_t m / (sqrt(v) + eps)``.
* ``sparse_softmax_cross_entropy_with_logits``, ``argmax`` (first maximum), ``where`` / ``gather_nd``.

Python-2 idioms of the reference handled from outside, without touching its files: ``xrange`` is put
into ``builtins``; ``net.get_shape()[-1].value / 2`` (uresnet.py:67, true division under Python 3) yields
256.0, which hashes like 256 as a dict key and is cast by ``num_outputs = int(...)`` here.
"""
from __future__ import annotations

import builtins
import contextlib
import math
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REFERENCE_LIB = "/root/reference/lib"

float32 = "float32"
int64 = "int64"
int32 = "int32"


class _Graph(object):
    def __init__(self, float_dtype=torch.float64, seed=0):
        self.scopes = []
        self.variables = []          # creation order (tf.global_variables())
        self.names = set()
        self.summaries = []
        self.layers = {}             # scope -> dict(raw=<conv output>, out=<after normalizer / activation>): fetchable
        self.float = float_dtype
        self.rng = np.random.Generator(np.random.PCG64(seed))


_G = _Graph()


def reset_default_graph(float_dtype=torch.float64, seed=0):
    """tf.reset_default_graph(); ``float_dtype`` = what tf.float32 tensors are evaluated in (float64 to
    compare with the fp64 oracle, float32 for 'what TF-CPU would give')."""
    global _G
    _G = _Graph(float_dtype, seed)


def set_random_seed(seed):
    _G.rng = np.random.Generator(np.random.PCG64(seed))


def _tdt(tok):
    return {float32: _G.float, int64: torch.int64, int32: torch.int32}[tok]


def _scoped(name):
    return "/".join(_G.scopes + [name]) if name else "/".join(_G.scopes)


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    _G.scopes.append(name)
    try:
        yield name
    finally:
        _G.scopes.pop()


name_scope = variable_scope


# ------------------------------------------------------------------------------------------ tensors
class Dimension(object):
    def __init__(self, v):
        self.value = None if v is None else int(v)

    def __repr__(self):
        return "Dimension(%s)" % self.value


class TensorShape(object):
    def __init__(self, dims):
        self._d = [None if d is None else int(d) for d in dims]

    def __len__(self):
        return len(self._d)

    def __getitem__(self, i):
        return Dimension(self._d[i])

    def as_list(self):
        return list(self._d)

    def __repr__(self):
        return "(%s)" % ", ".join("?" if d is None else str(d) for d in self._d)


class _Ctx(object):
    """one Session.run: feeds, memoised values, autograd leaves of the trainable variables, deferred assigns"""

    def __init__(self, feeds):
        self.feeds, self.memo, self.leaves, self.deferred, self.grads = feeds, {}, {}, [], {}

    def leaf(self, var):
        if var not in self.leaves:
            v = var.value.detach().clone()
            self.leaves[var] = v.requires_grad_(True) if var.trainable else v
        return self.leaves[var]


class Tensor(object):
    def __init__(self, fn, inputs, shape, dtype=float32, name=""):
        self._fn, self._inputs = fn, tuple(inputs)
        self._shape, self.dtype, self.name = TensorShape(shape), dtype, name

    @property
    def shape(self):
        return self._shape

    def get_shape(self):
        return self._shape

    def _eval(self, ctx):
        if self not in ctx.memo:
            ctx.memo[self] = self._fn(ctx, *[i._eval(ctx) if isinstance(i, Tensor) else i for i in self._inputs])
        return ctx.memo[self]

    def __add__(self, other):
        return Tensor(lambda c, a, b: a + b, (self, other), self._shape.as_list(), self.dtype)

    __radd__ = __add__

    def __mul__(self, other):
        return Tensor(lambda c, a, b: a * b, (self, other), self._shape.as_list(), self.dtype)

    def __gt__(self, other):
        return Tensor(lambda c, a, b: a > b, (self, other), self._shape.as_list(), "bool")


def _const(v):
    return Tensor(lambda c: torch.tensor(v, dtype=_G.float), (), np.shape(v))


def to_float(x):
    return _const(float(x)) if not isinstance(x, Tensor) else cast(x, float32)


def placeholder(dtype, shape=None, name=None):
    t = Tensor(None, (), shape, dtype, _scoped(name or "Placeholder") + ":0")

    def fn(ctx):
        if t not in ctx.feeds:
            raise ValueError("You must feed a value for placeholder tensor '%s'" % t.name)
        v = torch.as_tensor(np.asarray(ctx.feeds[t]), dtype=_tdt(dtype))
        want = t.shape.as_list()
        if len(want) != v.dim() or any(w is not None and w != s for w, s in zip(want, v.shape)):
            raise ValueError("Cannot feed value of shape %s for Tensor '%s', which has shape %s"
                             % (tuple(v.shape), t.name, t.shape))
        return v
    t._fn = fn
    return t


class Variable(Tensor):
    def __init__(self, initial_value=None, trainable=True, name=None, shape=None, _init=None):
        if isinstance(initial_value, Tensor):
            shape = initial_value.shape.as_list()
            _init = lambda: initial_value._eval(_Ctx({})).detach().clone()         # noqa: E731
        full, n = _scoped(name or "Variable"), 0
        while (full if n == 0 else "%s_%d" % (full, n)) in _G.names:     # TF uniquifies: Variable, Variable_1, ...
            n += 1
        full = full if n == 0 else "%s_%d" % (full, n)
        _G.names.add(full)
        Tensor.__init__(self, None, (), shape, float32, full + ":0")
        self.trainable, self._init, self.value, self._initial = bool(trainable), _init, None, None
        self.op = types.SimpleNamespace(name=self.name[:-2])
        _G.variables.append(self)

    def _eval(self, ctx):
        if self.value is None:
            raise RuntimeError("Attempting to use uninitialized value %s" % self.name)
        return ctx.leaf(self)

    def _initialize(self):
        self._initial = self._init().to(_G.float)
        self.value = self._initial.clone()

    def initialized_value(self):
        def fn(ctx):
            if self._initial is None:
                self._initialize()
            return self._initial
        return Tensor(fn, (), self.shape.as_list())

    def _assign(self, value, add):
        value = value if isinstance(value, Tensor) else _const(value)

        def fn(ctx, v):
            v = v.detach()

            def apply():
                self.value = (self.value + v) if add else v.clone().to(_G.float)
            ctx.deferred.append(apply)
            return None
        return Tensor(fn, (value,), self.shape.as_list())

    def assign(self, value):
        return self._assign(value, False)

    def assign_add(self, value):
        return self._assign(value, True)

    def load(self, value, session=None):
        v = torch.as_tensor(np.asarray(value), dtype=_G.float)
        assert list(v.shape) == self.shape.as_list(), (self.name, tuple(v.shape), self.shape)
        self.value = v.clone()


def get_variable(name, shape, init, trainable=True):
    if _scoped(name) in _G.names:       # tf.get_variable (what slim calls), unlike tf.Variable, never renames
        raise ValueError("Variable %s already exists, disallowed. Did you mean to set reuse=True in VarScope?"
                         % _scoped(name))
    v = Variable(name=name, shape=shape, trainable=trainable, _init=init)
    v._model = True                  # slim puts layer variables into GraphKeys.MODEL_VARIABLES
    return v


class GraphKeys(object):
    MODEL_VARIABLES = "model_variables"
    TRAINABLE_VARIABLES = "trainable_variables"
    GLOBAL_VARIABLES = "variables"


def get_collection(key):
    if key == GraphKeys.MODEL_VARIABLES:
        return [v for v in _G.variables if getattr(v, "_model", False)]
    if key == GraphKeys.TRAINABLE_VARIABLES:
        return trainable_variables()
    if key == GraphKeys.GLOBAL_VARIABLES:
        return global_variables()
    return []


def trainable_variables():
    return [v for v in _G.variables if v.trainable]


def global_variables():
    return list(_G.variables)


def global_variables_initializer():
    def fn(ctx):
        for v in _G.variables:
            v._initialize()
    return Tensor(fn, (), ())


def _infer(shape, new):
    new = [int(d) for d in new]
    if -1 in new and all(d is not None for d in shape):
        known = -int(np.prod(new))
        new[new.index(-1)] = int(np.prod(shape)) // known
    return [None if d == -1 else d for d in new]


def reshape(t, shape, name=None):
    shp = [int(d) for d in np.asarray(shape).reshape(-1)]
    return Tensor(lambda c, v: v.reshape(shp), (t,), _infer(t.shape.as_list(), shp), t.dtype)


def cast(t, dtype):
    return Tensor(lambda c, v: v.to(_tdt(dtype)), (t,), t.shape.as_list(), dtype)


def argmax(t, axis=None, **k):
    s = t.shape.as_list()
    return Tensor(lambda c, v: torch.argmax(v, dim=int(axis)), (t,), s[:axis] + s[axis + 1:], int64)


def equal(a, b):
    return Tensor(lambda c, x, y: x == y, (a, b), a.shape.as_list(), "bool")


def reduce_mean(t, axis=None):
    if axis is None:
        return Tensor(lambda c, v: v.mean(), (t,), ())
    s = t.shape.as_list()
    return Tensor(lambda c, v: v.mean(dim=axis), (t,), s[:axis] + s[axis + 1:])


def reduce_sum(t, axis=None):
    if axis is None:
        return Tensor(lambda c, v: v.sum(), (t,), ())
    s = t.shape.as_list()
    return Tensor(lambda c, v: v.sum(dim=axis), (t,), s[:axis] + s[axis + 1:])


def where(cond):
    return Tensor(lambda c, v: torch.nonzero(v), (cond,), [None, len(cond.shape)], int64)


def gather_nd(params, indices):
    return Tensor(lambda c, p, i: p[tuple(i.t())], (params, indices), [None], params.dtype)


def multiply(a, b):
    return Tensor(lambda c, x, y: x * y, (a, b), a.shape.as_list())


def zeros_like(t):
    return Tensor(lambda c, v: torch.zeros_like(v), (t,), t.shape.as_list())


def concat(values, axis, name=None):
    s = values[0].shape.as_list()
    cs = [v.shape.as_list()[axis] for v in values]
    s[axis] = None if any(c is None for c in cs) else sum(cs)
    return Tensor(lambda c, *v: torch.cat(v, dim=axis), tuple(values), s, values[0].dtype,
                  _scoped(name or "concat") + ":0")


# ------------------------------------------------------------------------------------------ tf.nn
def _relu(t):
    return Tensor(lambda c, v: F.relu(v), (t,), t.shape.as_list())


def _softmax(logits=None, **k):
    return Tensor(lambda c, v: torch.softmax(v, dim=-1), (logits,), logits.shape.as_list())


def _sparse_ce(labels=None, logits=None, **k):
    def fn(c, lab, lg):
        if int(lab.min()) < 0 or int(lab.max()) >= lg.shape[-1]:
            raise ValueError("Received a label value outside the valid range [0, %d)" % lg.shape[-1])
        return -torch.log_softmax(lg, dim=-1).gather(-1, lab.unsqueeze(-1)).squeeze(-1)
    return Tensor(fn, (labels, logits), labels.shape.as_list())


nn = types.SimpleNamespace(relu=_relu, softmax=_softmax, sparse_softmax_cross_entropy_with_logits=_sparse_ce)


# ------------------------------------------------------------------------------------------ slim layers
def _same_pads(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _to_cf(v):       # N..C -> NC..
    nd = v.dim() - 2
    return v.permute(0, nd + 1, *range(1, nd + 1))


def _to_cl(v):       # NC.. -> N..C
    nd = v.dim() - 2
    return v.permute(0, *range(2, nd + 2), 1)


def _conv_same_cf(x, w_tf, stride):
    """TF Conv2D / Conv3D, padding SAME, on a channel-first tensor; w_tf [k.., Cin, Cout]."""
    nd = x.dim() - 2
    pads = []
    for d in reversed(range(nd)):
        pads += list(_same_pads(x.shape[2 + d], w_tf.shape[d], stride))
    w = w_tf.permute(nd + 1, nd, *range(nd))
    return (F.conv2d, F.conv3d)[nd - 2](F.pad(x, pads), w, stride=stride)


def _conv_backprop_input_cf(x, w_tf, stride):
    """TF Conv{2,3}DBackpropInput(input_sizes = stride x the size of x, filter = w_tf [k.., Cout, Cin],
    out_backprop = x, SAME): the vector-Jacobian product of the SAME stride-s convolution of a tensor z of
    that size with w_tf read as [k.., in = Cout, out = Cin].  Differentiable in x and w_tf."""
    z = torch.zeros((x.shape[0], w_tf.shape[-2]) + tuple(n * stride for n in x.shape[2:]),
                    dtype=x.dtype, requires_grad=True)
    y = _conv_same_cf(z, w_tf, stride)
    assert y.shape == x.shape
    (dz,) = torch.autograd.grad(y, [z], grad_outputs=x, create_graph=True)
    return dz


def xavier_uniform(shape):
    rf = int(np.prod(shape[:-2]))
    lim = math.sqrt(6.0 / (rf * shape[-2] + rf * shape[-1]))
    return lambda: torch.as_tensor(_G.rng.uniform(-lim, lim, size=shape), dtype=_G.float)


def _zeros(shape):
    return lambda: torch.zeros(shape, dtype=_G.float)


def _ones(shape):
    return lambda: torch.ones(shape, dtype=_G.float)


def batch_norm(inputs, decay=0.999, center=True, scale=False, epsilon=0.001, activation_fn=None,
               is_training=True, trainable=True, scope=None, **unused):
    assert center and not scale and is_training, "only slim.batch_norm's defaults are emulated"
    c = inputs.shape.as_list()[-1]
    with variable_scope(scope or "BatchNorm"):
        beta = get_variable("beta", [c], _zeros([c]), trainable)
        get_variable("moving_mean", [c], _zeros([c]), False)
        get_variable("moving_variance", [c], _ones([c]), False)

    def fn(ctx, x, b):
        axes = tuple(range(x.dim() - 1))
        mean = x.mean(dim=axes)
        var = ((x - mean.detach()) ** 2).mean(dim=axes)          # SquaredDifference(x, StopGradient(mean))
        inv = torch.rsqrt(var + epsilon)
        return x * inv + (b - mean * inv)
    out = Tensor(fn, (inputs, beta), inputs.shape.as_list())
    return activation_fn(out) if activation_fn is not None else out


_DEFAULT = object()


def _conv_layer(nd, transpose):
    def layer(inputs, num_outputs, kernel_size, stride=1, padding="SAME", activation_fn=_relu,
              normalizer_fn=None, normalizer_params=None, weights_initializer=None, biases_initializer=_DEFAULT,
              trainable=True, scope=None, **unused):
        assert str(padding).upper() == "SAME", padding
        s = inputs.shape.as_list()
        assert len(s) == nd + 2, "slim.conv%dd on a rank-%d tensor" % (nd, len(s))
        cin, cout = s[-1], int(num_outputs)
        k = [int(kernel_size)] * nd if np.isscalar(kernel_size) else [int(v) for v in kernel_size]
        stride = int(stride)
        with variable_scope(scope or ("Conv%dd%s" % (nd, "_transpose" if transpose else ""))):
            wshape = k + ([cout, cin] if transpose else [cin, cout])
            w = get_variable("weights", wshape, xavier_uniform(wshape), trainable)
            b = None
            if normalizer_fn is None and biases_initializer is not None:
                b = get_variable("biases", [cout], _zeros([cout]), trainable)
            if transpose:
                out = Tensor(lambda c, x, wv: _to_cl(_conv_backprop_input_cf(_to_cf(x), wv, stride)), (inputs, w),
                             [s[0]] + [None if n is None else n * stride for n in s[1:-1]] + [cout])
            else:
                out = Tensor(lambda c, x, wv: _to_cl(_conv_same_cf(_to_cf(x), wv, stride)), (inputs, w),
                             [s[0]] + [None if n is None else -(-n // stride) for n in s[1:-1]] + [cout])
            raw = out
            if normalizer_fn is not None:
                out = normalizer_fn(out, **(normalizer_params or {}))
            elif b is not None:
                out = out + b
            out = activation_fn(out) if activation_fn is not None else out
            _G.layers[_scoped("")] = dict(raw=raw, out=out)      # TF: any tensor can be fetched by name
            return out
    return layer


def gradients(ys, xs):
    """tf.gradients(ys, xs) for a scalar ``ys``: d ys / d x for every tensor x of the graph (variables or
    intermediates; the values autograd differentiates are the very ones the forward evaluation memoised)."""
    xs = list(xs)

    def grad_of(i):
        def fn(ctx):
            key = ("gradients", ys, tuple(xs))
            if key not in ctx.grads:
                xv = [x._eval(ctx) for x in xs]
                yv = ys._eval(ctx)
                gs = torch.autograd.grad(yv, xv, allow_unused=True, retain_graph=True)
                ctx.grads[key] = [torch.zeros_like(x) if g is None else g.detach() for g, x in zip(gs, xv)]
            return ctx.grads[key][i]
        return Tensor(fn, (), xs[i].shape.as_list())
    return [grad_of(i) for i in range(len(xs))]


def layer_tensors():
    """{variable scope of a slim layer: {'raw': conv output, 'out': layer output}} of the current graph"""
    return dict(_G.layers)


# ------------------------------------------------------------------------------------------ optimizer
class AdamOptimizer(object):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **unused):
        self._lr, self._beta1, self._beta2, self._epsilon = learning_rate, beta1, beta2, epsilon
        self._slots, self._t = {}, 0

    def compute_gradients(self, loss, var_list=None):
        tvs = var_list or trainable_variables()

        def grad_of(i):
            def fn(ctx):
                if loss not in ctx.grads:
                    lv = loss._eval(ctx)
                    leaves = [ctx.leaf(v) for v in tvs]
                    gs = torch.autograd.grad(lv, leaves, allow_unused=True, retain_graph=True)
                    ctx.grads[loss] = [torch.zeros_like(l) if g is None else g.detach() for g, l in zip(gs, leaves)]
                return ctx.grads[loss][i]
            return Tensor(fn, (), tvs[i].shape.as_list())
        return [(grad_of(i), v) for i, v in enumerate(tvs)]

    def get_slot(self, var, name):
        return self._slots[var][name]

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        gv = list(grads_and_vars)

        def fn(ctx):
            gvals = [g._eval(ctx).detach() for g, _ in gv]

            def apply():
                self._t += 1
                b1, b2 = self._beta1, self._beta2
                lr_t = self._lr * math.sqrt(1.0 - b2 ** self._t) / (1.0 - b1 ** self._t)
                for g, (_, var) in zip(gvals, gv):
                    sl = self._slots.setdefault(var, dict(m=torch.zeros_like(var.value), v=torch.zeros_like(var.value)))
                    sl["m"] = sl["m"] + (g - sl["m"]) * (1.0 - b1)
                    sl["v"] = sl["v"] + (g * g - sl["v"]) * (1.0 - b2)
                    var.value = var.value - lr_t * sl["m"] / (torch.sqrt(sl["v"]) + self._epsilon)
            ctx.deferred.append(apply)
            return None
        return Tensor(fn, (), ())


# ------------------------------------------------------------------------------------------ summaries, session
def _summary_scalar(name, tensor):
    _G.summaries.append((name, tensor))


def _summary_image(name, tensor, max_outputs=3):
    _G.summaries.append((name, None))


def _merge_all():
    items = [(n, t) for n, t in _G.summaries if t is not None]

    def fn(ctx):
        return {n: float(t._eval(ctx).detach()) for n, t in items}
    return Tensor(fn, (), ())


def _grayscale_to_rgb(t, name=None):
    return Tensor(lambda c, v: v.expand(*v.shape[:-1], 3), (t,), t.shape.as_list()[:-1] + [3])


class FileWriter(object):
    """tf.summary.FileWriter: keeps what it is given (``records`` of this writer and, by directory, in
    ``FileWriter.by_dir``) instead of writing event files."""
    by_dir = {}

    def __init__(self, logdir, graph=None, **k):
        self.logdir, self.records, self.graphs = logdir, [], 0
        FileWriter.by_dir[logdir] = self

    def add_graph(self, graph):
        self.graphs += 1

    def add_summary(self, summary, global_step=None):
        self.records.append((None if global_step is None else int(global_step), summary))

    def flush(self):
        pass

    close = flush


class Saver(object):
    """tf.train.Saver: ``save`` writes ``<path>-<global_step>.npz`` (variable values keyed by ``var.op.name``)
    and returns ``<path>-<global_step>`` like TF; retention follows TF's rule (the newest ``max_to_keep``
    files; one that drops out of that window survives only if ``keep_checkpoint_every_n_hours`` have passed
    since the last one kept for good)."""

    def __init__(self, var_list=None, max_to_keep=5, keep_checkpoint_every_n_hours=10000.0, **k):
        import time
        self._vars, self._max, self._nhour = var_list, max_to_keep, keep_checkpoint_every_n_hours
        self._last, self._next_keep = [], time.time() + keep_checkpoint_every_n_hours * 3600.0

    def save(self, sess, save_path, global_step=None):
        import os
        import time
        prefix = save_path if global_step is None else "%s-%d" % (save_path, int(global_step))
        vs = self._vars if self._vars is not None else global_variables()
        np.savez(prefix + ".npz", **{v.op.name: v.value.detach().numpy() for v in vs})
        self._last.append((prefix, time.time()))
        while self._max and len(self._last) > self._max:
            old, t = self._last.pop(0)
            if t > self._next_keep:
                self._next_keep += self._nhour * 3600.0
            elif os.path.exists(old + ".npz"):
                os.remove(old + ".npz")
        return prefix

    def restore(self, sess, save_path):
        z = np.load(save_path if save_path.endswith(".npz") else save_path + ".npz")
        for v in (self._vars if self._vars is not None else global_variables()):
            v.load(z[v.op.name], sess)


def _max_bytes_in_use():
    return Tensor(lambda c: 0.0, (), ())


class Session(object):
    graph = "default-graph"

    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass

    def run(self, fetches, feed_dict=None):
        ctx = _Ctx(dict(feed_dict or {}))

        def ev(f):
            if isinstance(f, (list, tuple)):
                return [ev(x) for x in f]
            v = f._eval(ctx)
            return v.detach().numpy() if isinstance(v, torch.Tensor) else v
        # float32: torch's oneDNN path aborts in the double backward of the conv-transpose definition above
        was = torch.backends.mkldnn.enabled
        torch.backends.mkldnn.enabled = False
        try:
            out = ev(fetches)
        finally:
            torch.backends.mkldnn.enabled = was
        for f in ctx.deferred:          # assigns take effect after every fetch has read its inputs
            f()
        return out


InteractiveSession = Session


# ------------------------------------------------------------------------------------------ installation
def install(reference_lib=REFERENCE_LIB):
    """Registers this module as ``tensorflow`` (+ the sub-modules the reference imports), adds ``xrange``
    to builtins and puts the reference's ``lib/`` directory on sys.path."""
    me = sys.modules[__name__]
    tf = types.ModuleType("tensorflow")
    for k in ("float32", "int64", "int32", "variable_scope", "name_scope", "placeholder", "reshape", "cast", "argmax",
              "equal", "reduce_mean", "reduce_sum", "where", "to_float", "gather_nd", "multiply", "zeros_like", "concat",
              "Variable", "trainable_variables", "global_variables", "global_variables_initializer", "Session",
              "InteractiveSession", "reset_default_graph", "set_random_seed", "nn", "gradients"):
        setattr(tf, k, getattr(me, k))
    tf.train = types.SimpleNamespace(AdamOptimizer=AdamOptimizer, Saver=Saver)
    tf.summary = types.SimpleNamespace(scalar=_summary_scalar, image=_summary_image, merge_all=_merge_all,
                                       FileWriter=FileWriter)
    tf.GraphKeys, tf.get_collection = GraphKeys, get_collection
    tf.image = types.SimpleNamespace(grayscale_to_rgb=_grayscale_to_rgb)
    tf.__version__ = "1.x-shim"
    slim = types.ModuleType("tensorflow.contrib.slim")
    slim.conv2d, slim.conv3d = _conv_layer(2, False), _conv_layer(3, False)
    slim.conv2d_transpose, slim.conv3d_transpose = _conv_layer(2, True), _conv_layer(3, True)
    slim.batch_norm = batch_norm
    layers = types.ModuleType("tensorflow.contrib.layers")
    layers.batch_norm = batch_norm
    contrib = types.ModuleType("tensorflow.contrib")
    contrib.slim, contrib.layers = slim, layers
    contrib.memory_stats = types.SimpleNamespace(MaxBytesInUse=_max_bytes_in_use)
    python = types.ModuleType("tensorflow.python")
    platform = types.ModuleType("tensorflow.python.platform")
    python.platform = platform
    tf.contrib, tf.python = contrib, python
    sys.modules.update({"tensorflow": tf, "tensorflow.contrib": contrib, "tensorflow.contrib.slim": slim,
                        "tensorflow.contrib.layers": layers, "tensorflow.python": python,
                        "tensorflow.python.platform": platform})
    if not hasattr(builtins, "xrange"):
        builtins.xrange = range
    if reference_lib not in sys.path:
        sys.path.insert(0, reference_lib)
    return tf


def uninstall(reference_lib=REFERENCE_LIB):
    if getattr(sys.modules.get("tensorflow"), "__version__", None) == "1.x-shim":
        for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.")]:
            del sys.modules[k]
    for k in ("uresnet", "ssnet", "resnet_module", "ssnet_trainval", "config"):
        m = sys.modules.get(k)
        if m is not None and (getattr(m, "__file__", None) or "").startswith(reference_lib):
            del sys.modules[k]
    if getattr(sys.modules.get("larcv"), "__version__", None) == "shim":
        del sys.modules["larcv"], sys.modules["larcv.dataloader2"]
    if reference_lib in sys.path:
        sys.path.remove(reference_lib)
    if getattr(builtins, "xrange", None) is range:
        del builtins.xrange


def install_larcv(make_threadio):
    """Registers a stand-in ``larcv`` package whose ``larcv.dataloader2.larcv_threadio`` is ``make_threadio``
    (a class with the larcv_threadio call surface, ssnet_trainval.py:64-91), so that the reference's DRIVER
    (lib/ssnet_trainval.py) imports.  The stand-in ``larcv.larcv`` has no IOManager: ANA_OUTPUT_CONFIG must
    stay empty."""
    pkg = types.ModuleType("larcv")
    pkg.__version__ = "shim"
    pkg.larcv = types.SimpleNamespace()
    dl = types.ModuleType("larcv.dataloader2")
    dl.larcv_threadio = make_threadio
    pkg.dataloader2 = dl
    sys.modules.update({"larcv": pkg, "larcv.dataloader2": dl})


def load_reference_driver(make_threadio, reference_lib=REFERENCE_LIB):
    """Imports the UNMODIFIED reference driver (lib/ssnet_trainval.py, with lib/config.py and the model code)
    over the shim and the stand-in larcv; returns its ``ssnet_trainval`` class."""
    install(reference_lib)
    install_larcv(make_threadio)
    import importlib
    for k in ("uresnet", "ssnet", "resnet_module", "ssnet_trainval", "config"):
        sys.modules.pop(k, None)
    mod = importlib.import_module("ssnet_trainval")
    assert mod.__file__.startswith(reference_lib), mod.__file__
    return mod.ssnet_trainval


def load_reference(reference_lib=REFERENCE_LIB):
    """Imports the UNMODIFIED reference model code over the shim; returns its ``uresnet`` class."""
    install(reference_lib)
    import importlib
    for k in ("uresnet", "ssnet", "resnet_module"):
        sys.modules.pop(k, None)
    mod = importlib.import_module("uresnet")
    assert mod.__file__.startswith(reference_lib), mod.__file__
    return mod.uresnet
