"""Holds the oracle -- and the library's plan -- to the one artefact the reference itself produced: the
TensorBoard GraphDef dump under ``/root/reference/libs/uresnet_graph/`` (an OLDER revision of the model,
written by TensorFlow 1.x).  ``tests/golden/reference_graph.json`` is the committed extract of it
(``tests/golden/make_reference_graph_fixture.py``); where the reference tree is present the extract is
re-derived from the dump and compared.

What the dump can pin, and does here:
  * filter variable names / shapes / Xavier limits (TensorFlow's own float32 constants) for the 56 of 58
    filters whose kernel size the two revisions share, and the initializer formula for all 58;
  * the wiring of all 58 conv-type ops (which tensor feeds which, strides, SAME, NHWC), the skip map and
    the concat operand order, the residual adds;
  * slim's lowering of conv2d_transpose (Conv2DBackpropInput of a SAME stride-2 conv, 2 x output, filter
    [k,k,Cout,Cin]) -- evaluated against the oracle's conv_transpose_same;
  * the expansion of batch_norm (no gamma, batch moments over N,H,W, biased variance,
    x * rsqrt(var + eps) + (beta - mean * rsqrt(var + eps))) -- the dumped expression is EXECUTED on
    random data and compared with oracle.batch_norm;
  * the accuracy expressions (argmax over the channel axis, where / gather_nd on the thresholded image)
    -- executed and compared with oracle.loss_and_metrics.
What it cannot pin (the revisions differ): 7x7 conv0 / conv1, conv biases, the
pre-activation placement of BN + ReLU inside the residual branch, eps = 1e-5, RMSProp, the loss
reduction, the nonzero threshold.  No tensor VALUES of the reference exist anywhere: numerics stay
unpinned (oracle/__init__.py).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from oracle import uresnet_oracle as O
from oracle.steps import StepOracle

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "reference_graph.json")
DUMP = "/root/reference/libs/uresnet_graph/events.out.tfevents.1509651705.lee"

# the filters whose kernel size changed between the dumped revision (7x7) and the current code
# (uresnet.py:37-45,103-111: kernel_size = 3)
RESIZED = {"UResNet/conv0/weights", "UResNet/conv1/weights"}


@pytest.fixture(scope="module")
def fx():
    with open(FIXTURE) as f:
        return json.load(f)


@pytest.mark.skipif(not os.path.exists(DUMP), reason="reference tree not on this box")
def test_fixture_is_what_the_dump_says(fx):
    import importlib.util
    pytest.importorskip("tensorboard.compat.proto.graph_pb2")          # GraphDef protos only; no TensorFlow needed
    spec = importlib.util.spec_from_file_location("mk", os.path.join(HERE, "golden", "make_reference_graph_fixture.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    fresh = json.loads(json.dumps(mk.extract(DUMP), sort_keys=True))
    assert fresh == fx
    assert fx["num_nodes"] == 10303 and fx["producer"] == 24


def test_filter_names_shapes_and_xavier_limits(fx):
    names = dict(O.param_names(2, base=16, num_strides=5, num_class=3))
    ours = {k: v for k, v in names.items() if k.endswith("/weights")}
    assert set(ours) == set(fx["weights"])                    # the same 58 TF variable names
    # creation order of the filters is the order of the dump's VariableV2 nodes
    assert [k for k in names if k.endswith("/weights")] == sorted(
        fx["weights"], key=lambda k: list(names).index(k))
    for k, w in fx["weights"].items():
        # TensorFlow's constant reproduces from the dump's own shape: pins the initializer formula
        assert np.float32(O.xavier_limit(w["shape"])) == pytest.approx(w["uniform_limit"], rel=2e-7), k
        if k in RESIZED:
            assert list(ours[k][2:]) == w["shape"][2:] and w["shape"][:2] == [7, 7] and ours[k][:2] == (3, 3)
        else:
            assert list(ours[k]) == w["shape"], k              # incl. the [k,k,Cout,Cin] conv-transpose layout
            assert np.float32(O.xavier_limit(ours[k])) == pytest.approx(w["uniform_limit"], rel=2e-7)
    # init_params draws inside those limits and close to them
    P = O.init_params(2, base=16, num_strides=5, seed=3)
    for k, w in fx["weights"].items():
        if k not in RESIZED:
            m = np.abs(P[k]).max()
            assert m <= np.float32(w["uniform_limit"]) * (1 + 1e-6) and m > 0.9 * w["uniform_limit"], k


@pytest.fixture(scope="module")
def steps():
    rng = np.random.default_rng(0)
    P = O.init_params(2, base=2, num_strides=5, seed=1)
    return StepOracle(P, rng.random((1, 32, 32, 1)), rng.integers(0, 3, (1, 32, 32)).astype(np.float64),
                      None, base=2, num_strides=5)


def test_wiring_of_every_conv_op(fx, steps):
    assert set(steps.layers) == set(fx["convs"])
    for scope, L in steps.layers.items():
        g = fx["convs"][scope]
        assert g["source"] == [L["x"]], (scope, g["source"], L["x"])
        assert g["strides"] == [1, L["stride"], L["stride"], 1], scope
        assert g["padding"] == "SAME" and g["data_format"] == "NHWC"
        assert (g["op"] == "Conv2DBackpropInput") == bool(L["transpose"]), scope
    table = {l["scope"]: l for l in O.layer_table(2, base=16, num_strides=5)}
    for scope, g in fx["convs"].items():
        assert (g["op"] == "Conv2DBackpropInput") == (table[scope]["op"] == "convT")
        if "output_shape" in g:                                # slim.conv2d_transpose: 2 x input, Cout channels
            assert g["output_shape"] == {"spatial_multiplier": [2, 2], "channels": table[scope]["cout"],
                                         "batch_from_input": True}


def test_skip_map_concat_order_and_residual_adds(fx, steps):
    cats = {r[1]: [r[2], r[3]] for r in steps.recs if r[0] == "concat"}
    assert cats == {k: v["inputs"] for k, v in fx["concats"].items()}        # deconv channels first
    assert all(v["axis"] == 3 for v in fx["concats"].values())               # channel axis of NHWC
    mods = {r[1]: [r[2] if r[2] else r[3], r[1] + "/resnet_conv2"] for r in steps.recs if r[0] == "module"}
    assert mods == {k: v["add_inputs"] for k, v in fx["modules"].items()}    # relu(shortcut + residual)
    assert all(v["relu_of"] == "add" for v in fx["modules"].values())


@pytest.mark.parametrize("hw", [(4, 6), (5, 3)])
def test_conv_transpose_is_conv2d_backprop_input(fx, hw):
    """The dump lowers slim.conv2d_transpose(k=3, s=2, 'same') to
    Conv2DBackpropInput(output_shape=[B,2H,2W,Cout], filter=[3,3,Cout,Cin], out_backprop=x, strides 2, SAME):
    by definition of that op, the gradient with respect to z of Conv2D(z [B,2H,2W,Cout], filter, 2, SAME)
    contracted with x.  The oracle's conv_transpose_same must be exactly that function of (x, filter)."""
    g = fx["convs"]["deconv0"]
    assert g["op"] == "Conv2DBackpropInput" and g["strides"] == [1, 2, 2, 1] and g["padding"] == "SAME"
    rng = np.random.default_rng(5)
    cin, cout = 3, 2
    x = torch.tensor(rng.standard_normal((2, cin) + hw))
    w = torch.tensor(rng.standard_normal((3, 3, cout, cin)))                # [k,k,Cout,Cin] as in the dump
    z = torch.zeros((2, cout, 2 * hw[0], 2 * hw[1]), dtype=torch.float64, requires_grad=True)
    y = O.conv_same(z, w, 2)                      # the filter read as [k,k,in=Cout,out=Cin]
    assert y.shape == x.shape
    (dz,) = torch.autograd.grad(y, [z], x)
    assert torch.allclose(O.conv_transpose_same(x, w, 2), dz, atol=1e-12, rtol=0)


class _Eval(object):
    """Evaluates a dumped sub-graph {name: [op, inputs]} with numpy."""

    def __init__(self, form, feeds, const_override=()):
        self.form, self.val, self.ov = form, dict(feeds), dict(const_override)

    def __call__(self, name):
        if name in self.val:
            return self.val[name]
        op, arg = self.form[name]
        if op == "Const":
            v = np.asarray(self.ov.get(name, arg), dtype=np.float64)
            out = v if v.size > 1 or name.endswith("reduction_indices") or name.startswith("metrics/Const") else float(v.reshape(-1)[0])
        else:
            a = [self(i) for i in arg]
            axes = lambda v: tuple(int(i) for i in np.atleast_1d(v))
            out = {
                "Mean": lambda: a[0].mean(axis=axes(a[1]), keepdims=name.startswith("moments/")),
                "StopGradient": lambda: a[0], "Squeeze": lambda: np.squeeze(a[0]), "Cast": lambda: a[0].astype(np.float64),
                "SquaredDifference": lambda: (a[0] - a[1]) ** 2, "Add": lambda: a[0] + a[1], "Sub": lambda: a[0] - a[1],
                "Mul": lambda: a[0] * a[1], "Rsqrt": lambda: 1.0 / np.sqrt(a[0]), "Relu": lambda: np.maximum(a[0], 0),
                "ArgMax": lambda: np.argmax(a[0], axis=int(a[1])), "Equal": lambda: a[0] == a[1],
                "Greater": lambda: a[0] > a[1], "Where": lambda: np.argwhere(a[0]),
                "GatherNd": lambda: a[0][tuple(a[1].T)],
            }[op]()
        self.val[name] = out
        return out


def test_batch_norm_expansion_executed(fx):
    bn = fx["batch_norm"]
    assert bn["variables"] == ["beta", "moving_mean", "moving_variance"]      # no gamma (scale=False)
    assert bn["num_scopes"] == 40 and bn["moving_average_consumers"] == []    # nothing runs the updates
    form = bn["form"]
    assert form["moments/mean/reduction_indices"][1] == [0.0, 1.0, 2.0]       # N, H, W of NHWC
    assert form["batchnorm/add/y"][1][0] == pytest.approx(1e-5)              # the old revision passed epsilon=1e-5;
    assert O.BN_EPS == 1e-3                                                   # resnet_module.py:25-66 now takes slim's default
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3, 5, 4, 6)) * 3 + 1.5                           # NHWC
    beta = rng.standard_normal(6)
    ev = _Eval(form, {"<x>": x, "beta/read": beta}, {"batchnorm/add/y": [O.BN_EPS]})
    want = ev("batchnorm/add_1")                                             # the dumped expression
    got = O.batch_norm(torch.tensor(x).permute(0, 3, 1, 2), torch.tensor(beta)).permute(0, 2, 3, 1).numpy()
    assert np.abs(got - want).max() < 1e-12
    # the normalisation consumes the BATCH moments (is_training=True), not the moving averages
    assert form["batchnorm/mul_1"][1][0] == "moments/Squeeze" and form["batchnorm/add"][1][0] == "moments/Squeeze_1"
    # moving averages: decay 0.999 (slim default), update form moving -= (moving - batch) * 0.001
    assert form["AssignMovingAvg/decay"][1][0] == pytest.approx(1e-3)


def test_accuracy_expressions_executed(fx):
    """ssnet.py:57-63 today and the dump agree on everything but the threshold (10.0 then, 0 now)."""
    M = fx["metrics"]
    assert M["metrics/Greater/y"][1] == [10.0]
    rng = np.random.default_rng(3)
    B, H, W, Cc = 2, 6, 5, 3
    logits = rng.standard_normal((B, H, W, Cc))
    logits[0, 0, 0, :] = 0.25                                                  # a tie: first index wins
    label = rng.integers(0, 3, (B, H, W)).astype(np.float64)
    data = rng.random((B, H, W, 1)) * (rng.random((B, H, W, 1)) < 0.3)
    ev = _Eval(M, {"UResNet/conv2/Relu": logits, "input_prep/Cast": label.astype(np.int64),
                   "metrics/Reshape": data[..., 0]}, {"metrics/Greater/y": [0.0]})
    acc_all, acc_nz = ev("metrics/Mean"), ev("metrics/Mean_1")
    res = O.loss_and_metrics(torch.tensor(logits), torch.tensor(data), torch.tensor(label), None)
    assert float(res["acc_all"]) == pytest.approx(float(acc_all), abs=1e-7)       # tf.float32 means
    assert float(res["acc_nonzero"]) == pytest.approx(float(acc_nz), abs=1e-7)
    assert int(res["n_nz"]) == int((data > 0).sum()) > 0
    assert fx["placeholder_reshapes"]["input_prep/image_reshape"] == [-1, 512, 512, 1]


def test_library_plan_has_the_dumped_filters(fx, lib):
    """The C-ABI plan (host code, nothing launched) names and shapes its filters as the dump does."""
    from uresnet_b200 import _capi
    cfg = _capi.urn_net_cfg()
    cfg.nd = 2
    cfg.spatial[0], cfg.spatial[1], cfg.spatial[2] = 512, 512, 1
    cfg.in_channels, cfg.batch, cfg.base_filters, cfg.num_strides = 1, 1, 16, 5
    cfg.num_class, cfg.dtype, cfg.impl, cfg.training = 3, 0, 0, 1
    h = C.c_void_p()
    _capi.check(lib.urn_create(C.byref(cfg), C.byref(h)))
    seen = {}
    for i in range(lib.urn_param_count(h)):
        name, rank = C.c_char_p(), C.c_int()
        shp = (C.c_int64 * 6)()
        off, cnt = C.c_int64(), C.c_int64()
        _capi.check(lib.urn_param_info(h, i, C.byref(name), C.byref(rank), shp, C.byref(off), C.byref(cnt)))
        seen[name.value.decode()] = [int(shp[k]) for k in range(rank.value)]
    lib.urn_destroy(h)
    filt = {k: v for k, v in seen.items() if k.endswith("/weights")}
    assert set(filt) == set(fx["weights"])
    for k, w in fx["weights"].items():
        assert filt[k] == (([3, 3] + w["shape"][2:]) if k in RESIZED else w["shape"]), k
