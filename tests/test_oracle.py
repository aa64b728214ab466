"""CPU tests of the oracle itself: independent naive cross-check, finite differences,
and reproduction of the committed golden vectors (tests/golden/*.npz)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import naive as N
from oracle import uresnet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cf(x, nd):
    return torch.tensor(x).permute(0, nd + 1, *range(1, nd + 1))


def _cl(y, nd):
    return y.permute(0, *range(2, nd + 2), 1).numpy()


@pytest.mark.parametrize("nd,sp", [(2, (6, 8)), (2, (5, 7)), (3, (4, 6, 4))])
@pytest.mark.parametrize("k,s", [(3, 1), (3, 2), (1, 2), (1, 1)])
def test_conv_same_vs_naive(nd, sp, k, s):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2,) + sp + (3,))
    w = rng.standard_normal((k,) * nd + (3, 5))
    a = _cl(O.conv_same(_cf(x, nd), torch.tensor(w), s), nd)
    np.testing.assert_allclose(a, N.conv_same(x, w, s), atol=1e-12)


@pytest.mark.parametrize("nd,sp", [(2, (6, 8)), (3, (4, 3, 5))])
def test_conv_transpose_vs_naive(nd, sp):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2,) + sp + (3,))
    w = rng.standard_normal((3,) * nd + (5, 3))
    a = _cl(O.conv_transpose_same(_cf(x, nd), torch.tensor(w), 2), nd)
    assert a.shape[1:1 + nd] == tuple(2 * n for n in sp)
    np.testing.assert_allclose(a, N.conv_transpose_same(x, w, 2), atol=1e-12)


def test_conv_transpose_is_adjoint_of_strided_conv():
    """TF defines conv_transpose as the input-gradient of the SAME stride-2 conv."""
    rng = np.random.default_rng(2)
    x = torch.tensor(rng.standard_normal((1, 3, 8, 8)), requires_grad=True)   # big side
    w = torch.tensor(rng.standard_normal((3, 3, 3, 5)))                      # [k,k,Cbig,Csmall]
    y = O.conv_same(x, w, 2)
    g = torch.tensor(rng.standard_normal(tuple(y.shape)))
    (dx,) = torch.autograd.grad(y, x, g)
    ct = O.conv_transpose_same(g, w, 2)       # deconv input has Csmall, filter [k,k,Cout=Cbig,Cin=Csmall]
    np.testing.assert_allclose(dx.numpy(), ct.numpy(), atol=1e-12)


def test_bn_loss_acc_vs_naive():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 5, 6, 4))
    beta = rng.standard_normal(4)
    a = _cl(O.batch_norm(_cf(x, 2), torch.tensor(beta)), 2)
    np.testing.assert_allclose(a, N.batch_norm(x, beta), atol=1e-12)
    logits = rng.standard_normal((2, 5, 6, 3))
    label = rng.integers(0, 3, (2, 5, 6)).astype(np.float64)
    weight = rng.random((2, 5, 6))
    data = rng.standard_normal((2, 5, 6, 1)) * (rng.random((2, 5, 6, 1)) > 0.5)
    r = O.loss_and_metrics(torch.tensor(logits), torch.tensor(data), torch.tensor(label), torch.tensor(weight))
    assert abs(float(r["loss"]) - N.ce_loss(logits, label, weight)) < 1e-12
    aa, anz = N.accuracies(logits, label, data)
    assert abs(r["acc_all"] - aa) < 1e-7 and abs(r["acc_nonzero"] - anz) < 1e-12


def test_label_out_of_range_raises():
    with pytest.raises(ValueError):
        O.loss_and_metrics(torch.zeros(1, 2, 2, 3), torch.ones(1, 2, 2, 1), torch.full((1, 2, 2), 3.0))


def test_layer_table_counts():
    """53 convs + 5 deconvs, 116 trainable tensors, parameter totals."""
    for nd, total in ((2, 16858979), (3, 49862627)):
        L = O.layer_table(nd)
        assert len(L) == 58 and sum(l["op"] == "convT" for l in L) == 5
        assert sum(int(np.prod(s)) for _, s in O.param_names(nd)) == total
    names = [n for n, _ in O.param_names(2)]
    assert names[0] == "UResNet/conv0/weights" and names[1] == "UResNet/conv0/BatchNorm/beta"
    assert "UResNet/resnet_module5/module1/shortcut/weights" in names
    assert "UResNet/resnet_module0/module2/shortcut/weights" not in names
    assert dict(O.param_names(2))["UResNet/deconv0/weights"] == (3, 3, 256, 512)


@pytest.mark.parametrize("nd,dims", [(2, (8, 8)), (3, (4, 4, 4))])
def test_finite_difference_gradients(nd, dims):
    """fp64 central differences on every parameter tensor of a 2-level mini U-ResNet."""
    rng = np.random.default_rng(5)
    B, F, S = 2, 2, 2
    P = O.init_params(nd, base=F, num_strides=S, seed=7)
    data = (rng.random((B,) + dims + (1,)) > 0.6) * rng.random((B,) + dims + (1,)) * 5
    label = rng.integers(0, 3, (B,) + dims).astype(np.float64)
    weight = rng.random((B,) + dims)
    r = O.run(P, data, label, weight, base=F, num_strides=S)
    eps = 1e-5
    for name, g in r["grads"].items():
        flat = P[name].astype(np.float64).reshape(-1)
        for idx in rng.choice(flat.size, size=min(3, flat.size), replace=False):
            vals = []
            for sgn in (+1, -1):
                Q = {k: v.astype(np.float64) for k, v in P.items()}
                q = Q[name].reshape(-1).copy()
                q[idx] += sgn * eps
                Q[name] = q.reshape(P[name].shape)
                vals.append(O.run(Q, data, label, weight, base=F, num_strides=S, want_grads=False)["loss"])
            fd = (vals[0] - vals[1]) / (2 * eps)
            assert abs(fd - g.reshape(-1)[idx]) <= 1e-5 * max(1.0, abs(fd)), (name, idx, fd, g.reshape(-1)[idx])


def test_dlogits_formula():
    """dlogits = (softmax - onehot) * w / B."""
    rng = np.random.default_rng(6)
    P = O.init_params(2, base=2, num_strides=1, seed=3)
    data = rng.random((2, 4, 4, 1))
    label = rng.integers(0, 3, (2, 4, 4)).astype(np.float64)
    weight = rng.random((2, 4, 4))
    r = O.run(P, data, label, weight, base=2, num_strides=1)
    oh = np.eye(3)[label.astype(int)]
    np.testing.assert_allclose(r["dlogits"], (r["softmax"] - oh) * weight[..., None] / 2, atol=1e-12)


def test_adam_tf_formula():
    var, g = np.array([1.0, -2.0], np.float32), np.array([0.5, -0.25], np.float32)
    m, v = np.zeros(2, np.float32), np.zeros(2, np.float32)
    O.adam_tf(var, g, m, v, t=1, lr=1e-3)
    # first step: m = .1 g, v = .001 g^2, lr_t = lr*sqrt(.001)/.1 -> update ~ lr*sign(g)
    np.testing.assert_allclose(var, [1.0 - 1e-3, -2.0 + 1e-3], rtol=0, atol=2e-7)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*.npz"))))
def test_oracle_reproduces_golden(path):
    z = np.load(path)
    dims = [int(v) for v in z["dims"]]
    B, F, S, C = int(z["batch"]), int(z["base"]), int(z["strides"]), int(z["num_class"])
    nd = len(dims) - 1
    names = [str(n) for n in z["grad_names"]]
    if "p000" in z:
        P = {k: z["p%03d" % i] for i, k in enumerate(names)}
    else:
        P = O.init_params(nd, base=F, num_strides=S, num_class=C, seed=int(z["param_seed"]))
    shp = [B] + dims[:-1]
    w = z["weight"].reshape(shp) if bool(z["use_weight"]) else None
    r = O.run(P, z["data"].reshape(shp + [1]), z["label"].astype(np.float32).reshape(shp), w,
              base=F, num_strides=S)
    np.testing.assert_allclose(r["logits"], z["logits"], rtol=0, atol=2e-5)
    assert abs(r["loss"] - float(z["loss"])) <= 1e-9 * abs(float(z["loss"]))
    assert r["n_correct_all"] == int(z["n_correct_all"]) and r["n_nz"] == int(z["n_nz"])
    l2 = np.array([np.sqrt((r["grads"][k] ** 2).sum()) for k in names])
    np.testing.assert_allclose(l2, z["grad_l2"], rtol=1e-9)


# --------------------------------------------------------------------------------------------
# step-by-step oracle (oracle/steps.py) and gradient conditioning (oracle/conditioning.py)
# --------------------------------------------------------------------------------------------
def _case(dims, B, F, St, seed_d=7, seed_p=11):
    from uresnet_b200 import synthetic_io as S
    nd = len(dims) - 1
    data, label, weight = S.make_batch(B, dims, seed=seed_d)
    weight /= weight.sum(axis=1, keepdims=True)
    P = O.init_params(nd, base=F, num_strides=St, seed=seed_p)
    shp = [B] + dims[:-1]
    return P, data.reshape(shp + [1]), label.reshape(shp), weight.reshape(shp)


@pytest.mark.parametrize("dims,B,F,St", [([32, 32, 1], 2, 4, 3), ([8, 16, 8, 1], 1, 4, 2), ([32, 32, 1], 1, 4, 5)])
def test_step_oracle_equals_whole_graph_autograd(dims, B, F, St):
    """The layer-by-layer statement of the graph with per-step local backward (what the
    teacher-forced GPU parity tests compare against) reproduces the whole-graph autograd oracle:
    logits, loss, metrics and all gradients to fp64 round-off."""
    from oracle.steps import StepOracle
    P, data, label, weight = _case(dims, B, F, St)
    ref = O.run(P, data, label, weight, base=F, num_strides=St)
    so = StepOracle(P, data, label, weight, base=F, num_strides=St)
    np.testing.assert_allclose(so.logits.numpy(), ref["logits"], rtol=0, atol=1e-12)
    g = so.backward()
    assert abs(so.result["loss"] - ref["loss"]) < 1e-13 and so.result["n_correct_nz"] == ref["n_correct_nz"]
    assert list(g.keys()) == list(ref["grads"].keys())
    for k in g:
        scale = max(np.abs(ref["grads"][k]).max(), 1e-300)
        assert np.abs(g[k] - ref["grads"][k]).max() <= 1e-10 * scale, k
    # every schedule quantity has the layout the CUDA workspace uses: [rows, C], channel-last
    rows = B * int(np.prod(dims[:-1]))
    assert so.node("conv0").shape == (rows, F) and so.raw("conv2").shape == (rows, 3)
    assert so.node("concat%d" % (St - 1)).shape == (rows, 2 * F)
    assert so.stats("conv1").shape == (2 * F,) and so.draw("conv1").shape == (rows, F)
    m = "resnet_module%d/module2" % (St + 4)            # last decoder module: identity shortcut
    assert so.masked_gradient(m).shape == (rows, F)
    # concat gradient = [deconv | skip] contributions of the decoder module that reads it
    cat = "concat%d" % (St - 1)
    dec = "resnet_module%d/module1" % (St + 4)
    tot = so.input_gradient(dec + "/resnet_conv1") + so.input_gradient(dec + "/shortcut")
    np.testing.assert_allclose(so.node_gradient(cat), tot, rtol=0, atol=1e-12 * np.abs(tot).max())


def test_gradient_conditioning():
    """The measurement behind the 16-bit gradient tolerances: perturbing the
    fp64 oracle's OWN parameters by half an ulp of a 10-bit-mantissa format (2^-11 relative: fp16 /
    TF32 operand rounding, applied to the weights only) leaves logits within 2e-2 but moves most of
    its 116 gradient tensors by several per cent -- far beyond 2e-2, let alone 1e-3 -- while a
    2^-24 perturbation (fp32 rounding) moves them by ~1e-7.  The response is not linear (amplification
    ~3 at 2^-24, ~100 at 2^-11, then ~sqrt): ReLU-mask flips.  So no 10-bit-operand evaluation of
    this graph can track fp64 gradients end to end at 2e-2; fp32 can at 1e-3."""
    from oracle.conditioning import gradient_sensitivity
    P, data, label, weight = _case([64, 64, 1], 2, 8, 4, seed_d=99, seed_p=5)
    ref = O.run(P, data, label, weight, base=8, num_strides=4)
    s24 = gradient_sensitivity(P, data, label, weight, 8, 4, eps=2.0 ** -24, draws=1, ref=ref)
    s11 = gradient_sensitivity(P, data, label, weight, 8, 4, eps=2.0 ** -11, draws=2, ref=ref)
    s10 = gradient_sensitivity(P, data, label, weight, 8, 4, eps=2.0 ** -10, draws=2, ref=ref)
    print("\nconditioning 64x64 F=8 S=4 B=2 (per-tensor max|a-b|/max|b| of the 116->%d gradient tensors)" % len(P))
    for s in (s24, s11, s10):
        print("  eps=2^%d: logits %.2e loss %.2e grads median %.2e p90 %.2e worst %.2e"
              % (int(np.log2(s["eps"])), s["logits"], s["loss"], s["median"], s["p90"], s["worst"]))
    assert s24["worst"] < 1e-5 and s24["logits"] < 1e-5                  # fp32-level perturbation: benign
    assert s11["logits"] < 2e-2 and s11["loss"] < 2e-2                   # forward is well conditioned ...
    assert s11["logits"] > 1e-3 / 4                                      # ... but already at the 1e-3 scale
    assert s11["median"] > 2e-2 and s11["worst"] > 5e-2                  # gradients are not
    assert s11["median"] / s11["eps"] > 10 * (s24["median"] / s24["eps"])   # non-linear amplification
    assert s10["median"] < 1.8 * s11["median"]                           # sub-linear beyond (sqrt-like)


def test_step_oracle_storage_rounded_expectations():
    """storage='mix16': the chain (teacher) values are unchanged; the per-step expectations on
    storage-rounded inputs stay within the rounding they model (fp16 activations 2^-11, bf16
    gradients 2^-9) of the chain for well-conditioned steps."""
    from oracle.steps import StepOracle
    P, data, label, weight = _case([32, 32, 1], 2, 4, 3)
    a = StepOracle(P, data, label, weight, base=4, num_strides=3)
    b = StepOracle(P, data, label, weight, base=4, num_strides=3, storage="mix16")
    ga, gb = a.backward(), b.backward()
    for k in ga:                                       # chain not rounded (only the logits -> fp32 before the loss)
        assert np.abs(ga[k] - gb[k]).max() <= 1e-5 * max(np.abs(ga[k]).max(), 1e-300), k

    def rel(x, y):
        return np.abs(x - y).max() / np.abs(y).max()
    for scope in b.layers:
        assert rel(b.raw(scope, True), b.raw(scope)) < 4e-3, scope
        assert rel(b.draw(scope, True), b.draw(scope)) < 3e-2, scope
        assert rel(b.filter_gradient(scope, True), b.filter_gradient(scope)) < 3e-2, scope
    assert rel(b.node("conv1", True), b.node("conv1")) < 1e-5          # fp32 raw + statistics in, fp64 math
    m = "resnet_module7/module2"
    assert rel(b.masked_gradient(m, True), b.masked_gradient(m)) < 2e-2
    # fp32 storage: the closed-form BN backward on stored tensors IS the autograd backward of batch_norm
    c = StepOracle(P, data, label, weight, base=4, num_strides=3, storage="f32")
    c.backward()
    for scope in c.layers:
        assert rel(c.draw(scope, True), c.draw(scope)) < 1e-5, scope
        assert rel(c.dbeta(scope, True), c.dbeta(scope)) < 1e-4 or np.abs(c.dbeta(scope)).max() < 1e-12, scope
        assert rel(c.node(c.layers[scope]["out"], True), c.node(c.layers[scope]["out"])) < 1e-5 if c.layers[scope]["out"] else True
    # fp16 raw copies: only the named scopes are rounded, by ~2^-11 of the raw scale
    d = StepOracle(P, data, label, weight, base=4, num_strides=3, storage="mix16", raw16={"conv1"})
    e1 = rel(d.node("conv1", True), d.node("conv1"))
    assert 1e-5 < e1 < 5e-3
