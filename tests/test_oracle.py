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
