"""CPU restatement of the reference U-ResNet graph (TEST INFRASTRUCTURE ONLY).

Checked against the reference's own code executed over ``oracle/tf1_shim.py`` and against its GraphDef
dump; the numerics of TensorFlow's kernels themselves stay unpinned (see ``oracle/__init__.py``).
Follows, line by line:

* topology ............ /root/reference/lib/uresnet.py:22-123
* residual block ...... /root/reference/lib/resnet_module.py:10-87
* reshape / cast ...... /root/reference/lib/ssnet.py:29-40
* metrics ............. /root/reference/lib/ssnet.py:57-63
* loss ................ /root/reference/lib/ssnet.py:66-71
* accumulate + Adam ... /root/reference/lib/ssnet.py:72-79
* slim / TF-1.x op semantics: SURVEY.md section 3.4

Everything is expressed with PyTorch-CPU functional ops (oneDNN underneath,
the same class of CPU kernels TF-MKL would use); autograd supplies the
reference gradients.  Parameters live in a dict keyed by the TF variable
names, in TF layouts (conv ``[k..,Cin,Cout]``, deconv ``[k..,Cout,Cin]``).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # slim.batch_norm default epsilon


# --------------------------------------------------------------------------
# layer table (TF variable creation order)          uresnet.py:32-121
# --------------------------------------------------------------------------
def layer_table(nd, base=16, num_strides=5, num_class=3, in_ch=1):
    """List of dicts {scope, op, k, s, cin, cout} in TF creation order."""
    L = []

    def add(scope, op, k, s, cin, cout):
        L.append(dict(scope=scope, op=op, k=k, s=s, cin=cin, cout=cout))

    def double_resnet(idx, cin, cout, stride):          # resnet_module.py:70-87
        for m, (ci, st) in enumerate(((cin, stride), (cout, 1)), start=1):
            sc = "resnet_module%d/module%d" % (idx, m)
            if not (ci == cout and st == 1):            # resnet_module.py:22-33
                add(sc + "/shortcut", "conv", 1, st, ci, cout)
            add(sc + "/resnet_conv1", "conv", 3, st, ci, cout)
            add(sc + "/resnet_conv2", "conv", 3, 1, cout, cout)

    add("conv0", "conv", 3, 1, in_ch, base)              # uresnet.py:37-45
    c = base
    for step in range(num_strides):                      # uresnet.py:56-64
        double_resnet(step, c, 2 * c, 2)
        c *= 2
    for step in range(num_strides):                      # uresnet.py:66-101
        add("deconv%d" % step, "convT", 3, 2, c, c // 2)
        # scope hard-coded step+5 at uresnet.py:100
        double_resnet(step + 5, c, c // 2, 1)
        c //= 2
    add("conv1", "conv", 3, 1, c, base)                  # uresnet.py:103-111
    add("conv2", "conv", 3, 1, base, num_class)          # uresnet.py:113-121
    for l in L:
        kk = (l["k"],) * nd
        if l["op"] == "conv":
            l["wshape"] = kk + (l["cin"], l["cout"])
        else:                                            # deconv: [k..,Cout,Cin]
            l["wshape"] = kk + (l["cout"], l["cin"])
    return L


def param_names(nd, base=16, num_strides=5, num_class=3, in_ch=1):
    """tf.trainable_variables() order: weights then beta, per layer."""
    out = []
    for l in layer_table(nd, base, num_strides, num_class, in_ch):
        out.append(("UResNet/%s/weights" % l["scope"], l["wshape"]))
        out.append(("UResNet/%s/BatchNorm/beta" % l["scope"], (l["cout"],)))
    return out


def xavier_limit(wshape):
    """Limit of slim's default ``weights_initializer`` (``xavier_initializer()``, uniform):
    sqrt(6 / (fan_in + fan_out)) with fan = receptive field x channels, ``wshape`` in the TF layout
    (``[k..,Cin,Cout]``; the conv-transpose layout ``[k..,Cout,Cin]`` gives the same sum).  Pinned
    by the float32 constants TensorFlow itself wrote into the reference's graph dump
    (tests/test_reference_graph.py)."""
    rf = int(np.prod(wshape[:-2]))
    return math.sqrt(6.0 / (rf * wshape[-2] + rf * wshape[-1]))


def init_params(nd, base=16, num_strides=5, num_class=3, in_ch=1, seed=4321,
                beta_std=0.1):
    """Xavier-uniform filters (slim default), betas ~ N(0, beta_std).

    SURVEY 8d 'Weights for parity': PCG64(seed); non-zero betas so that the
    beta gradients are exercised.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    P = OrderedDict()
    for l in layer_table(nd, base, num_strides, num_class, in_ch):
        lim = xavier_limit(l["wshape"])
        P["UResNet/%s/weights" % l["scope"]] = rng.uniform(
            -lim, lim, size=l["wshape"]).astype(np.float32)
        P["UResNet/%s/BatchNorm/beta" % l["scope"]] = (
            beta_std * rng.standard_normal(l["cout"])).astype(np.float32)
    return P


# --------------------------------------------------------------------------
# primitives with TF-1.x semantics.  Tensors here are channel
# first (torch native); conversion from/to N(D)HWC happens at the boundary.
# --------------------------------------------------------------------------
def same_pads(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv_same(x, w_tf, stride):
    """slim.conv{2,3}d padding='same', no bias.  w_tf: [k..,Cin,Cout]."""
    nd = x.dim() - 2
    w = w_tf.permute(nd + 1, nd, *range(nd))            # -> [Cout,Cin,k..]
    pads = []
    for d in reversed(range(nd)):                        # F.pad: last dim first
        pb, pa = same_pads(x.shape[2 + d], w_tf.shape[d], stride)
        pads += [pb, pa]
    x = F.pad(x, pads)
    return (F.conv2d if nd == 2 else F.conv3d)(x, w, stride=stride)


def conv_transpose_same(x, w_tf, stride=2):
    """slim.conv{2,3}d_transpose k=3,s=2,'same': gradient-of-forward-conv form.

    out = conv_transpose(x, W, stride, padding=0) cropped by pad_before of the
    forward SAME conv that maps 2n -> n.  w_tf: [k..,Cout,Cin].
    """
    nd = x.dim() - 2
    w = w_tf.permute(nd + 1, nd, *range(nd))            # -> [Cin,Cout,k..]
    y = (F.conv_transpose2d if nd == 2 else F.conv_transpose3d)(x, w, stride=stride)
    sl = [slice(None), slice(None)]
    for d in range(nd):
        n_out = x.shape[2 + d] * stride
        pb, _ = same_pads(n_out, w_tf.shape[d], stride)
        sl.append(slice(pb, pb + n_out))
    return y[tuple(sl)]


def batch_norm(x, beta):
    """slim.batch_norm defaults: batch stats always, biased var, no gamma."""
    axes = [0] + list(range(2, x.dim()))
    mu = x.mean(dim=axes, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=axes, keepdim=True)
    shp = [1, -1] + [1] * (x.dim() - 2)
    return (x - mu) * torch.rsqrt(var + BN_EPS) + beta.view(shp)


# --------------------------------------------------------------------------
# optional bf16 storage model (mirrors WHERE the CUDA bf16 mode rounds, so that
# the two follow the same trajectory; the arithmetic in between stays fp64)
# --------------------------------------------------------------------------
def _bf16(x):
    return x.to(torch.float32).to(torch.bfloat16).to(x.dtype)


def _f16(x):
    return x.to(torch.float32).to(torch.float16).to(x.dtype)


class _RoundValueAndGrad(torch.autograd.Function):
    """A stored activation: value rounded to bf16; its gradient tensor is bf16 too."""

    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _RoundValueF16AndGrad(torch.autograd.Function):
    """mix16: activation stored in fp16, its gradient tensor in bf16."""

    @staticmethod
    def forward(ctx, x):
        return _f16(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _RoundValueF16STE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _f16(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """fp32 value (raw conv output / logits) whose gradient is stored in bf16."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _RoundValueSTE(torch.autograd.Function):
    """bf16 operand copy of an fp32 master weight; the gradient goes to the master."""

    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


# --------------------------------------------------------------------------
# the graph                                          uresnet.py:22-123
# --------------------------------------------------------------------------
def _t(P, name, dtype):
    v = P[name]
    if isinstance(v, np.ndarray):
        v = torch.from_numpy(v)
    return v.to(dtype) if v.dtype != dtype else v


def forward(P, data_nhwc, base=16, num_strides=5, keep=None, quant=None):
    """data_nhwc: torch tensor [B,H,W,(D,)C].  Returns logits, channel-last.

    P values may be numpy arrays or torch tensors (leaf tensors with
    requires_grad for autograd).  ``keep``: optional dict that receives every
    layer's post-BN/activation output (channel-last), keyed by TF scope.
    ``quant='bf16'`` / ``'mix16'`` insert rounding at the storage points of the CUDA 16-bit
    modes (input, conv weights, stored activations: bf16 resp. fp16; their gradients, d(raw)
    and d(logits): bf16).
    """
    nd = data_nhwc.dim() - 2
    dt = data_nhwc.dtype
    x = data_nhwc.permute(0, nd + 1, *range(1, nd + 1))  # -> channel first
    q = quant in ("bf16", "mix16")
    if quant == "bf16":
        x = _bf16(x)
        store, wround = _RoundValueAndGrad.apply, _RoundValueSTE.apply
    elif quant == "mix16":
        x = _f16(x)
        store, wround = _RoundValueF16AndGrad.apply, _RoundValueF16STE.apply
    else:
        store = (lambda t: t)

    def layer(scope, x, stride, relu, transpose=False, node=True):
        w = _t(P, "UResNet/%s/weights" % scope, dt)
        b = _t(P, "UResNet/%s/BatchNorm/beta" % scope, dt)
        if q:
            w = wround(w)
        y = conv_transpose_same(x, w, stride) if transpose else conv_same(x, w, stride)
        if q:
            y = _RoundGrad.apply(y)                        # d(raw) is a bf16 tensor
        y = batch_norm(y, b)                               # conv -> BN -> act
        if relu:
            y = F.relu(y)
        if node:
            y = store(y)
        if keep is not None:
            keep[scope] = y.permute(0, *range(2, nd + 2), 1)
        return y

    def resnet_module(scope, x, cout, stride):           # resnet_module.py:10-68
        if x.shape[1] == cout and stride == 1:
            sc = x
        else:
            sc = layer(scope + "/shortcut", x, stride, False, node=False)
        r = layer(scope + "/resnet_conv1", x, stride, False)   # no ReLU between
        r = layer(scope + "/resnet_conv2", r, 1, False, node=False)
        return store(F.relu(sc + r))

    def double_resnet(idx, x, cout, stride):             # resnet_module.py:70-87
        x = resnet_module("resnet_module%d/module1" % idx, x, cout, stride)
        return resnet_module("resnet_module%d/module2" % idx, x, cout, 1)

    net = layer("conv0", x, 1, True)
    fmap = {net.shape[1]: net}                           # keyed by channel count
    for step in range(num_strides):
        net = double_resnet(step, net, net.shape[1] * 2, 2)
        fmap[net.shape[1]] = net
    for step in range(num_strides):
        cout = net.shape[1] // 2
        net = layer("deconv%d" % step, net, 2, True, transpose=True)
        net = torch.cat([net, fmap[cout]], dim=1)        # deconv channels first
        net = double_resnet(step + 5, net, cout, 1)
    net = layer("conv1", net, 1, True)
    net = layer("conv2", net, 1, False, node=False)     # logits stay fp32 ...
    if q:
        net = _RoundGrad.apply(net)                      # ... their gradient is bf16
    return net.permute(0, *range(2, nd + 2), 1).contiguous()


# --------------------------------------------------------------------------
# loss + metrics                                     ssnet.py:57-71
# --------------------------------------------------------------------------
def loss_and_metrics(logits, data, label, weight=None):
    """logits [B,...,C] ; data [B,...,1] ; label/weight [B,...] (float class ids).

    Returns dict(loss, acc_all, acc_nonzero, softmax, n_correct_all,
    n_correct_nz, n_nz).
    """
    B = logits.shape[0]
    C = logits.shape[-1]
    lab = label.to(torch.int64)                          # ssnet.py:40 truncation
    if lab.min() < 0 or lab.max() >= C:
        raise ValueError("label out of range")
    lg = logits.reshape(B, -1, C)
    lab2 = lab.reshape(B, -1)
    ce = torch.logsumexp(lg, dim=-1) - torch.gather(lg, 2, lab2[..., None])[..., 0]
    if weight is not None:
        ce = ce * weight.reshape(B, -1).to(ce.dtype)
    loss = ce.sum(dim=1).mean()                          # mean_b( sum_p )
    pred = lg.argmax(dim=-1)                             # first max on ties
    eq = pred == lab2
    nz = data.reshape(B, -1) > 0                         # ssnet.py:59 (assumes C_in=1)
    n_nz = int(nz.sum())
    n_cnz = int((eq & nz).sum())
    return dict(loss=loss, acc_all=float(eq.float().mean()),
                acc_nonzero=(n_cnz / n_nz) if n_nz else float("nan"),
                softmax=torch.softmax(logits, dim=-1),
                n_correct_all=int(eq.sum()), n_correct_nz=n_cnz, n_nz=n_nz)


def run(P_np, data, label, weight=None, base=16, num_strides=5, dtype=torch.float64,
        want_grads=True, keep=None, quant=None):
    """Full forward(+backward).  numpy in / numpy out.

    data [B,...,Cin], label/weight [B,...].  Returns dict with logits, loss,
    acc_all, acc_nonzero, softmax and (optionally) grads keyed like P_np.
    """
    P = OrderedDict((k, torch.tensor(v, dtype=dtype, requires_grad=want_grads))
                    for k, v in P_np.items())
    d = torch.tensor(data, dtype=dtype)
    lab = torch.tensor(label, dtype=dtype)
    w = None if weight is None else torch.tensor(weight, dtype=dtype)
    with torch.set_grad_enabled(want_grads):
        logits = forward(P, d, base=base, num_strides=num_strides, keep=keep, quant=quant)
        res = loss_and_metrics(logits, d, lab, w)
    out = dict(logits=logits.detach().numpy(), loss=float(res["loss"].detach()),
               acc_all=res["acc_all"], acc_nonzero=res["acc_nonzero"],
               softmax=res["softmax"].detach().numpy(),
               n_correct_all=res["n_correct_all"], n_correct_nz=res["n_correct_nz"],
               n_nz=res["n_nz"])
    if want_grads:
        gl = torch.autograd.grad(res["loss"], list(P.values()) + [logits],
                                 allow_unused=False)
        out["grads"] = OrderedDict((k, g.numpy()) for k, g in zip(P.keys(), gl))
        out["dlogits"] = gl[-1].numpy()
    if keep is not None:
        for k in list(keep.keys()):
            keep[k] = keep[k].detach().numpy()
    return out


# --------------------------------------------------------------------------
# TF AdamOptimizer                                   ssnet.py:72-79, SURVEY 3.4
# --------------------------------------------------------------------------
def ssnet_decision(softmax, data, threshold=1.0):
    """lib/ssnet_trainval.py:269-287: shower (class 1) vs track (class 2) map of the analysis
    step, zeroed where data <= threshold.  softmax [..., num_class] numpy, data [...] numpy."""
    shower_score = softmax[..., 1]
    track_score = softmax[..., 2]
    ssnet_result = (shower_score > track_score).astype(np.float32) + \
                   (track_score >= shower_score).astype(np.float32) * 2.0
    nonzero_map = (data > threshold).astype(np.int32)
    return (ssnet_result * nonzero_map).astype(np.float32)


def adam_tf(var, grad, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """One TF-1.x ``ApplyAdam`` step in float32, in place; t is the 1-based step count.

    TF's kernel form (epsilon OUTSIDE the bias correction):
      alpha = lr*sqrt(1-b2^t)/(1-b1^t);  m += (g-m)*(1-b1);  v += (g*g-v)*(1-b2);
      var -= m*alpha/(sqrt(v)+eps)     with (1-b) evaluated in float32 like TF does."""
    f = np.float32
    alpha = f(lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t))
    omb1, omb2 = f(1) - f(b1), f(1) - f(b2)
    m[...] = m + (grad - m) * omb1
    v[...] = v + (grad * grad - v) * omb2
    var[...] = var - (m * alpha) / (np.sqrt(v) + f(eps))
    return var, m, v
