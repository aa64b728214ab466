"""Step-by-step form of the CPU oracle (TEST INFRASTRUCTURE ONLY; what is and is not pinned by the
reference: ``oracle/__init__.py``.  Every forward and backward intermediate kept here is compared with
the tensors of the reference's own graph in tests/test_reference_run.py).

``uresnet_oracle.forward`` states the graph of /root/reference/lib/uresnet.py:22-123 and
lib/resnet_module.py:10-87 as one autograd expression.  ``StepOracle`` states the SAME graph one
slim layer at a time, with the SAME primitives (``conv_same``, ``conv_transpose_same``,
``batch_norm``), keeps every intermediate by name, and derives the backward pass one schedule
step at a time with *local* autograd on those primitives:

  forward, per layer ``<scope>``        raw = conv(x, W);  (mean, rstd) of raw;  BN(+ReLU) output
  forward, per residual module          relu(shortcut + BN(raw of resnet_conv2))   (resnet_module.py:68)
  backward ``("bn", scope)``            d(raw), d(beta) [+ the shortcut's, + the masked gradient that
                                        flows to an identity shortcut] from the gradient of the node
  backward ``("wgrad", scope)``         d(W) from (input node, d(raw))
  backward ``("dgrad", scope)``         contribution of the layer to the gradient of its input node

Node names are the TF scopes under ``UResNet/`` that produce them: ``conv0``, ``deconv<k>``,
``conv1``, ``conv2`` (simple layers), ``resnet_module<i>/module<j>`` (module outputs),
``resnet_module<i>/module<j>/resnet_conv1`` (the BN output between the two 3x3 convs),
``concat<k>`` (uresnet.py:91-93, deconv channels first) and ``data``.  The CUDA library names its
plan nodes the same way (``urn_node_info``), which is what lets tests/test_parity_gpu.py feed
each of its schedule steps the oracle's tensors ("teacher forcing") and compare what the step
produces with what is recorded here.

``storage`` ("mix16" / "f32") adds, next to every chain value, the value the SAME step yields when
its inputs are first rounded to the dtype the CUDA library stores them in (activations fp16, their
bf16 shadow for the filter gradient, gradients bf16, raw conv outputs / statistics fp32): the
``expected=True`` accessors.  The chain itself is never rounded -- every step is taught the fp64
trajectory -- so a deviation between the library and ``expected`` is the arithmetic of that one
step on identical inputs, which is what north_star's per-tensor tolerance can be asked of (sums
that cancel, e.g. the beta gradient behind a zero-mean d(raw), amplify INPUT rounding without
bound; see oracle/conditioning.py).

``tests/test_oracle.py`` checks that the gradients assembled from these per-step records equal the
whole-graph autograd gradients of ``uresnet_oracle.run`` to fp64 round-off, so the two statements
of the graph cannot drift apart.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import uresnet_oracle as O


def _rounder(tdt):
    return lambda t: t.to(torch.float32).to(tdt).to(t.dtype)


STORAGE = {
    None: None,
    "f32": dict(act=_rounder(torch.float32), shadow=_rounder(torch.float32), grad=_rounder(torch.float32)),
    "mix16": dict(act=_rounder(torch.float16), shadow=_rounder(torch.bfloat16), grad=_rounder(torch.bfloat16)),
    "bf16": dict(act=_rounder(torch.bfloat16), shadow=_rounder(torch.bfloat16), grad=_rounder(torch.bfloat16)),
}
_r32 = _rounder(torch.float32)
_r16 = _rounder(torch.float16)


def _cl(t):
    """channel-first torch tensor -> channel-last numpy [rows, C]"""
    nd = t.dim() - 2
    return t.detach().permute(0, *range(2, nd + 2), 1).reshape(-1, t.shape[1]).numpy()


class StepOracle(object):

    def __init__(self, P_np, data, label, weight=None, base=16, num_strides=5, dtype=torch.float64,
                 storage=None, raw16=()):
        """data [B,...,Cin], label / weight [B,...] numpy (as for ``uresnet_oracle.run``).
        ``raw16``: scopes whose raw conv output the library stores as fp16 (``urn_layer_tensors``);
        all others are stored as fp32."""
        self.dt = dtype
        self.q = STORAGE[storage]
        self.raw16 = set(raw16)
        self.nodes_e = OrderedDict()        # name -> the producing step's output on storage-rounded inputs
        self.P = OrderedDict((k, torch.tensor(v, dtype=dtype)) for k, v in P_np.items())
        d = torch.tensor(data, dtype=dtype)
        self.nd = d.dim() - 2
        self.data_cl = d
        self.label = torch.tensor(label, dtype=dtype)
        self.weight = None if weight is None else torch.tensor(weight, dtype=dtype)
        self.base, self.S = base, num_strides
        self.nodes = OrderedDict()          # name -> channel-first tensor
        self.layers = OrderedDict()         # scope -> dict(x, raw, mean, rstd, stride, transpose, relu, out)
        self.recs = []                      # forward records: ("layer", scope) / ("module", scope, sc|None) / ("concat", k)
        self._forward()

    # ------------------------------------------------------------------ forward
    def _w(self, scope):
        return self.P["UResNet/%s/weights" % scope], self.P["UResNet/%s/BatchNorm/beta" % scope]

    def _conv(self, scope, x):
        L = self.layers[scope]
        w, _ = self._w(scope)
        return O.conv_transpose_same(x, w, L["stride"]) if L["transpose"] else O.conv_same(x, w, L["stride"])

    def _layer(self, scope, x_name, stride, relu, transpose=False, out=None):
        self.layers[scope] = L = dict(x=x_name, stride=stride, relu=relu, transpose=transpose, out=out)
        raw = self._conv(scope, self.nodes[x_name])
        axes = [0] + list(range(2, raw.dim()))
        mu = raw.mean(dim=axes)
        var = ((raw - mu.view([1, -1] + [1] * self.nd)) ** 2).mean(dim=axes)     # biased
        L.update(raw=raw, mean=mu, rstd=torch.rsqrt(var + O.BN_EPS))
        if self.q:                                                     # the same conv on the stored input
            raw_e = self._conv(scope, self.q["act"](self.nodes[x_name]))
            mu_e = raw_e.mean(dim=axes)
            var_e = ((raw_e - mu_e.view([1, -1] + [1] * self.nd)) ** 2).mean(dim=axes)
            L.update(raw_e=raw_e, mean_e=mu_e, rstd_e=torch.rsqrt(var_e + O.BN_EPS))
        if out is not None:
            y = O.batch_norm(raw, self._w(scope)[1])
            self.nodes[out] = F.relu(y) if relu else y
            if self.q:
                y = self._bn_stored(scope)
                self.nodes_e[out] = F.relu(y) if relu else y
            self.recs.append(("layer", scope))

    def _raw_stored(self, scope):
        """the raw conv output as the library stores it (fp32, or fp16 for the ``raw16`` scopes)"""
        return (_r16 if scope in self.raw16 else _r32)(self.layers[scope]["raw"])

    def _bn_stored(self, scope):
        """BN output from the STORED raw tensor and statistics (fp32), as the apply kernels see them."""
        L = self.layers[scope]
        shp = [1, -1] + [1] * self.nd
        return (self._raw_stored(scope) - _r32(L["mean"]).view(shp)) * _r32(L["rstd"]).view(shp) + self._w(scope)[1].view(shp)

    def _module(self, scope, x_name, cout, stride):                    # resnet_module.py:10-68
        x = self.nodes[x_name]
        sc = None
        if not (x.shape[1] == cout and stride == 1):
            sc = scope + "/shortcut"
            self._layer(sc, x_name, stride, False)
        c1, c2 = scope + "/resnet_conv1", scope + "/resnet_conv2"
        self._layer(c1, x_name, stride, False, out=c1)
        self.recs.pop()                                                # recorded as part of the module
        self._layer(c2, c1, 1, False)
        self.nodes[scope] = self._merge(scope, sc, self.layers[c2]["raw"],
                                        self.layers[sc]["raw"] if sc else None, x if sc is None else None)
        if self.q:
            s_e = self._bn_stored(sc) if sc else self.q["act"](x)
            self.nodes_e[scope] = F.relu(s_e + self._bn_stored(c2))
        self.recs.append(("module", scope, sc, x_name))
        return scope

    def _merge(self, scope, sc, raw2, raw_sc, x_res):
        r = O.batch_norm(raw2, self._w(scope + "/resnet_conv2")[1])
        s = O.batch_norm(raw_sc, self._w(sc)[1]) if sc else x_res
        return F.relu(s + r)                                            # resnet_module.py:68

    def _double(self, idx, x_name, cout, stride):                       # resnet_module.py:70-87
        m1 = self._module("resnet_module%d/module1" % idx, x_name, cout, stride)
        return self._module("resnet_module%d/module2" % idx, m1, cout, 1)

    def _forward(self):
        nd = self.nd
        self.nodes["data"] = self.data_cl.permute(0, nd + 1, *range(1, nd + 1))
        self._layer("conv0", "data", 1, True, out="conv0")
        net = "conv0"
        skips = {self.nodes[net].shape[1]: net}                         # keyed by channel count (uresnet.py:47,64)
        for step in range(self.S):
            net = self._double(step, net, self.nodes[net].shape[1] * 2, 2)
            skips[self.nodes[net].shape[1]] = net
        for step in range(self.S):
            cout = self.nodes[net].shape[1] // 2
            dn = "deconv%d" % step
            self._layer(dn, net, 2, True, transpose=True, out=dn)
            cat = "concat%d" % step
            self.nodes[cat] = torch.cat([self.nodes[dn], self.nodes[skips[cout]]], dim=1)
            self.recs.append(("concat", cat, dn, skips[cout]))
            net = self._double(step + 5, cat, cout, 1)                   # scope step+5 (uresnet.py:100)
        self._layer("conv1", net, 1, True, out="conv1")
        self._layer("conv2", "conv1", 1, False, out="conv2")
        self.logits = self.nodes["conv2"].permute(0, *range(2, nd + 2), 1).contiguous()

    # ------------------------------------------------------------------ loss
    def loss(self):
        lg = (_r32(self.logits) if self.q else self.logits).detach().requires_grad_(True)
        res = O.loss_and_metrics(lg, self.data_cl, self.label, self.weight)
        (dl,) = torch.autograd.grad(res["loss"], [lg])
        self.result = dict(loss=float(res["loss"].detach()), acc_all=res["acc_all"], acc_nonzero=res["acc_nonzero"],
                           n_correct_all=res["n_correct_all"], n_correct_nz=res["n_correct_nz"], n_nz=res["n_nz"])
        nd = self.nd
        return dl.permute(0, nd + 1, *range(1, nd + 1))                 # channel first

    # ------------------------------------------------------------------ backward
    def _bn_bwd(self, scope, dout, relu, stored=False):
        """Local autograd through y = [relu](BN(raw, beta)).  Returns (draw, dbeta)."""
        L = self.layers[scope]
        if stored:
            return self._bn_bwd_stored(scope, self.q["grad"](dout), relu=relu)[:2]
        raw = L["raw"].detach().requires_grad_(True)
        beta = self._w(scope)[1].detach().requires_grad_(True)
        y = O.batch_norm(raw, beta)
        if relu:
            y = F.relu(y)
        return torch.autograd.grad(y, [raw, beta], dout)

    def _bn_bwd_stored(self, scope, g, relu=False, mask=None):
        """slim.batch_norm backward in closed form on the STORED tensors, as the kernels evaluate it:
        xhat from the stored raw output and the stored statistics (which were taken from the fp32
        accumulators -- with an fp16 raw copy they are NOT the statistics of the stored values, so
        differentiating batch_norm(stored raw) would be a different function):
            g' = g * mask;  dbeta = sum g';  d(raw) = rstd * (g' - mean(g') - xhat * mean(g' * xhat)).
        Returns (d(raw), dbeta, g')."""
        L = self.layers[scope]
        shp = [1, -1] + [1] * self.nd
        axes = [0] + list(range(2, g.dim()))
        mu, rstd = _r32(L["mean"]).view(shp), _r32(L["rstd"]).view(shp)
        xhat = (self._raw_stored(scope) - mu) * rstd
        if relu:
            mask = (xhat + self._w(scope)[1].view(shp)) > 0
        if mask is not None:
            g = g * mask.to(g.dtype)
        mg = g.mean(dim=axes, keepdim=True)
        mgx = (g * xhat).mean(dim=axes, keepdim=True)
        return rstd * (g - mg - xhat * mgx), g.sum(dim=axes), g

    def _conv_bwd(self, scope, draw, stored=False):
        """Local autograd through raw = conv(x, W).  Returns (dx or None for the stem, dW)."""
        L = self.layers[scope]
        x = self.nodes[L["x"]]
        if stored:       # filter gradient reads the bf16 shadow of the activation (the fp16 data node for the stem)
            x = self.q["act" if L["x"] == "data" else "shadow"](x)
            draw = self.q["grad"](draw)
        x = x.detach().requires_grad_(L["x"] != "data")
        w = self._w(scope)[0].detach().requires_grad_(True)
        y = O.conv_transpose_same(x, w, L["stride"]) if L["transpose"] else O.conv_same(x, w, L["stride"])
        if L["x"] == "data":
            (dw,) = torch.autograd.grad(y, [w], draw)
            return None, dw
        dx, dw = torch.autograd.grad(y, [x, w], draw)
        return dx, dw

    def backward(self):
        """Fills ``node_grad`` (total gradient of every node), ``bn[scope]`` = dict(dout, draw,
        dbeta[, g]), ``dw[scope]``, ``dx[scope]`` (channel-first torch tensors) and ``grads``
        (numpy, keyed like the parameter dict)."""
        G = {}
        self.bn, self.dw, self.dx = {}, {}, {}
        self.bn_e, self.dw_e, self.dx_e = {}, {}, {}      # the same steps on storage-rounded inputs
        q = self.q

        def add(name, g):
            G[name] = g if name not in G else G[name] + g

        def conv_part(scope):
            dx, dw = self._conv_bwd(scope, self.bn[scope]["draw"])
            self.dw[scope] = dw
            if dx is not None:
                self.dx[scope] = dx
                add(self.layers[scope]["x"], dx)
            if q:
                self.dx_e[scope], self.dw_e[scope] = self._conv_bwd(scope, self.bn[scope]["draw"], stored=True)

        G["conv2"] = self.loss()
        for rec in reversed(self.recs):
            if rec[0] == "layer":
                scope = rec[1]
                L = self.layers[scope]
                dout = G[L["out"]]
                draw, dbeta = self._bn_bwd(scope, dout, L["relu"])
                self.bn[scope] = dict(dout=dout, draw=draw, dbeta=dbeta)
                if q:
                    d_e, db_e = self._bn_bwd(scope, dout, L["relu"], stored=True)
                    self.bn_e[scope] = dict(draw=d_e, dbeta=db_e)
                conv_part(scope)
            elif rec[0] == "concat":
                _, cat, dn, skip = rec
                c = self.nodes[dn].shape[1]
                add(dn, G[cat][:, :c])
                add(skip, G[cat][:, c:])
            else:
                _, scope, sc, x_name = rec
                c1, c2 = scope + "/resnet_conv1", scope + "/resnet_conv2"
                dout = G[scope]
                raw2 = self.layers[c2]["raw"].detach().requires_grad_(True)
                b2 = self._w(c2)[1].detach().requires_grad_(True)
                if sc:
                    raws = self.layers[sc]["raw"].detach().requires_grad_(True)
                    bs = self._w(sc)[1].detach().requires_grad_(True)
                    y = self._merge_local(raw2, b2, O.batch_norm(raws, bs))
                    d2, db2, ds, dbs = torch.autograd.grad(y, [raw2, b2, raws, bs], dout)
                    self.bn[c2] = dict(dout=dout, draw=d2, dbeta=db2)
                    self.bn[sc] = dict(dout=dout, draw=ds, dbeta=dbs)
                else:
                    xr = self.nodes[x_name].detach().requires_grad_(True)
                    y = self._merge_local(raw2, b2, xr)
                    d2, db2, g = torch.autograd.grad(y, [raw2, b2, xr], dout)
                    self.bn[c2] = dict(dout=dout, draw=d2, dbeta=db2, g=g)
                    add(x_name, g)
                if q:
                    # the kernels take the ReLU mask of a merge from the STORED module output; the same masked
                    # gradient flows through both batch-norms (and, unchanged, to an identity shortcut)
                    mask = q["act"](self.nodes[scope]) > 0
                    d2, db2, g = self._bn_bwd_stored(c2, q["grad"](dout), mask=mask)
                    self.bn_e[c2] = dict(draw=d2, dbeta=db2, g=g)
                    if sc:
                        ds, dbs, _ = self._bn_bwd_stored(sc, q["grad"](dout), mask=mask)
                        self.bn_e[sc] = dict(draw=ds, dbeta=dbs)
                conv_part(c2)
                draw1, db1 = self._bn_bwd(c1, G[c1], False)
                self.bn[c1] = dict(dout=G[c1], draw=draw1, dbeta=db1)
                if q:
                    d_e, db_e = self._bn_bwd(c1, G[c1], False, stored=True)
                    self.bn_e[c1] = dict(draw=d_e, dbeta=db_e)
                conv_part(c1)
                if sc:
                    conv_part(sc)
        self.node_grad = G
        self.grads = OrderedDict()
        for scope in self.layers:
            self.grads["UResNet/%s/weights" % scope] = self.dw[scope].numpy()
            self.grads["UResNet/%s/BatchNorm/beta" % scope] = self.bn[scope]["dbeta"].numpy()
        return self.grads

    @staticmethod
    def _merge_local(raw2, b2, s, relu=True):
        y = s + O.batch_norm(raw2, b2)
        return F.relu(y) if relu else y

    # ------------------------------------------------------------------ views for the tests
    # ``expected=True``: the producing step's output on storage-rounded inputs (needs ``storage``)
    def node(self, name, expected=False):
        """[rows, C] float64 numpy, channel-last (the layout of the CUDA workspace)."""
        return _cl((self.nodes_e if expected else self.nodes)[name])

    def node_gradient(self, name):
        return _cl(self.node_grad[name])

    def raw(self, scope, expected=False):
        return _cl(self.layers[scope]["raw_e" if expected else "raw"])

    def stats(self, scope, expected=False):
        L = self.layers[scope]
        e = "_e" if expected else ""
        return np.concatenate([L["mean" + e].numpy(), L["rstd" + e].numpy()])    # mean | rstd, as urn_bn_stats stores them

    def draw(self, scope, expected=False):
        return _cl((self.bn_e if expected else self.bn)[scope]["draw"])

    def dbeta(self, scope, expected=False):
        return (self.bn_e if expected else self.bn)[scope]["dbeta"].numpy()

    def filter_gradient(self, scope, expected=False):
        return (self.dw_e if expected else self.dw)[scope].numpy()

    def masked_gradient(self, scope, expected=False):
        """gradient handed to the identity shortcut by the merge of module ``scope``"""
        return _cl((self.bn_e if expected else self.bn)[scope + "/resnet_conv2"]["g"])

    def input_gradient(self, scope, expected=False):
        return _cl((self.dx_e if expected else self.dx)[scope])
