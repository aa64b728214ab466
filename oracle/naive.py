"""Independent naive NumPy statements of each primitive (TEST INFRASTRUCTURE ONLY).

Written directly from the index formulas of SURVEY.md section 3.4 with explicit
tap loops -- no torch, no shared code with ``uresnet_oracle.py`` -- so that the
two can be checked against each other at tiny shapes (tests/test_oracle.py).
All arrays channel-last, float64.
"""
import itertools

import numpy as np


def same_pad_before(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2


def conv_same(x, w, s):
    """x [B,*sp,Cin], w [*k,Cin,Cout]; out[o] = sum_k x[s*o + k - pb] w[k]."""
    nd = x.ndim - 2
    sp = x.shape[1:1 + nd]
    osp = tuple(-(-n // s) for n in sp)
    pb = [same_pad_before(sp[d], w.shape[d], s) for d in range(nd)]
    out = np.zeros((x.shape[0],) + osp + (w.shape[-1],))
    for o in itertools.product(*[range(n) for n in osp]):
        for k in itertools.product(*[range(w.shape[d]) for d in range(nd)]):
            i = tuple(s * o[d] + k[d] - pb[d] for d in range(nd))
            if any(i[d] < 0 or i[d] >= sp[d] for d in range(nd)):
                continue
            out[(slice(None),) + o] += x[(slice(None),) + i] @ w[k]
    return out


def conv_transpose_same(x, w, s=2):
    """x [B,*sp,Cin], w [*k,Cout,Cin]; scatter form:
    out[s*i + k - pb] += x[i] . w[k]^T, outputs outside [0, s*n) dropped."""
    nd = x.ndim - 2
    sp = x.shape[1:1 + nd]
    osp = tuple(s * n for n in sp)
    pb = [same_pad_before(osp[d], w.shape[d], s) for d in range(nd)]
    out = np.zeros((x.shape[0],) + osp + (w.shape[-2],))
    for i in itertools.product(*[range(n) for n in sp]):
        for k in itertools.product(*[range(w.shape[d]) for d in range(nd)]):
            y = tuple(s * i[d] + k[d] - pb[d] for d in range(nd))
            if any(y[d] < 0 or y[d] >= osp[d] for d in range(nd)):
                continue
            out[(slice(None),) + y] += x[(slice(None),) + i] @ w[k].T
    return out


def batch_norm(x, beta, eps=1e-3):
    ax = tuple(range(x.ndim - 1))
    mu = x.mean(axis=ax)
    var = (x * x).mean(axis=ax) - mu * mu
    return (x - mu) / np.sqrt(var + eps) + beta


def ce_loss(logits, label, weight=None):
    """mean over batch of per-image sums of (weighted) softmax cross entropy."""
    B, C = logits.shape[0], logits.shape[-1]
    lg = logits.reshape(B, -1, C)
    lab = label.reshape(B, -1).astype(np.int64)
    tot = 0.0
    for b in range(B):
        for p in range(lg.shape[1]):
            z = lg[b, p]
            m = z.max()
            ce = m + np.log(np.exp(z - m).sum()) - z[lab[b, p]]
            if weight is not None:
                ce *= weight.reshape(B, -1)[b, p]
            tot += ce
    return tot / B


def accuracies(logits, label, data):
    C = logits.shape[-1]
    pred = logits.reshape(-1, C).argmax(axis=1)
    lab = label.reshape(-1).astype(np.int64)
    nz = data.reshape(-1) > 0
    eq = pred == lab
    return eq.mean(), (eq[nz].mean() if nz.any() else float("nan"))
