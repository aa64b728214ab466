"""Synthetic LArTPC-like batches behind a ``larcv_threadio``-shaped interface.

The reference reads batches through larcv (``lib/ssnet_trainval.py:64-91,167-171``:
``configure`` / ``start_manager`` / ``next`` / ``fetch_data(name).data()`` /
``.dim()`` / ``fetch_entries`` / ``reset``).  larcv and its ROOT files are not
available, so this module supplies the same call surface over a deterministic
generator: sparse images with straight *tracks* (label 2) and
blob-like *showers* (label 1) on a zero background (label 0).

Arrays are host numpy float32, flattened per event exactly like larcv's
``BatchFillerImage2D`` / ``BatchFillerTensor3D`` output: data ``[B, prod(dims)]``,
label / weight ``[B, prod(dims[:-1])]``.
"""
from __future__ import annotations

import numpy as np


def _draw_event(rng, spatial):
    """One event: returns (data, label) as float32 arrays of shape ``spatial``."""
    nd = len(spatial)
    sp = np.asarray(spatial)
    data = np.zeros(spatial, np.float32)
    label = np.zeros(spatial, np.float32)
    scale = float(sp.min())

    def put(pts, vals, cls):
        pts = np.rint(pts).astype(np.int64)
        ok = np.all((pts >= 0) & (pts < sp), axis=1)
        idx = tuple(pts[ok].T)
        data[idx] = vals[ok]
        label[idx] = cls

    # showers first so that tracks overwrite them where they cross
    for _ in range(int(rng.integers(1, 4))):
        n = int(rng.integers(200, 2001)) if nd == 2 else int(rng.integers(500, 5001))
        n = max(8, int(n * min(1.0, (scale / 192.0) ** nd if nd == 3 else (scale / 512.0) ** 2)))
        start = rng.uniform(0.15, 0.85, nd) * sp
        axis = rng.standard_normal(nd)
        axis /= np.linalg.norm(axis)
        t = rng.exponential(0.12 * scale, n)                  # depth along the axis
        spread = 0.02 * scale + 0.25 * t[:, None]             # cone opening
        pts = start + t[:, None] * axis + rng.standard_normal((n, nd)) * spread * 0.3
        put(pts, (rng.exponential(30.0, n) + 1.0).astype(np.float32), 1.0)
    for _ in range(int(rng.integers(3, 7))):
        a = rng.uniform(0.05, 0.95, nd) * sp
        b = rng.uniform(0.05, 0.95, nd) * sp
        n = int(np.abs(b - a).max()) + 1
        s = np.linspace(0.0, 1.0, n)[:, None]
        put(a + s * (b - a), rng.uniform(10.0, 100.0, n).astype(np.float32), 2.0)
    return data, label


def make_batch(batch, dims, seed, num_class=3, out=None):
    """Return (data [B,prod(dims)], label [B,P], weight [B,P]) float32.

    weight = per-event inverse class frequency at each pixel (NOT yet divided
    by the per-event sum -- the driver does that, ``ssnet_trainval.py:173``).
    ``out``: optional (data, label, weight) arrays to fill (e.g. page-locked buffers).
    """
    spatial = tuple(int(d) for d in dims[:-1])
    if int(dims[-1]) != 1:
        raise NotImplementedError("synthetic source produces single-channel events")
    # seed: an int, or a sequence (seed, rank, step) hashed by SeedSequence into non-overlapping
    # streams (consecutive ints as PCG64 seeds would make rank r replay rank r+1's batches)
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed) if not np.isscalar(seed) else int(seed)))
    P = int(np.prod(spatial))
    if out is not None:
        data, label, weight = out
        assert data.shape == label.shape == weight.shape == (batch, P) and data.dtype == np.float32
    else:
        data = np.empty((batch, P), np.float32)
        label = np.empty((batch, P), np.float32)
        weight = np.empty((batch, P), np.float32)
    for b in range(batch):
        d, l = _draw_event(rng, spatial)
        data[b] = d.reshape(-1)
        label[b] = l.reshape(-1)
        cnt = np.bincount(label[b].astype(np.int64), minlength=num_class).astype(np.float64)
        inv = np.where(cnt > 0, 1.0 / np.maximum(cnt, 1.0), 0.0)
        weight[b] = inv[label[b].astype(np.int64)].astype(np.float32)
    return data, label, weight


class _batch_data(object):
    """What ``larcv_threadio.fetch_data(name)`` returns: ``.data()`` and ``.dim()``."""

    def __init__(self, arr, dim):
        self._arr, self._dim = arr, list(dim)

    def data(self):
        return self._arr

    def dim(self):
        return self._dim


class synthetic_threadio(object):
    """Stand-in for ``larcv.dataloader2.larcv_threadio`` (``ssnet_trainval.py:64-91``).

    ``configure`` accepts the reference's ``{'filler_name','verbosity','filler_cfg'}``
    dict plus the synthetic-only keys ``dims`` (e.g. ``[512,512,1]``), ``seed``,
    ``rank`` and the three stream keywords.  Batch ``i`` of rank ``r`` is drawn from
    ``PCG64(SeedSequence([seed, r, i]))``: streams of different ranks / steps / seeds never
    overlap.  Every rank draws a full ``MINIBATCH_SIZE`` batch (the global batch is
    ``world * MINIBATCH_SIZE``, the reference's ``NUM_MINIBATCHES = world`` accumulation).
    """

    def __init__(self):
        self._cfg = None
        self._batch = 0
        self._step = -1
        self._cur = None

    def configure(self, cfg):
        self._cfg = dict(cfg)
        self._dims = [int(v) for v in self._cfg.get("dims", [512, 512, 1])]
        self._seed = int(self._cfg.get("seed", 1234))
        self._rank = int(self._cfg.get("rank", 0))
        # optional allocator(shape) -> float32 array (``ssnet_base.pinned_empty``): batches are then
        # produced in a ring of three page-locked buffer sets and reach the device without a
        # pageable -> pinned staging copy (larcv's filler threads hand over buffers the same way)
        self._alloc = self._cfg.get("allocator")
        self._ring = None
        self._names = {"data": self._cfg.get("keyword_data", "data"),
                       "label": self._cfg.get("keyword_label", "label"),
                       "weight": self._cfg.get("keyword_weight", "weight")}

    def start_manager(self, batch_size):
        self._batch = int(batch_size)

    def next(self, store_entries=False, store_event_ids=False):
        self._step += 1
        out = None
        if self._alloc is not None:
            if self._ring is None:
                P = int(np.prod(self._dims[:-1]))
                self._ring = [tuple(self._alloc((self._batch, P)) for _ in range(3)) for _ in range(3)]
            out = self._ring[self._step % 3]
        d, l, w = make_batch(self._batch, self._dims, (self._seed, self._rank, self._step), out=out)
        self._cur = {self._names["data"]: (d, [self._batch] + self._dims),
                     self._names["label"]: (l, [self._batch] + self._dims[:-1] + [1]),
                     self._names["weight"]: (w, [self._batch] + self._dims[:-1] + [1])}
        self._entries = np.arange(self._step * self._batch, (self._step + 1) * self._batch)

    def fetch_data(self, name):
        arr, dim = self._cur[name]
        return _batch_data(arr, dim)

    def fetch_entries(self):
        return self._entries

    def fetch_event_ids(self):
        return [(0, 0, int(e)) for e in self._entries]

    def reset(self):
        self._cur = None
