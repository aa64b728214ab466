"""CPU checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/uresnet_b200.h declares; the plan (pure host code) matches the oracle's layer table."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import uresnet_oracle as O
from uresnet_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported_and_bound(lib):
    hdr = open(os.path.join(ROOT, "include", "uresnet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(urn_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in the header but not exported" % n
        assert n in _capi.SIGNATURES, "symbol %s has no ctypes signature" % n
    assert set(_capi.SIGNATURES) <= names


def _cfg(nd, sp, B, F=16, S=5, C_=3, dt=0, tr=1):
    cfg = _capi.urn_net_cfg()
    cfg.nd = nd
    for i in range(3):
        cfg.spatial[i] = sp[i] if i < len(sp) else 1
    cfg.in_channels, cfg.batch, cfg.base_filters, cfg.num_strides = 1, B, F, S
    cfg.num_class, cfg.dtype, cfg.impl, cfg.training = C_, dt, 0, tr
    return cfg


@pytest.mark.parametrize("nd,sp,B,F,S", [(2, (512, 512), 32, 16, 5), (3, (192, 192, 192), 4, 16, 5),
                                         (2, (32, 32), 2, 4, 3), (3, (16, 16, 16), 1, 8, 2)])
def test_plan_matches_oracle_layer_table(lib, nd, sp, B, F, S):
    h = C.c_void_p()
    cfg = _cfg(nd, sp, B, F, S)
    _capi.check(lib.urn_create(C.byref(cfg), C.byref(h)))
    names = O.param_names(nd, base=F, num_strides=S)
    assert lib.urn_param_count(h) == len(names) and lib.urn_layer_count(h) == len(names) // 2
    last_end = 0
    for i, (nm, shape) in enumerate(names):
        name, rank = C.c_char_p(), C.c_int()
        shp = (C.c_int64 * 6)()
        off, cnt = C.c_int64(), C.c_int64()
        _capi.check(lib.urn_param_info(h, i, C.byref(name), C.byref(rank), shp, C.byref(off), C.byref(cnt)))
        assert name.value.decode() == nm
        assert tuple(shp[k] for k in range(rank.value)) == tuple(shape)
        assert cnt.value == int(np.prod(shape)) and off.value >= last_end and off.value % 4 == 0
        last_end = off.value + cnt.value
    assert lib.urn_param_total(h) >= last_end
    assert lib.urn_workspace_bytes(h) > 0
    # layer geometry follows TF 'same' rules
    tab = O.layer_table(nd, base=F, num_strides=S)
    for i, l in enumerate(tab):
        scope, d, tr = C.c_char_p(), _capi.urn_conv_desc(), C.c_int()
        _capi.check(lib.urn_layer_info(h, i, C.byref(scope), C.byref(d), C.byref(tr), None, None, None, None))
        assert scope.value.decode() == l["scope"] and bool(tr.value) == (l["op"] == "convT")
        assert d.s == l["s"] and d.kh == l["k"] and d.kd == (l["k"] if nd == 3 else 1)
        if l["op"] == "conv":
            assert (d.cx, d.cy) == (l["cin"], l["cout"])
        else:
            assert (d.cx, d.cy) == (l["cout"], l["cin"])
        if l["k"] == 3 and l["s"] == 2:
            assert (d.pad_h, d.pad_w) == (0, 0)          # asymmetric SAME: pad only at the high end
        if l["k"] == 3 and l["s"] == 1:
            assert (d.pad_h, d.pad_w) == (1, 1)
    lib.urn_destroy(h)


def test_full_size_workspace_fits_one_b200(lib):
    for nd, sp, B in ((2, (512, 512), 32), (3, (192, 192, 192), 4)):
        for dt in (0, 1):
            h = C.c_void_p()
            cfg = _cfg(nd, sp, B, dt=dt)
            _capi.check(lib.urn_create(C.byref(cfg), C.byref(h)))
            assert lib.urn_workspace_bytes(h) < 120 * 2**30
            lib.urn_destroy(h)


def test_create_rejects_bad_configs(lib):
    h = C.c_void_p()
    for bad in (_cfg(4, (8, 8), 1), _cfg(2, (30, 32), 1), _cfg(2, (32, 32), 0), _cfg(2, (32, 32), 1, C_=99)):
        assert lib.urn_create(C.byref(bad), C.byref(h)) != 0
        assert len(lib.urn_last_error()) > 0


def test_same_padding_formula(lib):
    for n in range(1, 40):
        for k, s in ((3, 1), (3, 2), (1, 2), (1, 1)):
            d = _capi.conv_desc_same(2, 1, (n, n), 1, 1, k, s)
            assert d.yh == -(-n // s)
            assert d.pad_h == O.same_pads(n, k, s)[0]


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/liburesnet_b200.so")
    with pytest.raises(ImportError):
        _capi.load()
