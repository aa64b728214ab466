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


def test_create_rejects_more_than_five_strides(lib):
    """uresnet.py:100 hard-codes the decoder scope as step+5: with num_strides > 5 encoder and
    decoder modules would share variable names (TF raises); the plan must refuse, not overwrite."""
    h = C.c_void_p()
    bad = _cfg(2, (64, 64), 1, F=4, S=6)
    assert lib.urn_create(C.byref(bad), C.byref(h)) != 0
    assert b"num_strides" in lib.urn_last_error()
    from uresnet_b200.uresnet import uresnet
    with pytest.raises(ValueError):
        uresnet(dims=[64, 64, 1], num_class=3, num_strides=6, base_num_outputs=4)._build()


@pytest.mark.parametrize("nd,sp,B,F,S,dt", [(2, (64, 64), 2, 16, 5, 2), (3, (16, 16, 16), 1, 16, 3, 2), (2, (32, 32), 1, 4, 2, 0)])
def test_plan_introspection_matches_step_oracle(lib, nd, sp, B, F, S, dt):
    """The parity hooks (urn_plan_counts / urn_node_info / urn_layer_tensors / urn_*_step_info) describe
    the same graph as oracle/steps.py: same node names and shapes, every forward tensor inside the
    workspace, every schedule step well formed.  (Host-only: nothing is launched.)"""
    from oracle.steps import StepOracle
    h = C.c_void_p()
    cfg = _cfg(nd, sp, B, F, S, dt=dt)
    _capi.check(lib.urn_create(C.byref(cfg), C.byref(h)))
    rng = np.random.default_rng(0)
    shp = [B] + list(sp)
    so = StepOracle(O.init_params(nd, base=F, num_strides=S, seed=1), rng.random(shp + [1]),
                    rng.integers(0, 3, shp).astype(np.float64), None, base=F, num_strides=S)
    n, nf, nb = C.c_int(), C.c_int(), C.c_int()
    _capi.check(lib.urn_plan_counts(h, C.byref(n), C.byref(nf), C.byref(nb)))
    ws_bytes = lib.urn_workspace_bytes(h)
    names = []
    for i in range(n.value):
        name = C.c_char_p()
        v, sh, g = _capi.urn_tensor_ref(), _capi.urn_tensor_ref(), _capi.urn_tensor_ref()
        _capi.check(lib.urn_node_info(h, i, C.byref(name), C.byref(v), C.byref(sh), C.byref(g)))
        nm = name.value.decode()
        names.append(nm)
        assert nm in so.nodes, nm
        assert (v.rows, v.c) == so.node(nm).shape
        es = 4 if v.dtype == 0 else 2
        assert 0 <= v.off_bytes and v.off_bytes + ((v.rows - 1) * v.ld + v.c) * es <= ws_bytes
        assert (sh.off_bytes >= 0) == (dt == 2 and nm != "conv2")           # bf16 shadow of fp16 activations
        if nm != "data":
            assert g.off_bytes >= 0 and g.dtype == (1 if dt == 2 else dt) and (g.rows, g.c, g.ld) == (v.rows, v.c, v.ld)
    assert set(names) == set(so.nodes)
    nl = lib.urn_layer_count(h)
    scopes = []
    for li in range(nl):
        raw, draw = _capi.urn_tensor_ref(), _capi.urn_tensor_ref()
        so_, wo, bo, inn, outn = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
        _capi.check(lib.urn_layer_tensors(h, li, C.byref(raw), C.byref(draw), C.byref(so_), C.byref(wo), C.byref(bo),
                                          C.byref(inn), C.byref(outn)))
        scope = C.c_char_p()
        _capi.check(lib.urn_layer_info(h, li, C.byref(scope), None, None, None, None, None, None))
        sc = scope.value.decode()
        scopes.append(sc)
        assert (raw.rows, raw.c) == so.raw(sc).shape
        # mix16: raw conv outputs stored as fp16 except the stem, the head and non-power-of-two / narrow widths
        want16 = dt == 2 and sc not in ("conv0", "conv2") and raw.c % 16 == 0 and (raw.c & (raw.c - 1)) == 0
        assert raw.dtype == (2 if want16 else 0), (sc, raw.dtype)
        assert names[inn.value] == so.layers[sc]["x"]
        assert (names[outn.value] if outn.value >= 0 else None) == so.layers[sc]["out"]
        assert draw.c == raw.c and draw.ld >= draw.c
    assert scopes == list(so.layers)
    kinds = []
    for i in range(nf.value):
        st = _capi.urn_step_info()
        _capi.check(lib.urn_fwd_step_info(h, i, C.byref(st)))
        kinds.append(st.kind)
        if st.kind == 2:                       # merge: conv shortcut xor identity residual
            assert (st.layer2 >= 0) != (st.node_b >= 0)
            assert names[st.node_a] == scopes[st.layer][:-len("/resnet_conv2")]
    assert kinds.count(0) == nl and kinds.count(2) == 4 * S and kinds.count(1) == nl - 4 * S - (sum(1 for s_ in scopes if s_.endswith("shortcut")))
    nw = nd_ = 0
    for j in range(nb.value):
        st = _capi.urn_step_info()
        _capi.check(lib.urn_bwd_step_info(h, j, C.byref(st)))
        nw += st.kind == 1
        nd_ += st.kind == 2
        if st.kind == 0 and st.node_b >= 0:
            assert names[st.node_b] == so.layers[scopes[st.layer][:-1] + "1"]["x"]   # identity shortcut = module input
    assert nw == nl and nd_ == nl - 1          # every layer has a filter gradient; conv0 has no input gradient
    lib.urn_destroy(h)


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
