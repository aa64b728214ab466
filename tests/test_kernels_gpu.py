"""Per-kernel parity (GPU): every K-row of SURVEY.md 8a through the C ABI vs the CPU oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import uresnet_oracle as O
from uresnet_b200 import _capi

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
F32, BF16, F16 = _capi.URN_F32, _capi.URN_BF16, _capi.URN_F16
TOL = {F32: 2e-5, BF16: 1.5e-2, F16: 1.5e-2}   # F16 = mix16: fp16 activations, bf16 gradients
TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16, F16: torch.float16}     # max|a-b| / max|b| per tensor


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def dev(x, dt=F32):
    t = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).to(DEV)
    return t.to(TORCH_DT[dt])


def gdt(dt):
    """gradient storage dtype of a network mode (mix16 keeps gradients in bf16)"""
    return BF16 if dt == F16 else dt


def host(t):
    return t.float().cpu().numpy()


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def cf(x):
    nd = x.ndim - 2
    return torch.tensor(x, dtype=torch.float64).permute(0, nd + 1, *range(1, nd + 1))


def cl(y):
    nd = y.dim() - 2
    return y.permute(0, *range(2, nd + 2), 1).contiguous().numpy()


CONV_CASES = [
    # nd, n, spatial, cin, cout, k, s
    (2, 2, (12, 20), 16, 16, 3, 1),
    (2, 1, (16, 16), 1, 16, 3, 1),       # stem (Cin = 1)
    (2, 2, (9, 14), 16, 3, 3, 1),        # head (Cout = 3), odd size
    (2, 2, (16, 24), 16, 32, 3, 2),      # stride-2, asymmetric SAME pad
    (2, 1, (15, 17), 8, 12, 3, 2),       # odd sizes: pad_before changes
    (2, 2, (16, 16), 32, 64, 1, 2),      # 1x1 shortcut s2
    (2, 2, (8, 8), 64, 32, 1, 1),        # 1x1 shortcut s1
    (2, 1, (8, 8), 128, 80, 3, 1),       # two cout tiles, ragged
    (3, 2, (6, 8, 10), 16, 16, 3, 1),
    (3, 1, (8, 8, 8), 16, 32, 3, 2),
    (3, 1, (8, 6, 4), 32, 64, 1, 2),
    (3, 1, (4, 4, 4), 1, 16, 3, 1),
]


@pytest.mark.parametrize("dt", [F32, BF16, F16])
@pytest.mark.parametrize("nd,n,sp,cin,cout,k,s", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(lib, dt, nd, n, sp, cin, cout, k, s):
    _conv_case(lib, dt, 1, nd, n, sp, cin, cout, k, s)


STEM_HEAD_CASES = [
    # dedicated stem (Cin = 1) / head (Cout <= 4) kernels of stemhead.cu, reached with impl = auto
    (2, 2, (16, 24), 1, 16, 3, 1),
    (2, 1, (9, 20), 1, 8, 3, 1),
    (2, 3, (7, 36), 1, 16, 3, 1),        # rows not a multiple of the warp tile
    (3, 1, (6, 5, 12), 1, 16, 3, 1),
    (3, 2, (4, 4, 8), 1, 8, 3, 1),
    (2, 2, (9, 14), 16, 3, 3, 1),
    (2, 1, (12, 11), 8, 3, 3, 1),
    (2, 1, (8, 8), 16, 2, 3, 1),
    (2, 1, (8, 8), 16, 4, 3, 1),
    (3, 1, (5, 6, 7), 16, 3, 3, 1),
    (3, 2, (4, 6, 6), 8, 3, 3, 1),
]


@pytest.mark.parametrize("dt", [F32, BF16, F16])
@pytest.mark.parametrize("nd,n,sp,cin,cout,k,s", STEM_HEAD_CASES)
def test_stem_head_dedicated(lib, dt, nd, n, sp, cin, cout, k, s):
    _conv_case(lib, dt, 0, nd, n, sp, cin, cout, k, s)


def _conv_case(lib, dt, impl, nd, n, sp, cin, cout, k, s):
    rng = np.random.default_rng(hash((nd, n, sp, cin, cout, k, s)) % 2**32)
    x = rng.standard_normal((n,) + sp + (cin,)).astype(np.float32)
    w = (rng.standard_normal((k,) * nd + (cin, cout)) / np.sqrt(k**nd * cin)).astype(np.float32)
    if dt != F32:   # feed the oracle the same rounded activations
        x = host(dev(x, dt))
    d = _capi.conv_desc_same(nd, n, sp, cin, cout, k, s)
    osp = (d.yd, d.yh, d.yw)[3 - nd:]
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    yo = O.conv_same(xt.permute(0, nd + 1, *range(1, nd + 1)), wt, s)
    g = rng.standard_normal(tuple(yo.shape)).astype(np.float32)          # channel first
    g_cl = np.ascontiguousarray(np.moveaxis(g, 1, -1))
    if dt != F32:
        g_cl = host(dev(g_cl, gdt(dt)))
    dxo, dwo = torch.autograd.grad(yo, [xt, wt], torch.tensor(np.moveaxis(g_cl, -1, 1), dtype=torch.float64))
    # forward
    xd, wd = dev(x, dt), dev(w)
    y = torch.empty((n,) + osp + (cout,), dtype=torch.float32, device=DEV)
    _capi.check(lib.urn_conv_fwd(C.byref(d), dt, impl, xd.data_ptr(), cin, wd.data_ptr(), y.data_ptr(), cout, F32, st()))
    assert relerr(host(y), cl(yo.detach())) < 2e-5
    # dgrad (write, then accumulate)
    gd = dev(g_cl, gdt(dt))
    dx = torch.full((n,) + sp + (cin,), 7.0, dtype=TORCH_DT[gdt(dt)], device=DEV)
    _capi.check(lib.urn_conv_dgrad(C.byref(d), gdt(dt), impl, gd.data_ptr(), cout, wd.data_ptr(), dx.data_ptr(), cin, gdt(dt), 0, st()))
    assert relerr(host(dx), dxo.numpy()) < TOL[dt]
    _capi.check(lib.urn_conv_dgrad(C.byref(d), gdt(dt), impl, gd.data_ptr(), cout, wd.data_ptr(), dx.data_ptr(), cin, gdt(dt), 1, st()))
    assert relerr(host(dx), 2 * dxo.numpy()) < 2 * TOL[dt]
    # wgrad
    wsb = lib.urn_conv_wgrad_workspace(C.byref(d))
    ws = torch.empty(max(int(wsb), 16), dtype=torch.uint8, device=DEV)
    dw = torch.full(w.shape, 3.0, dtype=torch.float32, device=DEV)
    _capi.check(lib.urn_conv_wgrad(C.byref(d), dt, impl, xd.data_ptr(), cin, gd.data_ptr(), cout, gdt(dt), dw.data_ptr(), 0, ws.data_ptr(), st()))
    assert relerr(host(dw), dwo.numpy()) < 2e-5
    _capi.check(lib.urn_conv_wgrad(C.byref(d), dt, impl, xd.data_ptr(), cin, gd.data_ptr(), cout, gdt(dt), dw.data_ptr(), 1, ws.data_ptr(), st()))
    assert relerr(host(dw), 2 * dwo.numpy()) < 4e-5


TC_CASES = [
    # nd, n, spatial, cin, cout, k, s
    (2, 2, (32, 48), 16, 16, 3, 1),      # kc=16 (SWIZZLE_32B)
    (2, 2, (24, 40), 32, 32, 3, 1),      # kc=32 (SWIZZLE_64B)
    (2, 1, (20, 24), 64, 64, 3, 1),      # kc=64 (SWIZZLE_128B)
    (2, 1, (16, 16), 128, 256, 3, 1),    # two K chunks, N=256
    (2, 1, (8, 8), 256, 512, 3, 1),      # two N tiles
    (2, 2, (32, 32), 16, 32, 3, 2),      # stride 2 via TMA elementStrides, asymmetric SAME pad
    (2, 1, (17, 23), 32, 48, 3, 2),      # odd sizes (pad_before = 1), ragged tiles
    (2, 2, (16, 16), 32, 64, 1, 2),      # 1x1 shortcut s2
    (2, 1, (16, 24), 64, 32, 1, 1),
    (3, 1, (12, 16, 20), 16, 16, 3, 1),  # 27 taps
    (3, 1, (8, 8, 16), 32, 64, 3, 2),
    (3, 1, (6, 6, 6), 64, 64, 3, 1),     # small grid: tile box larger than the grid
    (3, 1, (8, 8, 8), 32, 64, 1, 2),
    (2, 1, (17, 23), 16, 32, 3, 2),      # odd sizes, stride 2, Cin=16 (halo kernels: pad_before = 1)
    (2, 2, (40, 72), 32, 16, 1, 1),      # 1x1 stride 1, several super-tiles
    (2, 1, (34, 70), 16, 32, 1, 2),      # 1x1 stride 2
    (3, 1, (10, 12, 16), 16, 32, 3, 2),  # 27 taps, stride 2
    (3, 1, (7, 9, 11), 16, 16, 3, 2),    # odd 3D sizes, stride 2
]


@pytest.mark.parametrize("dt", [F16, BF16])
@pytest.mark.parametrize("nd,n,sp,cin,cout,k,s", TC_CASES)
def test_conv_tcgen05_fwd_dgrad(lib, dt, nd, n, sp, cin, cout, k, s):
    """The tcgen05 implicit-GEMM path against the oracle on bit-identical (pre-rounded) operands:
    fp32 outputs must agree to accumulation-order noise."""
    rng = np.random.default_rng(hash((nd, n, sp, cin, cout, k, s, 1)) % 2**32)
    x = host(dev(rng.standard_normal((n,) + sp + (cin,)).astype(np.float32), dt))
    w = (rng.standard_normal((k,) * nd + (cin, cout)) / np.sqrt(k**nd * cin)).astype(np.float32)
    wq = host(dev(w, dt))                                     # operand rounding of the filter
    d = _capi.conv_desc_same(nd, n, sp, cin, cout, k, s)
    osp = (d.yd, d.yh, d.yw)[3 - nd:]
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    yo = O.conv_same(xt.permute(0, nd + 1, *range(1, nd + 1)), torch.tensor(wq, dtype=torch.float64), s)
    xd, wd = dev(x, dt), dev(w)
    wpack = torch.empty(2 * w.size + 64, dtype=torch.uint8, device=DEV)
    y = torch.full((n,) + osp + (cout,), 5.0, dtype=torch.float32, device=DEV)
    _capi.check(lib.urn_conv_fwd_tc(C.byref(d), dt, xd.data_ptr(), cin, wd.data_ptr(), wpack.data_ptr(), y.data_ptr(), cout, F32, 0, st()))
    assert relerr(host(y), cl(yo.detach())) < 2e-5
    # input gradient: gradients are bf16 operands in both 16-bit modes
    g_cl = host(dev(rng.standard_normal((n,) + osp + (cout,)).astype(np.float32), BF16))
    wqb = host(dev(w, BF16))
    yb = O.conv_same(xt.permute(0, nd + 1, *range(1, nd + 1)), torch.tensor(wqb, dtype=torch.float64), s)
    (dxo,) = torch.autograd.grad(yb, xt, torch.tensor(np.moveaxis(g_cl, -1, 1), dtype=torch.float64))
    gd = dev(g_cl, BF16)
    dx = torch.full((n,) + sp + (cin,), 3.0, dtype=torch.float32, device=DEV)
    _capi.check(lib.urn_conv_dgrad_tc(C.byref(d), BF16, gd.data_ptr(), cout, wd.data_ptr(), wpack.data_ptr(), dx.data_ptr(), cin, F32, 0, st()))
    assert relerr(host(dx), dxo.numpy()) < 2e-5
    _capi.check(lib.urn_conv_dgrad_tc(C.byref(d), BF16, gd.data_ptr(), cout, wd.data_ptr(), wpack.data_ptr(), dx.data_ptr(), cin, F32, 1, st()))
    assert relerr(host(dx), 2 * dxo.numpy()) < 4e-5
    dxb = torch.zeros((n,) + sp + (cin,), dtype=torch.bfloat16, device=DEV)
    _capi.check(lib.urn_conv_dgrad_tc(C.byref(d), BF16, gd.data_ptr(), cout, wd.data_ptr(), wpack.data_ptr(), dxb.data_ptr(), cin, BF16, 0, st()))
    assert relerr(host(dxb), dxo.numpy()) < 1e-2


@pytest.mark.parametrize("xdt,ydt", [(BF16, BF16), (F16, F16)])   # kind::f16 needs one format for both operands
@pytest.mark.parametrize("nd,n,sp,cin,cout,k,s", TC_CASES)
def test_conv_tcgen05_wgrad(lib, xdt, ydt, nd, n, sp, cin, cout, k, s, tol=2e-5):
    """Filter gradient on tcgen05 (MN-major operands, split-K, deterministic reduce)."""
    rng = np.random.default_rng(hash((nd, n, sp, cin, cout, k, s, 2)) % 2**32)
    x = host(dev(rng.standard_normal((n,) + sp + (cin,)).astype(np.float32), xdt))
    d = _capi.conv_desc_same(nd, n, sp, cin, cout, k, s)
    osp = (d.yd, d.yh, d.yw)[3 - nd:]
    g_cl = host(dev(rng.standard_normal((n,) + osp + (cout,)).astype(np.float32), ydt))
    wt = torch.zeros((k,) * nd + (cin, cout), dtype=torch.float64, requires_grad=True)
    yo = O.conv_same(cf(x), wt, s)
    (dwo,) = torch.autograd.grad(yo, wt, torch.tensor(np.moveaxis(g_cl, -1, 1), dtype=torch.float64))
    xd, gd = dev(x, xdt), dev(g_cl, ydt)
    ws = torch.empty(max(int(lib.urn_conv_wgrad_workspace(C.byref(d))), 16), dtype=torch.uint8, device=DEV)
    dw = torch.full(tuple(wt.shape), 2.0, dtype=torch.float32, device=DEV)
    _capi.check(lib.urn_conv_wgrad(C.byref(d), xdt, 2, xd.data_ptr(), cin, gd.data_ptr(), cout, ydt, dw.data_ptr(), 0, ws.data_ptr(), st()))
    assert relerr(host(dw), dwo.numpy()) < tol
    _capi.check(lib.urn_conv_wgrad(C.byref(d), xdt, 2, xd.data_ptr(), cin, gd.data_ptr(), cout, ydt, dw.data_ptr(), 1, ws.data_ptr(), st()))
    assert relerr(host(dw), 2 * dwo.numpy()) < 2 * tol


GUARD_CASES = [c for c in TC_CASES if c[3] <= 64] + [(2, 1, (16, 16), 128, 128, 3, 1)]


@pytest.mark.parametrize("nd,n,sp,cin,cout,k,s", GUARD_CASES)
def test_conv_kernels_stay_inside_their_buffers(lib, nd, n, sp, cin, cout, k, s):
    """No sanitizer on the GPU pool, so: operands sit in the middle of NaN-filled arenas and results
    in the middle of sentinel-filled ones.  Any read outside the operand that reaches the arithmetic
    turns the result NaN / changes bits against the run on tight buffers; any write outside the
    result damages the sentinels.  Covers forward, input gradient (store + accumulate, fp32 + bf16)
    and filter gradient of the halo-tile (cp.async address arithmetic by hand) and tap-box kernels."""
    G = 8192                                                   # guard elements on each side
    rng = np.random.default_rng(hash((nd, n, sp, cin, cout, k, s, 9)) % 2**32)
    d = _capi.conv_desc_same(nd, n, sp, cin, cout, k, s)
    osp = (d.yd, d.yh, d.yw)[3 - nd:]
    xs, ys = (n,) + sp + (cin,), (n,) + osp + (cout,)
    w = dev((rng.standard_normal((k,) * nd + (cin, cout)) / np.sqrt(k**nd * cin)).astype(np.float32))
    wpack = torch.empty(2 * w.numel() + 64, dtype=torch.uint8, device=DEV)

    def arena(shape, tdt, fill, value=None):
        a = torch.full((int(np.prod(shape)) + 2 * G,), fill, dtype=tdt, device=DEV)
        mid = a[G:G + int(np.prod(shape))].view(shape)
        if value is not None:
            mid.copy_(value)
        return a, mid

    def intact(a, fill):
        g = torch.cat([a[:G], a[-G:]]).float()
        return bool(torch.isnan(g).all()) if fill != fill else bool((g == fill).all())

    nan = float("nan")
    for dt in (F16, BF16):
        tdt = TORCH_DT[dt]
        x = dev(rng.standard_normal(xs).astype(np.float32), dt)
        g = dev(rng.standard_normal(ys).astype(np.float32), BF16)
        gx = dev(rng.standard_normal(ys).astype(np.float32), dt)    # filter gradient: one format for both
        xa, xm = arena(xs, tdt, nan, x)
        ga, gm = arena(ys, torch.bfloat16, nan, g)
        gxa, gxm = arena(ys, tdt, nan, gx)
        # forward
        y0 = torch.zeros(ys, dtype=torch.float32, device=DEV)
        _capi.check(lib.urn_conv_fwd_tc(C.byref(d), dt, x.data_ptr(), cin, w.data_ptr(), wpack.data_ptr(), y0.data_ptr(), cout, F32, 0, st()))
        ya, ym = arena(ys, torch.float32, 7.25)
        _capi.check(lib.urn_conv_fwd_tc(C.byref(d), dt, xm.data_ptr(), cin, w.data_ptr(), wpack.data_ptr(), ym.data_ptr(), cout, F32, 0, st()))
        assert torch.equal(ym, y0) and not torch.isnan(ym).any() and intact(ya, 7.25)
        # input gradient, fp32 store then accumulate, bf16 store
        for odt, acc in ((F32, 0), (F32, 1), (BF16, 0), (BF16, 1)):
            todt = TORCH_DT[odt]
            d0 = torch.ones(xs, dtype=todt, device=DEV)
            _capi.check(lib.urn_conv_dgrad_tc(C.byref(d), BF16, g.data_ptr(), cout, w.data_ptr(), wpack.data_ptr(), d0.data_ptr(), cin, odt, acc, st()))
            da, dm = arena(xs, todt, 7.25, torch.ones(xs, dtype=todt, device=DEV))
            _capi.check(lib.urn_conv_dgrad_tc(C.byref(d), BF16, gm.data_ptr(), cout, w.data_ptr(), wpack.data_ptr(), dm.data_ptr(), cin, odt, acc, st()))
            assert torch.equal(dm, d0) and not torch.isnan(dm.float()).any() and intact(da, 7.25), (dt, odt, acc)
        # filter gradient
        ws = torch.empty(max(int(lib.urn_conv_wgrad_workspace(C.byref(d))), 16), dtype=torch.uint8, device=DEV)
        w0 = torch.zeros(tuple(w.shape), dtype=torch.float32, device=DEV)
        _capi.check(lib.urn_conv_wgrad(C.byref(d), dt, 2, x.data_ptr(), cin, gx.data_ptr(), cout, dt, w0.data_ptr(), 0, ws.data_ptr(), st()))
        wa, wm = arena(tuple(w.shape), torch.float32, 7.25)
        _capi.check(lib.urn_conv_wgrad(C.byref(d), dt, 2, xm.data_ptr(), cin, gxm.data_ptr(), cout, dt, wm.data_ptr(), 0, ws.data_ptr(), st()))
        assert torch.equal(wm, w0) and not torch.isnan(wm).any() and intact(wa, 7.25)
        assert intact(xa, nan) and intact(ga, nan) and intact(gxa, nan)
    torch.cuda.synchronize()


MANY_TILE_CASES = [TC_CASES[i] for i in (0, 1, 2, 3, 5, 6, 7, 9, 10, 14, 16)] + [(2, 3, (64, 96), 16, 16, 3, 1), (2, 2, (48, 160), 32, 32, 3, 1)]


@pytest.mark.parametrize("ctas", [1, 3])
@pytest.mark.parametrize("nd,n,sp,cin,cout,k,s", MANY_TILE_CASES)
def test_conv_many_tiles_per_cta(lib, monkeypatch, ctas, nd, n, sp, cin, cout, k, s):
    """The persistent kernels capped to 1 / 3 CTAs (URN_MAX_CTAS): every CTA walks many tiles, so the
    ring wrap-arounds, accumulator recycling, look-ahead barrier waits and partial merges that the
    full-size network relies on are checked against the oracle on small shapes too."""
    monkeypatch.setenv("URN_MAX_CTAS", str(ctas))
    test_conv_tcgen05_fwd_dgrad(lib, F16, nd, n, sp, cin, cout, k, s)
    # one CTA sums every position in a single fp32 chain instead of <= 148 partials: slightly more rounding
    test_conv_tcgen05_wgrad(lib, BF16, BF16, nd, n, sp, cin, cout, k, s, tol=6e-5)
    if k == 3 and s == 1 or (nd, cin, cout, k, s) == (2, 32, 16, 1, 1):
        test_conv_fused_bn_statistics(lib, nd, n, sp, cin, cout, k, s)


@pytest.mark.parametrize("rows", ["1", "0"])
@pytest.mark.parametrize("nd,n,sp,cin,cout,k,s", [(2, 12, (32, 48), 16, 16, 3, 1), (2, 18, (24, 40), 32, 32, 3, 1), (2, 16, (40, 72), 16, 32, 3, 1),
                                                  (2, 33, (20, 24), 32, 16, 3, 1)])
def test_conv_images_as_rows_plan(lib, monkeypatch, rows, nd, n, sp, cin, cout, k, s):
    """2D stride-1 3x3 layers of a batch of >= 12 images run through the 3D concat path with relabelled
    axes (halo_plan_rows: planes := image rows, tile rows := images; default on, URN_HALO_ROWS=0 keeps
    the 16x8 tiles).  Same oracle checks for both plans, with batches below, at and above one 16-image
    tile row group."""
    monkeypatch.setenv("URN_HALO_ROWS", rows)
    test_conv_tcgen05_fwd_dgrad(lib, F16, nd, n, sp, cin, cout, k, s)
    test_conv_tcgen05_fwd_dgrad(lib, BF16, nd, n, sp, cin, cout, k, s)
    test_conv_fused_bn_statistics(lib, nd, n, sp, cin, cout, k, s)
    test_conv_kernels_stay_inside_their_buffers(lib, nd, n, sp, cin, cout, k, s)


ACC_CASES = [
    # nd, n, spatial, cin (gradient written), cout (gradient read), k   -- stride 1
    (2, 2, (33, 19), 16, 16, 3),          # ragged tiles
    (2, 2, (40, 72), 32, 16, 1),          # 1x1 shortcut into a 32-channel gradient (64-byte rows)
    (2, 13, (24, 40), 16, 16, 3),         # batch >= 12: images-as-rows plan, images not a multiple of the 16-row tile
    (2, 1, (16, 24), 64, 32, 1),          # 128-byte rows
    (3, 1, (9, 20, 12), 16, 16, 3),       # 27 taps, depth tile with a ragged last tile
    (3, 1, (6, 16, 16), 32, 16, 1),
]


@pytest.mark.parametrize("red,tma", [("1", "2"), ("0", "2"), ("1", "1"), ("1", "0")])
@pytest.mark.parametrize("nd,n,sp,cin,cout,k", ACC_CASES)
def test_conv_dgrad_bf16_accumulate_variants(lib, monkeypatch, red, tma, nd, n, sp, cin, cout, k):
    """Accumulating bf16 input gradient (dX += dgrad) into a CHANNEL SLICE of a wider gradient buffer (the
    concat / skip layout): through the TMA reduce unit (URN_HALO_RED=1, default) or the read-modify-write
    epilogue (=0), with the input halo staged as one TMA box per plane (URN_HALO_TMA=2, default), one box
    per kw shift (=1) or by cp.async (=0).  Expected: old + oracle gradient to bf16 storage precision; the
    other half of the buffer must not be touched."""
    monkeypatch.setenv("URN_HALO_RED", red)
    monkeypatch.setenv("URN_HALO_TMA", tma)
    rng = np.random.default_rng(hash((nd, n, sp, cin, cout, k, 77)) % 2**32)
    w = (rng.standard_normal((k,) * nd + (cin, cout)) / np.sqrt(k**nd * cin)).astype(np.float32)
    wqb = host(dev(w, BF16))
    d = _capi.conv_desc_same(nd, n, sp, cin, cout, k, 1)
    g_cl = host(dev(rng.standard_normal((n,) + sp + (cout,)).astype(np.float32), BF16))
    xt = torch.zeros((n, cin) + sp, dtype=torch.float64, requires_grad=True)
    yb = O.conv_same(xt, torch.tensor(wqb, dtype=torch.float64), 1)
    (dxo,) = torch.autograd.grad(yb, xt, torch.tensor(np.moveaxis(g_cl, -1, 1), dtype=torch.float64))
    dxo = np.moveaxis(dxo.numpy(), 1, -1)
    old = host(dev(rng.standard_normal((n,) + sp + (2 * cin,)).astype(np.float32), BF16))
    buf = dev(old, BF16)
    gd, wd = dev(g_cl, BF16), dev(w)
    wpack = torch.empty(2 * w.size + 64, dtype=torch.uint8, device=DEV)
    for off in (cin, 0):                                      # second, then first half of the channels
        _capi.check(lib.urn_conv_dgrad_tc(C.byref(d), BF16, gd.data_ptr(), cout, wd.data_ptr(), wpack.data_ptr(),
                                          buf.data_ptr() + 2 * off, 2 * cin, BF16, 1, st()))
        got = host(buf)
        want = old.copy()
        want[..., off:off + cin] += dxo
        assert relerr(got[..., off:off + cin], want[..., off:off + cin]) < 1e-2
        other = slice(0, cin) if off else slice(cin, 2 * cin)
        assert np.array_equal(got[..., other], old[..., other])
        old = got


HALO_WGRAD_CASES = [
    # nd, n, spatial, cin, cout, ldx factor  (stride-1 3x3(x3): the halo-tile filter-gradient kernel)
    (2, 3, (40, 24), 16, 16, 1),         # several tiles per image, ragged in h
    (2, 2, (33, 19), 16, 32, 1),         # ragged in both dims
    (2, 1, (32, 16), 32, 16, 2),         # concat slice as X (pitch 2C)
    (2, 2, (16, 24), 32, 32, 1),
    (2, 1, (20, 12), 64, 32, 1),
    (2, 1, (16, 8), 64, 64, 1),
    (2, 1, (18, 10), 16, 64, 1),
    (3, 1, (5, 20, 12), 16, 16, 1),
    (3, 2, (3, 16, 8), 32, 32, 1),
    (3, 1, (1, 16, 8), 16, 16, 1),       # depth 1: only the central depth tap sees data
    (3, 1, (52, 16, 8), 16, 16, 1),      # two depth chunks
    (3, 1, (4, 9, 9), 64, 32, 1),
    (3, 1, (6, 16, 16), 32, 16, 2),
]


@pytest.mark.parametrize("xdt", [BF16, F16])
@pytest.mark.parametrize("nd,n,sp,cin,cout,ldf", HALO_WGRAD_CASES)
def test_conv_halo_wgrad(lib, xdt, nd, n, sp, cin, cout, ldf):
    """Halo-tile tcgen05 filter gradient (wgrad_halo.cu) against the oracle on pre-rounded operands."""
    rng = np.random.default_rng(hash((nd, n, sp, cin, cout, ldf, 3)) % 2**32)
    z = host(dev(rng.standard_normal((n,) + sp + (ldf * cin,)).astype(np.float32), xdt))
    off = (ldf - 1) * cin
    x = np.ascontiguousarray(z[..., off:])
    d = _capi.conv_desc_same(nd, n, sp, cin, cout, 3, 1)
    g_cl = host(dev(rng.standard_normal((n,) + sp + (cout,)).astype(np.float32), xdt))
    wt = torch.zeros((3,) * nd + (cin, cout), dtype=torch.float64, requires_grad=True)
    yo = O.conv_same(cf(x), wt, 1)
    (dwo,) = torch.autograd.grad(yo, wt, torch.tensor(np.moveaxis(g_cl, -1, 1), dtype=torch.float64))
    zd, gd = dev(z, xdt), dev(g_cl, xdt)
    ws = torch.empty(max(int(lib.urn_conv_wgrad_workspace(C.byref(d))), 16), dtype=torch.uint8, device=DEV)
    dw = torch.full(tuple(wt.shape), 2.0, dtype=torch.float32, device=DEV)
    _capi.check(lib.urn_conv_wgrad(C.byref(d), xdt, 2, zd.data_ptr() + 2 * off, ldf * cin, gd.data_ptr(), cout, xdt, dw.data_ptr(), 0, ws.data_ptr(), st()))
    assert relerr(host(dw), dwo.numpy()) < 2e-5
    _capi.check(lib.urn_conv_wgrad(C.byref(d), xdt, 2, zd.data_ptr() + 2 * off, ldf * cin, gd.data_ptr(), cout, xdt, dw.data_ptr(), 1, ws.data_ptr(), st()))
    assert relerr(host(dw), 2 * dwo.numpy()) < 4e-5
    torch.cuda.synchronize()


@pytest.mark.parametrize("rows", ["1", "0"])
@pytest.mark.parametrize("n,sp,cin,cout,ldf", [(12, (20, 24), 16, 16, 1), (33, (12, 16), 32, 16, 2), (16, (24, 40), 16, 32, 1),
                                               (20, (9, 11), 64, 32, 1), (17, (70, 40), 16, 16, 1), (13, (5, 8), 32, 32, 1)])
def test_conv_halo_wgrad_images_as_rows(lib, monkeypatch, rows, n, sp, cin, cout, ldf):
    """Filter gradient of 2D 3x3 stride-1 layers with >= 12 images: the images-as-rows plan (planes :=
    image rows, the three kh taps concatenated along N; default) and the 16x8-tile plan
    (URN_WHALO_ROWS=0) against the oracle: ragged sizes, pitched X, several depth chunks, batches that
    are not multiples of the 16-image tile."""
    monkeypatch.setenv("URN_WHALO_ROWS", rows)
    for xdt in (BF16, F16):
        test_conv_halo_wgrad(lib, xdt, 2, n, sp, cin, cout, ldf)


@pytest.mark.parametrize("nd,n,sp,cin,cout,k,s", [
    (2, 3, (40, 72), 16, 16, 3, 1), (2, 2, (33, 19), 16, 32, 3, 1), (2, 1, (64, 64), 32, 64, 3, 1),
    (2, 2, (32, 48), 16, 32, 3, 2), (2, 2, (16, 40), 32, 16, 1, 1), (3, 1, (9, 20, 12), 16, 16, 3, 1),
    (3, 1, (8, 16, 16), 16, 32, 3, 2),
    (2, 2, (16, 24), 128, 128, 3, 1), (2, 1, (12, 20), 128, 256, 3, 2), (2, 3, (8, 8), 256, 128, 1, 1),
    (3, 1, (4, 8, 8), 128, 64, 3, 1)])                 # tap-box kernel (Cin >= 128)
def test_conv_fused_bn_statistics(lib, nd, n, sp, cin, cout, k, s):
    """BN statistics fused into the conv epilogue (per-warp butterfly + Chan merges) against
    float64 mean / biased variance of the kernel's own fp32 output."""
    rng = np.random.default_rng(hash((nd, n, sp, cin, cout, k, s, 9)) % 2**32)
    x = host(dev(rng.standard_normal((n,) + sp + (cin,)).astype(np.float32) + 0.7, F16))
    w = (rng.standard_normal((k,) * nd + (cin, cout)) / np.sqrt(k**nd * cin)).astype(np.float32)
    w[..., 0] *= 1e-3                       # a nearly constant channel: variance << mean^2 + eps matters
    d = _capi.conv_desc_same(nd, n, sp, cin, cout, k, s)
    osp = (d.yd, d.yh, d.yw)[3 - nd:]
    xd, wd = dev(x, F16), dev(w)
    wpack = torch.empty(2 * w.size + 64, dtype=torch.uint8, device=DEV)
    y = torch.empty((n,) + osp + (cout,), dtype=torch.float32, device=DEV)
    stats = torch.empty(2 * cout, dtype=torch.float32, device=DEV)
    ws = torch.empty(int(lib.urn_bn_workspace(cout)), dtype=torch.uint8, device=DEV)
    _capi.check(lib.urn_conv_fwd_tc_stats(C.byref(d), F16, xd.data_ptr(), cin, wd.data_ptr(), wpack.data_ptr(),
                                          y.data_ptr(), cout, 1e-3, stats.data_ptr(), ws.data_ptr(), st()))
    yh = host(y).reshape(-1, cout).astype(np.float64)
    mu, var = yh.mean(0), yh.var(0)
    np.testing.assert_allclose(host(stats)[:cout], mu, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(host(stats)[cout:], 1 / np.sqrt(var + 1e-3), rtol=2e-5)
    # and the separate-pass kernel agrees
    stats2 = torch.empty(2 * cout, dtype=torch.float32, device=DEV)
    _capi.check(lib.urn_bn_stats(y.data_ptr(), cout, yh.shape[0], cout, 1e-3, stats2.data_ptr(), ws.data_ptr(), st()))
    np.testing.assert_allclose(host(stats), host(stats2), rtol=2e-5, atol=1e-6)


def test_conv_tcgen05_pitched_concat_view(lib):
    """TMA reads a channel slice of a concat buffer (pitch 2C, offset C)."""
    rng = np.random.default_rng(5)
    n, sp, c = 1, (16, 16), 32
    z = host(dev(rng.standard_normal((n,) + sp + (2 * c,)).astype(np.float32), F16))
    w = rng.standard_normal((3, 3, c, c)).astype(np.float32) / 17
    d = _capi.conv_desc_same(2, n, sp, c, c, 3, 1)
    zd, wd = dev(z, F16), dev(w)
    wpack = torch.empty(2 * w.size + 64, dtype=torch.uint8, device=DEV)
    y = torch.empty((n,) + sp + (c,), dtype=torch.float32, device=DEV)
    _capi.check(lib.urn_conv_fwd_tc(C.byref(d), F16, zd.data_ptr() + 2 * c, 2 * c, wd.data_ptr(), wpack.data_ptr(), y.data_ptr(), c, F32, 0, st()))
    yo = cl(O.conv_same(cf(z[..., c:]), torch.tensor(host(dev(w, F16)), dtype=torch.float64), 1))
    assert relerr(host(y), yo) < 2e-5


@pytest.mark.parametrize("nd,n,sp,cin,cout", [(2, 2, (8, 12), 32, 16), (3, 1, (4, 6, 4), 32, 16), (2, 1, (5, 7), 8, 4)])
def test_conv_transpose_forward(lib, nd, n, sp, cin, cout):
    """slim.conv*_transpose forward == urn_conv_dgrad on the deconv's base relation."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal((n,) + sp + (cin,)).astype(np.float32)
    w = rng.standard_normal((3,) * nd + (cout, cin)).astype(np.float32) / 10
    yo = cl(O.conv_transpose_same(cf(x), torch.tensor(w, dtype=torch.float64), 2))
    osp = tuple(2 * v for v in sp)
    d = _capi.conv_desc_same(nd, n, osp, cout, cin, 3, 2)
    y = torch.empty((n,) + osp + (cout,), dtype=torch.float32, device=DEV)
    xd, wd = dev(x), dev(w)          # keep the device tensors alive across the call
    _capi.check(lib.urn_conv_dgrad(C.byref(d), F32, 1, xd.data_ptr(), cin, wd.data_ptr(), y.data_ptr(), cout, F32, 0, st()))
    assert relerr(host(y), yo) < 2e-5


def test_conv_pitched_views(lib):
    """Channel slices of a concat buffer as conv input (ldx > cin) and dgrad output."""
    rng = np.random.default_rng(12)
    n, sp, cin, cout = 2, (8, 8), 16, 16
    z = rng.standard_normal((n,) + sp + (2 * cin,)).astype(np.float32)
    w = rng.standard_normal((3, 3, cin, cout)).astype(np.float32) / 12
    d = _capi.conv_desc_same(2, n, sp, cin, cout, 3, 1)
    zd, wd = dev(z), dev(w)
    y = torch.empty((n,) + sp + (cout,), dtype=torch.float32, device=DEV)
    _capi.check(lib.urn_conv_fwd(C.byref(d), F32, 1, zd.data_ptr() + 4 * cin, 2 * cin, wd.data_ptr(), y.data_ptr(), cout, F32, st()))
    yo = cl(O.conv_same(cf(z[..., cin:]), torch.tensor(w, dtype=torch.float64), 1))
    assert relerr(host(y), yo) < 2e-5


@pytest.mark.parametrize("c,rows", [(16, 5000), (3, 777), (512, 64), (32, 1)])
def test_bn_stats_apply(lib, c, rows):
    rng = np.random.default_rng(c)
    x = (rng.standard_normal((rows, c)) * rng.random(c) * 3 + rng.standard_normal(c)).astype(np.float32)
    beta = rng.standard_normal(c).astype(np.float32)
    xd, bd = dev(x), dev(beta)
    stats = torch.empty(2 * c, dtype=torch.float32, device=DEV)
    ws = torch.empty(int(lib.urn_bn_workspace(c)), dtype=torch.uint8, device=DEV)
    _capi.check(lib.urn_bn_stats(xd.data_ptr(), c, rows, c, 1e-3, stats.data_ptr(), ws.data_ptr(), st()))
    mu, var = x.astype(np.float64).mean(0), x.astype(np.float64).var(0)
    np.testing.assert_allclose(host(stats)[:c], mu, atol=1e-5)
    np.testing.assert_allclose(host(stats)[c:], 1 / np.sqrt(var + 1e-3), rtol=2e-5)
    ref = (x - mu) / np.sqrt(var + 1e-3) + beta
    for relu in (0, 1):
        out = torch.empty((rows, c), dtype=torch.float32, device=DEV)
        _capi.check(lib.urn_bn_apply(xd.data_ptr(), c, rows, c, stats.data_ptr(), bd.data_ptr(), None, 0, 0, None, 0,
                                     None, None, relu, out.data_ptr(), c, F32, None, st()))
        assert relerr(host(out), np.maximum(ref, 0) if relu else ref) < 2e-5
    # residual forms: identity tensor, and a second BN (conv shortcut)
    r = rng.standard_normal((rows, c)).astype(np.float32)
    out = torch.empty((rows, c), dtype=torch.float32, device=DEV)
    rd = dev(r)
    _capi.check(lib.urn_bn_apply(xd.data_ptr(), c, rows, c, stats.data_ptr(), bd.data_ptr(), rd.data_ptr(), c, F32,
                                 None, 0, None, None, 1, out.data_ptr(), c, F32, None, st()))
    assert relerr(host(out), np.maximum(ref + r, 0)) < 2e-5
    _capi.check(lib.urn_bn_apply(xd.data_ptr(), c, rows, c, stats.data_ptr(), bd.data_ptr(), None, 0, 0,
                                 xd.data_ptr(), c, stats.data_ptr(), bd.data_ptr(), 1, out.data_ptr(), c, F32, None, st()))
    assert relerr(host(out), np.maximum(2 * ref, 0)) < 2e-5


@pytest.mark.parametrize("dt", [F32, BF16, F16])
@pytest.mark.parametrize("c,rows,relu,mode", [(16, 3000, 1, "conv_sc"), (16, 3000, 1, "identity"), (3, 500, 0, "plain"),
                                              (64, 129, 1, "plain"), (256, 40, 0, "plain"),
                                              (16, 70001, 1, "plain"), (32, 5003, 0, "plain"), (8, 999, 1, "identity"),
                                              (512, 33, 1, "conv_sc"), (16, 3000, 1, "plain+beta"), (32, 4099, 0, "plain+beta"),
                                              (128, 300, 1, "plain+beta")])
def test_bn_backward(lib, dt, c, rows, relu, mode):
    with_beta = mode.endswith("+beta")      # urn_bn_bwd_beta: xhat from out / mask from raw
    mode = mode.split("+")[0]
    rng = np.random.default_rng(rows + c)
    x = rng.standard_normal((rows, c)).astype(np.float32) * 2 + 0.5
    x2 = rng.standard_normal((rows, c)).astype(np.float32)
    res = rng.standard_normal((rows, c)).astype(np.float32)
    beta = rng.standard_normal(c).astype(np.float32) * 0.3
    dout = rng.standard_normal((rows, c)).astype(np.float32)
    if dt != F32:
        dout, res = host(dev(dout, gdt(dt))), host(dev(res, dt))
    xt, x2t, rt = (torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in (x, x2, res))
    bt = torch.tensor(beta, dtype=torch.float64, requires_grad=True)

    def bn(v, b):
        mu = v.mean(0)
        var = ((v - mu) ** 2).mean(0)
        return (v - mu) * torch.rsqrt(var + 1e-3) + b
    y = bn(xt, bt)
    if mode == "conv_sc":
        y = y + bn(x2t, bt * 0)
    elif mode == "identity":
        y = y + rt
    out = torch.relu(y) if relu else y
    ins = [xt, bt] + ([x2t] if mode == "conv_sc" else []) + ([rt] if mode == "identity" else [])
    grads = torch.autograd.grad(out, ins, torch.tensor(dout, dtype=torch.float64))
    # device side
    def stats_of(v):
        vd = dev(v)
        s = torch.empty(2 * c, dtype=torch.float32, device=DEV)
        _capi.check(lib.urn_bn_stats(vd.data_ptr(), c, rows, c, 1e-3, s.data_ptr(), ws.data_ptr(), st()))
        return vd, s
    ws = torch.empty(int(lib.urn_bn_workspace(c)), dtype=torch.uint8, device=DEV)
    xd, s1 = stats_of(x)
    x2d, s2 = stats_of(x2)
    outd = dev(out.detach().numpy(), dt)
    if with_beta:     # the fast path recomputes mask / xhat from what the forward kernel stored
        bd = dev(beta)
        outd = torch.empty((rows, c), dtype=TORCH_DT[dt], device=DEV)
        _capi.check(lib.urn_bn_apply(xd.data_ptr(), c, rows, c, s1.data_ptr(), bd.data_ptr(), None, 0, 0, None, 0,
                                     None, None, relu, outd.data_ptr(), c, dt, None, st()))
    doutd = dev(dout, gdt(dt))
    tdt = TORCH_DT[gdt(dt)]
    draw = torch.empty((rows, c), dtype=tdt, device=DEV)
    draw2 = torch.empty((rows, c), dtype=tdt, device=DEV)
    dres = torch.ones((rows, c), dtype=tdt, device=DEV)
    dbeta = torch.zeros(c, dtype=torch.float32, device=DEV)
    dbeta2 = torch.zeros(c, dtype=torch.float32, device=DEV)
    fn = lib.urn_bn_bwd_beta if with_beta else lib.urn_bn_bwd
    extra = (bd.data_ptr(),) if with_beta else ()
    _capi.check(fn(
        doutd.data_ptr(), c, outd.data_ptr(), c, gdt(dt), dt, relu, rows, c, xd.data_ptr(), c, s1.data_ptr(), *extra, dbeta.data_ptr(),
        draw.data_ptr(), c,
        x2d.data_ptr() if mode == "conv_sc" else None, c, s2.data_ptr() if mode == "conv_sc" else None,
        dbeta2.data_ptr() if mode == "conv_sc" else None, draw2.data_ptr() if mode == "conv_sc" else None, c,
        dres.data_ptr() if mode == "identity" else None, c, 1, 0, ws.data_ptr(), st()))
    tol = TOL[dt]
    assert relerr(host(draw), grads[0].numpy()) < tol
    assert relerr(host(dbeta), grads[1].numpy()) < 1e-4
    if mode == "conv_sc":
        assert relerr(host(draw2), grads[2].numpy()) < tol
        assert relerr(host(dbeta2), grads[1].numpy()) < 1e-4
    if mode == "identity":
        assert relerr(host(dres), 1 + grads[2].numpy()) < tol   # accumulated onto the ones


@pytest.mark.parametrize("c,rows,relu,mode", [(16, 70001, 1, "plain"), (32, 4099, 0, "plain"), (16, 3000, 1, "conv_sc"),
                                              (64, 129, 1, "identity"), (512, 33, 1, "conv_sc"), (8, 999, 1, "identity")])
def test_bn_fp16_raw(lib, c, rows, relu, mode):
    """mix16 stores the raw conv output as fp16 (urn_bn_*_t with raw_dtype = URN_F16): statistics, apply
    (plain / identity residual / second BN) and backward against float64 on the same fp16 values."""
    rng = np.random.default_rng(rows * 7 + c)
    x = host(dev(rng.standard_normal((rows, c)).astype(np.float32) * 2 + 0.5, F16))
    x2 = host(dev(rng.standard_normal((rows, c)).astype(np.float32), F16))
    res = host(dev(rng.standard_normal((rows, c)).astype(np.float32), F16))
    beta = rng.standard_normal(c).astype(np.float32) * 0.3
    dout = host(dev(rng.standard_normal((rows, c)).astype(np.float32), BF16))
    xt, x2t, rt = (torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in (x, x2, res))
    bt = torch.tensor(beta, dtype=torch.float64, requires_grad=True)

    def bn(v, b):
        mu = v.mean(0)
        var = ((v - mu) ** 2).mean(0)
        return (v - mu) * torch.rsqrt(var + 1e-3) + b
    y = bn(xt, bt)
    if mode == "conv_sc":
        y = y + bn(x2t, bt * 0)
    elif mode == "identity":
        y = y + rt
    out = torch.relu(y) if relu else y
    ins = [xt, bt] + ([x2t] if mode == "conv_sc" else []) + ([rt] if mode == "identity" else [])
    grads = torch.autograd.grad(out, ins, torch.tensor(dout, dtype=torch.float64))
    ws = torch.empty(int(lib.urn_bn_workspace(c)), dtype=torch.uint8, device=DEV)

    def stats_of(v):
        vd = dev(v, F16)
        s = torch.empty(2 * c, dtype=torch.float32, device=DEV)
        _capi.check(lib.urn_bn_stats_t(vd.data_ptr(), F16, c, rows, c, 1e-3, s.data_ptr(), ws.data_ptr(), st()))
        return vd, s
    xd, s1 = stats_of(x)
    x2d, s2 = stats_of(x2)
    mu, var = x.astype(np.float64).mean(0), x.astype(np.float64).var(0)
    np.testing.assert_allclose(host(s1)[:c], mu, atol=1e-5)
    np.testing.assert_allclose(host(s1)[c:], 1 / np.sqrt(var + 1e-3), rtol=2e-5)
    bd, zd, rd = dev(beta), dev(0 * beta), dev(res, F16)
    outd = torch.empty((rows, c), dtype=torch.float16, device=DEV)
    sh = torch.empty((rows, c), dtype=torch.bfloat16, device=DEV)
    _capi.check(lib.urn_bn_apply_t(
        xd.data_ptr(), F16, c, rows, c, s1.data_ptr(), bd.data_ptr(), rd.data_ptr() if mode == "identity" else None, c, F16,
        x2d.data_ptr() if mode == "conv_sc" else None, c, s2.data_ptr() if mode == "conv_sc" else None,
        zd.data_ptr() if mode == "conv_sc" else None, relu, outd.data_ptr(), c, F16, sh.data_ptr(), st()))
    assert relerr(host(outd), out.detach().numpy()) < 1e-3 and relerr(host(sh), out.detach().numpy()) < 8e-3
    draw = torch.empty((rows, c), dtype=torch.bfloat16, device=DEV)
    draw2 = torch.empty((rows, c), dtype=torch.bfloat16, device=DEV)
    dres = torch.ones((rows, c), dtype=torch.bfloat16, device=DEV)
    dbeta = torch.zeros(c, dtype=torch.float32, device=DEV)
    dbeta2 = torch.zeros(c, dtype=torch.float32, device=DEV)
    doutd = dev(dout, BF16)
    for with_beta in ((False, True) if mode == "plain" else (False,)):
        _capi.check(lib.urn_bn_bwd_t(
            doutd.data_ptr(), c, outd.data_ptr(), c, BF16, F16, relu, rows, c, xd.data_ptr(), F16, c, s1.data_ptr(),
            bd.data_ptr() if with_beta else None, dbeta.data_ptr(), draw.data_ptr(), c,
            x2d.data_ptr() if mode == "conv_sc" else None, c, s2.data_ptr() if mode == "conv_sc" else None,
            dbeta2.data_ptr() if mode == "conv_sc" else None, draw2.data_ptr() if mode == "conv_sc" else None, c,
            dres.data_ptr() if mode == "identity" else None, c, 1, 0, ws.data_ptr(), st()))
        assert relerr(host(draw), grads[0].numpy()) < TOL[BF16], with_beta
        assert relerr(host(dbeta), grads[1].numpy()) < 1e-4
    if mode == "conv_sc":
        assert relerr(host(draw2), grads[2].numpy()) < TOL[BF16]
        assert relerr(host(dbeta2), grads[1].numpy()) < 1e-4
    if mode == "identity":
        assert relerr(host(dres), 1 + grads[2].numpy()) < TOL[BF16]
    # a shape outside the fast path must be refused, not mis-read
    bad = torch.zeros((8, 3), dtype=torch.float16, device=DEV)
    sb = torch.empty(6, dtype=torch.float32, device=DEV)
    assert lib.urn_bn_stats_t(bad.data_ptr(), F16, 3, 8, 3, 1e-3, sb.data_ptr(), ws.data_ptr(), st()) != 0


@pytest.mark.parametrize("nd,n,sp,cin,cout,k,s", [
    (2, 3, (40, 72), 16, 16, 3, 1), (2, 2, (32, 48), 16, 32, 3, 2), (2, 2, (16, 40), 32, 16, 1, 1),
    (3, 1, (9, 20, 12), 16, 16, 3, 1), (2, 2, (16, 24), 128, 128, 3, 1), (2, 3, (8, 8), 256, 128, 1, 1)])
def test_conv_fp16_raw_output(lib, nd, n, sp, cin, cout, k, s):
    """mix16 conv epilogues store the raw output as fp16 while the fused BN statistics are still those of
    the fp32 accumulators: identical statistics as the fp32-output launch, output = its rounding."""
    rng = np.random.default_rng(hash((nd, n, sp, cin, cout, k, s, 5)) % 2**32)
    x = host(dev(rng.standard_normal((n,) + sp + (cin,)).astype(np.float32) + 0.3, F16))
    w = (rng.standard_normal((k,) * nd + (cin, cout)) / np.sqrt(k**nd * cin)).astype(np.float32)
    d = _capi.conv_desc_same(nd, n, sp, cin, cout, k, s)
    osp = (d.yd, d.yh, d.yw)[3 - nd:]
    xd, wd = dev(x, F16), dev(w)
    wpack = torch.empty(2 * w.size + 64, dtype=torch.uint8, device=DEV)
    ws = torch.empty(int(lib.urn_bn_workspace(cout)), dtype=torch.uint8, device=DEV)
    res = {}
    for ydt in (F32, F16):
        y = torch.full((n,) + osp + (cout,), 7.0, dtype=TORCH_DT[ydt], device=DEV)
        stats = torch.empty(2 * cout, dtype=torch.float32, device=DEV)
        _capi.check(lib.urn_conv_fwd_tc_stats_t(C.byref(d), F16, xd.data_ptr(), cin, wd.data_ptr(), wpack.data_ptr(),
                                                y.data_ptr(), ydt, cout, 1e-3, stats.data_ptr(), ws.data_ptr(), st()))
        res[ydt] = (y, host(stats))
    assert np.array_equal(res[F32][1], res[F16][1])
    assert torch.equal(res[F32][0].to(torch.float16), res[F16][0])
    # and without statistics (plain forward entry point)
    y2 = torch.empty_like(res[F16][0])
    _capi.check(lib.urn_conv_fwd_tc(C.byref(d), F16, xd.data_ptr(), cin, wd.data_ptr(), wpack.data_ptr(), y2.data_ptr(), cout,
                                    F16, 0, st()))
    assert torch.equal(y2, res[F16][0])


@pytest.mark.parametrize("c,rows,relu,mode", [(16, 70001, 1, "plain"), (32, 5003, 0, "plain"), (64, 129, 1, "identity"),
                                              (128, 300, 1, "conv_sc"), (512, 33, 0, "plain"), (8, 999, 1, "identity"),
                                              (16, 4096, 1, "conv_sc"), (3, 500, 0, "plain")])
def test_bn_kernels_stay_inside_their_buffers(lib, c, rows, relu, mode):
    """Same guard-arena check as the conv kernels for the batch-norm family (statistics, apply with
    bf16 shadow, backward with and without beta): the fast kernels keep several packed 16-byte
    loads of consecutive rows in flight, so the last rows are where a stray access would happen."""
    G = 8192
    rng = np.random.default_rng(rows * 7 + c)
    nan = float("nan")

    def arena(tdt, fill, value=None, shape=(rows, c)):
        a = torch.full((int(np.prod(shape)) + 2 * G,), fill, dtype=tdt, device=DEV)
        mid = a[G:G + int(np.prod(shape))].view(shape)
        if value is not None:
            mid.copy_(value)
        return a, mid

    def intact(a, fill):
        g = torch.cat([a[:G], a[-G:]]).float()
        return bool(torch.isnan(g).all()) if fill != fill else bool((g == fill).all())

    ws = torch.empty(int(lib.urn_bn_workspace(c)), dtype=torch.uint8, device=DEV)
    raw, raw2 = (dev(rng.standard_normal((rows, c)).astype(np.float32) * 2 + 0.5) for _ in range(2))
    res = dev(rng.standard_normal((rows, c)).astype(np.float32), F16)
    beta = dev(rng.standard_normal(c).astype(np.float32) * 0.3)
    dout = dev(rng.standard_normal((rows, c)).astype(np.float32), BF16)
    ra, rm = arena(torch.float32, nan, raw)
    r2a, r2m = arena(torch.float32, nan, raw2)
    rsa, rsm = arena(torch.float16, nan, res)
    doa, dom = arena(torch.bfloat16, nan, dout)

    def run(raw, raw2, res, dout, mk):
        """mk(tdt) -> (arena or None, tensor) for every result."""
        s1 = torch.empty(2 * c, dtype=torch.float32, device=DEV)
        s2 = torch.empty(2 * c, dtype=torch.float32, device=DEV)
        _capi.check(lib.urn_bn_stats(raw.data_ptr(), c, rows, c, 1e-3, s1.data_ptr(), ws.data_ptr(), st()))
        _capi.check(lib.urn_bn_stats(raw2.data_ptr(), c, rows, c, 1e-3, s2.data_ptr(), ws.data_ptr(), st()))
        out, shadow = mk(torch.float16), mk(torch.bfloat16)
        _capi.check(lib.urn_bn_apply(
            raw.data_ptr(), c, rows, c, s1.data_ptr(), beta.data_ptr(),
            res.data_ptr() if mode == "identity" else None, c, F16,
            raw2.data_ptr() if mode == "conv_sc" else None, c, s2.data_ptr() if mode == "conv_sc" else None,
            beta.data_ptr() if mode == "conv_sc" else None, relu, out[1].data_ptr(), c, F16, shadow[1].data_ptr(), st()))
        results = [out, shadow]
        for with_beta in (0, 1):
            draw, draw2, dres = mk(torch.bfloat16), mk(torch.bfloat16), mk(torch.bfloat16)
            dres[1].fill_(1.0)
            dbeta = torch.zeros(2 * c, dtype=torch.float32, device=DEV)
            fn = lib.urn_bn_bwd_beta if with_beta else lib.urn_bn_bwd
            extra = (beta.data_ptr(),) if with_beta else ()
            _capi.check(fn(
                dout.data_ptr(), c, out[1].data_ptr(), c, BF16, F16, relu, rows, c, raw.data_ptr(), c, s1.data_ptr(), *extra,
                dbeta.data_ptr(), draw[1].data_ptr(), c,
                raw2.data_ptr() if mode == "conv_sc" else None, c, s2.data_ptr() if mode == "conv_sc" else None,
                dbeta[c:].data_ptr() if mode == "conv_sc" else None, draw2[1].data_ptr() if mode == "conv_sc" else None, c,
                dres[1].data_ptr() if mode == "identity" else None, c, 1, 0, ws.data_ptr(), st()))
            results += [draw] + ([draw2] if mode == "conv_sc" else []) + ([dres] if mode == "identity" else []) + [(None, dbeta)]
        return [s1, s2], results

    tight = run(raw, raw2, res, dout, lambda tdt: (None, torch.full((rows, c), 7.25, dtype=tdt, device=DEV)))
    guarded = run(rm, r2m, rsm, dom, lambda tdt: arena(tdt, 7.25))
    for a, b in zip(tight[0], guarded[0]):
        assert torch.equal(a, b) and not torch.isnan(a).any()
    for (_, a), (ar, b) in zip(tight[1], guarded[1]):
        assert torch.equal(a, b) and not torch.isnan(b.float()).any()
        assert ar is None or intact(ar, 7.25)
    assert all(intact(a, nan) for a in (ra, r2a, rsa, doa))
    torch.cuda.synchronize()


@pytest.mark.parametrize("B,P,c,use_w", [(2, 1000, 3, True), (3, 257, 3, False), (1, 64, 5, True)])
def test_ce_metrics(lib, B, P, c, use_w):
    rng = np.random.default_rng(P)
    logits = (rng.standard_normal((B, P, c)) * 3).astype(np.float32)
    logits[0, :5] = 1.25                                       # exact ties -> first index wins
    label = rng.integers(0, c, (B, P)).astype(np.float32)
    weight = rng.random((B, P)).astype(np.float32)
    data = (rng.random((B, P)) > 0.7).astype(np.float32) * 3
    r = O.loss_and_metrics(torch.tensor(logits, dtype=torch.float64).requires_grad_(True), torch.tensor(data)[..., None],
                           torch.tensor(label), torch.tensor(weight) if use_w else None)
    lt = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    r2 = O.loss_and_metrics(lt, torch.tensor(data)[..., None], torch.tensor(label), torch.tensor(weight) if use_w else None)
    (dl_ref,) = torch.autograd.grad(r2["loss"], lt)
    ws = torch.zeros(int(lib.urn_loss_workspace(B * P)), dtype=torch.uint8, device=DEV)
    res = torch.zeros(4, dtype=torch.float64, device=DEV)
    dl = torch.empty((B, P, c), dtype=torch.float32, device=DEV)
    sm = torch.empty((B, P, c), dtype=torch.float32, device=DEV)
    lgd, labd, wgd, dad = dev(logits), dev(label), dev(weight), dev(data)
    _capi.check(lib.urn_ce_fwd_bwd(lgd.data_ptr(), c, labd.data_ptr(), wgd.data_ptr() if use_w else None,
                                   dad.data_ptr(), B, P, c, dl.data_ptr(), c, F32, sm.data_ptr(), res.data_ptr(),
                                   ws.data_ptr(), st()))
    out = res.cpu().numpy()
    assert abs(out[0] - float(r["loss"])) <= 2e-6 * abs(float(r["loss"]))
    assert (int(out[1]), int(out[2]), int(out[3])) == (r["n_correct_all"], r["n_correct_nz"], r["n_nz"])
    assert relerr(host(dl), dl_ref.numpy()) < 1e-5
    assert relerr(host(sm), r["softmax"].detach().numpy()) < 1e-5
    # out-of-range label -> NaN loss
    label[0, 0] = c
    labd = dev(label)
    _capi.check(lib.urn_ce_fwd_bwd(lgd.data_ptr(), c, labd.data_ptr(), None, dad.data_ptr(), B, P, c,
                                   None, c, F32, None, res.data_ptr(), ws.data_ptr(), st()))
    assert np.isnan(res.cpu().numpy()[0])


@pytest.mark.parametrize("total,c,ld,dld", [(5000, 3, 3, 1), (777, 5, 8, 2), (0, 3, 3, 1)])
def test_ssnet_decision(lib, total, c, ld, dld):
    """lib/ssnet_trainval.py:269-287 on the device: equal, value for value, to the host expression
    applied to the softmax this library returns (ties -> track) and to the oracle's softmax away
    from ties."""
    rng = np.random.default_rng(total + c)
    logits = np.zeros((max(total, 1), ld), np.float32)
    logits[:, :c] = (rng.standard_normal((max(total, 1), c)) * 2).astype(np.float32)
    logits[:50, 2] = logits[:50, 1]                             # exact shower/track ties
    data = np.zeros((max(total, 1), dld), np.float32)
    data[:, 0] = rng.choice([0.0, 0.5, 1.0, 1.5, 30.0], size=max(total, 1)).astype(np.float32)
    lgd, dad = dev(logits), dev(data)
    dec = torch.full((max(total, 1),), -1.0, dtype=torch.float32, device=DEV)
    _capi.check(lib.urn_ssnet_decision(lgd.data_ptr(), ld, total, c, dad.data_ptr(), dld, 1.0, dec.data_ptr(), st()))
    if total == 0:
        assert float(dec[0]) == -1.0
        return
    # the library's own softmax of the same logits
    ws = torch.zeros(int(lib.urn_loss_workspace(total)), dtype=torch.uint8, device=DEV)
    res = torch.zeros(4, dtype=torch.float64, device=DEV)
    sm = torch.empty((total, c), dtype=torch.float32, device=DEV)
    lab = torch.zeros(total, dtype=torch.float32, device=DEV)
    _capi.check(lib.urn_ce_fwd_bwd(lgd.data_ptr(), ld, lab.data_ptr(), None, dad.data_ptr(), 1, total, c, None, c, F32,
                                   sm.data_ptr(), res.data_ptr(), ws.data_ptr(), st()))
    smh = host(sm)
    want = O.ssnet_decision(smh, data[:, 0])
    got = host(dec)
    assert np.array_equal(got, want)
    assert np.all(got[:50][data[:50, 0] > 1.0] == 2.0)
    ref_sm = torch.softmax(torch.tensor(logits[:, :c], dtype=torch.float64), -1).numpy()
    clear = np.abs(ref_sm[:, 1] - ref_sm[:, 2]) > 1e-6
    assert np.array_equal(got[clear], O.ssnet_decision(ref_sm, data[:, 0])[clear])
    assert lib.urn_ssnet_decision(lgd.data_ptr(), ld, total, 2, dad.data_ptr(), dld, 1.0, dec.data_ptr(), st()) != 0


def test_adam_matches_tf_formula(lib):
    rng = np.random.default_rng(0)
    n = 1003
    p, g = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    pd, gd, md, vd = dev(p), dev(g), dev(m), dev(v)
    for t in (1, 2, 3):
        O.adam_tf(p, g, m, v, t, lr=1e-3)
        _capi.check(lib.urn_adam_step(pd.data_ptr(), gd.data_ptr(), md.data_ptr(), vd.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, t, st()))
    np.testing.assert_allclose(host(pd), p, rtol=0, atol=1e-6)
    np.testing.assert_allclose(host(vd), v, rtol=1e-5, atol=1e-12)


def test_error_paths(lib):
    d = _capi.urn_conv_desc()
    assert lib.urn_conv_desc_same(C.byref(d), 4, 1, 1, 8, 8, 1, 1, 3, 1) != 0
    assert b"nd" in lib.urn_last_error()
    d = _capi.conv_desc_same(2, 1, (8, 8), 4, 4, 3, 1)
    assert lib.urn_conv_fwd(C.byref(d), F32, 1, None, 2, None, None, 4, F32, None) != 0   # pitch < channels
