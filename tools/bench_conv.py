"""Micro-benchmark of single conv kernels through the C ABI (GPU): time, TFLOP/s, GB/s."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uresnet_b200 import _capi  # noqa: E402

lib = _capi.load()
F32, BF16, F16 = 0, 1, 2


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def time_it(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def run(nd, n, sp, cin, cout, k, s, what="fwd"):
    d = _capi.conv_desc_same(nd, n, sp, cin, cout, k, s)
    osp = (d.yd, d.yh, d.yw)[3 - nd:]
    x = torch.randn((n,) + tuple(sp) + (cin,), device="cuda").half()
    w = torch.randn((k,) * nd + (cin, cout), device="cuda") * 0.05
    wpack = torch.empty(2 * w.numel() + 64, dtype=torch.uint8, device="cuda")
    y = torch.empty((n,) + tuple(osp) + (cout,), dtype=torch.float32, device="cuda")
    pos = y.numel() // cout
    flops = 2.0 * pos * (k ** nd) * cin * cout
    if what == "fwd":
        fn = lambda: _capi.check(lib.urn_conv_fwd_tc(C.byref(d), F16, x.data_ptr(), cin, w.data_ptr(), wpack.data_ptr(), y.data_ptr(), cout, F32, 0, st()))
        byts = x.numel() * 2 + y.numel() * 4
    else:
        g = torch.randn((n,) + tuple(osp) + (cout,), device="cuda").bfloat16()
        dx = torch.zeros((n,) + tuple(sp) + (cin,), dtype=torch.bfloat16, device="cuda")
        fn = lambda: _capi.check(lib.urn_conv_dgrad_tc(C.byref(d), BF16, g.data_ptr(), cout, w.data_ptr(), wpack.data_ptr(), dx.data_ptr(), cin, BF16, int(os.environ.get('ACC', '0')), st()))
        byts = g.numel() * 2 + dx.numel() * 2
    ms = time_it(fn)
    print("%s nd=%d n=%d sp=%s %d->%d k%d s%d : %.3f ms  %.1f TFLOP/s  %.0f GB/s (algorithmic)" % (
        what, nd, n, "x".join(map(str, sp)), cin, cout, k, s, ms, flops / ms / 1e9, byts / ms / 1e6), flush=True)


if __name__ == "__main__":
    cases = [(2, 32, (512, 512), 16, 16, 3, 1), (2, 32, (256, 256), 32, 32, 3, 1), (2, 32, (128, 128), 64, 64, 3, 1),
             (2, 32, (64, 64), 128, 128, 3, 1), (2, 32, (32, 32), 256, 256, 3, 1), (2, 32, (16, 16), 512, 512, 3, 1),
             (3, 2, (192, 192, 192), 16, 16, 3, 1), (3, 4, (96, 96, 96), 32, 32, 3, 1), (3, 4, (48, 48, 48), 64, 64, 3, 1),
             (3, 4, (24, 24, 24), 128, 128, 3, 1)]
    if os.environ.get("CASE"):        # e.g. CASE="2,32,(512,512),16,32,3,2"
        cases = [eval("(" + os.environ["CASE"] + ")")]
    sel = os.environ.get("CASES")
    if sel:
        cases = [cases[int(i)] for i in sel.split(",")]
    for c in cases:
        run(*c, what=os.environ.get("WHAT", "fwd"))
