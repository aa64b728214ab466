"""Micro-benchmark of the BN kernels through the C ABI (GPU): time and achieved GB/s."""
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


def run(rows, c):
    pad = int(os.environ.get("PAD_BYTES", "0"))
    if pad:
        _keep = [torch.empty(pad * (i + 1), dtype=torch.uint8, device="cuda") for i in range(4)]   # shift later allocations
    raw = torch.randn(rows, c, device="cuda")
    raw2 = torch.randn(rows, c, device="cuda")
    stats = torch.empty(2 * c, device="cuda")
    beta = torch.zeros(c, device="cuda")
    ws = torch.empty(int(lib.urn_bn_workspace(c)), dtype=torch.uint8, device="cuda")
    if pad:
        _k2 = torch.empty(pad * 3 + 512, dtype=torch.uint8, device="cuda")
    out = torch.empty(rows, c, device="cuda", dtype=torch.float16)
    if pad:
        _k3 = torch.empty(pad * 5 + 1024, dtype=torch.uint8, device="cuda")
    out2 = torch.empty(rows, c, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(rows, c, device="cuda").half()
    dout = torch.randn(rows, c, device="cuda").bfloat16()
    draw = torch.empty(rows, c, device="cuda", dtype=torch.bfloat16)
    draw2 = torch.empty(rows, c, device="cuda", dtype=torch.bfloat16)
    dres = torch.zeros(rows, c, device="cuda", dtype=torch.bfloat16)
    dbeta = torch.zeros(c, device="cuda")
    n = rows * c
    t = time_it(lambda: _capi.check(lib.urn_bn_stats(raw.data_ptr(), c, rows, c, 1e-3, stats.data_ptr(), ws.data_ptr(), st())))
    print("ptrs raw %x out %x dout %x draw %x" % (raw.data_ptr(), out.data_ptr(), dout.data_ptr(), draw.data_ptr()))
    print("stats              rows=%d c=%d : %.3f ms  %.0f GB/s" % (rows, c, t, n * 4 / t / 1e6))
    for relu, resk, name, byts in ((1, 0, "apply relu+shadow ", 8), (0, 0, "apply plain+shadow", 8), (1, 1, "apply +res        ", 10), (1, 2, "apply +bn(res)    ", 12)):
        t = time_it(lambda: _capi.check(lib.urn_bn_apply(raw.data_ptr(), c, rows, c, stats.data_ptr(), beta.data_ptr(),
                                                         res.data_ptr() if resk == 1 else None, c, F16,
                                                         raw2.data_ptr() if resk == 2 else None, c, stats.data_ptr() if resk == 2 else None,
                                                         beta.data_ptr() if resk == 2 else None, relu, out.data_ptr(), c, F16, out2.data_ptr(), st())))
        print("%s rows=%d c=%d : %.3f ms  %.0f GB/s" % (name, rows, c, t, n * byts / t / 1e6))
    cases = (("bwd plain (xhat=out-beta)", 0, False, False, True, 10), ("bwd relu (mask from raw) ", 1, False, False, True, 14),
             ("bwd relu + dres acc      ", 1, False, True, False, 22), ("bwd relu + dual BN       ", 1, True, False, False, 28))
    for name, relu, dual, dr, wb, byts in cases:
        t = time_it(lambda: _capi.check(lib.urn_bn_bwd_beta(
            dout.data_ptr(), c, out.data_ptr(), c, BF16, F16, relu, rows, c, raw.data_ptr(), c, stats.data_ptr(),
            beta.data_ptr() if wb else None, dbeta.data_ptr(), draw.data_ptr(), c,
            raw2.data_ptr() if dual else None, c, stats.data_ptr() if dual else None, dbeta.data_ptr() if dual else None,
            draw2.data_ptr() if dual else None, c, dres.data_ptr() if dr else None, c, 1, 0, ws.data_ptr(), st())))
        print("%s rows=%d c=%d : %.3f ms  %.0f GB/s (both passes)" % (name, rows, c, t, n * byts / t / 1e6))


if __name__ == "__main__":
    for rows, c in ((32 * 512 * 512, 16), (32 * 256 * 256, 32), (32 * 128 * 128, 64), (32 * 32 * 32, 256)):
        run(rows, c)
