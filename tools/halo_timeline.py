import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
buf = torch.zeros(3 * 64 * 8 + 256, dtype=torch.int64, device="cuda")
os.environ["URN_HALO_TIMELINE"] = str(buf.data_ptr())
from uresnet_b200 import _capi
lib = _capi.load()
nd, n, sp, cin, cout = 2, 32, (512, 512), 16, 16
if len(sys.argv) > 1 and sys.argv[1] == "3d":
    nd, n, sp = 3, 2, (192, 192, 192)
what = sys.argv[2] if len(sys.argv) > 2 else "fwd"        # fwd | dgrad | dgrad_acc
d = _capi.conv_desc_same(nd, n, sp, cin, cout, 3, 1)
x = torch.randn((n,) + sp + (cin,), device="cuda").half()
w = torch.randn((3,) * nd + (cin, cout), device="cuda") * 0.05
wpack = torch.empty(2 * w.numel() + 64, dtype=torch.uint8, device="cuda")
ydt = int(os.environ.get("YDT", "2"))                       # raw output dtype: 2 = fp16 (mix16 product path), 0 = fp32
y = torch.empty((n,) + sp + (cout,), dtype=torch.float16 if ydt == 2 else torch.float32, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
g = torch.randn((n,) + sp + (cout,), device="cuda").bfloat16()
dx = torch.zeros((n,) + sp + (cin,), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    if what == "fwd":
        _capi.check(lib.urn_conv_fwd_tc(C.byref(d), 2, x.data_ptr(), cin, w.data_ptr(), wpack.data_ptr(), y.data_ptr(), cout, ydt, 0, st))
    else:
        _capi.check(lib.urn_conv_dgrad_tc(C.byref(d), 1, g.data_ptr(), cout, w.data_ptr(), wpack.data_ptr(), dx.data_ptr(), cin, 1,
                                          int(what == "dgrad_acc"), st))
torch.cuda.synchronize()
full = buf.cpu().numpy()
t = full[:3 * 64 * 8].reshape(3, 64, 8)
cta = full[3 * 64 * 8:3 * 64 * 8 + 148]
smid = cta & 255
cta = cta >> 8
print('per-CTA total cycles: min %d median %d max %d  (cta0 %d)' % (cta.min(), sorted(cta)[74], cta.max(), cta[0]))
print('cta:smid:kcycles ' + ' '.join('%d:%d:%d' % (i, smid[i], cta[i] // 1000) for i in range(148)))
t0 = t[t > 0].min()
names = {0: ["P.empty_ok", "P.issued", "P.arrived"], 1: ["M.start", "M.tempty_ok", "M.full_ok", "M.fenced", "M.mma_issued", "M.committed"],
         2: ["E.start", "E.tfull_ok", "E.stored", "E.arrived"]}
for k in list(range(0, 4)) + list(range(30, 40)):
    print("tile %2d" % k, " | ".join("%s" % " ".join("%s=%d" % (nm.split('.')[1], t[r, k, i] - t0) for i, nm in enumerate(names[r])) for r in (0, 1, 2)))
print("steady-state per-tile period (cycles): producer %.0f  mma %.0f  epilogue %.0f" % tuple((t[r, 60, 0] - t[r, 30, 0]) / 30.0 for r in (0, 1, 2)))

# where each role waits, steady state (tiles 20..59): mean cycles per tile
import numpy as np
sl = slice(20, 60)
P, M, E = t[0], t[1], t[2]
print("producer: wait for empty slot %.0f | issue loads %.0f | until arrive %.0f" % (
    np.mean(P[sl, 0] - np.roll(P[:, 2], 1)[sl]), np.mean(P[sl, 1] - P[sl, 0]), np.mean(P[sl, 2] - P[sl, 1])))
print("mma     : wait tempty (epilogue) %.0f | wait full (producer) %.0f | fence %.0f | issue %.0f | commit %.0f" % (
    np.mean(M[sl, 1] - M[sl, 0]), np.mean(M[sl, 2] - M[sl, 1]), np.mean(M[sl, 3] - M[sl, 2]), np.mean(M[sl, 4] - M[sl, 3]),
    np.mean(M[sl, 5] - M[sl, 4])))
print("epilogue: wait tfull (mma) %.0f | ld + convert + store %.0f | arrive %.0f" % (
    np.mean(E[sl, 1] - E[sl, 0]), np.mean(E[sl, 2] - E[sl, 1]), np.mean(E[sl, 3] - E[sl, 2])))
print("mma commit -> epilogue sees tfull: %.0f cycles" % np.mean(E[sl, 1] - M[sl, 5]))
