import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
buf = torch.zeros(3 * 64 * 8 + 256, dtype=torch.int64, device="cuda")
os.environ["URN_HALO_TIMELINE"] = str(buf.data_ptr())
from uresnet_b200 import _capi
lib = _capi.load()
nd, n, sp, cin, cout = 2, 32, (512, 512), 16, 16
if len(sys.argv) > 1 and sys.argv[1] == "3d":
    nd, n, sp = 3, 2, (192, 192, 192)
d = _capi.conv_desc_same(nd, n, sp, cin, cout, 3, 1)
x = torch.randn((n,) + sp + (cin,), device="cuda").half()
w = torch.randn((3,) * nd + (cin, cout), device="cuda") * 0.05
wpack = torch.empty(2 * w.numel() + 64, dtype=torch.uint8, device="cuda")
y = torch.empty((n,) + sp + (cout,), dtype=torch.float32, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    _capi.check(lib.urn_conv_fwd_tc(C.byref(d), 2, x.data_ptr(), cin, w.data_ptr(), wpack.data_ptr(), y.data_ptr(), cout, 0, 0, st))
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
