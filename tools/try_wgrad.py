import ctypes as C, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from uresnet_b200 import _capi
from oracle import uresnet_oracle as O
lib = _capi.load()
TD = {0: torch.float32, 1: torch.bfloat16, 2: torch.float16}
xdt, ydt = int(sys.argv[1]), int(sys.argv[2])
nd, n, sp, cin, cout, k, s = eval(sys.argv[3])
rng = np.random.default_rng(0)
x = torch.tensor(rng.standard_normal((n,) + sp + (cin,)), dtype=torch.float32).cuda().to(TD[xdt])
d = _capi.conv_desc_same(nd, n, sp, cin, cout, k, s)
osp = (d.yd, d.yh, d.yw)[3 - nd:]
g = torch.tensor(rng.standard_normal((n,) + osp + (cout,)), dtype=torch.float32).cuda().to(TD[ydt])
wt = torch.zeros((k,) * nd + (cin, cout), dtype=torch.float64, requires_grad=True)
xc = x.float().cpu().double(); gc = g.float().cpu().double()
yo = O.conv_same(xc.permute(0, nd + 1, *range(1, nd + 1)), wt, s)
(dwo,) = torch.autograd.grad(yo, wt, gc.permute(0, nd + 1, *range(1, nd + 1)))
ws = torch.empty(max(int(lib.urn_conv_wgrad_workspace(C.byref(d))), 16), dtype=torch.uint8, device="cuda")
dw = torch.zeros(tuple(wt.shape), dtype=torch.float32, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
_capi.check(lib.urn_conv_wgrad(C.byref(d), xdt, 2, x.data_ptr(), cin, g.data_ptr(), cout, ydt, dw.data_ptr(), 0, ws.data_ptr(), st))
torch.cuda.synchronize()
a, b = dw.cpu().double().numpy(), dwo.numpy()
print("xdt", xdt, "ydt", ydt, sys.argv[3], "relerr %.3e" % (np.abs(a - b).max() / np.abs(b).max()), "max ref %.3e" % np.abs(b).max())
