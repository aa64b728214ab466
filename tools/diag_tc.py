"""tcgen05 path vs CUDA-core path on identical operands (GPU): per-layer raw conv outputs and
final gradients of two nets that differ only in `impl`.  Both round operands identically, so
they must agree to fp32 accumulation-order noise."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uresnet_b200 import _capi, synthetic_io as S  # noqa: E402
from uresnet_b200.uresnet import uresnet  # noqa: E402


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def layer_raws(net, B):
    h, ws = net._handle(B)
    out = []
    for i in range(net._num_layers):
        scope, d, tr = C.c_char_p(), _capi.urn_conv_desc(), C.c_int()
        ro = C.c_int64()
        _capi.check(net._lib.urn_layer_info(h, i, C.byref(scope), C.byref(d), C.byref(tr), C.byref(ro), None, None, None))
        cout = d.cx if tr.value else d.cy
        rows = d.n * (d.xd * d.xh * d.xw if tr.value else d.yd * d.yh * d.yw)
        out.append((scope.value.decode(), ws[ro.value: ro.value + 4 * rows * cout].view(torch.float32).cpu().numpy()))
    return out


def main():
    cases = os.environ.get("CASES", "64,64,1:2:16:3;32,32,32,1:1:16:2;96,80,1:1:16:5").split(";")
    dtype = os.environ.get("DTYPE", "mix16")
    worst = 0.0
    for case in cases:
        dims_s, B, F, St = case.split(":")
        dims, B, F, St = [int(v) for v in dims_s.split(",")], int(B), int(F), int(St)
        data, label, weight = S.make_batch(B, dims, seed=7)
        weight /= weight.sum(axis=1, keepdims=True)
        nets = {}
        for impl in ("ffma", "tc"):
            net = uresnet(dims=dims, num_class=3, num_strides=St, base_num_outputs=F)
            net.construct(trainable=True, use_weight=True, dtype=dtype, impl=impl, seed=3)
            net.zero_gradients(None)
            res, _ = net.accum_gradients(None, data, label, weight)
            nets[impl] = (net, res)
        print("== case", case, "loss ffma %.6f tc %.6f" % (nets["ffma"][1][1], nets["tc"][1][1]))
        ra, rb = layer_raws(nets["ffma"][0], B), layer_raws(nets["tc"][0], B)
        shown = 0
        for (sc, a), (_, b) in zip(ra, rb):
            e = relerr(b, a)
            if e > 1e-4 and shown < 6:
                print("   raw %-44s rel %.3e  (max ref %.3e) nan=%d" % (sc, e, np.abs(a).max(), int(np.isnan(b).sum())))
                shown += 1
        ga, gb = nets["ffma"][0].get_variables("grads"), nets["tc"][0].get_variables("grads")
        ge = {k: relerr(gb[k], ga[k]) for k in ga}
        kmax = max(ge, key=ge.get)
        fe = max(relerr(b, a) for (_, a), (_, b) in zip(ra, rb))
        print("   worst fwd raw rel %.3e ; worst grad rel %.3e (%s)" % (fe, ge[kmax], kmax))
        worst = max(worst, fe, ge[kmax])
    print("WORST", worst)


if __name__ == "__main__":
    main()
