"""Per-layer parity diagnostics (GPU): device intermediates vs the oracle's, per dtype/impl."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import uresnet_oracle as O  # noqa: E402
from uresnet_b200 import _capi, synthetic_io as S  # noqa: E402
from uresnet_b200.uresnet import uresnet  # noqa: E402


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def rmserr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.sqrt(((a - b) ** 2).mean()) / max(np.sqrt((b ** 2).mean()), 1e-30)


def main():
    dims = [int(v) for v in os.environ.get("DIMS", "64,64,1").split(",")]
    B, F, St = int(os.environ.get("B", 2)), int(os.environ.get("F", 8)), int(os.environ.get("S", 4))
    nd = len(dims) - 1
    data, label, weight = S.make_batch(B, dims, seed=99)
    weight /= weight.sum(axis=1, keepdims=True)
    P = O.init_params(nd, base=F, num_strides=St, seed=5)
    shp = [B] + dims[:-1]
    keep = {}
    ref = O.run(P, data.reshape(shp + [1]), label.reshape(shp), weight.reshape(shp), base=F, num_strides=St, keep=keep)
    ref64, keep64 = ref, keep
    vs64 = os.environ.get("VS64", "1") == "1"
    for dtype, impl in [tuple(v.split(":")) for v in os.environ.get("MODES", "f32:ffma,mix16:ffma").split(",")]:
        ref, keep = ref64, keep64
        if dtype != "f32" and not vs64:
            keep = {}
            ref = O.run(P, data.reshape(shp + [1]), label.reshape(shp), weight.reshape(shp), base=F, num_strides=St, keep=keep, quant=dtype)
        net = uresnet(dims=dims, num_class=3, num_strides=St, base_num_outputs=F)
        net.construct(trainable=True, use_weight=True, dtype=dtype, impl=impl)
        net.set_variables(P)
        net.zero_gradients(None)
        res, _ = net.accum_gradients(None, data, label, weight)
        lg = net.logits(B).cpu().numpy()
        print("== %s/%s loss %.6f (ref %.6f) logits max-rel %.3e rms %.3e" % (dtype, impl, res[1], ref["loss"], relerr(lg, ref["logits"]), rmserr(lg, ref["logits"])))
        h, ws = net._handle(B)
        lib = net._lib
        if os.environ.get("LAYERS", "1") == "1":
            for i in range(net._num_layers):
                scope, d, tr = C.c_char_p(), _capi.urn_conv_desc(), C.c_int()
                ro, oo, old, so = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
                _capi.check(lib.urn_layer_info(h, i, C.byref(scope), C.byref(d), C.byref(tr), C.byref(ro), C.byref(oo), C.byref(old), C.byref(so)))
                sc = scope.value.decode()
                if oo.value < 0:
                    continue
                r = keep[sc]
                cout = r.shape[-1]
                rows = r.size // cout
                es = 4 if (dtype == "f32" or sc == "conv2") else 2
                tdt = torch.float32 if es == 4 else (torch.bfloat16 if dtype == "bf16" else torch.float16)
                buf = ws[oo.value: oo.value + es * ((rows - 1) * old.value + cout)].view(tdt)
                v = torch.as_strided(buf, (rows, cout), (old.value, 1)).float().cpu().numpy().reshape(r.shape)
                print("  fwd %-42s max-rel %.3e rms %.3e" % (sc, relerr(v, r), rmserr(v, r)))
        g = net.get_variables("grads")
        for k in g:
            print("  grad %-52s max-rel %.3e rms %.3e" % (k, relerr(g[k], ref["grads"][k]), rmserr(g[k], ref["grads"][k])))


if __name__ == "__main__":
    main()
