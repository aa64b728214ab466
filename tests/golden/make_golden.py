"""Generate the golden fixtures in this directory from the CPU oracle.

    python tests/golden/make_golden.py

The reference ships no golden vectors and its TensorFlow cannot run here, so these
vectors (round 1) freeze the *oracle* (oracle/uresnet_oracle.py, fp64), which is
cross-checked against oracle/naive.py and finite differences in tests/test_oracle.py.  Inputs come from the seeded synthetic generator and the seeded
Xavier initialiser; both inputs and outputs are stored so that the fixtures do not
depend on those generators staying bit-stable.

Since the end of round 2 the same cases also exist as gref*.npz, made by executing the
reference's own lib/*.py over oracle/tf1_shim.py (make_reference_run_golden.py);
tests/test_reference_run.py asserts that the files written here hold the same numbers.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import uresnet_oracle as O  # noqa: E402
from uresnet_b200 import synthetic_io as S  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (dims, batch, base, strides, num_class, use_weight, store_params)
    "g2d_f4s3": ([32, 32, 1], 2, 4, 3, 3, True, True),
    "g3d_f4s2": ([16, 16, 16, 1], 2, 4, 2, 3, True, True),
    "g2d_f16s5": ([64, 64, 1], 1, 16, 5, 3, False, False),
}


def make(name):
    dims, B, F, St, C, use_w, store_params = CASES[name]
    nd = len(dims) - 1
    data, label, weight = S.make_batch(B, dims, seed=1234 + sum(map(ord, name)))
    weight = weight / weight.sum(axis=1, keepdims=True)          # ssnet_trainval.py:173
    P = O.init_params(nd, base=F, num_strides=St, num_class=C, seed=4321)
    shp = [B] + dims[:-1]
    r = O.run(P, data.reshape(shp + [1]), label.reshape(shp), weight.reshape(shp) if use_w else None,
              base=F, num_strides=St)
    out = dict(dims=np.array(dims), batch=B, base=F, strides=St, num_class=C, use_weight=use_w,
               data=data, label=label.astype(np.int8), weight=weight,
               logits=r["logits"].astype(np.float32), loss=np.float64(r["loss"]),
               acc_all=np.float64(r["acc_all"]), acc_nonzero=np.float64(r["acc_nonzero"]),
               n_correct_all=r["n_correct_all"], n_correct_nz=r["n_correct_nz"], n_nz=r["n_nz"],
               param_seed=4321)
    names = list(P.keys())
    out["grad_names"] = np.array(names)
    out["grad_l2"] = np.array([np.sqrt((r["grads"][k] ** 2).sum()) for k in names])
    out["grad_sum"] = np.array([r["grads"][k].sum() for k in names])
    if store_params:
        for i, k in enumerate(names):
            out["p%03d" % i] = P[k]
            out["g%03d" % i] = r["grads"][k].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", r["loss"], "acc", r["acc_all"], r["acc_nonzero"],
          os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    for n in CASES:
        make(n)
