"""Golden vectors made by EXECUTING THE REFERENCE'S OWN MODEL CODE: /root/reference/lib/uresnet.py,
resnet_module.py and ssnet.py are imported unmodified over ``oracle/tf1_shim.py`` (an emulation of the
TensorFlow-1.x ops they call, evaluated with torch-CPU float64) and driven through the reference's public
API -- ``uresnet(...)``, ``construct``, ``zero_gradients``, ``accum_gradients``, ``apply_gradients``,
``inference`` (lib/ssnet.py:20-139).

    python tests/golden/make_reference_run_golden.py        # in the build container (needs /root/reference)

Writes, next to this file, for the same seeded inputs as make_golden.py's cases:
  gref2d_f4s3.npz, gref3d_f4s2.npz   same keys as g2d_f4s3 / g3d_f4s2 (inputs, parameters, logits, loss,
                                     accuracies, all gradients) + the parameters after one Adam apply and the
                                     softmax of ``inference``; tests/test_net_gpu.py::test_golden_small picks
                                     them up by name, so the CUDA path is compared with them in ``-m gpu``
  gref2d_f16s5.npz                   the default topology (F=16, S=5) at 64x64, summary keys only
tests/test_reference_run.py compares the oracle (and the oracle-made goldens) with these files, and, where
/root/reference exists, re-runs the reference code and compares with the committed files.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from uresnet_b200 import synthetic_io as S  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LR = 1e-3

# the cases of make_golden.py: name -> (dims, batch, base, strides, num_class, use_weight, store_params)
CASES = {
    "g2d_f4s3": ([32, 32, 1], 2, 4, 3, 3, True, True),
    "g3d_f4s2": ([16, 16, 16, 1], 2, 4, 2, 3, True, True),
    "g2d_f16s5": ([64, 64, 1], 1, 16, 5, 3, False, False),
}


def case_inputs(name):
    """inputs and parameters exactly as make_golden.py draws them (the oracle is used for nothing else here)"""
    from oracle import uresnet_oracle as O
    dims, B, F, St, C, use_w, _ = CASES[name]
    data, label, weight = S.make_batch(B, dims, seed=1234 + sum(map(ord, name)))
    weight = weight / weight.sum(axis=1, keepdims=True)          # ssnet_trainval.py:173
    P = O.init_params(len(dims) - 1, base=F, num_strides=St, num_class=C, seed=4321)
    return P, data, label, weight


def run_reference(dims, F, St, C, use_w, P, minibatches, lr=LR, float64=True):
    """Drives the reference's classes over the shim.  ``minibatches``: list of (data, label, weight) host
    arrays in the reference's flattened layout, accumulated into one Adam apply (ssnet_trainval.py:164-191).
    Returns per-minibatch [loss, acc_all, acc_nonzero], logits / softmax of the first minibatch, the
    accumulated gradients and the parameters after ``apply_gradients``, keyed by TF variable name."""
    import torch
    from oracle import tf1_shim as shim
    shim.reset_default_graph(torch.float64 if float64 else torch.float32)
    uresnet = shim.load_reference()
    import tensorflow as tf                                       # the shim
    net = uresnet(dims=dims, num_class=C, num_strides=St, base_num_outputs=F)
    net.construct(trainable=True, use_weight=use_w, learning_rate=lr)
    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    tvs = tf.trainable_variables()
    names = [v.op.name for v in tvs]
    if P is not None:
        assert names == list(P.keys()), "variable creation order differs from the oracle's parameter table"
        for v in tvs:
            v.load(P[v.op.name], sess)
    out = dict(names=names, shapes=[v.shape.as_list() for v in tvs], metrics=[],
               all_variables=[v.name for v in tf.global_variables()])
    logits_t = net._softmax._inputs[0]                              # the tensor _build returned (ssnet.py:45,63)
    net.zero_gradients(sess)
    for i, (d, l, w) in enumerate(minibatches):
        res, doc = net.accum_gradients(sess, d, l, w if use_w else None)
        assert doc == ['', 'loss', 'acc. all', 'acc. nonzero']
        out["metrics"].append([float(res[1]), float(res[2]), float(res[3])])
        if i == 0:
            out["logits"] = sess.run(logits_t, feed_dict=net.feed_dict(d, l, w if use_w else None))
            if use_w:       # inference() of a weighted net raises TypeError in the reference (ssnet.py:141-145)
                out["softmax"] = sess.run(net._softmax, feed_dict=net.feed_dict(d, l, w))
            else:
                inf = net.inference(sess, d, l)
                out["softmax"], out["inference_acc"] = inf[0], [float(inf[1]), float(inf[2])]
            out["run_test"] = [float(x) for x in net.run_test(sess, d, l, w if use_w else None)[0]]
    out["grads"] = dict(zip(names, sess.run(net._accum_vars)))
    net.apply_gradients(sess)
    out["params_after"] = dict(zip(names, sess.run(tvs)))
    out["lr"] = net._opt._lr
    shim.uninstall()
    return out


def make(name):
    dims, B, F, St, C, use_w, store_params = CASES[name]
    P, data, label, weight = case_inputs(name)
    r = run_reference(dims, F, St, C, use_w, P, [(data, label, weight)])
    lg = r["logits"]
    nz = data.reshape(B, -1) > 0
    pred = lg.reshape(B, -1, C).argmax(-1)
    lab = label.reshape(B, -1).astype(np.int64)
    loss, acc_all, acc_nz = r["metrics"][0]
    out = dict(dims=np.array(dims), batch=B, base=F, strides=St, num_class=C, use_weight=use_w,
               data=data, label=label.astype(np.int8), weight=weight,
               logits=lg.astype(np.float32), loss=np.float64(loss), acc_all=np.float64(acc_all),
               acc_nonzero=np.float64(acc_nz), n_correct_all=int((pred == lab).sum()),
               n_correct_nz=int(((pred == lab) & nz).sum()), n_nz=int(nz.sum()), param_seed=4321,
               adam_lr=np.float64(LR), made_by="reference lib/*.py executed over oracle/tf1_shim.py (float64)")
    names = r["names"]
    out["grad_names"] = np.array(names)
    out["grad_l2"] = np.array([np.sqrt((r["grads"][k] ** 2).sum()) for k in names])
    out["grad_sum"] = np.array([r["grads"][k].sum() for k in names])
    if store_params:
        out["softmax"] = r["softmax"].astype(np.float32)
        for i, k in enumerate(names):
            out["p%03d" % i] = P[k]
            out["g%03d" % i] = r["grads"][k].astype(np.float32)
            out["a%03d" % i] = r["params_after"][k].astype(np.float32)
    dst = os.path.join(HERE, name.replace("g", "gref", 1) + ".npz")
    np.savez_compressed(dst, **out)
    print(os.path.basename(dst), "loss", loss, "acc", acc_all, acc_nz, os.path.getsize(dst) // 1024, "KiB")


if __name__ == "__main__":
    for n in CASES:
        make(n)
