"""The oracle against THE REFERENCE'S OWN CODE RUNNING.

``/root/reference/lib/{uresnet,resnet_module,ssnet}.py`` are imported unmodified over ``oracle/tf1_shim.py``
(an emulation of the TensorFlow-1.x ops they call: the reference's graph composition, loss / accuracy
expressions and accumulate / zero / Adam protocol execute as written; the ops are evaluated with torch-CPU)
and driven through the reference's public API.  ``tests/golden/gref*.npz`` are outputs of such runs
(``tests/golden/make_reference_run_golden.py``).

* always (no reference tree needed -- also on the GPU box): the oracle reproduces the committed reference-run
  vectors (logits, softmax, loss, accuracies, every gradient, the parameters after one Adam apply), and the
  oracle-made goldens the GPU tests use are the same numbers;
* where ``/root/reference`` exists: the reference code is re-run live -- committed vectors regenerate; a
  3-minibatch accumulation + two Adam applies; variable names / shapes / creation order of the 2D and 3D default
  topologies against the oracle's table and the library's plan; the debug shape sequence; the float32 evaluation;
  the reference's error behaviour.
``tests/test_net_gpu.py::test_golden_small`` picks ``gref*_f4s*.npz`` up by name, so the CUDA path is compared
with reference-run vectors in ``-m gpu``.
"""
import ctypes as C
import glob
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import uresnet_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
HAVE_REF = os.path.exists("/root/reference/lib/uresnet.py")
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="reference tree not on this box")


def _mk():
    spec = importlib.util.spec_from_file_location("mkref", os.path.join(GOLD, "make_reference_run_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _oracle_on(z):
    dims = [int(v) for v in z["dims"]]
    B, F, St = int(z["batch"]), int(z["base"]), int(z["strides"])
    names = [str(n) for n in z["grad_names"]]
    if "p000" in z:
        P = {k: z["p%03d" % i] for i, k in enumerate(names)}
    else:
        P = O.init_params(len(dims) - 1, base=F, num_strides=St, num_class=int(z["num_class"]), seed=int(z["param_seed"]))
    shp = [B] + dims[:-1]
    w = z["weight"].reshape(shp) if bool(z["use_weight"]) else None
    return P, names, O.run(P, z["data"].reshape(shp + [1]), z["label"].astype(np.float32).reshape(shp), w,
                           base=F, num_strides=St)


@pytest.mark.parametrize("name", ["gref2d_f4s3", "gref3d_f4s2", "gref2d_f16s5"])
def test_oracle_reproduces_reference_run_vectors(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    assert "reference lib" in str(z["made_by"])
    P, names, r = _oracle_on(z)
    assert list(P.keys()) == names
    assert rel(r["logits"], z["logits"]) < 2e-7                     # the vectors are stored as float32
    assert abs(r["loss"] - float(z["loss"])) <= 1e-12 * abs(float(z["loss"]))
    assert r["acc_all"] == pytest.approx(float(z["acc_all"]), abs=1e-7)
    assert r["acc_nonzero"] == pytest.approx(float(z["acc_nonzero"]), abs=1e-7)
    assert (r["n_correct_all"], r["n_correct_nz"], r["n_nz"]) == (int(z["n_correct_all"]), int(z["n_correct_nz"]), int(z["n_nz"]))
    l2 = np.array([np.sqrt((r["grads"][k] ** 2).sum()) for k in names])
    np.testing.assert_allclose(l2, z["grad_l2"], rtol=1e-10)
    np.testing.assert_allclose([r["grads"][k].sum() for k in names], z["grad_sum"], rtol=1e-7,
                               atol=1e-10 * np.abs(z["grad_l2"]).max())
    if "g000" in z:
        assert rel(r["softmax"], z["softmax"]) < 2e-7
        for i, k in enumerate(names):
            assert rel(r["grads"][k], z["g%03d" % i]) < 2e-7, k
            # one TF-Adam apply on the accumulated gradients (ssnet.py:72-79): the oracle's float32 ApplyAdam
            # against the reference run's (float64) update
            var = P[k].astype(np.float32).copy()
            g = r["grads"][k].astype(np.float32)
            O.adam_tf(var, g, np.zeros_like(var), np.zeros_like(var), 1, lr=float(z["adam_lr"]))
            assert np.abs(var - z["a%03d" % i]).max() <= 2e-6 * max(1.0, np.abs(var).max()) + 1e-7, k


@pytest.mark.parametrize("name", ["g2d_f4s3", "g3d_f4s2", "g2d_f16s5"])
def test_oracle_made_goldens_are_the_reference_run_numbers(name):
    """The fixtures the GPU tests have used since round 1 (made by the oracle) against the reference-run ones."""
    a = np.load(os.path.join(GOLD, name + ".npz"))
    b = np.load(os.path.join(GOLD, name.replace("g", "gref", 1) + ".npz"))
    for k in ("data", "label", "weight", "dims", "batch", "base", "strides", "num_class", "use_weight"):
        assert np.array_equal(a[k], b[k]), k
    assert list(a["grad_names"]) == list(b["grad_names"])
    assert np.abs(a["logits"].astype(np.float64) - b["logits"]).max() <= 1e-6 * np.abs(b["logits"]).max()
    assert abs(float(a["loss"]) - float(b["loss"])) <= 1e-12 * abs(float(b["loss"]))
    for k in ("acc_all", "acc_nonzero", "n_correct_all", "n_correct_nz", "n_nz"):
        assert float(a[k]) == pytest.approx(float(b[k]), abs=1e-7), k
    np.testing.assert_allclose(a["grad_l2"], b["grad_l2"], rtol=1e-10)
    if "g000" in a:
        for i in range(len(a["grad_names"])):
            assert np.array_equal(a["p%03d" % i], b["p%03d" % i])
            assert rel(a["g%03d" % i], b["g%03d" % i]) < 2e-7, i


def test_gpu_suite_picks_up_the_reference_run_vectors():
    got = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "g*_f4s*.npz")))
    assert got == ["g2d_f4s3.npz", "g3d_f4s2.npz", "gref2d_f4s3.npz", "gref3d_f4s2.npz"]


# ------------------------------------------------------------------------------------------------
# live runs of the reference code (build container only)
# ------------------------------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("name", ["g2d_f4s3", "g3d_f4s2"])
def test_live_reference_run_regenerates_committed_vectors(name):
    mk = _mk()
    dims, B, F, St, Cc, use_w, _ = mk.CASES[name]
    P, data, label, weight = mk.case_inputs(name)
    r = mk.run_reference(dims, F, St, Cc, use_w, P, [(data, label, weight)])
    z = np.load(os.path.join(GOLD, name.replace("g", "gref", 1) + ".npz"))
    assert r["names"] == [str(n) for n in z["grad_names"]]
    assert rel(r["logits"], z["logits"]) < 2e-7 and abs(r["metrics"][0][0] - float(z["loss"])) < 1e-12
    assert r["run_test"] == pytest.approx(r["metrics"][0], abs=1e-12)           # ssnet.py:121-128 == :103-115
    for i, k in enumerate(r["names"]):
        assert rel(r["grads"][k], z["g%03d" % i]) < 2e-7
        assert np.abs(r["params_after"][k] - z["a%03d" % i]).max() < 1e-6


@needs_ref
def test_live_accumulation_over_minibatches_and_two_adam_applies():
    """ssnet_trainval.py:164-191: zero, NUM_MINIBATCHES x accum (a SUM, not a mean), one apply -- twice."""
    import uresnet_b200.synthetic_io as S
    from oracle import tf1_shim as shim
    dims, B, F, St = [16, 32, 1], 2, 4, 2
    P = O.init_params(2, base=F, num_strides=St, seed=9)
    mbs = []
    for j in range(3):
        d, l, w = S.make_batch(B, dims, seed=100 + j)
        mbs.append((d, l, w / w.sum(axis=1, keepdims=True)))
    shim.reset_default_graph(torch.float64)
    uresnet = shim.load_reference()
    import tensorflow as tf
    net = uresnet(dims=dims, num_class=3, num_strides=St, base_num_outputs=F)
    net.construct(trainable=True, use_weight=True, learning_rate=2e-3)
    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    tvs = tf.trainable_variables()
    # the accumulators start as COPIES OF THE VARIABLES (tv.initialized_value(), ssnet.py:53-55): zero_gradients is
    # not optional
    for a, v in zip(sess.run(net._accum_vars), sess.run(tvs)):
        assert np.array_equal(a, v)
    for v in tvs:
        v.load(P[v.op.name], sess)
    shp = [B] + dims[:-1]
    cur = {k: v.astype(np.float32).copy() for k, v in P.items()}
    m = {k: np.zeros_like(v) for k, v in cur.items()}
    vv = {k: np.zeros_like(v) for k, v in cur.items()}
    for t in (1, 2):
        net.zero_gradients(sess)
        want = {k: 0.0 for k in P}
        for d, l, w in mbs:
            res, _ = net.accum_gradients(sess, d, l, w)
            r = O.run(cur, d.reshape(shp + [1]), l.reshape(shp), w.reshape(shp), base=F, num_strides=St)
            assert abs(res[1] - r["loss"]) < 1e-5 * abs(r["loss"])          # cur is float32-rounded after step 1
            want = {k: want[k] + r["grads"][k] for k in P}
        got = dict(zip([v.op.name for v in tvs], sess.run(net._accum_vars)))
        for k in P:
            assert rel(got[k], want[k]) < (1e-10 if t == 1 else 2e-3), (t, k)
        net.apply_gradients(sess)
        new = dict(zip([v.op.name for v in tvs], sess.run(tvs)))
        for k in P:
            O.adam_tf(cur[k], want[k].astype(np.float32), m[k], vv[k], t, lr=2e-3)
            assert np.abs(cur[k] - new[k]).max() < 2e-5, (t, k)
            cur[k] = new[k].astype(np.float32)                                  # follow the reference's trajectory
    assert net._opt._lr == 2e-3                                                 # read by ssnet_trainval.py:209
    shim.uninstall()


@needs_ref
@pytest.mark.parametrize("dims", [[512, 512, 1], [192, 192, 192, 1]])
def test_live_default_topology_variables_and_shapes(dims, lib, capsys):
    """Constructing the reference's default net (F=16, S=5; nothing is evaluated): variable names, shapes and
    creation order against the oracle's table and the C-ABI plan; the debug shape prints (uresnet.py:24-122)
    against the plan's layer geometry."""
    from oracle import tf1_shim as shim
    from uresnet_b200 import _capi
    nd = len(dims) - 1
    shim.reset_default_graph(torch.float64)
    uresnet = shim.load_reference()
    import tensorflow as tf
    net = uresnet(dims=dims, num_class=3, debug=True)
    net.construct(trainable=True, use_weight=True, learning_rate=-1)           # <= 0: AdamOptimizer() default lr
    assert net._opt._lr == 0.001
    printed = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("(")]
    tvs = tf.trainable_variables()
    want = O.param_names(nd, base=16, num_strides=5, num_class=3)
    assert [(v.op.name, tuple(v.shape.as_list())) for v in tvs] == [(n, tuple(s)) for n, s in want]
    assert sum(int(np.prod(s)) for _, s in want) == {2: 16858979, 3: 49862627}[nd]
    every = [v.op.name for v in tf.global_variables()]
    for l in O.layer_table(nd):
        for leaf in ("weights", "BatchNorm/beta", "BatchNorm/moving_mean", "BatchNorm/moving_variance"):
            assert "UResNet/%s/%s" % (l["scope"], leaf) in every
    assert not any(n.endswith("/biases") or n.endswith("/gamma") for n in every)
    assert net._softmax.shape.as_list() == [None] + dims[:-1] + [3]
    # the library's plan: same names, shapes, order
    cfg = _capi.urn_net_cfg()
    cfg.nd = nd
    for i in range(3):
        cfg.spatial[i] = dims[i] if i < nd else 1
    cfg.in_channels, cfg.batch, cfg.base_filters, cfg.num_strides = 1, 1, 16, 5
    cfg.num_class, cfg.dtype, cfg.impl, cfg.training = 3, 0, 0, 1
    h = C.c_void_p()
    _capi.check(lib.urn_create(C.byref(cfg), C.byref(h)))
    assert lib.urn_param_count(h) == len(tvs)
    for i, v in enumerate(tvs):
        name, rank = C.c_char_p(), C.c_int()
        shp = (C.c_int64 * 6)()
        off, cnt = C.c_int64(), C.c_int64()
        _capi.check(lib.urn_param_info(h, i, C.byref(name), C.byref(rank), shp, C.byref(off), C.byref(cnt)))
        assert name.value.decode() == v.op.name and [shp[k] for k in range(rank.value)] == v.shape.as_list()
    # debug prints: input, conv0, 5 x resnet_module, 5 x (deconv, concat, resnet_module), conv1, conv2
    assert len(printed) == 2 + 5 + 15 + 2
    n, c = dims[0], 16
    seq = [(n, 1), (n, 16)]
    for _ in range(5):
        n, c = n // 2, c * 2
        seq.append((n, c))
    for _ in range(5):
        n, c = n * 2, c // 2
        seq += [(n, c), (n, 2 * c), (n, c)]
    seq += [(n, 16), (n, 3)]
    for ln, (side, ch) in zip(printed, seq):
        assert ln.startswith("(?, " + ", ".join([str(side)] * nd) + ", %d)" % ch), (ln, side, ch)
    # ... which is the geometry the plan executes
    sides = {}
    for i, l in enumerate(O.layer_table(nd)):
        scope, d, tr = C.c_char_p(), _capi.urn_conv_desc(), C.c_int()
        _capi.check(lib.urn_layer_info(h, i, C.byref(scope), C.byref(d), C.byref(tr), None, None, None, None))
        sides[scope.value.decode()] = (d.xh, d.cx) if tr.value else (d.yh, d.cy)
    assert sides["conv0"] == seq[1] and sides["resnet_module4/module2/resnet_conv2"] == seq[6]
    assert sides["deconv0"] == seq[7] and sides["conv1"] == seq[-2] and sides["conv2"] == seq[-1]
    lib.urn_destroy(h)
    shim.uninstall()


@needs_ref
def test_live_float32_evaluation_brackets_the_exact_mode():
    """The reference computes in float32 (ssnet.py:30-32).  Its code evaluated in float32 over the shim deviates
    from the float64 evaluation by what 'fp32 parity' can mean on this graph -- the band the f32 GPU tests use."""
    mk = _mk()
    dims, B, F, St, Cc, use_w, _ = mk.CASES["g2d_f4s3"]
    P, data, label, weight = mk.case_inputs("g2d_f4s3")
    r32 = mk.run_reference(dims, F, St, Cc, use_w, P, [(data, label, weight)], float64=False)
    z = np.load(os.path.join(GOLD, "gref2d_f4s3.npz"))
    assert r32["logits"].dtype == np.float32
    assert rel(r32["logits"], z["logits"]) < 1e-4
    assert abs(r32["metrics"][0][0] - float(z["loss"])) < 1e-5 * float(z["loss"])
    e = [rel(r32["grads"][k], z["g%03d" % i]) for i, k in enumerate(r32["names"])]
    assert np.median(e) < 1e-4 and max(e) < 1e-3


@needs_ref
def test_live_error_behaviour_of_the_reference():
    from oracle import tf1_shim as shim
    shim.reset_default_graph(torch.float64)
    uresnet = shim.load_reference()
    import tensorflow as tf
    with pytest.raises(NotImplementedError):                                    # ssnet.py:12-14
        uresnet(dims=[32, 32], num_class=3)
    # uresnet.py:100 hard-codes the decoder scopes as step+5: a sixth stride re-creates resnet_module5 (TF refuses;
    # so does the product: tests/test_capi_cpu.py::test_create_rejects_more_than_five_strides)
    with pytest.raises(ValueError, match="resnet_module5/module1/shortcut/weights already exists"):
        uresnet(dims=[64, 64, 1], num_class=3, num_strides=6, base_num_outputs=2).construct(True, True, 1e-3)
    shim.reset_default_graph(torch.float64)
    net = uresnet(dims=[8, 8, 1], num_class=3, num_strides=1, base_num_outputs=2)
    net.construct(trainable=True, use_weight=True, learning_rate=1e-3)
    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    d = np.zeros((1, 64), np.float32)
    with pytest.raises(TypeError):                                              # ssnet.py:141-145
        net.accum_gradients(sess, d, d)
    with pytest.raises(TypeError):                                              # inference() of a weighted net
        net.inference(sess, d, d)
    bad = d.copy()
    bad[0, 3] = 7.0                                                             # label outside [0, 3)
    with pytest.raises(ValueError):
        net.accum_gradients(sess, d, bad, d)
    shim.uninstall()


@needs_ref
@pytest.mark.parametrize("dims,B,F,St", [([32, 32, 32, 1], 1, 2, 5), ([64, 32, 1], 2, 3, 5)])
def test_live_five_stride_topologies_numerics(dims, B, F, St):
    """S=5 (the default depth: decoder scopes resnet_module5..9, five skips) in 3D and on a non-square 2D image,
    odd base width: the reference code's run against the oracle, every gradient."""
    import uresnet_b200.synthetic_io as S
    mk = _mk()
    nd = len(dims) - 1
    d, l, w = S.make_batch(B, dims, seed=77)
    w = w / w.sum(axis=1, keepdims=True)
    P = O.init_params(nd, base=F, num_strides=St, seed=5)
    r = mk.run_reference(dims, F, St, 3, True, P, [(d, l, w)])
    shp = [B] + dims[:-1]
    o = O.run(P, d.reshape(shp + [1]), l.reshape(shp), w.reshape(shp), base=F, num_strides=St)
    assert rel(r["logits"], o["logits"]) < 1e-10 and abs(r["metrics"][0][0] - o["loss"]) < 1e-12 * abs(o["loss"])
    assert r["metrics"][0][1] == pytest.approx(o["acc_all"], abs=1e-7)
    assert r["metrics"][0][2] == pytest.approx(o["acc_nonzero"], abs=1e-7)
    assert len(r["grads"]) == 116
    for k in P:
        assert rel(r["grads"][k], o["grads"][k]) < 1e-9, k


@needs_ref
def test_live_inference_network():
    """The analysis configuration (ssnet_trainval.py:101-103,237: trainable=False, no weights): softmax and the
    two accuracies of ``inference`` against the oracle; no optimizer, no accumulators' ops."""
    import uresnet_b200.synthetic_io as S
    from oracle import tf1_shim as shim
    dims, B, F, St = [32, 32, 1], 2, 4, 3
    d, l, _ = S.make_batch(B, dims, seed=3)
    P = O.init_params(2, base=F, num_strides=St, seed=8)
    shim.reset_default_graph(torch.float64)
    uresnet = shim.load_reference()
    import tensorflow as tf
    net = uresnet(dims=dims, num_class=3, num_strides=St, base_num_outputs=F)
    net.construct(trainable=False, use_weight=False)
    assert net._loss is None and not hasattr(net, "_opt")
    # trainable=False reaches the filters; slim does not forward it to normalizer_fn, so the betas stay "trainable"
    assert [v.op.name for v in tf.trainable_variables()] == [k for k in P if k.endswith("/beta")]
    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    byname = {v.op.name: v for v in tf.global_variables()}
    for k, v in P.items():
        byname[k].load(v, sess)
    sm, acc_all, acc_nz = net.inference(sess, d, l)
    only = net.inference(sess, d)
    shp = [B] + dims[:-1]
    o = O.run(P, d.reshape(shp + [1]), l.reshape(shp), None, base=F, num_strides=St, want_grads=False)
    assert sm.shape == tuple(shp + [3]) and len(only) == 1 and np.array_equal(only[0], sm)
    assert rel(sm, o["softmax"]) < 1e-12
    assert float(acc_all) == pytest.approx(o["acc_all"], abs=1e-7) and float(acc_nz) == pytest.approx(o["acc_nonzero"], abs=1e-7)
    # the shower / track map the analysis step derives from it (ssnet_trainval.py:269-287)
    dec = O.ssnet_decision(sm, d.reshape(shp))
    assert set(np.unique(dec)) <= {0.0, 1.0, 2.0} and np.all(dec[d.reshape(shp) <= 1.0] == 0)
    shim.uninstall()


@needs_ref
def test_live_baseline_config0_inference_512():
    """BASELINE configs[0] as the reference states it: the default net (F=16, S=5, 3 classes) on one 512 x 512
    image, inference -- the reference code's softmax and accuracies against the oracle's."""
    import time
    import uresnet_b200.synthetic_io as S
    from oracle import tf1_shim as shim
    dims, B = [512, 512, 1], 1
    d, l, _ = S.make_batch(B, dims, seed=1234)
    P = O.init_params(2, seed=4321)
    shim.reset_default_graph(torch.float64)
    uresnet = shim.load_reference()
    import tensorflow as tf
    net = uresnet(dims=dims, num_class=3)
    net.construct(trainable=False, use_weight=False)
    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    byname = {v.op.name: v for v in tf.global_variables()}
    for k, v in P.items():
        byname[k].load(v, sess)
    t0 = time.time()
    sm, acc_all, acc_nz = net.inference(sess, d, l)
    t1 = time.time()
    o = O.run(P, d.reshape(1, 512, 512, 1), l.reshape(1, 512, 512), None, want_grads=False)
    assert sm.shape == (1, 512, 512, 3) and rel(sm, o["softmax"]) < 1e-12
    assert float(acc_all) == pytest.approx(o["acc_all"], abs=1e-7) and float(acc_nz) == pytest.approx(o["acc_nonzero"], abs=1e-7)
    assert (sm.argmax(-1) == o["softmax"].argmax(-1)).all()
    print("reference inference over the shim, 512x512 B=1, fp64: %.1f s" % (t1 - t0))
    shim.uninstall()


@needs_ref
def test_live_baseline_geometry_train_step_512():
    """BASELINE configs[1] geometry (512 x 512, F=16, S=5, weighted CE; one image): the reference code's training
    step -- logits, loss, all 116 accumulated gradients -- against the oracle, the very numbers
    tests/test_parity_gpu.py::test_baseline_size_2d_512 holds the CUDA path to."""
    import uresnet_b200.synthetic_io as S
    mk = _mk()
    dims, B = [512, 512, 1], 1
    d, l, w = S.make_batch(B, dims, seed=1234)
    w = w / w.sum(axis=1, keepdims=True)
    P = O.init_params(2, seed=4321)
    r = mk.run_reference(dims, 16, 5, 3, True, P, [(d, l, w)])
    o = O.run(P, d.reshape(1, 512, 512, 1), l.reshape(1, 512, 512), w.reshape(1, 512, 512))
    assert rel(r["logits"], o["logits"]) < 1e-12 and abs(r["metrics"][0][0] - o["loss"]) < 1e-12 * abs(o["loss"])
    assert len(r["grads"]) == 116
    worst = max(rel(r["grads"][k], o["grads"][k]) for k in P)
    assert worst < 1e-10, worst


@needs_ref
@pytest.mark.parametrize("dims,B,F,St", [([32, 32, 1], 2, 4, 3), ([16, 16, 16, 1], 1, 4, 2)])
def test_live_every_intermediate_tensor_of_the_step_oracle(dims, B, F, St):
    """oracle/steps.py keeps every intermediate by TF scope name and hands them to the CUDA schedule steps in the
    teacher-forced GPU tests.  Here the same tensors are FETCHED FROM THE REFERENCE'S GRAPH (raw conv output and
    layer output of every slim layer, by variable scope) and compared, layer by layer."""
    import uresnet_b200.synthetic_io as S
    from oracle import tf1_shim as shim
    from oracle.steps import StepOracle
    nd = len(dims) - 1
    d, l, w = S.make_batch(B, dims, seed=12)
    w = w / w.sum(axis=1, keepdims=True)
    P = O.init_params(nd, base=F, num_strides=St, seed=6)
    shp = [B] + dims[:-1]
    so = StepOracle(P, d.reshape(shp + [1]), l.reshape(shp), w.reshape(shp), base=F, num_strides=St)
    shim.reset_default_graph(torch.float64)
    uresnet = shim.load_reference()
    import tensorflow as tf
    net = uresnet(dims=dims, num_class=3, num_strides=St, base_num_outputs=F)
    net.construct(trainable=True, use_weight=True, learning_rate=1e-3)
    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    for v in tf.trainable_variables():
        v.load(P[v.op.name], sess)
    lt = shim.layer_tensors()
    assert set(lt) == {"UResNet/" + s for s in so.layers}
    scopes = list(so.layers)
    vals = sess.run([[lt["UResNet/" + s]["raw"], lt["UResNet/" + s]["out"]] for s in scopes],
                    feed_dict=net.feed_dict(d, l, w))
    cl = lambda t: t.detach().permute(0, *range(2, nd + 2), 1).numpy()        # noqa: E731
    # ... and the backward intermediates: tf.gradients of the reference's loss with respect to the raw conv output
    # of every layer (what the BN-backward steps produce and the filter- / input-gradient steps consume) and to
    # every layer output that is a node of the schedule
    so.backward()
    node_scopes = [s for s in scopes if so.layers[s]["out"] is not None]
    gr = sess.run(tf.gradients(net._loss, [lt["UResNet/" + s]["raw"] for s in scopes]
                               + [lt["UResNet/" + s]["out"] for s in node_scopes]), feed_dict=net.feed_dict(d, l, w))
    for s, g in zip(scopes, gr[:len(scopes)]):
        assert rel(g, cl(so.bn[s]["draw"])) < 1e-9, ("d(raw)", s)
    for s, g in zip(node_scopes, gr[len(scopes):]):
        assert rel(g, cl(so.node_grad[so.layers[s]["out"]])) < 1e-9, ("node gradient", s)
    for s, (raw, out) in zip(scopes, vals):
        L = so.layers[s]
        assert rel(raw, cl(L["raw"])) < 1e-11, s
        if L["out"] is not None:                       # simple layers and resnet_conv1: BN(+ReLU) output is a node
            assert rel(out, cl(so.nodes[L["out"]])) < 1e-11, s
        else:                                          # shortcut / resnet_conv2: BN output feeds the residual add
            want = O.batch_norm(L["raw"], so.P["UResNet/%s/BatchNorm/beta" % s])
            assert rel(out, cl(want)) < 1e-11, s
    shim.uninstall()


def test_reference_run_vectors_detect_semantic_mutations(monkeypatch):
    """The comparison has teeth: change one TF semantic in the oracle -- batch-norm epsilon, which side the
    asymmetric SAME padding goes, the concat operand order -- and the reference-run vectors reject it."""
    z = np.load(os.path.join(GOLD, "gref2d_f4s3.npz"))

    def deviation():
        _, _, r = _oracle_on(z)
        return rel(r["logits"], z["logits"])
    assert deviation() < 2e-7
    with monkeypatch.context() as m:
        m.setattr(O, "BN_EPS", 1e-5)                      # the dumped older revision's epsilon
        assert deviation() > 1e-3
    with monkeypatch.context() as m:
        sp = O.same_pads
        m.setattr(O, "same_pads", lambda n, k, s: sp(n, k, s)[::-1])
        assert deviation() > 1e-1
    with monkeypatch.context() as m:
        cat = torch.cat
        m.setattr(torch, "cat", lambda ts, dim=0: cat(list(ts)[::-1], dim=dim))
        assert deviation() > 1e-1
    assert deviation() < 2e-7
