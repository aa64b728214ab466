"""Whole-path parity (GPU): the reference-facing classes (uresnet / ssnet_base API) driving the
CUDA library vs the CPU oracle and the committed golden vectors.

Tolerances (BASELINE.json north_star): logits and gradients within 1e-3 relative in fp32,
2e-2 in bf16 -- measured per tensor as max|a-b| / max|b| (SURVEY.md 7 'Hard parts') --
and argmax agreement on >= 99.9 % of nonzero pixels.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import uresnet_oracle as O
from uresnet_b200 import synthetic_io as S
from uresnet_b200.uresnet import uresnet

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {"f32": 1e-3, "bf16": 2e-2, "mix16": 2e-2}
# Gradients in the 16-bit modes.  The gradient of this network is an ill-conditioned function of
# its parameters: perturbing the fp64 oracle's OWN weights by 2^-11 relative (half an ulp of fp16 /
# TF32 operands) moves most of its gradient tensors by 5-16 % while logits move by 1.5e-3 -- measured,
# not asserted from prose: oracle/conditioning.py, tests/test_oracle.py::test_gradient_conditioning.
# No 10-bit-operand evaluation can therefore track fp64 gradients END TO END at 2e-2.  What is
# asserted for mix16:  forward (logits, loss, argmax) against the fp64 oracle at 2e-2;  every step of
# the composed schedule at 2e-2 on identical inputs (teacher forcing, tests/test_parity_gpu.py);  and
# end-to-end gradients within GRAD_SLACK x the oracle's measured sensitivity for the very inputs of
# the test (per tensor: max over draws of max|g(p(1+2^-11 u)) - g(p)| / max|g(p)|; median likewise).
# GRAD_SLACK: the sensitivity run perturbs the parameters alone by 2^-11, mix16 also stores every
# gradient tensor in bf16 (2^-9: a factor 2 by the measured ~sqrt(eps) response) and the per-tensor
# maximum over 3 draws is itself a noisy estimate (x1.5).
GRAD_SLACK = 3.0


def oracle(P, data, label, weight, dims, B, F, St, quant=None):
    shp = [B] + list(dims[:-1])
    return O.run(P, data.reshape(shp + [1]), label.reshape(shp),
                 None if weight is None else weight.reshape(shp), base=F, num_strides=St, quant=quant)


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def cosine(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def make_net(dims, F, St, C, dtype, use_weight=True, trainable=True, impl=None, lr=None):
    net = uresnet(dims=dims, num_class=C, num_strides=St, base_num_outputs=F)
    net.construct(trainable=trainable, use_weight=use_weight, learning_rate=lr, dtype=dtype, impl=impl)
    return net


def sensitivity(P, data, label, weight, dims, B, F, St, ref):
    from oracle.conditioning import gradient_sensitivity
    shp = [B] + list(dims[:-1])
    return gradient_sensitivity(P, data.reshape(shp + [1]), label.reshape(shp),
                                None if weight is None else weight.reshape(shp), base=F, num_strides=St,
                                eps=2.0 ** -11, draws=3, ref=ref)


def check_against(net, P, data, label, weight, ref, dtype, B, dims, sens=None):
    """ref: fp64 oracle result; sens: the oracle's gradient sensitivity (needed for the 16-bit modes)."""
    net.set_variables(P)
    net.zero_gradients(None)
    res, doc = net.accum_gradients(None, data, label, weight)
    assert doc == ['', 'loss', 'acc. all', 'acc. nonzero']
    tol = TOL[dtype]
    logits = net.logits(B).cpu().numpy()
    assert relerr(logits, ref["logits"]) < tol, ("logits", relerr(logits, ref["logits"]))
    assert abs(res[1] - ref["loss"]) <= tol * abs(ref["loss"]), (res[1], ref["loss"])
    # argmax agreement on nonzero pixels >= 99.9 %
    nz = data.reshape(B, -1) > 0
    pa = logits.reshape(B, -1, logits.shape[-1]).argmax(-1)[nz]
    pb = np.asarray(ref["logits"]).reshape(B, -1, logits.shape[-1]).argmax(-1)[nz]
    assert (pa == pb).mean() >= 0.999 or (pa != pb).sum() <= 2, (pa != pb).sum()
    g = net.get_variables("grads")
    if dtype == "f32":
        assert abs(res[2] - ref["acc_all"]) < 2e-3 and abs(res[3] - ref["acc_nonzero"]) < 2e-2
        for k, gr in ref["grads"].items():
            assert relerr(g[k], gr) < tol, (k, relerr(g[k], gr))
    else:
        e = {k: relerr(g[k], gr) for k, gr in ref["grads"].items()}
        med = np.median(list(e.values()))
        assert med < max(tol, GRAD_SLACK * sens["median"]), (med, sens["median"])
        for k in ref["grads"]:
            assert e[k] < max(tol, GRAD_SLACK * max(sens["grads"][k], sens["median"])), (k, e[k], sens["grads"][k])


@pytest.mark.parametrize("dtype", ["f32", "mix16"])
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "g*_f4s*.npz"))))
def test_golden_small(path, dtype):
    z = np.load(path)
    dims = [int(v) for v in z["dims"]]
    B, F, St, C = int(z["batch"]), int(z["base"]), int(z["strides"]), int(z["num_class"])
    names = [str(n) for n in z["grad_names"]]
    P = {k: z["p%03d" % i] for i, k in enumerate(names)}
    ref = dict(logits=z["logits"], loss=float(z["loss"]), acc_all=float(z["acc_all"]),
               acc_nonzero=float(z["acc_nonzero"]), grads={k: z["g%03d" % i] for i, k in enumerate(names)})
    net = make_net(dims, F, St, C, dtype, use_weight=bool(z["use_weight"]))
    assert [n for n, _ in net.trainable_variables()] == names
    label = z["label"].astype(np.float32)
    sens = None
    if dtype != "f32":
        w = z["weight"] if bool(z["use_weight"]) else None
        full = oracle(P, z["data"], label, w, dims, B, F, St)
        sens = sensitivity(P, z["data"], label, w, dims, B, F, St, full)
    check_against(net, P, z["data"], label, z["weight"], ref, dtype, B, dims, sens)


@pytest.mark.parametrize("dtype", ["f32", "mix16"])
def test_golden_default_topology(dtype):
    """F=16, S=5 (the reference's default net) on 64x64, unweighted loss; weights from the seed."""
    z = np.load(os.path.join(GOLD, "g2d_f16s5.npz"))
    dims = [int(v) for v in z["dims"]]
    B = int(z["batch"])
    P = O.init_params(2, seed=int(z["param_seed"]))
    net = make_net(dims, 16, 5, 3, dtype, use_weight=False)
    net.set_variables(P)
    net.zero_gradients(None)
    res, _ = net.accum_gradients(None, z["data"], z["label"].astype(np.float32))
    tol = TOL[dtype]
    lg = net.logits(B).cpu().numpy()
    assert relerr(lg, z["logits"]) < tol
    assert abs(res[1] - float(z["loss"])) <= tol * float(z["loss"])
    g = net.get_variables("grads")
    names = [str(n) for n in z["grad_names"]]
    l2 = np.array([np.sqrt((g[k].astype(np.float64) ** 2).sum()) for k in names])
    if dtype == "f32":
        np.testing.assert_allclose(l2, z["grad_l2"], rtol=5 * tol, atol=1e-6 * z["grad_l2"].max())
    else:   # gradient norms within the measured sensitivity of the oracle's own gradients (GRAD_SLACK note above)
        shp = [B] + dims[:-1]
        lab = z["label"].astype(np.float32)
        full = O.run(P, z["data"].reshape(shp + [1]), lab.reshape(shp), None)
        sens = sensitivity(P, z["data"], lab, None, dims, B, 16, 5, full)
        band = GRAD_SLACK * max(sens["worst"], 2e-2)
        np.testing.assert_allclose(l2, z["grad_l2"], rtol=band, atol=1e-3 * z["grad_l2"].max())


@pytest.mark.parametrize("dtype", ["f32", "mix16"])
@pytest.mark.parametrize("dims,B,F,St", [([64, 64, 1], 2, 8, 4), ([32, 32, 32, 1], 1, 4, 3), ([48, 80, 1], 3, 4, 2)])
def test_live_oracle(dims, B, F, St, dtype):
    nd = len(dims) - 1
    data, label, weight = S.make_batch(B, dims, seed=99)
    weight /= weight.sum(axis=1, keepdims=True)
    P = O.init_params(nd, base=F, num_strides=St, seed=5)
    ref = oracle(P, data, label, weight, dims, B, F, St)
    sens = sensitivity(P, data, label, weight, dims, B, F, St, ref) if dtype != "f32" else None
    net = make_net(dims, F, St, 3, dtype)
    check_against(net, P, data, label, weight, ref, dtype, B, dims, sens)


def test_all_bf16_mode_forward():
    """The all-bf16 storage mode: forward held to a looser bound (8-bit mantissa storage
    decorrelates from fp64 by ~2 % over 58 layers on these small nets)."""
    dims, B, F, St = [64, 64, 1], 2, 8, 4
    data, label, weight = S.make_batch(B, dims, seed=99)
    weight /= weight.sum(axis=1, keepdims=True)
    P = O.init_params(2, base=F, num_strides=St, seed=5)
    ref = oracle(P, data, label, weight, dims, B, F, St)
    net = make_net(dims, F, St, 3, "bf16")
    net.set_variables(P)
    net.zero_gradients(None)
    res, _ = net.accum_gradients(None, data, label, weight)
    assert relerr(net.logits(B).cpu().numpy(), ref["logits"]) < 4e-2
    assert abs(res[1] - ref["loss"]) < 2e-2 * ref["loss"]
    g = net.get_variables("grads")
    assert cosine(g["UResNet/conv1/weights"], ref["grads"]["UResNet/conv1/weights"]) > 0.98


def test_accumulate_zero_apply_and_adam():
    """ssnet.py:76-79: accum += grad per minibatch (a SUM), Adam on (accum, var); TF-Adam formula."""
    dims, B, F, St = [32, 32, 1], 2, 4, 2
    data, label, weight = S.make_batch(2 * B, dims, seed=3)
    weight /= weight.sum(axis=1, keepdims=True)
    P = O.init_params(2, base=F, num_strides=St, seed=9)
    net = make_net(dims, F, St, 3, "f32", lr=0.01)
    net.set_variables(P)
    net.zero_gradients(None)
    shp = [B] + dims[:-1]
    gsum = None
    for j in range(2):
        sl = slice(j * B, (j + 1) * B)
        net.accum_gradients(None, data[sl], label[sl], weight[sl])
        r = O.run(P, data[sl].reshape(shp + [1]), label[sl].reshape(shp), weight[sl].reshape(shp), base=F, num_strides=St)
        gsum = r["grads"] if gsum is None else {k: gsum[k] + r["grads"][k] for k in gsum}
    g = net.get_variables("grads")
    for k in gsum:
        assert relerr(g[k], gsum[k]) < 1e-3, k
    net.apply_gradients(None)
    after = net.get_variables()
    for k in gsum:
        var, m, v = P[k].copy(), np.zeros_like(P[k]), np.zeros_like(P[k])
        O.adam_tf(var, g[k], m, v, t=1, lr=0.01)
        np.testing.assert_allclose(after[k], var, rtol=0, atol=2e-6)
    net.zero_gradients(None)
    assert all(np.all(v == 0) for v in net.get_variables("grads").values())


def test_inference_and_run_test_api():
    dims, B, F, St = [32, 32, 1], 2, 4, 2
    data, label, weight = S.make_batch(B, dims, seed=4)
    P = O.init_params(2, base=F, num_strides=St, seed=9)
    shp = [B] + dims[:-1]
    ref = O.run(P, data.reshape(shp + [1]), label.reshape(shp), None, base=F, num_strides=St, want_grads=False)
    net = make_net(dims, F, St, 3, "f32", use_weight=False, trainable=False)
    net.set_variables(P)
    out = net.inference(None, data)
    assert len(out) == 1 and out[0].shape == (B, 32, 32, 3)
    assert relerr(out[0], ref["softmax"]) < 1e-3
    sm, acc_all, acc_nz = net.inference(None, data, label)
    assert abs(acc_all - ref["acc_all"]) < 2e-3
    # shower/track map of the analysis step (ssnet_trainval.py:269-287), computed on the device
    sm2, a2, nz2, dec = net.inference(None, data, label, with_decision=True)
    want = O.ssnet_decision(sm2, data.reshape(shp))
    assert dec.shape == tuple(shp) and dec.dtype == np.float32 and np.array_equal(dec, want)
    assert set(np.unique(dec)) <= {0.0, 1.0, 2.0} and np.all(dec[data.reshape(shp) <= 1.0] == 0)
    dec2, a3, nz3 = net.decision(None, data, label)
    assert np.array_equal(dec2, dec) and a3 == a2 and (nz3 == nz2 or (np.isnan(nz3) and np.isnan(nz2)))
    assert np.array_equal(net.decision(None, data)[0], dec)
    (loss, a, nz), doc = net.run_test(None, data, label)
    assert doc == ['loss', 'acc. all', 'acc. nonzero'] and abs(loss - ref["loss"]) < 1e-3 * ref["loss"]
    # weighted net refuses to run without weights (ssnet.py:143-146)
    netw = make_net(dims, F, St, 3, "f32", use_weight=True, trainable=False)
    with pytest.raises(TypeError):
        netw.inference(None, data, label)
    bad = label.copy()
    bad[0, 0] = 7
    with pytest.raises(ValueError):
        net.run_test(None, data, bad)
    # the C ABI reports the same condition through urn_check_labels (device flag of the last loss call)
    h, ws = net._handle(B)
    assert net._lib.urn_check_labels(h, ws.data_ptr(), net._stream()) != 0
    assert b"label out of range" in net._lib.urn_last_error()
    net.run_test(None, data, label)
    assert net._lib.urn_check_labels(h, ws.data_ptr(), net._stream()) == 0


def test_dims_validation():
    with pytest.raises(NotImplementedError):
        uresnet(dims=[32, 32], num_class=3)
    net = uresnet(dims=[30, 32, 1], num_class=3)
    with pytest.raises(ValueError):
        net.construct(trainable=False, use_weight=False, dtype="f32")


def test_full_size_properties_2d():
    """BASELINE configs[0/1] shape (512x512, F=16, S=5) at small batch: size-independent properties --
    softmax rows sum to 1, per-channel BN outputs of the head have zero mean / unit variance
    up to beta, dlogits sums to ~0 per pixel, gradient accumulation is linear."""
    dims, B = [512, 512, 1], 2
    data, label, weight = S.make_batch(B, dims, seed=1234)
    weight /= weight.sum(axis=1, keepdims=True)
    net = make_net(dims, 16, 5, 3, "mix16")
    net.zero_gradients(None)
    r1, _ = net.accum_gradients(None, data, label, weight)
    g1 = net._accum.clone()
    r2, _ = net.accum_gradients(None, data, label, weight)
    assert np.isfinite(r1[1]) and r1[1] == r2[1]                      # deterministic
    assert torch.allclose(net._accum, 2 * g1, rtol=1e-5, atol=1e-8)   # accum += grad
    lg = net.logits(B).reshape(-1, 3).double()
    beta = torch.tensor(net.get_variables()["UResNet/conv2/BatchNorm/beta"], dtype=torch.float64, device=lg.device)
    assert torch.allclose(lg.mean(0), beta, atol=1e-3)               # beta is zero-initialised
    var = lg.var(0, unbiased=False)
    assert torch.all(var <= 1.0 + 1e-3)
    net2 = make_net(dims, 16, 5, 3, "mix16", use_weight=False, trainable=False)
    sm = net2.inference(None, data)[0]
    assert np.allclose(sm.sum(-1), 1.0, atol=1e-5)


def test_trainval_driver_train_checkpoint_and_ana(tmp_path, capsys):
    """The drop-in driver (lib/ssnet_trainval.py mirror): cfg override, train_step loop with the
    reference's report format, npz checkpoint with TF variable names, LOAD_FILE / iteration from
    the file-name suffix, ana_step result dictionary."""
    from uresnet_b200.ssnet_trainval import ssnet_trainval
    cfg = tmp_path / "train.cfg"
    cfg.write_text("\n".join([
        "NUM_CLASS 3", "BASE_NUM_FILTERS 4", "MAIN_INPUT_CONFIG 'config/input_train2d.cfg'",
        "LOGDIR '%s'" % (tmp_path / "log"), "SAVE_FILE '%s'" % (tmp_path / "w" / "uresnet"),
        "ITERATIONS 4", "MINIBATCH_SIZE 2", "NUM_MINIBATCHES 2", "TRAIN True", "USE_WEIGHTS True",
        "CHECKPOINT_STEPS 2", "SUMMARY_STEPS 2", "REPORT_STEPS 1", "LEARNING_RATE 0.001",
        "KEYWORD_DATA 'main_data'", "KEYWORD_LABEL 'main_label'", "KEYWORD_WEIGHT 'main_weight'",
        "SYNTHETIC_DIMS [64,64,1]", "DTYPE 'f32'"]))
    t = ssnet_trainval()
    t.override_config(str(cfg))
    t.initialize()
    t.batch_process()
    out = capsys.readouterr().out
    assert "@ iteration 0 LR 0.001 Mem" in out and "Train set: loss=" in out and "acc. nonzero=" in out
    assert t.current_iteration() == 3
    ck = sorted((tmp_path / "w").glob("uresnet-*.npz"))
    assert [p.name for p in ck] == ["uresnet-1.npz", "uresnet-3.npz"]
    z = np.load(ck[-1])
    assert z["UResNet/conv0/weights"].shape == (3, 3, 1, 4) and "UResNet/conv0/BatchNorm/moving_variance" in z
    assert (tmp_path / "log" / "train" / "scalars.jsonl").read_text().count("\n") == 2
    ea = pytest.importorskip("tensorboard.backend.event_processing.event_accumulator")
    acc = ea.EventAccumulator(str(tmp_path / "log" / "train"))
    acc.Reload()
    assert set(acc.Tags()["scalars"]) == {"accuracy_all", "accuracy_nonzero", "loss"}      # ssnet.py:84-86
    assert [e.step for e in acc.Scalars("loss")] == [0, 2] and "data_example" in acc.Tags()["images"]
    saved = t._net.get_variables()
    t.reset()
    # inference from the checkpoint
    cfg2 = tmp_path / "ana.cfg"
    cfg2.write_text("\n".join([
        "NUM_CLASS 3", "BASE_NUM_FILTERS 4", "MAIN_INPUT_CONFIG 'config/input_ana2d.cfg'", "LOGDIR ''", "SAVE_FILE ''",
        "LOAD_FILE '%s'" % str(ck[-1])[:-4], "AVOID_LOAD_PARAMS ['UResNet/conv2/BatchNorm/beta:0']",
        "MINIBATCH_SIZE 2", "ITERATIONS 2", "TRAIN False", "USE_WEIGHTS False",
        "KEYWORD_DATA 'main_data'", "KEYWORD_LABEL 'main_label'", "SYNTHETIC_DIMS [64,64,1]", "DTYPE 'f32'"]))
    a = ssnet_trainval()
    a.override_config(str(cfg2))
    a.initialize()
    assert a.current_iteration() == 3                       # recovered from the '-3' suffix
    v = a._net.get_variables()
    assert np.array_equal(v["UResNet/conv1/weights"], saved["UResNet/conv1/weights"])
    assert np.all(v["UResNet/conv2/BatchNorm/beta"] == 0)   # skipped by AVOID_LOAD_PARAMS
    res = a.ana_step()
    assert res["softmax"].shape == (2, 64, 64, 3) and res["input"].shape == (2, 64, 64, 1)
    assert 0.0 <= res["acc_all"] <= 1.0 and set(res) == {"entries", "input", "label", "softmax", "acc_all", "acc_nonzero"}
    a.reset()
    # analysis with an output: per-batch npz with the shower/track map (larcv output stand-in)
    cfg3 = tmp_path / "ana_out.cfg"
    cfg3.write_text(cfg2.read_text() + "\nANA_OUTPUT_CONFIG '%s'" % (tmp_path / "ana_result.cfg"))
    b = ssnet_trainval()
    b.override_config(str(cfg3))
    b.initialize()
    r1 = b.ana_step()                                        # interactive: softmax + map
    z1 = np.load(tmp_path / "ana_result_out" / ("ssnet_%06d.npz" % b.current_iteration()))
    assert np.array_equal(z1["ssnet"], O.ssnet_decision(r1["softmax"], r1["input"][..., 0]))
    b.ana_step(batch_mode=True)                              # batch mode: map only, no softmax download
    z2 = np.load(tmp_path / "ana_result_out" / ("ssnet_%06d.npz" % b.current_iteration()))
    assert z2["ssnet"].shape == (2, 64, 64) and set(np.unique(z2["ssnet"])) <= {0.0, 1.0, 2.0}
    assert len(z2["entries"]) == 2
    b.reset()


@pytest.mark.parametrize("dtype", ["f32", "mix16"])
def test_driver_reproduces_reference_driver_transcript(dtype, capsys):
    """tests/golden/refdriver_train.json is the transcript of the REFERENCE'S OWN DRIVER (lib/ssnet_trainval.py run
    unmodified over oracle/tf1_shim.py, float64) training the default-depth net on the synthetic source: 6
    iterations x 2 minibatches, a test stream evaluated at every report.  The product's driver with the same
    configuration -- same TF_RANDOM_SEED, hence the same initial filters (tests/test_host_cpu.py) -- must print
    the same report: every loss within 5e-3 (f32) / 2e-2 (mix16), accuracies within a few pixels' worth."""
    import importlib.util
    import json
    from uresnet_b200.ssnet_trainval import ssnet_trainval
    spec = importlib.util.spec_from_file_location("mkdrv", os.path.join(GOLD, "make_reference_driver_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = json.load(open(os.path.join(GOLD, "refdriver_train.json")))
    t = ssnet_trainval()
    for k, v in g["cfg"].items():
        setattr(t._cfg, k, v)
    t._cfg.SYNTHETIC_DIMS, t._cfg.SYNTHETIC_SEED, t._cfg.DTYPE = list(g["dims"]), int(g["synthetic_seed"]), dtype
    t.initialize()
    v0 = t._net.get_variables()
    assert list(v0.keys()) == g["names"]
    for i, k in enumerate(g["names"]):
        assert np.sqrt((v0[k].astype(np.float64) ** 2).sum()) == pytest.approx(g["init_l2"][i], rel=1e-6, abs=1e-12), k
    capsys.readouterr()
    t.batch_process()
    reps = mk.parse_transcript(capsys.readouterr().out)
    v1 = t._net.get_variables()
    t.reset()
    # measured on B200: f32 2.1e-3 / 1.2e-3 / 8.3e-3, mix16 5.5e-3 / 3.3e-3 / 1.2e-2 (the 2 x 2 x 2-value batch norms at
    # the bottom of a five-stride net on 64 x 64 images amplify rounding; Adam's first steps are sign-like)
    tol_loss, tol_all, tol_nz = {"f32": (5e-3, 5e-3, 2e-2), "mix16": (2e-2, 2e-2, 6e-2)}[dtype]
    assert [r["iteration"] for r in reps] == [r["iteration"] for r in g["reports"]] == list(range(6))
    worst = [0.0, 0.0, 0.0]
    for r, w in zip(reps, g["reports"]):
        assert r["lr"] == w["lr"]
        for key in ("train", "test"):
            d = [abs(r[key][0] - w[key][0]) / w[key][0], abs(r[key][1] - w[key][1]), abs(r[key][2] - w[key][2])]
            worst = [max(a, b) for a, b in zip(worst, d)]
    upd = np.array([np.sqrt(((v1[k].astype(np.float64) - v0[k]) ** 2).sum()) for k in g["names"]])
    print("\ndriver vs reference-driver transcript [%s]: worst loss deviation %.2e, acc. all %.2e, acc. nonzero %.2e; "
          "update norms vs reference: median ratio %.4f" % (dtype, worst[0], worst[1], worst[2],
                                                            float(np.median(upd / np.array(g["update_l2"])))))
    assert worst[0] < tol_loss and worst[1] < tol_all and worst[2] < tol_nz, worst
    # six Adam applies moved every tensor about as far as they moved the reference's
    np.testing.assert_allclose(upd, g["update_l2"], rtol=0.25)


@pytest.mark.parametrize("dims,B", [([32, 64, 1], 2), ([8, 16, 16, 1], 1), ([24, 40, 1], 13)])   # batch >= 12: images-as-rows plan
def test_padded_head_matches_cuda_core_head(dims, B, monkeypatch):
    """The head conv2 (Cout = 3) runs on the tensor cores with its filter zero-padded to 16 output
    channels (forward stores 3; d(raw) is kept at pitch 16 with zero pad channels so that dgrad /
    wgrad are plain 16 -> 16 halo launches).  It must reproduce the dedicated CUDA-core head
    kernels (URN_NO_PADDED_HEAD=1) on the same 16-bit operands: logits to accumulation-order noise,
    every gradient to bf16 storage precision."""
    F, St = 16, 2
    nd = len(dims) - 1
    data, label, weight = S.make_batch(B, dims, seed=21)
    weight /= weight.sum(axis=1, keepdims=True)
    P = O.init_params(nd, base=F, num_strides=St, seed=3)
    out = {}
    for tag, env in (("cuda_core", "1"), ("padded", None)):
        if env:
            monkeypatch.setenv("URN_NO_PADDED_HEAD", env)
        else:
            monkeypatch.delenv("URN_NO_PADDED_HEAD", raising=False)
        net = make_net(dims, F, St, 3, "mix16")
        net.set_variables(P)
        net.zero_gradients(None)
        res, _ = net.accum_gradients(None, data, label, weight)
        out[tag] = (net.logits(B).cpu().numpy(), res, net.get_variables("grads"))
    la, ra, ga = out["cuda_core"]
    lb, rb, gb = out["padded"]
    assert relerr(lb, la) < 1e-5
    assert abs(ra[1] - rb[1]) <= 1e-5 * abs(ra[1])
    # two 16-bit realisations of an ill-conditioned backward pass (see GRAD_SLACK above): a 1e-6
    # difference in the head's statistics moves single beta gradients by a few per cent
    for k in ga:
        assert relerr(gb[k], ga[k]) < 8e-2, (k, relerr(gb[k], ga[k]))
    # the tensor-core filter gradient multiplies the bf16 shadow of the activations (2^-9 rounding),
    # the CUDA-core one their fp16 copy
    assert relerr(gb["UResNet/conv2/weights"], ga["UResNet/conv2/weights"]) < 8e-3


def test_full_size_properties_3d():
    """BASELINE configs[3] shape (192^3, F=16, S=5) at batch 1: size-independent properties of the
    full-size 3D path (27-tap halo kernels with N-concatenated depth taps, 8-class conv-transposes):
    determinism, linear gradient accumulation, BN output statistics of the head, finite values."""
    dims, B = [192, 192, 192, 1], 1
    data, label, weight = S.make_batch(B, dims, seed=1234)
    weight /= weight.sum(axis=1, keepdims=True)
    net = make_net(dims, 16, 5, 3, "mix16")
    net.zero_gradients(None)
    r1, _ = net.accum_gradients(None, data, label, weight)
    g1 = net._accum.clone()
    r2, _ = net.accum_gradients(None, data, label, weight)
    assert np.isfinite(r1[1]) and r1[1] == r2[1]
    assert torch.isfinite(g1).all()
    assert torch.allclose(net._accum, 2 * g1, rtol=1e-5, atol=1e-8)
    lg = net.logits(B).reshape(-1, 3).double()
    assert torch.allclose(lg.mean(0), torch.zeros(3, dtype=torch.float64, device=lg.device), atol=1e-3)
    assert torch.all(lg.var(0, unbiased=False) <= 1.0 + 1e-3)
    assert 0.0 <= r1[2] <= 1.0


def test_cuda_graph_replay_is_bit_identical():
    """The forward and loss+backward launch sequences are replayed as CUDA graphs from the third
    call on: accumulators, loss and Adam-updated variables must equal the eager run bit for bit."""
    dims, B, F, St = [64, 96, 1], 2, 16, 3
    data, label, weight = S.make_batch(B, dims, seed=5)
    weight /= weight.sum(axis=1, keepdims=True)
    P = O.init_params(2, base=F, num_strides=St, seed=2)
    out = []
    for graphs in (False, True):
        net = make_net(dims, F, St, 3, "mix16", lr=1e-3)
        net._use_graphs = graphs
        net.set_variables(P)
        losses = []
        for _ in range(5):
            net.zero_gradients(None)
            res, _ = net.accum_gradients(None, data, label, weight)
            net.apply_gradients(None)
            losses.append(float(res[1]))
        assert (len(net._graphs) == 2) == graphs
        out.append((losses, net._accum.clone(), net._params.clone()))
    assert out[0][0] == out[1][0]
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
