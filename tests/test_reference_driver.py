"""Differential test of the drop-in DRIVER against the reference's own driver running.

``/root/reference/lib/ssnet_trainval.py`` (with ``lib/config.py`` and the model code) is imported unmodified
over ``oracle/tf1_shim.py`` and a stand-in ``larcv`` whose ``larcv_threadio`` is this repo's seeded synthetic
source; ``uresnet_b200.ssnet_trainval`` is run on the SAME source with its network class replaced by an
adapter that serves the product's network API from the same reference net (so no GPU is needed and any
difference is a difference of the DRIVERS' control flow): iteration numbering, the NUM_MINIBATCHES loop and the
mean over minibatches in the report, per-event weight normalisation, when the test stream advances and is
evaluated, report / summary / checkpoint cadence, checkpoint naming and retention, resume from ``LOAD_FILE``
(iteration from the suffix, AVOID_LOAD_PARAMS), "Finished training", and the dict ``ana_step`` returns.

Known, intended difference: checkpoints are ``.npz`` files (suffix masked in the comparison).
Build container only: the reference tree is not on the GPU box.
"""
import json
import os
import re

import numpy as np
import pytest
import torch

from uresnet_b200.synthetic_io import synthetic_threadio

HAVE_REF = os.path.exists("/root/reference/lib/ssnet_trainval.py")
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="reference tree not on this box")

DIMS = [32, 32, 1]
SEED = 4242


class _threadio(synthetic_threadio):
    """what ``larcv.dataloader2.larcv_threadio`` resolves to for the reference driver: the product's synthetic
    source, configured the way ``uresnet_b200.ssnet_trainval._make_io`` configures it"""

    def configure(self, cfg):
        c = dict(cfg)
        test = c.get("filler_name") == "TestIO"
        c.update(dims=DIMS, seed=SEED + (500000 if test else 0), rank=0,
                 keyword_data="test_data" if test else "data", keyword_label="test_label" if test else "label",
                 keyword_weight="test_weight" if test else "weight")
        synthetic_threadio.configure(self, c)


class _ref_backed_net(object):
    """The network API the product driver uses (u-resnet_b200/ssnet.py), served by the reference's net over
    the shim."""

    def __init__(self, dims, num_class, num_strides=5, base_num_outputs=16, debug=False):
        self._args = dict(dims=list(dims), num_class=num_class, num_strides=num_strides,
                          base_num_outputs=base_num_outputs, debug=debug)
        self._dims = np.array(dims, np.int32)

    def construct(self, trainable=True, use_weight=True, learning_rate=None, dtype=None, impl=None, seed=0):
        from oracle import tf1_shim as shim
        shim.reset_default_graph(torch.float64)
        ref_uresnet = shim.load_reference()
        import tensorflow as tf
        self._tf = tf
        tf.set_random_seed(seed)
        self._net = ref_uresnet(**self._args)
        if trainable:
            self._net.construct(trainable=trainable, use_weight=use_weight, learning_rate=learning_rate)
            self._opt = self._net._opt
        else:
            self._net.construct(trainable=trainable, use_weight=use_weight)
        self._sess = tf.Session()
        self._sess.run(tf.global_variables_initializer())
        self._last = None

    def pinned_empty(self, shape):
        return np.empty(shape, np.float32)

    def stage_next(self, *a):
        pass

    def zero_gradients(self, sess=None):
        return self._net.zero_gradients(self._sess)

    def accum_gradients(self, sess, input_data, input_label, input_weight=None, fetch=True):
        self._last = self._net.accum_gradients(self._sess, input_data, input_label, input_weight)
        return self._last if fetch else None

    def fetch_metrics(self, batch):
        return self._last

    def apply_gradients(self, sess=None):
        return self._net.apply_gradients(self._sess)

    def run_test(self, sess, d, l, w=None):
        return self._net.run_test(self._sess, d, l, w)

    def make_summary(self, sess, d, l, w=None):
        return self._net.make_summary(self._sess, d, l, w)

    def inference(self, sess, input_data, input_label=None, with_decision=False):
        return self._net.inference(self._sess, input_data, input_label)

    def _model_variables(self):
        return [v for v in self._tf.trainable_variables()]

    def trainable_variables(self):
        return [(v.op.name, tuple(v.shape.as_list())) for v in self._model_variables()]

    def get_variables(self, what="params"):
        return {v.op.name: v.value.detach().numpy().copy() for v in self._model_variables()}

    def set_variables(self, values, strict=True):
        by = {v.op.name: v for v in self._model_variables()}
        for k, v in values.items():
            by[k].load(v)


def _configure(cfg, tmp, tag, **over):
    base = dict(MAIN_INPUT_CONFIG="main.cfg", TEST_INPUT_CONFIG="test.cfg", BASE_NUM_FILTERS=2, MINIBATCH_SIZE=2,
                NUM_MINIBATCHES=3, TEST_BATCH_SIZE=3, ITERATIONS=7, LEARNING_RATE=2e-3, TF_RANDOM_SEED=77,
                TRAIN=True, USE_WEIGHTS=True, REPORT_STEPS=2, SUMMARY_STEPS=3, CHECKPOINT_STEPS=2, CHECKPOINT_NMAX=2,
                CHECKPOINT_NHOUR=0.4, KEYWORD_TEST_DATA="test_data", KEYWORD_TEST_LABEL="test_label",
                KEYWORD_TEST_WEIGHT="test_weight", LOGDIR=str(tmp / ("log_" + tag)),
                SAVE_FILE=str(tmp / ("ckpt_" + tag) / "uresnet"), LOAD_FILE="", AVOID_LOAD_PARAMS=[])
    base.update(over)
    for k, v in base.items():
        assert hasattr(cfg, k), k
        setattr(cfg, k, v)


def _run_reference(tmp, capsys, steps=None, **over):
    from oracle import tf1_shim as shim
    shim.reset_default_graph(torch.float64)
    drv = shim.load_reference_driver(_threadio)()
    _configure(drv._cfg, tmp, "ref", **over)
    drv.initialize()
    out = {}
    if steps is None:
        drv.batch_process()
    else:
        out["ana"] = [drv.ana_step() for _ in range(steps)]
    import tensorflow as tf
    out["iteration"] = drv.current_iteration()
    out["variables"] = {v.op.name: v.value.detach().numpy().copy() for v in tf.global_variables()}
    out["summaries"] = {k: list(w.records) for k, w in shim.FileWriter.by_dir.items()}
    out["stdout"] = capsys.readouterr().out
    drv.reset()
    shim.uninstall()
    return out


def _run_product(tmp, capsys, monkeypatch, steps=None, **over):
    from oracle import tf1_shim as shim
    from uresnet_b200 import ssnet_trainval as mod
    monkeypatch.setattr(mod, "uresnet", _ref_backed_net)
    drv = mod.ssnet_trainval()
    monkeypatch.setattr(drv, "report_memory", lambda: 0.0)
    _configure(drv._cfg, tmp, "our", SYNTHETIC_DIMS=DIMS, SYNTHETIC_SEED=SEED, **over)
    drv.initialize()
    out = {}
    if steps is None:
        drv.batch_process()
    else:
        out["ana"] = [drv.ana_step() for _ in range(steps)]
    out["iteration"] = drv.current_iteration()
    out["variables"] = drv._net.get_variables()
    out["stdout"] = capsys.readouterr().out
    drv.reset()
    shim.uninstall()
    return out


def _lines(text, tag):
    """report lines with what legitimately differs masked: wall-clock stamp, device memory, checkpoint suffix and
    directory tag"""
    keep = []
    for ln in text.splitlines():
        ln = re.sub(r"Mem \S+ @ .*$", "Mem * @ *", ln)
        ln = ln.replace(".npz", "").replace("_" + tag, "_X")
        keep.append(ln)
    return keep


def test_training_drivers_agree(tmp_path, capsys, monkeypatch):
    ref = _run_reference(tmp_path, capsys)
    our = _run_product(tmp_path, capsys, monkeypatch)
    a, b = _lines(ref["stdout"], "ref"), _lines(our["stdout"], "our")
    assert a == b, "\n".join(["--- reference driver"] + a + ["--- product driver"] + b)
    # what the common transcript must contain (so that an empty == empty cannot pass)
    assert [ln for ln in a if ln.startswith("@ iteration")] == ["@ iteration %d LR 0.002 Mem * @ *" % i for i in (0, 2, 4, 6)]
    assert sum(ln.startswith("Train set: loss=") for ln in a) == 4 and sum(ln.startswith("Test set: loss=") for ln in a) == 4
    assert [ln for ln in a if ln.startswith("saved @")] == ["saved @ %s-%d" % (tmp_path / "ckpt_X" / "uresnet", i) for i in (1, 3, 5)]
    assert ref["iteration"] == our["iteration"] == 6
    # same parameters after 7 iterations x 3 minibatches
    for k, v in our["variables"].items():
        assert np.array_equal(v, ref["variables"][k]), k
    # summaries: same steps, same scalars (reference: FileWriter records; product: scalars.jsonl)
    for sub in ("train", "test"):
        want = ref["summaries"][str(tmp_path / "log_ref" / sub)]
        got = [json.loads(ln) for ln in open(tmp_path / "log_our" / sub / "scalars.jsonl")]
        assert [s for s, _ in want] == [g["step"] for g in got] == [0, 3, 6]
        for (_, w), g in zip(want, got):
            assert set(w) == {"loss", "accuracy_all", "accuracy_nonzero"}
            for tag, val in w.items():
                assert g[tag] == pytest.approx(val, rel=1e-12), (sub, tag)
    # retention: CHECKPOINT_NMAX = 2 newest survive in both
    left = lambda d: sorted(int(f.split("-")[-1].split(".")[0]) for f in os.listdir(tmp_path / d))     # noqa: E731
    assert left("ckpt_ref") == left("ckpt_our") == [3, 5]
    # checkpoint contents: every tensor the product stores is in the reference's checkpoint under the same name
    zr, zo = np.load(tmp_path / "ckpt_ref" / "uresnet-5.npz"), np.load(tmp_path / "ckpt_our" / "uresnet-5.npz")
    assert any(k.endswith("/moving_variance") for k in zo.files)
    for k in zo.files:
        assert np.array_equal(zo[k], zr[k]), k

    # ---- resume from the checkpoint: iteration from the file name, one variable skipped.  (SUMMARY_STEPS = 0: without a
    # LOGDIR the reference's summary step dereferences a writer that was never made, ssnet_trainval.py:218; the
    # product tolerates that combination)
    skip = "UResNet/conv2/weights"
    ref2 = _run_reference(tmp_path, capsys, LOAD_FILE=str(tmp_path / "ckpt_ref" / "uresnet-5"), ITERATIONS=8,
                          AVOID_LOAD_PARAMS=[skip + ":0"], LOGDIR="", SUMMARY_STEPS=0)
    our2 = _run_product(tmp_path, capsys, monkeypatch, LOAD_FILE=str(tmp_path / "ckpt_our" / "uresnet-5"),
                        ITERATIONS=8, AVOID_LOAD_PARAMS=[skip + ":0"], LOGDIR="", SUMMARY_STEPS=0)
    a, b = _lines(ref2["stdout"], "ref"), _lines(our2["stdout"], "our")
    assert a == b, "\n".join(["--- reference driver"] + a + ["--- product driver"] + b)
    assert any("Skipping" in ln and skip + ":0" in ln for ln in a)
    assert sum("Loading" in ln for ln in a) == 4 * 58 - 1       # weights, beta, moving_mean, moving_variance per layer
    assert "Finished training (iteration 8)" in a
    assert [ln for ln in a if ln.startswith("@ iteration")] == ["@ iteration %d LR 0.002 Mem * @ *" % i for i in (6, 8)]
    assert ref2["iteration"] == our2["iteration"] == 8
    for k, v in our2["variables"].items():
        assert np.array_equal(v, ref2["variables"][k]), k
    # the reference restores the moving statistics too; they are 0 / 1 forever (nothing updates them)
    for k, v in ref2["variables"].items():
        if k.endswith("moving_mean"):
            assert not v.any()
        if k.endswith("moving_variance"):
            assert np.all(v == 1.0)


def test_analysis_drivers_agree(tmp_path, capsys, monkeypatch):
    over = dict(TRAIN=False, USE_WEIGHTS=False, TEST_INPUT_CONFIG="", LOGDIR="", SAVE_FILE="", ITERATIONS=2)
    ref = _run_reference(tmp_path, capsys, steps=2, **over)
    our = _run_product(tmp_path, capsys, monkeypatch, steps=2, **over)
    assert ref["iteration"] == our["iteration"] == 1
    for r, o in zip(ref["ana"], our["ana"]):
        assert set(r) == set(o) == {"entries", "input", "label", "softmax", "acc_all", "acc_nonzero"}
        assert r["softmax"].shape == (2, 32, 32, 3) and r["input"].shape == (2, 32, 32, 1)
        for k in r:
            assert np.array_equal(np.asarray(r[k]), np.asarray(o[k])), k
    assert not np.array_equal(ref["ana"][0]["input"], ref["ana"][1]["input"])          # the stream advanced
    assert list(ref["ana"][1]["entries"]) == [2, 3]


def test_config_schema_defaults_are_the_reference_s():
    """lib/config.py:6-36 imported as is: every key of the reference schema exists in the product's schema with
    the same default value and type (``override`` / ``dump`` themselves rely on Python-2 ``exec`` semantics --
    assignments to function locals -- and cannot run under Python 3; tests/test_host_cpu.py covers the
    product's reader on the reference's own .cfg files)."""
    from oracle import tf1_shim as shim
    from uresnet_b200.config import ssnet_config as ours
    shim.load_reference_driver(_threadio)
    import config as ref_config
    assert ref_config.__file__.startswith("/root/reference/lib")
    ref = {k: v for k, v in ref_config.ssnet_config.__dict__.items() if k == k.upper() and not k.startswith("_")}
    assert len(ref) == 29
    mine = ours()
    for k, v in ref.items():
        assert hasattr(mine, k), k
        assert getattr(mine, k) == v and type(getattr(mine, k)) is type(v), (k, getattr(mine, k), v)
    extra = set(ours._keys()) - set(ref)
    assert extra == {"DTYPE", "KERNEL_IMPL", "SYNTHETIC_DIMS", "SYNTHETIC_SEED"}, extra
    shim.uninstall()


# ------------------------------------------------------------------------------------------------
# the committed transcript of the reference driver (tests/golden/refdriver_train.json): consumed by
# tests/test_net_gpu.py::test_driver_reproduces_reference_driver_transcript on the B200
# ------------------------------------------------------------------------------------------------
def _mkdrv():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "mkdrv", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_driver_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_driver_transcript_regenerates():
    mk = _mkdrv()
    want = json.load(open(mk.OUT))
    got = json.loads(json.dumps(mk.run()))
    assert got["reports"] == want["reports"] and got["names"] == want["names"] and got["cfg"] == want["cfg"]
    assert len(got["reports"]) == 6 and all(set(r) == {"iteration", "lr", "train", "test"} for r in got["reports"])
    for k in ("init_l2", "init_sum", "final_l2", "update_l2"):
        np.testing.assert_allclose(got[k], want[k], rtol=1e-9, atol=1e-12)
    losses = [r["train"][0] for r in want["reports"]]
    assert losses[-1] < 0.7 * losses[0]                       # it trains


class _io_manager(object):
    """stand-in for ``larcv.IOManager`` (ssnet_trainval.py:81-83,254-302): hands out the producers the analysis
    step asks for and keeps what it stores"""
    stored = []

    def __init__(self, cfg):
        self.cfg, self.entry = cfg, None

    def initialize(self):
        pass

    def finalize(self):
        pass

    def read_entry(self, e):
        self.entry = int(e)

    def save_entry(self):
        pass

    def get_data(self, kind, name):
        mgr = self

        class product(object):
            def as_vector(self):
                return self

            def front(self):
                return self

            def meta(self):
                return ("meta", mgr.entry)

            def append(self, img):
                _io_manager.stored.append((mgr.entry, kind, name, img))

            def set(self, vs, meta):
                _io_manager.stored.append((mgr.entry, kind, name, vs))
        return product()


@pytest.mark.parametrize("dims", [[32, 32, 1], [32, 32, 32, 1]])
def test_analysis_output_map_is_the_oracle_s_decision(dims, tmp_path, capsys, monkeypatch):
    """The shower / track map the reference's analysis step writes to its larcv output (ssnet_trainval.py:254-302,
    executed here with a stand-in IOManager) equals ``oracle.ssnet_decision`` of the same softmax -- the function
    the device kernel ``urn_decision`` is held to bit for bit in tests/test_net_gpu.py / test_kernels_gpu.py."""
    import sys
    from oracle import tf1_shim as shim
    from oracle import uresnet_oracle as O
    monkeypatch.setattr(sys.modules[__name__], "DIMS", dims)
    shim.reset_default_graph(torch.float64)
    cls = shim.load_reference_driver(_threadio)
    larcv = sys.modules["larcv"].larcv
    larcv.IOManager = _io_manager
    larcv.as_image2d_meta = lambda arr, meta: np.array(arr)
    larcv.as_tensor3d = lambda arr: np.array(arr)
    _io_manager.stored = []
    drv = cls()
    _configure(drv._cfg, tmp_path, "ref", TRAIN=False, USE_WEIGHTS=False, TEST_INPUT_CONFIG="", LOGDIR="", SAVE_FILE="",
               ANA_OUTPUT_CONFIG="out.cfg", BASE_NUM_FILTERS=2, MINIBATCH_SIZE=2)
    drv.initialize()
    res = drv.ana_step()
    out = capsys.readouterr().out
    drv.reset()
    shim.uninstall()
    assert len(_io_manager.stored) == 2 and out.count("Entry") == 2
    nd = len(dims) - 1
    for b, (entry, kind, name, img) in enumerate(_io_manager.stored):
        assert (entry, kind, name) == (b, "image2d" if nd == 2 else "sparse3d", "ssnet")
        want = O.ssnet_decision(res["softmax"][b], res["input"][b, ..., 0])
        assert img.dtype == np.float32 and img.shape == tuple(dims[:-1]) and np.array_equal(img, want)
        assert set(np.unique(img)) <= {0.0, 1.0, 2.0} and (img > 0).sum() == (res["input"][b, ..., 0] > 1.0).sum()
