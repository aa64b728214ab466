"""CPU tests of the host-side mirror of the reference interface: config reader, synthetic IO
source, bucket planning, class-level validation (no GPU, no compute calls)."""
import os
import time

import numpy as np
import pytest

from uresnet_b200 import dp, synthetic_io as S
from uresnet_b200.config import ssnet_config
from uresnet_b200.uresnet import uresnet

REF_CFG = "/root/reference/config"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write(tmp_path, text, name="a.cfg"):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_config_defaults_match_reference_schema():
    c = ssnet_config()
    assert (c.NUM_CLASS, c.BASE_NUM_FILTERS, c.MINIBATCH_SIZE, c.NUM_MINIBATCHES) == (3, 16, 10, 5)
    assert c.LEARNING_RATE == -1 and c.TF_RANDOM_SEED == 1234 and c.USE_WEIGHTS is True
    assert c.CHECKPOINT_NHOUR == 0.4 and c.KEYWORD_DATA == 'data'
    ref_keys = {"NUM_CLASS", "BASE_NUM_FILTERS", "MAIN_INPUT_CONFIG", "TEST_INPUT_CONFIG", "ANA_OUTPUT_CONFIG",
                "LOGDIR", "SAVE_FILE", "LOAD_FILE", "AVOID_LOAD_PARAMS", "LEARNING_RATE", "MINIBATCH_SIZE",
                "NUM_MINIBATCHES", "TEST_BATCH_SIZE", "ITERATIONS", "TF_RANDOM_SEED", "TRAIN", "DEBUG",
                "USE_WEIGHTS", "REPORT_STEPS", "SUMMARY_STEPS", "CHECKPOINT_STEPS", "CHECKPOINT_NMAX",
                "CHECKPOINT_NHOUR", "KEYWORD_DATA", "KEYWORD_LABEL", "KEYWORD_WEIGHT", "KEYWORD_TEST_DATA",
                "KEYWORD_TEST_LABEL", "KEYWORD_TEST_WEIGHT"}
    assert ref_keys <= set(ssnet_config._keys())


def test_config_override_rules(tmp_path, capsys):
    c = ssnet_config()
    c.override(write(tmp_path, "NUM_CLASS 4   # comment\nLEARNING_RATE 0.001\nthree tokens here\nUNKNOWN_KEY 1\n"
                               "SAVE_FILE './w'\nAVOID_LOAD_PARAMS ['a:0']\n"))
    assert c.NUM_CLASS == 4 and c.LEARNING_RATE == 0.001 and c.SAVE_FILE == './w' and c.AVOID_LOAD_PARAMS == ['a:0']
    out = capsys.readouterr().out
    assert "Ignoring a line: three tokens here" in out and "Ignoring a parameter in file: UNKNOWN_KEY" in out
    with pytest.raises(TypeError):
        c.override(write(tmp_path, "NUM_CLASS 4.5\n", "b.cfg"))        # type must match the default's
    with pytest.raises(IOError):
        c.override(str(tmp_path / "missing.cfg"))
    with pytest.raises(TypeError):
        c.override(write(tmp_path, "LOGDIR __import__('os')\n", "c.cfg"))   # literals only, no exec


@pytest.mark.skipif(not os.path.isdir(REF_CFG), reason="reference configs not present on this box")
@pytest.mark.parametrize("name", ["train2d.cfg", "train3d.cfg", "ana2d.cfg", "ana3d.cfg", "default2d.cfg", "ana2d_no_out.cfg"])
def test_reference_cfg_files_load(name):
    c = ssnet_config()
    c.override(os.path.join(REF_CFG, name))          # ana3d.cfg's unknown DUMP_IMAGE is tolerated
    assert c.NUM_CLASS == 3


def test_config_dump_format(capsys):
    ssnet_config().dump()
    lines = capsys.readouterr().out.splitlines()
    assert lines[0] == "NUM_CLASS........... 3"


def test_synthetic_source_interface():
    io = S.synthetic_threadio()
    io.configure({'filler_name': 'MainIO', 'verbosity': 0, 'filler_cfg': 'x.cfg', 'dims': [32, 32, 1],
                  'keyword_data': 'main_data', 'keyword_label': 'main_label', 'keyword_weight': 'main_weight'})
    io.start_manager(3)
    io.next()
    d = io.fetch_data('main_data')
    assert d.dim() == [3, 32, 32, 1] and d.data().shape == (3, 1024) and d.data().dtype == np.float32
    lab, w = io.fetch_data('main_label').data(), io.fetch_data('main_weight').data()
    assert set(np.unique(lab)) <= {0.0, 1.0, 2.0} and (w > 0).all()
    assert ((d.data() > 0) == (lab > 0)).all()              # nonzero pixels carry a class
    first = d.data().copy()
    io.next()
    assert not np.array_equal(first, io.fetch_data('main_data').data())
    assert list(io.fetch_entries()) == [3, 4, 5]
    a = S.make_batch(2, [16, 16, 16, 1], seed=5)
    b = S.make_batch(2, [16, 16, 16, 1], seed=5)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))  # deterministic


def test_dims_validation_matches_reference():
    with pytest.raises(NotImplementedError):       # ssnet.py:12-14
        uresnet(dims=[32, 32], num_class=3)
    with pytest.raises(NotImplementedError):
        uresnet(dims=[2, 3, 4, 5, 6], num_class=3)
    n = uresnet(dims=[64, 64, 1], num_class=3, num_strides=5, base_num_outputs=16, debug=True)
    assert n._build(None) == dict(nd=2, spatial=[64, 64], in_channels=1, base_filters=16, num_strides=5)


def test_bucket_plan_covers_everything_in_backward_order():
    ranges, off = [], 0
    for n in (100, 3000, 50, 70000, 10, 10, 500000, 64):
        ranges.append((off, off + n))
        off += n
    buckets = dp.plan_buckets(ranges, 4000)
    seen = [l for _, _, ls in buckets for l in ls]
    assert seen == list(range(len(ranges) - 1, -1, -1))          # last layer first
    assert buckets[0][1] == off and buckets[-1][0] == 0
    for (lo, hi, ls), nxt in zip(buckets, buckets[1:]):
        assert nxt[1] == lo                                        # contiguous, no gaps
    assert all(hi - lo >= 4000 for lo, hi, _ in buckets[:-1])


def test_summary_log_writes_reference_tags(tmp_path):
    """tf.summary stand-in (ssnet.py:81-89): scalars accuracy_all / accuracy_nonzero / loss, image
    data_example (first 10 of the minibatch), one record per summary step."""
    import json
    from uresnet_b200.ssnet_trainval import _summary_log
    w = _summary_log(str(tmp_path / "train"))
    w.add_summary({"loss": 1.0, "accuracy_all": 0.5, "accuracy_nonzero": 0.25}, 0, np.random.rand(12, 16, 16, 1))
    w.add_summary({"loss": 0.5, "accuracy_all": 0.75, "accuracy_nonzero": 0.5}, 20, None)
    w.close()
    recs = [json.loads(l) for l in (tmp_path / "train" / "scalars.jsonl").read_text().splitlines()]
    assert [r["step"] for r in recs] == [0, 20] and recs[1]["loss"] == 0.5
    ea = pytest.importorskip("tensorboard.backend.event_processing.event_accumulator")
    acc = ea.EventAccumulator(str(tmp_path / "train"))
    acc.Reload()
    assert set(acc.Tags()["scalars"]) == {"accuracy_all", "accuracy_nonzero", "loss"}
    assert [e.step for e in acc.Scalars("accuracy_all")] == [0, 20]
    assert acc.Tags()["images"] == ["data_example"]


def test_run_ssnet_usage_without_cfg(capsys):
    import importlib.util
    spec = importlib.util.spec_from_file_location("run_ssnet_entry", os.path.join(ROOT, "run_ssnet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(["not_a_config.txt"]) == 2
    assert "run_ssnet.py" in capsys.readouterr().err


def test_iteration_from_checkpoint_name():
    """lib/ssnet_trainval.py:41-42: the iteration is the integer after the last '-' of the file name."""
    from uresnet_b200.ssnet_trainval import ssnet_trainval
    t = ssnet_trainval()
    assert t.iteration_from_file_name("weights/uresnet-1200") == 1200
    assert t.iteration_from_file_name("weights/uresnet-37.npz") == 37


def test_oracle_decision_map_matches_per_pixel_loops():
    """lib/ssnet_trainval.py:269-287 restated per pixel: 1 = shower, 2 = track (ties go to track),
    0 where the input is not above the threshold."""
    from oracle import uresnet_oracle as O
    rng = np.random.default_rng(5)
    sm = rng.random((2, 6, 7, 3)).astype(np.float32)
    sm[0, 0, :3, 2] = sm[0, 0, :3, 1]                     # ties
    data = rng.choice([0.0, 1.0, 1.5, 40.0], size=(2, 6, 7)).astype(np.float32)
    got = O.ssnet_decision(sm, data)
    for idx in np.ndindex(2, 6, 7):
        want = 0.0
        if data[idx] > 1.0:
            want = 1.0 if sm[idx + (1,)] > sm[idx + (2,)] else 2.0
        assert got[idx] == want
    assert got.dtype == np.float32


def test_synthetic_streams_do_not_overlap():
    """Batch seeds are SeedSequence([seed, rank, step]): rank r never replays rank r+1's batches and the
    test stream (seed + 500000) never collides with the train stream."""
    from uresnet_b200.synthetic_io import synthetic_threadio
    seen = {}
    for seed in (1234, 1234 + 500000):
        for rank in (0, 1, 2):
            io = synthetic_threadio()
            io.configure({"dims": [16, 16, 1], "seed": seed, "rank": rank})
            io.start_manager(1)
            for step in range(4):
                io.next()
                key = io.fetch_data("data").data().tobytes()
                assert key not in seen, (seed, rank, step, seen.get(key))
                seen[key] = (seed, rank, step)
    a, b = synthetic_threadio(), synthetic_threadio()
    for io in (a, b):
        io.configure({"dims": [16, 16, 1], "seed": 7, "rank": 3})
        io.start_manager(2)
        io.next()
    assert np.array_equal(a.fetch_data("data").data(), b.fetch_data("data").data())     # still deterministic


def test_checkpoint_retention_nmax_and_nhour(tmp_path):
    """tf.train.Saver(max_to_keep=CHECKPOINT_NMAX, keep_checkpoint_every_n_hours=CHECKPOINT_NHOUR)
    (ssnet_trainval.py:133-135): the newest NMAX checkpoints stay; one that drops out of the window
    is kept for good only if NHOUR hours passed since the last one kept for good."""
    from collections import OrderedDict
    from uresnet_b200.ssnet_trainval import ssnet_trainval

    class FakeNet(object):
        def get_variables(self):
            return OrderedDict([("UResNet/conv0/weights", np.zeros((3, 3, 1, 4), np.float32)),
                                ("UResNet/conv0/BatchNorm/beta", np.zeros(4, np.float32))])
    for nhour, want in ((1e6, ["u-3.npz", "u-4.npz"]), (0.0, ["u-3.npz", "u-4.npz"]),
                        (1e-12, ["u-0.npz", "u-1.npz", "u-2.npz", "u-3.npz", "u-4.npz"])):
        d = tmp_path / ("h%g" % nhour)
        d.mkdir()
        t = ssnet_trainval()
        t._cfg.SAVE_FILE, t._cfg.CHECKPOINT_NMAX, t._cfg.CHECKPOINT_NHOUR = str(d / "u"), 2, nhour
        t._net = FakeNet()
        for it in range(5):
            t._iteration = it
            t._save()
            time.sleep(0.02)                                   # strictly increasing mtimes
        assert sorted(p.name for p in d.glob("*.npz")) == want, (nhour, sorted(p.name for p in d.glob("*.npz")))
        t._net = None


def test_initialiser_draws_the_filters_the_reference_driver_transcript_started_from():
    """tests/golden/refdriver_train.json records the initial filters of the reference driver's run over the TF
    shim (PCG64(TF_RANDOM_SEED), Xavier-uniform filter by filter in creation order).  ``ssnet_base._init_variables``
    -- called here on a host-side stand-in for the network object, no GPU -- must draw the same ones, so that the
    product's driver started with that TF_RANDOM_SEED begins where the transcript began."""
    import json
    import types
    import torch
    from oracle import uresnet_oracle as O
    from uresnet_b200.ssnet import ssnet_base
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "refdriver_train.json")))
    names = O.param_names(2, base=g["cfg"]["BASE_NUM_FILTERS"], num_strides=5, num_class=g["cfg"]["NUM_CLASS"])
    assert [n for n, _ in names] == g["names"]
    table, off = {}, 0
    for n, s in names:
        cnt = int(np.prod(s))
        table[n] = (tuple(s), off, cnt)
        off += (cnt + 63) // 64 * 64
    fake = types.SimpleNamespace(_param_total=off, _param_table=table, _params=torch.zeros(off))
    ssnet_base._init_variables(fake, g["cfg"]["TF_RANDOM_SEED"])
    host = fake._params.numpy().astype(np.float64)
    for i, (n, (s, o, c)) in enumerate(table.items()):
        v = host[o:o + c]
        assert np.sqrt((v ** 2).sum()) == pytest.approx(g["init_l2"][i], rel=1e-6, abs=1e-12), n
        assert v.sum() == pytest.approx(g["init_sum"][i], rel=1e-5, abs=1e-5), n
