"""Transcript of THE REFERENCE'S DRIVER training on the synthetic source -> tests/golden/refdriver_train.json.

``/root/reference/lib/ssnet_trainval.py`` (+ ``config.py``, ``uresnet.py``, ``resnet_module.py``, ``ssnet.py``)
run unmodified over ``oracle/tf1_shim.py`` (float64) with a stand-in ``larcv`` whose ``larcv_threadio`` is this
repo's seeded synthetic source configured exactly as ``uresnet_b200.ssnet_trainval._make_io`` configures it.
The default topology depth (num_strides = 5) on 64x64 images, 4 base filters; 6 iterations of 2 minibatches of
2 images, a test stream of 2 images evaluated at every report.  The shim's Xavier initialiser draws from
``PCG64(TF_RANDOM_SEED)`` filter by filter in creation order -- the product's ``_init_variables`` does the same
-- so the product's driver started with the same TF_RANDOM_SEED begins from the same weights.

    python tests/golden/make_reference_driver_golden.py        # build container only (needs /root/reference)

Consumers: tests/test_reference_driver.py (CPU: regeneration where the reference exists; the product's
initialiser formula reproduces the recorded initial filters), tests/test_net_gpu.py (``-m gpu``: the CUDA
library behind the product's driver reproduces this transcript).
"""
import contextlib
import io
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from uresnet_b200.synthetic_io import synthetic_threadio  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "refdriver_train.json")

DIMS = [64, 64, 1]
SYNTHETIC_SEED = 99
CFG = dict(MAIN_INPUT_CONFIG="config/input_train2d.cfg", TEST_INPUT_CONFIG="config/input_train2d.cfg",
           NUM_CLASS=3, BASE_NUM_FILTERS=4, MINIBATCH_SIZE=2, NUM_MINIBATCHES=2, TEST_BATCH_SIZE=2, ITERATIONS=6,
           LEARNING_RATE=1e-3, TF_RANDOM_SEED=1234, TRAIN=True, USE_WEIGHTS=True, REPORT_STEPS=1, SUMMARY_STEPS=0,
           CHECKPOINT_STEPS=0, LOGDIR="", SAVE_FILE="", LOAD_FILE="",
           KEYWORD_DATA="main_data", KEYWORD_LABEL="main_label", KEYWORD_WEIGHT="main_weight",
           KEYWORD_TEST_DATA="test_data", KEYWORD_TEST_LABEL="test_label", KEYWORD_TEST_WEIGHT="test_weight")


class threadio(synthetic_threadio):
    """``larcv_threadio`` for the reference driver = the synthetic source as ``_make_io`` sets it up"""

    def configure(self, cfg):
        c = dict(cfg)
        test = c.get("filler_name") == "TestIO"
        pre = "test_" if test else "main_"
        c.update(dims=DIMS, seed=SYNTHETIC_SEED + (500000 if test else 0), rank=0, keyword_data=pre + "data",
                 keyword_label=pre + "label", keyword_weight=pre + "weight")
        synthetic_threadio.configure(self, c)


_NUM = r"([-+0-9.eE]+|nan)"
_LINE = re.compile(r"loss=%s\s+acc\. all=%s\s+acc\. nonzero=%s" % (_NUM, _NUM, _NUM))


def parse_transcript(text):
    """[{iteration, lr, train: [loss, acc_all, acc_nonzero], test: [...]}] from the driver's stdout"""
    reps = []
    for ln in text.splitlines():
        m = re.match(r"@ iteration (\d+) LR (\S+) Mem", ln)
        if m:
            reps.append(dict(iteration=int(m.group(1)), lr=float(m.group(2))))
            continue
        for tag, key in (("Train set: ", "train"), ("Test set: ", "test")):
            if ln.startswith(tag):
                reps[-1][key] = [float(x) for x in _LINE.search(ln).groups()]
    return reps


def run():
    import torch
    from oracle import tf1_shim as shim
    shim.reset_default_graph(torch.float64)
    drv = shim.load_reference_driver(threadio)()
    for k, v in CFG.items():
        assert hasattr(drv._cfg, k), k
        setattr(drv._cfg, k, v)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        drv.initialize()
        import tensorflow as tf
        init = {v.op.name: v.value.detach().numpy().copy() for v in tf.trainable_variables()}
        drv.batch_process()
    final = {v.op.name: v.value.detach().numpy().copy() for v in tf.trainable_variables()}
    drv.reset()
    shim.uninstall()
    names = list(init.keys())
    return dict(made_by="reference lib/ssnet_trainval.py executed over oracle/tf1_shim.py (float64)",
                dims=DIMS, synthetic_seed=SYNTHETIC_SEED, cfg=CFG, reports=parse_transcript(buf.getvalue()),
                names=names, init_l2=[float(np.sqrt((init[k] ** 2).sum())) for k in names],
                init_sum=[float(init[k].sum()) for k in names],
                final_l2=[float(np.sqrt((final[k] ** 2).sum())) for k in names],
                update_l2=[float(np.sqrt(((final[k] - init[k]) ** 2).sum())) for k in names])


if __name__ == "__main__":
    g = run()
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
        f.write("\n")
    for r in g["reports"]:
        print(r)
