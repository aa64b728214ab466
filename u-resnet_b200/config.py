"""``ssnet_config``: reader of the reference's flat ``KEY VALUE`` configuration files.

Behavioural mirror of /root/reference/lib/config.py:6-76: the same 29 keys with the same defaults,
``#`` starts a comment, a line that is not exactly ``KEY VALUE`` is reported and skipped, several
files may be applied one after the other, a value must have the type of the key's default (except
``LEARNING_RATE``, which the reference lets change from int to float), ``IOError`` for a missing
file and ``TypeError`` for a type mismatch, and ``dump()`` prints ``KEY....... value`` lines.

Differences, on purpose: values are parsed as Python LITERALS (``ast.literal_eval``) where the
reference ``exec``s the line; a key that is not in the schema is reported and skipped (the
reference prints "Ignoring" and then dies in its type check -- ``DUMP_IMAGE`` in its own
config/ana3d.cfg:11 does that).  The schema is a table here, not a list of class attributes; the
values still appear as attributes of the instance (``cfg.NUM_CLASS``), which is all the driver uses.
"""
from __future__ import annotations

import ast
import copy
import os
from collections import OrderedDict

# key -> default.  First block: the reference schema (lib/config.py:8-36), in its order.
_REFERENCE_SCHEMA = OrderedDict([
    ("NUM_CLASS", 3), ("BASE_NUM_FILTERS", 16),
    ("MAIN_INPUT_CONFIG", "config/input_train.cfg"), ("TEST_INPUT_CONFIG", ""), ("ANA_OUTPUT_CONFIG", ""),
    ("LOGDIR", "ssnet_train_log"), ("SAVE_FILE", "ssnet_checkpoint/uresnet"), ("LOAD_FILE", ""),
    ("AVOID_LOAD_PARAMS", []), ("LEARNING_RATE", -1),
    ("MINIBATCH_SIZE", 10), ("NUM_MINIBATCHES", 5), ("TEST_BATCH_SIZE", 10), ("ITERATIONS", 100000),
    ("TF_RANDOM_SEED", 1234), ("TRAIN", True), ("DEBUG", False), ("USE_WEIGHTS", True),
    ("REPORT_STEPS", 200), ("SUMMARY_STEPS", 20),
    ("CHECKPOINT_STEPS", 200), ("CHECKPOINT_NMAX", 10), ("CHECKPOINT_NHOUR", 0.4),
    ("KEYWORD_DATA", "data"), ("KEYWORD_LABEL", "label"), ("KEYWORD_WEIGHT", "weight"),
    ("KEYWORD_TEST_DATA", ""), ("KEYWORD_TEST_LABEL", ""), ("KEYWORD_TEST_WEIGHT", ""),
])
# Second block: keys the reference does not have -- numeric mode of this library and the
# synthetic batch source that stands in for larcv.
_EXTRA_SCHEMA = OrderedDict([
    ("DTYPE", "f32"),              # 'f32' (default: the reference computes in float32) | 'mix16' | 'bf16'
    ("KERNEL_IMPL", "auto"),       # 'auto' | 'ffma' | 'tc'
    ("SYNTHETIC_DIMS", []),        # e.g. [512,512,1]; [] = infer 2D/3D from the larcv filler cfg name
    ("SYNTHETIC_SEED", 1234),
])
_FREE_TYPE = frozenset(["LEARNING_RATE"])


def _tokens(line):
    """The words of a config line with its ``#`` comment removed."""
    return line.split("#", 1)[0].split()


class ssnet_config(object):

    def __init__(self):
        for table in (_REFERENCE_SCHEMA, _EXTRA_SCHEMA):
            for key, default in table.items():
                setattr(self, key, copy.copy(default))

    @classmethod
    def _keys(cls):
        return list(_REFERENCE_SCHEMA) + list(_EXTRA_SCHEMA)

    def _parse(self, key, text, line):
        try:
            value = ast.literal_eval(text)
        except (ValueError, SyntaxError):
            value = NotImplemented                 # not a literal: no type can match
        if key not in _FREE_TYPE and type(value) is not type(getattr(self, key)):
            print("Incompatible type: %s" % line)
            raise TypeError(line)
        if value is NotImplemented:
            print("Incompatible type: %s" % line)
            raise TypeError(line)
        return value

    def override(self, file_name):
        if not os.path.isfile(file_name):
            print("Config file not found", file_name)
            raise IOError(file_name)
        known = set(self._keys())
        with open(file_name, "r") as f:
            for line in f.read().split("\n"):
                words = _tokens(line)
                if not words:
                    continue
                if len(words) != 2:
                    print("Ignoring a line:", line)
                elif words[0] not in known:
                    print("Ignoring a parameter in file:", words[0])
                else:
                    setattr(self, words[0], self._parse(words[0], words[1], line))

    def dump(self):
        for key in self._keys():
            print("%s %s" % (key.ljust(20, "."), getattr(self, key)))


if __name__ == "__main__":
    import sys
    cfg = ssnet_config()
    cfg.dump()
    for name in sys.argv[1:]:
        cfg.override(name)
        print("\n")
        cfg.dump()
