"""``ssnet_config``: Python-3 re-reader of the reference's flat ``KEY VALUE`` config files.

Mirror of /root/reference/lib/config.py:6-76 -- same 29 keys and defaults (class attributes are
the schema), same parsing rules (``#`` comments, lines that do not have exactly two tokens are
ignored with a message, several files applied in order), same type rule (the value's type must
equal the default's type, except ``LEARNING_RATE``), ``IOError`` for a missing file,
``TypeError`` for a type mismatch.  Values are parsed with ``ast.literal_eval`` instead of the
reference's ``exec``; unknown keys are reported and skipped (the reference prints "Ignoring"
and then crashes in its type check, e.g. ``DUMP_IMAGE`` in config/ana3d.cfg:11).

Keys below the marked line do not exist in the reference; they select the numeric mode and
describe the synthetic batch source that stands in for larcv.
"""
from __future__ import annotations

import ast
import os


class ssnet_config:

    NUM_CLASS         = 3
    BASE_NUM_FILTERS  = 16
    MAIN_INPUT_CONFIG = 'config/input_train.cfg'
    TEST_INPUT_CONFIG = ''
    ANA_OUTPUT_CONFIG = ''
    LOGDIR            = 'ssnet_train_log'
    SAVE_FILE         = 'ssnet_checkpoint/uresnet'
    LOAD_FILE         = ''
    AVOID_LOAD_PARAMS = []
    LEARNING_RATE     = -1
    MINIBATCH_SIZE    = 10
    NUM_MINIBATCHES   = 5
    TEST_BATCH_SIZE   = 10
    ITERATIONS        = 100000
    TF_RANDOM_SEED    = 1234
    TRAIN             = True
    DEBUG             = False
    USE_WEIGHTS       = True
    REPORT_STEPS      = 200
    SUMMARY_STEPS     = 20
    CHECKPOINT_STEPS  = 200
    CHECKPOINT_NMAX   = 10
    CHECKPOINT_NHOUR  = 0.4
    KEYWORD_DATA      = 'data'
    KEYWORD_LABEL     = 'label'
    KEYWORD_WEIGHT    = 'weight'
    KEYWORD_TEST_DATA   = ''
    KEYWORD_TEST_LABEL  = ''
    KEYWORD_TEST_WEIGHT = ''
    # ---- not in the reference schema -------------------------------------------------
    DTYPE             = 'mix16'       # 'f32' | 'mix16' | 'bf16'   (uresnet_b200 numeric mode)
    KERNEL_IMPL       = 'auto'        # 'auto' | 'ffma' | 'tc'
    SYNTHETIC_DIMS    = []            # e.g. [512,512,1]; [] = infer 2D/3D from the larcv filler cfg
    SYNTHETIC_SEED    = 1234

    def __init__(self):
        pass

    @classmethod
    def _keys(cls):
        return [s for s in cls.__dict__.keys() if s == s.upper() and not s.startswith('_')]

    def override(self, file_name):
        keys = self._keys()
        if not os.path.isfile(file_name):
            print('Config file not found', file_name)
            raise IOError
        for line in open(file_name, 'r').read().split('\n'):
            valid_line = str(line)
            if line.find('#') >= 0:
                valid_line = line[0:line.find('#')]
            words = valid_line.split()
            if len(words) == 0:
                continue
            if not len(words) == 2:
                print('Ignoring a line:', line)
                continue
            if not words[0] in keys:
                print('Ignoring a parameter in file:', words[0])
                continue
            try:
                value = ast.literal_eval(words[1])
            except (ValueError, SyntaxError):
                print('Incompatible type: %s' % line)
                raise TypeError
            valid = type(getattr(self, words[0])) == type(value)
            if words[0] != 'LEARNING_RATE' and not valid:
                print('Incompatible type: %s' % line)
                raise TypeError
            setattr(self, words[0], value)

    def dump(self):
        for key in self._keys():
            msg = key
            while len(msg) < 20:
                msg += '.'
            print('%s %s' % (msg, str(getattr(self, key))))


if __name__ == '__main__':
    k = ssnet_config()
    k.dump()
    import sys
    if len(sys.argv) > 1:
        k.override(sys.argv[1])
        print('\n')
        k.dump()
