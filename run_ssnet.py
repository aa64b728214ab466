#!/usr/bin/env python
"""Train / analysis entry point, usable exactly like the reference's ``run_ssnet.py``:

    python run_ssnet.py config/train2d_synthetic.cfg [more.cfg ...]
    torchrun --nproc-per-node 8 run_ssnet.py config/train2d_synthetic.cfg      # data parallel

Flow of /root/reference/run_ssnet.py:11-19: every ``*.cfg`` argument overrides the configuration in
the order given, then ``initialize()`` -> ``batch_process()`` -> ``reset()``.  ``URESNETDIR`` is
not needed (the package is found next to this file).
"""
import os
import sys


def main(argv):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from uresnet_b200.ssnet_trainval import ssnet_trainval

    cfg_files = [a for a in argv if a.endswith('.cfg')]
    if not cfg_files:
        sys.stderr.write(__doc__)
        return 2
    driver = ssnet_trainval()
    for name in cfg_files:
        driver.override_config(name)
    driver.initialize()
    try:
        driver.batch_process()
    finally:
        driver.reset()
    return 0


if __name__ == '__main__':
    sys.exit(main(sys.argv[1:]))
