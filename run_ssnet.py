#!/usr/bin/env python
"""Drop-in for the reference's ``run_ssnet.py`` (train / ana entry point):

    python run_ssnet.py config/train2d.cfg [more.cfg ...]

Every ``*.cfg`` argument is applied in order, then ``initialize()``, ``batch_process()``,
``reset()`` -- /root/reference/run_ssnet.py:11-19.  ``URESNETDIR`` is not needed.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from uresnet_b200 import ssnet_trainval as api  # noqa: E402

t = api.ssnet_trainval()
for argv in sys.argv:
    if argv.endswith('.cfg'):
        t.override_config(argv)

t.initialize()

t.batch_process()

t.reset()
