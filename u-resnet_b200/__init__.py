"""uresnet_b200: B200-native hot path of DeepLearnPhysics/u-resnet.

Sub-modules (host-side mirror of the reference's ``lib/``):
  _capi            ctypes loader for liburesnet_b200.so (the C ABI in include/uresnet_b200.h)
  uresnet, ssnet   ``uresnet`` / ``ssnet_base`` classes (lib/uresnet.py, lib/ssnet.py)
  config           ``ssnet_config`` (lib/config.py)
  ssnet_trainval   train/ana driver (lib/ssnet_trainval.py)
  synthetic_io     larcv_threadio-shaped synthetic batch source
  dp               data-parallel gradient exchange (bucketed NCCL allreduce)
"""
__version__ = "0.1.0"
