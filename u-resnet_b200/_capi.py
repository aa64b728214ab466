"""ctypes binding of ``include/uresnet_b200.h`` (the C ABI of liburesnet_b200.so).

The library is built in-tree by ``uresnet_b200.build`` (nvcc, sm_100a).  There is no
CPU fallback: if the library is missing, importing the product classes raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liburesnet_b200.so")

URN_F32, URN_BF16, URN_F16 = 0, 1, 2
URN_IMPL_AUTO, URN_IMPL_FFMA, URN_IMPL_TC = 0, 1, 2
# network modes: f32 = exact CUDA-core path; mix16 = fp16 activations/filters + bf16 gradients
# (tensor cores); bf16 = all-bf16 storage
DTYPES = {"f32": URN_F32, "fp32": URN_F32, "float32": URN_F32, "bf16": URN_BF16, "bfloat16": URN_BF16,
          "mix16": URN_F16, "f16": URN_F16, "fp16": URN_F16}
IMPLS = {"auto": URN_IMPL_AUTO, "ffma": URN_IMPL_FFMA, "tc": URN_IMPL_TC}


class urn_conv_desc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n", "xd", "xh", "xw", "cx", "yd", "yh", "yw", "cy", "kd", "kh", "kw", "s",
        "pad_d", "pad_h", "pad_w")]


class urn_net_cfg(C.Structure):
    _fields_ = [("nd", C.c_int32), ("spatial", C.c_int32 * 3), ("in_channels", C.c_int32),
                ("batch", C.c_int32), ("base_filters", C.c_int32), ("num_strides", C.c_int32),
                ("num_class", C.c_int32), ("dtype", C.c_int32), ("impl", C.c_int32),
                ("training", C.c_int32)]


class urn_tensor_ref(C.Structure):
    _fields_ = [("off_bytes", C.c_int64), ("ld", C.c_int64), ("rows", C.c_int64), ("c", C.c_int32),
                ("dtype", C.c_int32)]


class urn_step_info(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("kind", "layer", "layer2", "node_a", "node_b", "relu", "accumulate")]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_tref = C.POINTER(urn_tensor_ref)
_descp = C.POINTER(urn_conv_desc)

# name -> (restype, argtypes); must list every symbol declared in include/uresnet_b200.h
SIGNATURES = {
    "urn_conv_desc_same": (_i, [_descp, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    "urn_last_error": (C.c_char_p, []),
    "urn_build_info": (C.c_char_p, []),
    "urn_conv_fwd": (_i, [_descp, _i, _i, _vp, _i64, _vp, _vp, _i64, _i, _vp]),
    "urn_conv_dgrad": (_i, [_descp, _i, _i, _vp, _i64, _vp, _vp, _i64, _i, _i, _vp]),
    "urn_conv_fwd_tc": (_i, [_descp, _i, _vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _vp]),
    "urn_conv_dgrad_tc": (_i, [_descp, _i, _vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _vp]),
    "urn_conv_fwd_tc_stats": (_i, [_descp, _i, _vp, _i64, _vp, _vp, _vp, _i64, _f, _vp, _vp, _vp]),
    "urn_conv_wgrad_workspace": (_i64, [_descp]),
    "urn_conv_wgrad": (_i, [_descp, _i, _i, _vp, _i64, _vp, _i64, _i, _vp, _i, _vp, _vp]),
    "urn_bn_workspace": (_i64, [_i]),
    "urn_bn_stats": (_i, [_vp, _i64, _i64, _i, _f, _vp, _vp, _vp]),
    "urn_bn_apply": (_i, [_vp, _i64, _i64, _i, _vp, _vp, _vp, _i64, _i, _vp, _i64, _vp, _vp, _i,
                          _vp, _i64, _i, _vp, _vp]),
    "urn_bn_bwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _i64, _i, _vp, _i64, _vp, _vp, _vp, _i64,
                        _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i, _i, _vp, _vp]),
    "urn_bn_bwd_beta": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _i64, _i, _vp, _i64, _vp, _vp, _vp, _vp, _i64,
                             _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i, _i, _vp, _vp]),
    "urn_bn_stats_t": (_i, [_vp, _i, _i64, _i64, _i, _f, _vp, _vp, _vp]),
    "urn_bn_apply_t": (_i, [_vp, _i, _i64, _i64, _i, _vp, _vp, _vp, _i64, _i, _vp, _i64, _vp, _vp, _i,
                            _vp, _i64, _i, _vp, _vp]),
    "urn_bn_bwd_t": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _i64, _i, _vp, _i, _i64, _vp, _vp, _vp, _vp, _i64,
                          _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i, _i, _vp, _vp]),
    "urn_conv_fwd_tc_stats_t": (_i, [_descp, _i, _vp, _i64, _vp, _vp, _vp, _i, _i64, _f, _vp, _vp, _vp]),
    "urn_loss_workspace": (_i64, [_i64]),
    "urn_ce_fwd_bwd": (_i, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _i, _vp, _i64, _i, _vp, _vp, _vp,
                            _vp]),
    "urn_adam_step": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i64, _vp]),
    "urn_create": (_i, [C.POINTER(urn_net_cfg), C.POINTER(_vp)]),
    "urn_destroy": (None, [_vp]),
    "urn_param_count": (_i, [_vp]),
    "urn_param_info": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_i), C.POINTER(_i64),
                            C.POINTER(_i64), C.POINTER(_i64)]),
    "urn_param_total": (_i64, [_vp]),
    "urn_layer_count": (_i, [_vp]),
    "urn_workspace_bytes": (_i64, [_vp]),
    "urn_logits_offset_bytes": (_i64, [_vp]),
    "urn_forward": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "urn_loss_metrics": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "urn_backward": (_i, [_vp, _vp, _vp, _i, _vp, C.POINTER(_vp), _vp]),
    "urn_check_labels": (_i, [_vp, _vp, _vp]),
    "urn_decision": (_i, [_vp, _vp, C.c_float, _vp, _vp, _vp]),
    "urn_ssnet_decision": (_i, [_vp, _i64, _i64, _i, _vp, _i64, C.c_float, _vp, _vp]),
    "urn_plan_counts": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "urn_node_info": (_i, [_vp, _i, C.POINTER(C.c_char_p), _tref, _tref, _tref]),
    "urn_layer_tensors": (_i, [_vp, _i, _tref, _tref, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64),
                               C.POINTER(_i), C.POINTER(_i)]),
    "urn_fwd_step_info": (_i, [_vp, _i, C.POINTER(urn_step_info)]),
    "urn_bwd_step_info": (_i, [_vp, _i, C.POINTER(urn_step_info)]),
    "urn_forward_range": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "urn_backward_range": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _vp]),
    "urn_launch_count": (_i64, []),
    "urn_profile_enable": (_i, [_vp, _i]),
    "urn_profile_collect": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "urn_layer_info": (_i, [_vp, _i, C.POINTER(C.c_char_p), _descp, C.POINTER(_i), C.POINTER(_i64),
                            C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
}

_lib = None


class UrnError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach signatures.  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "uresnet_b200: %s not found -- build it with `python -m uresnet_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise UrnError(load().urn_last_error().decode("utf-8", "replace"))


def conv_desc_same(nd, n, spatial, cx, cy, k, s):
    """TF 'same' geometry; spatial = (h,w) or (d,h,w) of the big (x) side."""
    d = urn_conv_desc()
    sp = (1,) + tuple(spatial) if nd == 2 else tuple(spatial)
    check(load().urn_conv_desc_same(C.byref(d), nd, n, sp[0], sp[1], sp[2], cx, cy, k, s))
    return d
