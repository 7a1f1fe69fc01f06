"""ctypes binding of libxgating.so (include/xgating.h).  No torch types cross this boundary.

The library is the product: if it is missing or cannot be loaded this module raises — there is
no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libxgating.so")

XG_NUM_PARAMS = 57
XG_OK = 0
XG_ERR_BAD_ARG, XG_ERR_BAD_SHAPE, XG_ERR_NULL_POINTER, XG_ERR_CUDA = 1, 2, 3, 4
XG_ERR_NOT_BOUND, XG_ERR_WORKSPACE, XG_ERR_UNSUPPORTED = 5, 6, 7
XG_ACT = {"ReLU": 1, "Tanh": 2, "Sigmoid": 3}
(XG_WS_ENCODE, XG_WS_DECODE_STEP, XG_WS_GREEDY, XG_WS_BEAM, XG_WS_TRAIN_SAVED, XG_WS_TRAIN_FWD,
 XG_WS_TRAIN_BWD) = range(7)
DROP_SITES = {"enc_emb_rgb": 1, "enc_emb_opfl": 2, "enc_gate_rgb": 3, "enc_gate_opfl": 4, "enc_fusion": 5,
              "dec_gate": 6, "dec_h1": 7, "dec_h2": 8, "cls": 9}


class XgDims(ctypes.Structure):
    _fields_ = [("feat_rgb", c_int), ("feat_opfl", c_int), ("rnn", c_int), ("embed", c_int), ("att", c_int),
                ("vocab", c_int), ("categories", c_int), ("cls_hidden", c_int), ("fusion_act", c_int),
                ("drop_prob", c_float), ("bn_eps", c_float), ("bn_momentum", c_float)]


class XgAdamTensor(ctypes.Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p), ("n", c_int64)]


PtrTable = c_void_p * XG_NUM_PARAMS
Ptr4 = c_void_p * 4

# name -> (restype, argtypes); must list every symbol declared in include/xgating.h
SIGNATURES = {
    "xg_abi_version": (c_int, []),
    "xg_status_string": (c_char_p, [c_int]),
    "xg_last_error": (c_char_p, [c_void_p]),
    "xg_create": (c_int, [POINTER(XgDims), c_int, POINTER(c_void_p)]),
    "xg_destroy": (c_int, [c_void_p]),
    "xg_param_shape": (c_int, [c_void_p, c_int, POINTER(c_int), POINTER(c_int)]),
    "xg_bind_params": (c_int, [c_void_p, POINTER(c_void_p), c_int]),
    "xg_bind_bn_buffers": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "xg_params_changed": (c_int, [c_void_p]),
    "xg_set_engine": (c_int, [c_void_p, c_int]),
    "xg_set_bwd_split_event": (c_int, [c_void_p, c_void_p]),
    "xg_set_strict": (c_int, [c_void_p, c_int]),
    "xg_path_counters": (c_int, [c_void_p, POINTER(c_uint64), POINTER(c_uint64)]),
    "xg_set_decode_dropout": (c_int, [c_void_p, c_int, c_uint64]),
    "xg_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "xg_encode_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_uint64,
                              c_void_p, c_void_p, POINTER(c_void_p), c_void_p, c_size_t, c_void_p]),
    "xg_init_hidden": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p), c_void_p, c_size_t,
                               c_void_p]),
    "xg_attend_precompute": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "xg_decode_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_void_p, c_int, c_int,
                               c_void_p, c_size_t, c_void_p]),
    "xg_sample_greedy": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p), c_int, c_int, c_int,
                                 c_int, c_float, c_uint64, c_void_p, c_void_p, POINTER(c_int),
                                 c_void_p, c_size_t, c_void_p]),
    "xg_scheduled_tokens": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_int, c_int,
                                    c_int, c_int, c_float, c_uint64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "xg_sample_beam": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_size_t, c_void_p]),
    "xg_seq_steps": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_int), c_void_p]),
    "xg_train_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int, c_int, c_int, c_int, c_int, c_uint64, c_void_p, c_void_p,
                             c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "xg_train_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int, c_int, c_int, c_int, c_int, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_size_t, POINTER(c_void_p), c_int, c_void_p, c_size_t, c_void_p]),
    "xg_nll_criterion_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "xg_nll_criterion_bwd": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "xg_profile_enable": (c_int, [c_void_p, c_int]),
    "xg_profile_report": (c_int, [c_void_p, c_char_p, c_size_t]),
    "xg_adam_step": (c_int, [POINTER(XgAdamTensor), c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_float, c_int,
                             c_int, c_void_p]),
    "xg_debug_dropout_mask": (c_int, [c_uint64, c_int, c_size_t, c_float, c_void_p, c_void_p]),
    "xg_debug_gemm": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
}

_lib = None


class XGatingError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str):
        self.status = status
        super().__init__("%s failed: status %d%s" % (where, status, (" — " + detail) if detail else ""))


def load() -> ctypes.CDLL:
    """dlopen libxgating.so and type every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libxgating.so not found at %s. Build it with `python -m controllable_xgating_b200.build` "
            "(needs nvcc). The CUDA extension is the only implementation; there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, where: str, handle=None):
    """Map a non-zero xg_status to a Python exception (the reference raises AssertionError for its
    shape asserts: sub_modules.py:69,673; SAModel.py:134)."""
    if status == XG_OK:
        return
    lib = load()
    detail = lib.xg_last_error(handle)
    detail = detail.decode() if detail else ""
    name = lib.xg_status_string(status).decode()
    msg = "%s: %s" % (name, detail) if detail else name
    if status == XG_ERR_BAD_SHAPE:
        raise AssertionError("%s: %s" % (where, msg))
    raise XGatingError(status, where, msg)
