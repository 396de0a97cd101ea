"""ctypes binding of the C ABI in include/coattn_b200.h (libcoattn_b200.so, built in-tree).

There is deliberately no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COATTN_B200_LIB", os.path.join(_HERE, "lib", "libcoattn_b200.so"))
ABI_VERSION = 2

_lock = threading.Lock()
_lib = None

_vp, _i, _i64, _u = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint

FLAG_BF16 = 1  # COATTN_FLAG_BF16
FLAG_UNFUSED_GATE = 2  # COATTN_FLAG_UNFUSED_GATE
FLAG_SINGLE_CTA = 4  # COATTN_FLAG_SINGLE_CTA
FLAG_A_ONLY = 8  # COATTN_FLAG_A_ONLY
FLAG_UNFUSED_PREP = 16  # COATTN_FLAG_UNFUSED_PREP
FLAG_GATED_ONLY = 32  # COATTN_FLAG_GATED_ONLY
FLAG_KMAJOR = 64  # COATTN_FLAG_KMAJOR
FLAG_SOFTMAX16 = 128  # COATTN_FLAG_SOFTMAX16
FLAG_SPLIT_KEYS = 256  # COATTN_FLAG_SPLIT_KEYS
FLAG_PLANES_READY = 512  # COATTN_FLAG_PLANES_READY
FLAG_UNFOLDED = 1024  # COATTN_FLAG_UNFOLDED
STATUS_WORDS = 8  # COATTN_STATUS_WORDS
STATUS_OVERFLOW_B, STATUS_OVERFLOW_A, STATUS_OVERFLOW_Q = 1, 2, 4

# name -> (restype, argtypes); mirrors include/coattn_b200.h one to one
SIGNATURES = {
    "coattn_b200_abi_version": (_i, []),
    "coattn_b200_strerror": (ctypes.c_char_p, [_i]),
    "coattn_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "coattn_status_clear": (_i, [_vp, _vp]),
    "coattn_status_read": (_i, [_vp, _vp, _vp]),
    "coattn_workspace_segment": (_i, [ctypes.c_char_p, _i, _i, _i, _i, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    "coattn_forward": (_i, [_vp] * 11 + [_i64, _i, _i, _i, _i, _u, _vp]),
    "coattn_stage_attend_gate": (_i, [_vp] * 10 + [_i64, _i, _i, _i, _i, _u, _vp]),
    "coattn_stage_passthrough": (_i, [_vp] * 4 + [_i, _i, _i, _i, _vp]),
    "coattn_stage_prep": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _u, _vp]),
    "coattn_stage_project": (_i, [_vp, _i64, _i, _i, _i, _i, _u, _vp]),
    "coattn_stage_prep_project": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _u, _vp]),
    "coattn_stage_tail": (_i, [_vp] * 6 + [_i64, _i, _i, _i, _i, _i, _u, _vp]),
    "coattn_forward_queries": (_i, [_vp] * 7 + [_i64, _i, _i, _i, _i, _i, _u, _vp]),
    "coattn_forward16": (_i, [_vp] * 10 + [_i64, _i, _i, _i, _i, _i, _u, _vp]),
    "coattn_stage_attend": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _u, _vp]),
    "coattn_stage_gate": (_i, [_vp] * 7 + [_i, _i, _i, _i, _vp]),
    "coattn_backward_workspace_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "coattn_backward": (_i, [_vp] * 15 + [_i64, _i, _i, _i, _i, _u, _vp]),
}


class CoattnError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise CoattnError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU / eager fallback for the co-attention path)")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            got = lib.coattn_b200_abi_version()
            if got != ABI_VERSION:
                raise CoattnError(f"libcoattn_b200 ABI {got} != expected {ABI_VERSION}: rebuild")
            _lib = lib
    return _lib


def check(code: int, what: str):
    if code != 0:
        msg = load().coattn_b200_strerror(code).decode()
        raise CoattnError(f"{what} failed with code {code}: {msg}")
