"""ctypes binding of the C-ABI in include/amb200.h (csrc/libamb200.so).

The library is the product: if it is missing or the device is not sm_100 this module raises — there is
no PyTorch / CPU fallback anywhere in the package.
"""
import ctypes
import os
import re
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
REPO_ROOT = os.path.dirname(PKG_ROOT)
LIB_PATH = os.path.join(PKG_ROOT, "csrc", "libamb200.so")
HEADER_PATH = os.path.join(REPO_ROOT, "include", "amb200.h")

P = c_void_p
I = c_int
_SIGS = {
    "am_version": (c_int, []),
    "am_check_device": (c_int, []),
    "am_launch_count": (c_int64, []),
    "am_last_error": (c_char_p, []),
    "am_randn": (c_int, [P, c_int64, I, c_int64, c_uint64, c_uint64, P]),
    "am_p_sample_update": (c_int, [P, P, P, P, P, P, P, P, I, I, c_int64, c_uint64, c_int64, P]),
    "am_ddim_update": (c_int, [P, P, P, P, P, P, P, P, c_float, P, I, I, c_int64, c_uint64, c_int64, P]),
    "am_q_sample": (c_int, [P, P, P, P, P, P, I, c_int64, P]),
    "am_masked_mse": (c_int, [P, P, P, P, I, I, I, P]),
    "am_add_i32": (c_int, [P, c_int32, I, P]),
    "am_linear_f32": (c_int, [P, I, P, I, P, I, I, I, I, P, I, P, I, I, I, I, I, I, I, I, P]),
    "am_layernorm": (c_int, [P, I, P, I, P, P, P, I, I, I, c_float, P, I, P]),
    "am_mha_fwd": (c_int, [P, P, P, I, I, I, I, c_float, P, P]),
    "am_gather_time_token": (c_int, [P, I, I, I, P, P, I, I, P, P]),
    "am_gather_rows": (c_int, [P, P, P, I, I, P]),
    "am_furthestsampling": (c_int, [I, I, P, P, P, P, P, P]),
    "am_knnquery": (c_int, [I, I, I, P, P, P, P, P, P, P]),
    "am_pt_layer_fwd": (c_int, [P] * 16 + [I, I, I, P]),
    "am_transition_down_fwd": (c_int, [P] * 7 + [I, I, I, I, P]),
    "am_cdm_encoder_partial": (c_int, [P, P, P, P, P, P, P, I, P, I, I, I, I, P]),
    "am_cdm_encoder_combine": (c_int, [P, P, I, I, P]),
    "am_cdm_decoder_point": (c_int, [P, P, P, P, P, P, P, I, P, P, P, P, P, P, P, I, I, I, P]),
    "am_linear_skinny": (c_int, [P, I, I, P, I, I, P, P, P, I, I, I, P]),
    "am_mha_tc_fwd": (c_int, [P, P, P, P, I, I, I, I, c_float, P]),
    "am_split_bf16": (c_int, [P, I, P, I, I, I, P]),
    "am_linear_tc": (c_int, [P, P, I, I, I, P, I, P, I, I, P, I, I, I, I, P, I, P]),
}

_lib = None


def declared_symbols():
    """Function names declared in include/amb200.h (the boundary a maintainer binds)."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(am_[a-z0-9_]+)\s*\(", txt)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"amb200: {LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C afford-motion_b200/csrc`). There is no CPU/PyTorch fallback for the hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class AmbError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = load().am_last_error()
        raise AmbError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().am_launch_count())
