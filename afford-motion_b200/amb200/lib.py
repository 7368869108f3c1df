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
# AMB200_LIB: A/B timing of another BUILD of this library (tools/ab_tc.py); never a different implementation
LIB_PATH = os.environ.get("AMB200_LIB") or os.path.join(PKG_ROOT, "csrc", "libamb200.so")
HEADER_PATH = os.path.join(REPO_ROOT, "include", "amb200.h")

_CT = {"double": ctypes.c_double, "int": c_int, "int32_t": c_int32, "int64_t": c_int64, "uint64_t": c_uint64, "uint32_t": ctypes.c_uint32, "float": c_float,
       "am_stream_t": c_void_p}


def _parse_header():
    """ctypes signatures straight from the prototypes in include/amb200.h, so the binding cannot drift from the boundary."""
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER_PATH).read(), flags=re.S)
    sigs = {}
    for ret, name, args in re.findall(r"\b(int64_t|int|const char\*|void)\s+(am_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        argt = []
        for a in [x.strip() for x in args.split(",") if x.strip() and x.strip() != "void"]:
            if "*" in a:
                argt.append(c_void_p)
            else:
                argt.append(_CT[a.replace("const ", "").split()[0]])
        sigs[name] = ({"int": c_int, "int64_t": c_int64, "const char*": c_char_p, "void": None}[ret], argt)
    return sigs


_SIGS = _parse_header()

_lib = None


class AmbError(RuntimeError):
    pass


PRECISIONS = {"parity": 0, "fast": 1}


def set_precision(mode: str) -> None:
    """'parity' (default; 3-term bf16 split = fp32-equivalent, all parity claims) or 'fast' (one bf16 pass, ~1e-2 relative error;
    outside the parity budget, reported separately by bench.py).  Captured CUDA graphs bake the mode they were captured in:
    the models' sampler handles are keyed on it."""
    check(load().am_set_precision(PRECISIONS[mode]), "am_set_precision")


def get_precision() -> str:
    return {v: k for k, v in PRECISIONS.items()}[int(load().am_get_precision())]


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
        if os.environ.get("AMB200_LIB") and not hasattr(lib, name):
            continue  # an older build under A/B test lacks the newer entry points
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    _check_device_once(lib)
    env = os.environ.get("AMB200_PRECISION")
    if env and hasattr(lib, "am_set_precision"):
        if env not in PRECISIONS:
            raise AmbError(f"AMB200_PRECISION must be one of {sorted(PRECISIONS)}, got {env!r}")
        lib.am_set_precision(PRECISIONS[env])
    return lib


def _check_device_once(lib):
    """Raise on a visible CUDA device that is not sm_100 (the kernels are sm_100a-only SASS).  On a box without a GPU (the CPU
    build / symbol checks) there is nothing to check: every compute entry point fails at launch there anyway."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu and lib.am_check_device() != 0:
        msg = lib.am_last_error()
        raise AmbError("amb200: the current CUDA device is not compute capability 10.x (B200, sm_100a): "
                       + (msg.decode() if msg else ""))



def check(rc: int, what: str):
    if rc != 0:
        msg = load().am_last_error()
        raise AmbError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().am_launch_count())
