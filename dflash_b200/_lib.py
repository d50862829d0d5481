"""ctypes binding of libdflash_b200.so (the C ABI in include/dflash_b200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing, or the device
is not sm_100, every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# DFLASH_LIB: another build of the same library (kernel-tuning experiments, scripts/sweep_*.sh); never a fallback
LIB_PATH = os.environ.get("DFLASH_LIB") or os.path.join(_HERE, "libdflash_b200.so")

OK = 0


class DFlashNativeError(RuntimeError):
    pass


_lib = None


def _declare(lib):
    lib.dflash_abi_version.restype = c_int
    lib.dflash_last_error.restype = c_char_p
    lib.dflash_device_check.restype = c_int
    lib.dflash_gemm_max_slots.restype = c_int
    lib.dflash_gemm_max_slots.argtypes = [c_int, c_int, c_int]
    lib.dflash_gemm_skinny.restype = c_int
    lib.dflash_gemm_skinny.argtypes = [
        c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
        c_longlong, c_void_p, c_longlong, c_int, c_int, c_void_p,
    ]
    lib.dflash_gemm_swiglu.restype = c_int
    lib.dflash_gemm_swiglu.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_longlong,
                                       c_void_p, c_void_p, c_int, c_int, c_void_p]
    lib.dflash_gemm_argmax.restype = c_int
    lib.dflash_gemm_argmax.argtypes = [
        c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
        c_void_p, c_longlong, c_void_p, c_int, c_int, c_void_p,
    ]


def load():
    """Load the native library or raise DFlashNativeError (never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DFlashNativeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). dflash_b200 has no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    _declare(lib)
    _lib = lib
    return lib


def check(rc: int, what: str = "dflash"):
    if rc < 0:
        msg = load().dflash_last_error().decode("utf-8", "replace")
        raise DFlashNativeError(f"{what} failed ({rc}): {msg}")
    return rc


def exported_symbols():
    """Names declared in include/dflash_b200.h (parsed), for the ABI export test."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "dflash_b200.h")
    txt = open(hdr).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dflash_[a-z0-9_]+)\s*\(", txt)))
