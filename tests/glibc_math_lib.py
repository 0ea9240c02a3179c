"""Builds tests/glibc_math_check.c (the host build of gapprober_b200/csrc/fb_glibc_math.h next to the host libm)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_lib = None


def host_has_fma():
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return " fma " in flags and " avx2 " in flags


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "_glibc_math_check.so")
        src = os.path.join(HERE, "glibc_math_check.c")
        hdrs = [os.path.join(ROOT, "gapprober_b200", "csrc", f) for f in ("fb_glibc_math.h", "fb_glibc_tables.h")]
        if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
            subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-o", so, src, "-lm"])
        _lib = C.CDLL(so)
        _lib.gm_check.restype = C.c_long
        _lib.gm_check.argtypes = [C.c_int, C.c_long, C.c_uint64]
        _lib.gm_check_edges.restype = C.c_long
        dp = C.POINTER(C.c_double)
        _lib.gm_inputs.argtypes = [C.c_int, dp, C.c_long, C.c_uint64]
        _lib.gm_libm.argtypes = [C.c_int, dp, dp, C.c_long]
    return _lib
