"""ORACLE — TEST INFRASTRUCTURE ONLY.  Builds oracle/csrc/ld_filters.c (long-double IIR
"truth", see the file header) into oracle/_c/libwfm_ld.so (git-ignored, travels to the
GPU box with the snapshot) and loads it through ctypes.

    python oracle/build_c.py
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / 'csrc' / 'ld_filters.c'
LIB = HERE / '_c' / 'libwfm_ld.so'


def build(force=False):
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    subprocess.run(['gcc', '-O2', '-fPIC', '-shared', '-o', str(LIB), str(SRC)], check=True)
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(build()))
        lib.wfm_ld_sosfilt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_long]
        lib.wfm_ld_lfilter.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_long,
                                       C.c_void_p]
        _lib = lib
    return _lib


def sosfilt_ld(sos, x):
    """scipy.signal.sosfilt(sos, x) with every operation in x87 long double."""
    sos = np.ascontiguousarray(np.asarray(sos, dtype=np.float64)).reshape(-1, 6)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    _load().wfm_ld_sosfilt(sos.ctypes.data, len(sos), x.ctypes.data, y.ctypes.data, len(x))
    return y


def lfilter_ld(b, a, x, zi=None):
    """scipy.signal.lfilter(b, a, x, zi=zi)[0] with every operation in x87 long double."""
    b = np.ascontiguousarray(b, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    z = None if zi is None else np.ascontiguousarray(zi, dtype=np.float64)
    _load().wfm_ld_lfilter(b.ctypes.data, len(b), a.ctypes.data, len(a), x.ctypes.data, y.ctypes.data, len(x),
                           None if z is None else z.ctypes.data)
    return y


if __name__ == '__main__':
    print(build(force=True))
