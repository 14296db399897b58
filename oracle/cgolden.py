"""ctypes binding of oracle/libgolden.so (C golden model).  TEST INFRASTRUCTURE
ONLY - same import rules as oracle/golden.py."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libgolden.so")
    src = os.path.join(_HERE, "golden.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libgolden.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def hann_rom(n=16384):
    out = np.empty(n, dtype=np.int16)
    lib().gold_hann_rom(_p(out, ctypes.c_int16), ctypes.c_int(n))
    return out


def window_iir(x, rom, mode, bank0, bank1, state=None, start=0):
    """x int16 [C, T] -> (y int16 [C, T], state int16 [C, 6, 4])."""
    x = np.ascontiguousarray(x, dtype=np.int16)
    c, t = x.shape
    y = np.empty_like(x)
    rom = np.ascontiguousarray(rom, dtype=np.int16)
    b0 = np.ascontiguousarray(bank0, dtype=np.int8)
    b1 = np.ascontiguousarray(bank1, dtype=np.int8)
    st = (np.zeros((c, 6, 4), dtype=np.int16) if state is None
          else np.ascontiguousarray(state, dtype=np.int16).copy())
    lib().gold_window_iir(_p(x, ctypes.c_int16), _p(y, ctypes.c_int16),
                          ctypes.c_size_t(c), ctypes.c_size_t(t),
                          _p(rom, ctypes.c_int16), ctypes.c_size_t(len(rom)),
                          ctypes.c_size_t(start), ctypes.c_int(mode),
                          _p(b0, ctypes.c_int8), _p(b1, ctypes.c_int8),
                          _p(st, ctypes.c_int16))
    return y, st


def window_iir_sections(x, rom, coeff6x6, state=None, start=0):
    """Window, then six independent sections (int8 [6][6]): x int16 [C, T] -> (y, state)."""
    x = np.ascontiguousarray(x, dtype=np.int16)
    c, t = x.shape
    w = np.empty_like(x)
    y = np.empty_like(x)
    rom = np.ascontiguousarray(rom, dtype=np.int16)
    k = np.ascontiguousarray(coeff6x6, dtype=np.int8).reshape(36)
    st = (np.zeros((c, 6, 4), dtype=np.int16) if state is None
          else np.ascontiguousarray(state, dtype=np.int16).copy())
    lib().gold_window(_p(x, ctypes.c_int16), _p(w, ctypes.c_int16), ctypes.c_size_t(c), ctypes.c_size_t(t),
                      _p(rom, ctypes.c_int16), ctypes.c_size_t(len(rom)), ctypes.c_size_t(start))
    lib().gold_iir_sections(_p(w, ctypes.c_int16), _p(y, ctypes.c_int16), ctypes.c_size_t(c), ctypes.c_size_t(t),
                            _p(k, ctypes.c_int8), _p(st, ctypes.c_int16))
    return y, st


def iir12(x, coeff12, state=None):
    x = np.ascontiguousarray(x, dtype=np.int16)
    c, t = x.shape
    y = np.empty_like(x)
    k = np.ascontiguousarray(coeff12, dtype=np.int8)
    st = (np.zeros((c, 6, 4), dtype=np.int16) if state is None
          else np.ascontiguousarray(state, dtype=np.int16).copy())
    lib().gold_iir12(_p(x, ctypes.c_int16), _p(y, ctypes.c_int16),
                     ctypes.c_size_t(c), ctypes.c_size_t(t),
                     _p(k, ctypes.c_int8), _p(st, ctypes.c_int16))
    return y, st


def quantize_pack(bins, log2_scale=-14, rounding=0):
    """complex128 [..., N] -> uint8 [..., 4N]."""
    bins = np.asarray(bins)
    flat = bins.reshape(-1, bins.shape[-1])
    out = np.empty((flat.shape[0], 4 * flat.shape[1]), dtype=np.uint8)
    for i in range(flat.shape[0]):
        re = np.ascontiguousarray(flat[i].real, dtype=np.float64)
        im = np.ascontiguousarray(flat[i].imag, dtype=np.float64)
        lib().gold_quantize_pack(_p(re, ctypes.c_double), _p(im, ctypes.c_double),
                                 ctypes.c_size_t(flat.shape[1]), ctypes.c_int(log2_scale),
                                 ctypes.c_int(rounding), _p(out[i], ctypes.c_uint8))
    return out.reshape(*bins.shape[:-1], -1)
