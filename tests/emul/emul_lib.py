"""Builds and drives tests/emul/libfra_emul.so: the UNCHANGED csrc/ sources
compiled as host C++ on top of cusim.h.  TEST INFRASTRUCTURE ONLY - a way to run
the kernel source on the CPU before spending GPU time; never loaded by the
product package (fpga_real_time_fft_analyzer_b200/_lib.py loads libfra.so only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from fpga_real_time_fft_analyzer_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "fpga_real_time_fft_analyzer_b200", "csrc")
_SO = os.path.join(_HERE, "libfra_emul.so")
_lib = None


def build():
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [os.path.join(_HERE, "cusim.h"),
                                                                 os.path.join(_ROOT, "include", "fra.h")]
    if os.path.exists(_SO) and all(os.path.getmtime(s) <= os.path.getmtime(_SO) for s in srcs):
        return _SO
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-frounding-math", "-fPIC", "-shared", "-DFRA_HOST_EMUL",
                           "-I" + _HERE, "-x", "c++", os.path.join(_CSRC, "fra_api.cu"), "-o", _SO, "-lpthread"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = _abi.declare(C.CDLL(build()))
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


class EmulFra:
    """numpy-in / numpy-out driver of the emulated library (host memory plays device memory)."""

    def __init__(self, channels, n=16384, flags=0):
        self.L = lib()
        self.h = C.c_void_p()
        rc = self.L.fra_create(C.byref(self.h), 0, channels, n, flags)
        assert rc == 0, rc
        self.c, self.n = channels, n

    def close(self):
        if self.h:
            self.L.fra_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def command(self, data: bytes):
        return self.L.fra_command(self.h, bytes(data), len(data))

    def process(self, x, continuous=False, log2_scale=_abi.FRA_SCALE_DEFAULT, want=("filtered", "frames", "iq"),
                host=False, out=None):
        x = np.ascontiguousarray(x, dtype=np.int16).reshape(self.c, self.n)
        if out is not None:
            want = ()
        out = {} if out is None else out
        if "filtered" in want:
            out["filtered"] = np.zeros((self.c, self.n), np.int16)
        if "frames" in want:
            out["frames"] = np.zeros((self.c, 4 * self.n), np.uint8)
        if "iq" in want:
            out["iq"] = np.zeros((self.c, self.n, 2), np.float32)
        if "mag" in want:
            out["mag"] = np.zeros((self.c, self.n), np.float32)
        if "phase" in want:
            out["phase"] = np.zeros((self.c, self.n), np.float32)
        o = _abi.FraOutputs(_ptr(out.get("filtered")), _ptr(out.get("frames")), _ptr(out.get("iq")),
                            _ptr(out.get("mag")), _ptr(out.get("phase")))
        if host:
            rc = self.L.fra_process_host(self.h, x.ctypes.data, int(continuous), log2_scale, C.byref(o))
        else:
            rc = self.L.fra_process(self.h, x.ctypes.data, int(continuous), log2_scale, C.byref(o), None)
        assert rc == 0, (rc, self.L.fra_last_cuda_error(self.h))
        return out

    def load_sections(self, coeff6x6):
        arr = (C.c_int8 * 36)(*[int(v) for v in np.asarray(coeff6x6).reshape(36)])
        assert self.L.fra_load_sections(self.h, arr) == 0

    def get_state(self):
        st = np.zeros((self.c, 6, 4), np.int16)
        assert self.L.fra_get_state(self.h, st.ctypes.data, None) == 0
        return st

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=np.int16)
        assert self.L.fra_set_state(self.h, st.ctypes.data, None) == 0

    def iir_stream(self, x, continuous=False, exact=False):
        x = np.ascontiguousarray(x, dtype=np.int16)
        y = np.zeros_like(x)
        st = _abi.FraStreamStats()
        rc = self.L.fra_iir_stream(self.h, x.ctypes.data, y.ctypes.data, x.size, int(continuous), int(exact),
                                   C.byref(st))
        assert rc == 0, rc
        return y, {k: getattr(st, k) for k, _ in st._fields_}

    def fft_only(self, x):
        x = np.ascontiguousarray(x, dtype=np.int16).reshape(-1, self.n)
        iq = np.zeros((x.shape[0], self.n, 2), np.float32)
        rc = self.L.fra_fft_only(self.h, x.ctypes.data, x.shape[0], iq.ctypes.data, None)
        assert rc == 0, rc
        return iq[..., 0] + 1j * iq[..., 1]
