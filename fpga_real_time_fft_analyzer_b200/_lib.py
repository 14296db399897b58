"""Loads the one product library, libfra.so, from this package directory.

There is deliberately no path override, no environment variable and no CPU
fallback: if the CUDA extension is missing or there is no CUDA device, the
failure is loud (ImportError / FraError)."""
from __future__ import annotations

import ctypes
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfra.so")
_lib = None


class FraError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        msg = f"{where}: {status_text(status)} ({status})"
        if detail:
            msg += f" - {detail}"
        super().__init__(msg)


def status_text(status):
    try:
        return lib().fra_strerror(status).decode()
    except Exception:
        return "?"


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        cdll = ctypes.CDLL(LIB_PATH)
        _abi.declare(cdll)
        if cdll.fra_abi_version() != _abi.FRA_ABI_VERSION:
            raise ImportError("libfra.so ABI version mismatch; rebuild")
        _lib = cdll
    return _lib
