"""Builds libfra.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "fra_api.cu")
OUT = os.path.join(_HERE, "libfra.so")
DEPS = ["fra_api.cu", "fra_common.cuh", "k1_window_iir.cuh", "k1b_stream.cuh", "k2_fft.cuh", "k2_fixed.cuh", "hann_rom_q15.inc"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    hdr = os.path.join(os.path.dirname(_HERE), "include", "fra.h")
    files = [os.path.join(_HERE, "csrc", d) for d in DEPS] + [hdr]
    return any(os.path.getmtime(f) > t for f in files)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: experiment builds (tools/build_variants.py) beside the product library."""
    if out is None and not force and not stale():
        return OUT
    out = out or OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-DFRA_USE_F32X2", *[f"-D{d}" for d in defines], "-shared", "-Xcompiler", "-fPIC", "--cudart", "static",
           "-o", out, SRC]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
