#!/usr/bin/env python3
"""Diagnostic (not product code): where in time do the window+IIR and FFT launches of a
FRA_PIPELINE loop sit?  Builds csrc/ with -DFRA_TIMELINE into tools/libfra_timeline.so (every
launch records its first-CTA start and last-CTA end in %globaltimer), repeats bench.py's entry
sequence (idle, W warm-up steps, sync, K timed steps) several times in one process and prints
the step time and the phase of the two kernels for each repetition.
usage: timeline_probe.py [--build-only] [--reps R] [--warmup W] [--steps K] [--channels C]"""
import argparse
import ctypes
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SO = os.path.join(ROOT, "tools", "libfra_timeline.so")
SRC = os.path.join(ROOT, "fpga_real_time_fft_analyzer_b200", "csrc", "fra_api.cu")

ap = argparse.ArgumentParser()
ap.add_argument("--build-only", action="store_true")
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--channels", type=int, default=4096)
a = ap.parse_args()
if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "-DFRA_USE_F32X2", "-DFRA_TIMELINE", "-shared", "-Xcompiler", "-fPIC", "--cudart", "static",
                           "-o", SO, SRC])
if a.build_only:
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from fpga_real_time_fft_analyzer_b200 import _abi, _lib, synth  # noqa: E402

L = _abi.declare(ctypes.CDLL(SO))
_lib._lib = L                                   # this process only: FraContext drives the diagnostic build
from fpga_real_time_fft_analyzer_b200 import FraContext  # noqa: E402

L.fra_debug_timeline.argtypes = [ctypes.c_void_p]
C, N = a.channels, 16384
ctx = FraContext(C, N, flags=_abi.FRA_PIPELINE)
ctx.command(0x00)
xs = [synth.tone_noise(C, N, "cuda", frame=i) for i in range(3)]
out = {"frames": torch.empty((C, 4 * N), dtype=torch.uint8, device="cuda")}
calls = 0
for rep in range(a.reps):
    L.fra_debug_timeline(None)
    ctx.sync(); torch.cuda.synchronize()
    time.sleep(0.05)
    first = calls
    for i in range(a.warmup):
        ctx.process(xs[i % 3], want=("frames",), out=out); calls += 1
    ctx.sync(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = calls
    e0.record()
    for i in range(a.steps):
        ctx.process(xs[i % 3], want=("frames",), out=out); calls += 1
    ctx.join(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    tl = np.zeros((4, 4096), dtype=np.uint64)
    L.fra_debug_timeline(tl.ctypes.data_as(ctypes.c_void_p))
    tl = tl.astype(np.int64)
    s = t0 + a.steps // 2                                       # a step in the middle of the timed loop
    if s + 3 < 4096:
        base = tl[0, s]
        rel = lambda v: (v - base) / 1e3                        # microseconds
        print(f"rep {rep}: {ms:.4f} ms/step   step {s}: K1 [{rel(tl[0, s]):.0f}, {rel(tl[1, s]):.0f}] us  "
              f"K2(prev) [{rel(tl[2, s - 1]):.0f}, {rel(tl[3, s - 1]):.0f}]  K2 [{rel(tl[2, s]):.0f}, {rel(tl[3, s]):.0f}]  "
              f"K1(next) [{rel(tl[0, s + 1]):.0f}, {rel(tl[1, s + 1]):.0f}]")
    else:
        print(f"rep {rep}: {ms:.4f} ms/step")
ctx.close()
