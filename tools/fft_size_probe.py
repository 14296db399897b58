#!/usr/bin/env python3
"""FFT timing at one frame size: fft_only (int16 in, complex64 out) and the bypass chain with int16 frames
out (window fused into the FFT's first pass), for a list of fra_create flag values.
usage: fft_size_probe.py N [flags ...]   (flags in hex, e.g. 0 0x2000 0x400)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpga_real_time_fft_analyzer_b200 import FraContext  # noqa: E402

n = int(sys.argv[1])
flag_list = [int(f, 16) for f in sys.argv[2:]] or [0]
b = (1 << 26) // n
x = torch.randint(-32768, 32767, (b, n), dtype=torch.int16, device="cuda")
ref = np.fft.fft(x[:2].cpu().numpy().astype(np.float64), axis=-1)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for flags in flag_list:
    with FraContext(b, n, flags=flags) as ctx:
        y = ctx.fft_only(x)
        got = y[:2].cpu().numpy()
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        ms = timed(lambda: ctx.fft_only(x))
        out = {"frames": torch.empty((b, 4 * n), dtype=torch.uint8, device="cuda")}
        ms2 = timed(lambda: ctx.process(x, out=out))
        print(f"N={n} batch={b} flags={flags:#x}: fft_only {ms:.4f} ms {b * n / ms / 1e6:.1f} Gsamples/s (rel L2 {err:.2e}) | "
              f"bypass chain -> frames {ms2:.4f} ms {b * n / ms2 / 1e6:.1f} Gsamples/s", flush=True)
