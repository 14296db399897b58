#!/usr/bin/env python3
"""A few launches of the FFT alone (bypass chain -> int16 frames, or fft_only) for ncu.
usage: prof_fft.py N flags_hex [fft_only]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpga_real_time_fft_analyzer_b200 import FraContext  # noqa: E402

n = int(sys.argv[1])
flags = int(sys.argv[2], 16)
b = (1 << 26) // n
x = torch.randint(-32768, 32767, (b, n), dtype=torch.int16, device="cuda")
with FraContext(b, n, flags=flags) as ctx:
    out = {"frames": torch.empty((b, 4 * n), dtype=torch.uint8, device="cuda")}
    for _ in range(3):
        if len(sys.argv) > 3:
            ctx.fft_only(x)
        else:
            ctx.process(x, out=out)
    torch.cuda.synchronize()
