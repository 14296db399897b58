#!/usr/bin/env python3
"""BASELINE config 5 timings: one long stream through K1b (exact systolic chain vs block scan)
and the FFT alone at N = 1K .. 32K with batch = 2^26 / N samples."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fpga_real_time_fft_analyzer_b200 import FraContext, synth  # noqa: E402

dev = "cuda"
with FraContext(1, 16384) as ctx:
    ctx.command(0x00)
    y_exact = None
    for log2len, modes in ((24, (True, False)), (26, (False,))):
        n = 1 << log2len
        x = synth.tone_noise(1, n, dev)[0].contiguous()
        for exact in modes:
            ctx.iir_stream(x[: 1 << 20], exact=exact)
            torch.cuda.synchronize()
            best = 1e9
            for rep in range(1 if exact else 3):
                t0 = time.perf_counter()
                y, st = ctx.iir_stream(x, exact=exact)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            note = ""
            if exact:
                y_exact = y
            elif y_exact is not None and y_exact.numel() == y.numel():
                d = (y.to(torch.int32) - y_exact.to(torch.int32)).abs()
                note = f"  max |output - exact stream| = {int(d.max())} LSB, mean {float(d.float().mean()):.3f}"
            print(f"stream 2^{log2len} samples exact={int(exact)}: {best * 1e3:.2f} ms  {n / best / 1e9:.3f} Gsamples/s  stats={st}{note}")
for log2n in range(10, 17):
    N = 1 << log2n
    batch = (1 << 26) // N
    with FraContext(batch, N) as ctx:
        xs = synth.full_range(batch, N, dev)
        for _ in range(3):                 # warm-up: module load, and the caching allocator gets both output blocks
            out = ctx.fft_only(xs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            out = ctx.fft_only(xs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"fft_only N={N:6d} batch={batch:6d}: {ms:.4f} ms  {batch * N / ms / 1e6:.1f} Gsamples/s (int16 in, complex64 out: "
              f"{batch * N * 10 / ms / 1e6:.0f} GB/s)")
