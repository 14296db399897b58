#!/usr/bin/env python3
"""e2e loop (process_host_async, two calls in flight) at fixed half-spectrum shares: step time, wait / mirror split.
usage: e2e_probe.py [--channels C] [--steps K] [--shares 0,0.5,1,-1]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fpga_real_time_fft_analyzer_b200 import FraContext, _abi, synth

ap = argparse.ArgumentParser()
ap.add_argument("--channels", type=int, default=65536)
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--shares", default="0,0.5,0.75,1,-1")
a = ap.parse_args()
ctx = FraContext(a.channels, a.n, flags=_abi.FRA_HOST_HALF_SPECTRUM)
ctx.command(0x00)
xh = [synth.tone_noise(a.channels, a.n, "cuda", frame=i).cpu().pin_memory() for i in range(2)]
samples = a.channels * a.n


def loop(steps):
    pend, log = None, []
    for i in range(steps):
        cur = ctx.process_host_async(xh[i % 2], continuous=True, want=("frames",))
        if pend is not None:
            ctx.host_wait(pend[1])
            sh = ctx.host_transfer()
            log.append((ctx.host_wait_s, ctx.host_mirror_s, sh[1], sh[2]))
        pend = cur
    ctx.host_wait(pend[1])
    return log


loop(3)
for sh in [float(v) for v in a.shares.split(",")]:
    ctx.set_host_half_share(sh)
    loop(2)
    t0 = time.perf_counter()
    log = loop(a.steps if sh >= 0 else 3 * a.steps)
    dt = (time.perf_counter() - t0) / (a.steps if sh >= 0 else 3 * a.steps)
    w = sum(l[0] for l in log) / len(log); m = sum(l[1] for l in log) / len(log)
    print(f"share {sh:5.2f}: {dt * 1e3:7.2f} ms/step  {samples / dt / 1e9:6.2f} Gs/s   wait {w * 1e3:6.2f} ms  mirror {m * 1e3:6.2f} ms  "
          f"d2h {log[-1][2] / 2**30:.2f} GiB  share now {log[-1][3]:.3f}", flush=True)
    if sh < 0:
        print("   adaptive trace:", " ".join(f"{l[3]:.2f}" for l in log))
