#!/usr/bin/env python3
"""Runs a few steps of the hot path for ncu / timing (not a benchmark).
usage: prof_step.py [--channels C] [--steps K] [--mode 0x00|0xB1] [--k1 auto|lane|split] [--time]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fpga_real_time_fft_analyzer_b200 import FraContext, _abi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--channels", type=int, default=4096)
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--mode", default="0x00")
ap.add_argument("--k1", default="auto")
ap.add_argument("--time", action="store_true")
ap.add_argument("--pipeline", action="store_true", help="FRA_PIPELINE context; reports whole-loop throughput")
ap.add_argument("--flags", type=lambda v: int(v, 0), default=0, help="extra fra_create flags (e.g. 0x80 = FRA_K1_NO_BIASED)")
ap.add_argument("--hiccup", type=float, default=0.0, help="host sleep (s) every 40 steps inside the timed loop")
a = ap.parse_args()
flags = {"auto": 0, "lane": _abi.FRA_K1_FORCE_LANE, "split": _abi.FRA_K1_FORCE_SPLIT,
         "spec": _abi.FRA_K1_SPECULATE | _abi.FRA_K1_FORCE_SPLIT, "duo": _abi.FRA_K1_FORCE_DUO}[a.k1]
if a.pipeline:
    flags |= _abi.FRA_PIPELINE
flags |= a.flags
ctx = FraContext(a.channels, a.n, flags=flags)
ctx.command(int(a.mode, 16))
xs = [synth.tone_noise(a.channels, a.n, "cuda", frame=i) for i in range(2)]
out = {"frames": torch.empty((a.channels, 4 * a.n), dtype=torch.uint8, device="cuda")}
if a.pipeline or a.time:
    # whole-loop throughput (events on torch's stream; join() makes it wait for the internal streams)
    xs3 = xs + [synth.tone_noise(a.channels, a.n, "cuda", frame=2)]
    for i in range(3):
        ctx.process(xs3[i % 3], want=("frames",), out=out)
    ctx.sync(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(a.steps, 20)
    import time
    if a.pipeline and not os.environ.get("FRA_NO_PROFILE"):
        ctx.profile(True)
    e0.record()
    for i in range(reps):
        ctx.process(xs3[i % 3], want=("frames",), out=out)
        if a.hiccup and i % 40 == 20:
            time.sleep(a.hiccup)
    ctx.join()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if a.pipeline and not os.environ.get("FRA_NO_PROFILE"):
        t = ctx.profile_last()
        print(f"  last step's kernel spans while overlapped: K1 {t[0]:.4f} ms  K2 {t[1]:.4f} ms")
    print(f"channels={a.channels} n={a.n} pipeline={int(a.pipeline)} k1={a.k1}: {ms:.4f} ms/step  "
          f"{a.channels * a.n / ms / 1e6:.1f} Gs/s over {reps} steps")
    if a.pipeline:
        print("done"); sys.exit(0)
ctx.profile(True)
k1, k2 = [], []
for i in range(a.steps):
    ctx.process(xs[i % 2], want=("frames",), out=out)
    if a.time:
        t = ctx.profile_last()
        k1.append(t[0]); k2.append(t[1])
torch.cuda.synchronize()
if a.time:
    s = a.channels * a.n / 1e9
    m1, m2 = min(k1[1:]), min(k2[1:])
    print(f"channels={a.channels} n={a.n} mode={a.mode} k1={a.k1}: K1 {m1:.4f} ms ({s / m1 * 1e3 if m1 else 0:.1f} Gs/s)  "
          f"K2 {m2:.4f} ms ({s / m2 * 1e3:.1f} Gs/s)  chain {s / (m1 + m2) * 1e3:.1f} Gs/s")
print("done")
