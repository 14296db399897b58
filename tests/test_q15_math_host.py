"""Exhaustive host check of the FP32-pipe identities behind the bit-exact biquad:
fma.rm / fma.rp accumulate floor(v*c/128) for every int16 x int8, the low-16-bit
wrap, and the window's round/resize for every int16 x int16 (tests/host/check_q15_math.cpp
includes csrc/fra_common.cuh itself, so the primitives checked are the shipped ones)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fp32_floor_identities_exhaustive(tmp_path):
    exe = str(tmp_path / "check_q15")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-frounding-math", "-DFRA_HOST_EMUL",
                           "-I" + os.path.join(ROOT, "tests", "emul"),
                           "-I" + os.path.join(ROOT, "fpga_real_time_fft_analyzer_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "check_q15_math.cpp"), "-o", exe, "-lpthread"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout
