#!/usr/bin/env python3
"""One small invocation of every kernel family, for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
Sizes are small because the sanitizer slows kernels down 10-100x; results are still checked."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from fpga_real_time_fft_analyzer_b200 import FraContext, _abi  # noqa: E402
from oracle import cgolden as cg  # noqa: E402
from oracle import golden as g  # noqa: E402

rom = np.fromfile(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "hann_rom.i16"), dtype="<i2")
B1 = np.array([32, 10, -33, 119, 35, 0, 52, -16, 11, 84, -10, 0], dtype=np.int8)
rng = np.random.default_rng(1)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


n = 2048
for name, flags, c in (("duo", _abi.FRA_K1_FORCE_DUO, 37), ("lane_biased", _abi.FRA_K1_FORCE_LANE, 70),
                       ("lane_general", _abi.FRA_K1_FORCE_LANE | _abi.FRA_K1_NO_BIASED, 33), ("split", _abi.FRA_K1_FORCE_SPLIT, 7)):
    x = rng.integers(-32768, 32768, (c, n)).astype(np.int16)
    with FraContext(c, n, flags=flags) as ctx:
        ctx.command(0x00)
        out = ctx.process(dev(x), want=("filtered", "frames", "mag"))
        y, _ = cg.window_iir(x, rom, 0, g.BANK0_COEFF, B1)
        assert np.array_equal(out["filtered"].cpu().numpy(), y), name
    print("ok", name, flush=True)
for nn, extra in ((16384, 0), (16384, _abi.FRA_K2_STAGED), (1024, 0), (32768, 0), (65536, 0), (65536, _abi.FRA_K2_64K_SPLIT), (4096, _abi.FRA_FFT_FIXED16)):
    c = 3
    x = rng.integers(-32768, 32768, (c, nn)).astype(np.int16)
    with FraContext(c, nn, flags=extra) as ctx:
        out = ctx.process(dev(x), want=("frames", "iq", "phase"))
        torch.cuda.synchronize()
        if not extra & _abi.FRA_FFT_FIXED16:
            ref = np.fft.fft(g.window(x, rom).astype(np.float64), axis=-1)
            iq = out["iq"].cpu().numpy()
            err = np.linalg.norm((iq[..., 0] + 1j * iq[..., 1]) - ref) / np.linalg.norm(ref)
            assert err < 1e-4, (nn, extra, err)
    print("ok fft", nn, hex(extra), flush=True)
with FraContext(1, 16384) as ctx:
    ctx.command(0x00)
    x = g.tone_noise([3], n=1 << 20, seed=2)[0]
    y, st = ctx.iir_stream(dev(x), exact=False)
    assert st["chunk"] == 512 and st["max_state_dev"] <= 32
    y, st = ctx.iir_stream(dev(x[: 1 << 14]), exact=True)
    print("ok stream", st, flush=True)
with FraContext(64, 16384, flags=_abi.FRA_PIPELINE) as ctx:
    ctx.command(0x00)
    xs = [dev(rng.integers(-2048, 2048, (64, 16384)).astype(np.int16)) for _ in range(3)]
    outs = [ctx.process(x, continuous=i > 0, want=("frames",)) for i, x in enumerate(xs)]
    ctx.join()
    torch.cuda.synchronize()
    print("ok pipeline", flush=True)
print("sanitize_case: all ok")
