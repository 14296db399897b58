"""The kernel SOURCE (csrc/*.cuh) run on the CPU under tests/emul/cusim.h and checked
against the oracle at small sizes - index arithmetic, systolic skew, barriers,
swizzles, the C-ABI host logic.  This is a pre-flight for the real thing
(tests/test_gpu_parity.py, -m gpu): the emulated library is test infrastructure and
is never reachable from the product package."""
import ctypes as C

import numpy as np
import pytest

from fpga_real_time_fft_analyzer_b200 import _abi
from oracle import cgolden as cg
from oracle import golden as g
from tests.emul.emul_lib import EmulFra, lib

B1 = np.array([32, 10, -33, 119, 35, 0, 52, -16, 11, 84, -10, 0], dtype=np.int8)


def adversarial(rng, c, n):
    x = rng.integers(-32768, 32768, size=(c, n)).astype(np.int16)
    x[0, :40] = -32768                      # the window's resize quirk at rom = -32768 (n = 0..14)
    x[-1, -40:] = -32768
    return x


def biased_edge_sections(rng):
    """Six sections on the edge of biased_order_ok() (fra_common.cuh): the coefficient sums of the
    product-order suffixes (y[n-2] | x[n-2], x[n-1], x[n], y[n-1]) reach +-60, the first product's
    coefficient (A0) is anything, including -128 / 127."""
    sec = np.zeros((6, 6), dtype=np.int8)
    pats = [(60, 0, -60, -128, -60),      # B0, B1, B2, A0, A1: suffixes -A1=60 | B2-A1=0 | B1+..=0 | B0+..=60
            (-60, 60, -60, 127, 0),       # suffixes 0 | -60 | 0 | -60
            (0, -120, 120, -1, 60),       # -60 | 60 | -60 | -60
            (-14, 0, 14, 107, 21),        # the reference's fixed bank
            (-15, 0, 15, 107, -21),
            (127, -127, 60, 64, 0)]       # 0 | 60 | -67?  replaced below if not eligible
    def ok(b0, b1, b2, a0, a1):
        t = [-a0, b0, b1, b2, -a1]
        s = 0
        for k in range(4, 0, -1):
            s += t[k]
            if abs(s) > 60:
                return False
        return True
    for i, p in enumerate(pats):
        if not ok(*p):
            while True:
                p = tuple(int(v) for v in rng.integers(-128, 128, 5))
                if ok(*p):
                    break
        sec[i, :5] = p
    assert all(ok(*[int(v) for v in sec[i, :5]]) for i in range(6))
    return sec


@pytest.mark.parametrize("flags,name", [(_abi.FRA_K1_FORCE_LANE, "lane"), (_abi.FRA_K1_FORCE_SPLIT, "split"),
                                        (_abi.FRA_K1_FORCE_DUO, "duo")])
@pytest.mark.parametrize("channels", [1, 6, 37])
def test_k1_bit_exact_with_state_and_reload(flags, name, channels, rom):
    rng = np.random.default_rng(channels)
    n = 1024
    f = EmulFra(channels, n, flags)
    try:
        f.command(bytes([0x00]))
        x = adversarial(rng, channels, n)
        y, st = cg.window_iir(x, rom, 0x00, g.BANK0_COEFF, B1)
        out = f.process(x, want=("filtered",))
        assert np.array_equal(out["filtered"], y)
        assert np.array_equal(f.get_state(), st)
        # 0xF1 upload + switch to bank 1 between frames: history is kept
        assert f.command(bytes([0xF1]) + B1.tobytes() + bytes([0xA1])) == 0
        x2 = adversarial(rng, channels, n)
        y2, st2 = cg.window_iir(x2, rom, 0xA1, g.BANK0_COEFF, B1, st)
        out2 = f.process(x2, continuous=True, want=("filtered",))
        assert np.array_equal(out2["filtered"], y2) and np.array_equal(f.get_state(), st2)
        # third frame after a gap: history restarts at zero
        y3, st3 = cg.window_iir(x2, rom, 0xA1, g.BANK0_COEFF, B1, None)
        out3 = f.process(x2, continuous=False, want=("filtered",))
        assert np.array_equal(out3["filtered"], y3) and np.array_equal(f.get_state(), st3)
    finally:
        f.close()


@pytest.mark.parametrize("flags", [_abi.FRA_K1_FORCE_LANE, _abi.FRA_K1_FORCE_SPLIT,
                                   _abi.FRA_K1_FORCE_DUO])
def test_six_independent_sections(flags, rom):
    """fra_load_sections: every stage its own coefficients."""
    rng = np.random.default_rng(21)
    c, n = 5, 1024
    f = EmulFra(c, n, flags)
    try:
        for trial in range(2):
            sec = rng.integers(-128, 128, (6, 6)).astype(np.int8)
            if trial == 1:
                sec[:, 4] = rng.integers(-60, 61, 6)          # k1_duo's two-instruction recurrence
                sec[:, 1] = 0                                 # and the skipped x[n-1] product
            f.load_sections(sec)
            f.command(bytes([0xA1]))
            st0 = rng.integers(-32768, 32768, (c, 6, 4)).astype(np.int16)
            f.set_state(st0)
            x = adversarial(rng, c, n)
            y, st = cg.window_iir_sections(x, rom, sec, st0)
            out = f.process(x, continuous=True, want=("filtered",))
            assert np.array_equal(out["filtered"], y) and np.array_equal(f.get_state(), st)
        # a 12-byte upload restores the RTL's two alternating sets
        f.command(bytes([0xF1]) + B1.tobytes())
        x = adversarial(rng, c, n)
        y, st = cg.window_iir(x, rom, 0xA1, g.BANK0_COEFF, B1)
        out = f.process(x, continuous=False, want=("filtered",))
        assert np.array_equal(out["filtered"], y)
    finally:
        f.close()


def test_magnitude_averaging_in_the_pack_stage(rom):
    """fra_set_mag_average: the caller's magnitude buffer holds an exponential moving average."""
    rng = np.random.default_rng(9)
    c, n, alpha = 3, 1024, 0.25
    f = EmulFra(c, n)
    try:
        assert f.L.fra_set_mag_average(f.h, 0.0) != 0 and f.L.fra_set_mag_average(f.h, 1.5) != 0
        assert f.L.fra_set_mag_average(f.h, alpha) == 0
        out = {"mag": np.zeros((c, n), np.float32), "frames": np.zeros((c, 4 * n), np.uint8)}
        avg = np.zeros((c, n), np.float32)
        for frame in range(4):
            x = adversarial(rng, c, n)
            f.process(x, out=out)
            _, _, mag = g.decode_frame(out["frames"])
            avg = (avg + np.float32(alpha) * (mag.astype(np.float32) - avg)).astype(np.float32)
            assert np.allclose(out["mag"], avg, rtol=1e-6, atol=1e-4)
        assert f.L.fra_set_mag_average(f.h, 1.0) == 0
        f.process(x, out=out)
        _, _, mag = g.decode_frame(out["frames"])
        assert np.array_equal(out["mag"].view(np.uint32), mag.astype(np.float32).view(np.uint32))
    finally:
        f.close()


def test_pipeline_flag_same_results(rom):
    """FRA_PIPELINE only changes which streams the kernels go to: the host logic around it
    (two scratch buffers, hand-over events) must leave results and state unchanged."""
    rng = np.random.default_rng(5)
    c, n = 6, 1024
    f = EmulFra(c, n, _abi.FRA_PIPELINE)
    try:
        f.command(bytes([0x00]))
        st = None
        for frame in range(3):
            x = adversarial(rng, c, n)
            y, st = cg.window_iir(x, rom, 0x00, g.BANK0_COEFF, B1, st)
            out = f.process(x, continuous=frame > 0, want=("filtered", "frames"))
            assert np.array_equal(out["filtered"], y)
            out2 = f.process(x, continuous=False, want=("frames",))       # library-owned scratch, alternating
            y0, st = cg.window_iir(x, rom, 0x00, g.BANK0_COEFF, B1, None)
            assert np.array_equal(f.get_state(), st)                      # joins: the held-back FFT is flushed
            seq = EmulFra(c, n)
            try:
                seq.command(bytes([0x00]))
                ref = seq.process(x, continuous=False, want=("frames",))
            finally:
                seq.close()
            assert np.array_equal(out2["frames"], ref["frames"])
            # the host-buffer call on a pipelined context drains the pipeline first
            out3 = f.process(x, continuous=False, want=("frames", "filtered"), host=True)
            assert np.array_equal(out3["frames"], ref["frames"]) and np.array_equal(out3["filtered"], y0)
    finally:
        f.close()


def test_pipeline_flag_64k_frames(rom):
    """FRA_PIPELINE with the three-kernel 64K FFT path as the held-back launch."""
    rng = np.random.default_rng(6)
    c, n = 2, 65536
    f, seq = EmulFra(c, n, _abi.FRA_PIPELINE), EmulFra(c, n)
    try:
        f.command(bytes([0x00])); seq.command(bytes([0x00]))
        outs, refs = [], []
        for frame in range(2):
            x = adversarial(rng, c, n)
            outs.append(f.process(x, continuous=frame > 0, want=("frames",)))
            refs.append(seq.process(x, continuous=frame > 0, want=("frames",)))
        assert f.L.fra_join(f.h, None) == 0
        for o, r in zip(outs, refs):
            assert np.array_equal(o["frames"], r["frames"])
        assert np.array_equal(f.get_state(), seq.get_state())
    finally:
        f.close(); seq.close()


def test_k1_random_coefficients_and_user_state(rom):
    rng = np.random.default_rng(7)
    n, c = 1024, 7
    for trial in range(5):
        coef = rng.integers(-128, 128, 12).astype(np.int8)
        if trial == 0:
            coef[:] = [-128, 127, -128, -128, 127, 0, 127, -128, 127, 127, -128, 0]
        if trial == 3:      # |A1| at the limit of the two-instruction recurrence (k1_duo FAST), the rest extreme
            coef[:] = [-128, 127, -128, -128, 60, 0, 127, -128, 127, 127, -60, 0]
        if trial == 4:
            coef[4], coef[10] = rng.integers(-60, 61, 2)
        for flags in (_abi.FRA_K1_FORCE_LANE, _abi.FRA_K1_FORCE_SPLIT, _abi.FRA_K1_FORCE_DUO,
                      _abi.FRA_K1_FORCE_SPLIT | _abi.FRA_K1_SPECULATE):      # speculation must roll back correctly
            f = EmulFra(c, n, flags)
            try:
                f.command(bytes([0xF1]) + coef.tobytes() + bytes([0xA1]))
                st0 = rng.integers(-32768, 32768, (c, 6, 4)).astype(np.int16)
                f.set_state(st0)
                x = adversarial(rng, c, n)
                y, st = cg.window_iir(x, rom, 0xA1, g.BANK0_COEFF, coef, st0)
                out = f.process(x, continuous=True, want=("filtered",))
                assert np.array_equal(out["filtered"], y) and np.array_equal(f.get_state(), st)
            finally:
                f.close()


@pytest.mark.parametrize("n", [1024, 2048, 4096, 8192, 16384, 32768, 65536])
def test_k2_fft_all_sizes_and_framing(n, rom):
    rng = np.random.default_rng(n)
    b = max(1, 16384 // n) + 1                      # more than one CTA, last CTA partly empty
    x = adversarial(rng, b, n)
    f = EmulFra(b, n)
    try:
        got = f.fft_only(x)
        ref = np.fft.fft(x.astype(np.float64), axis=-1)
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-6      # contract: 1e-4
        # bypass chain: window fused into the FFT load
        out = f.process(x, want=("filtered", "frames", "iq", "mag", "phase"))
        w = g.window(x, rom)
        assert np.array_equal(out["filtered"], w)
        ref = np.fft.fft(w.astype(np.float64), axis=-1)
        got = out["iq"][..., 0] + 1j * out["iq"][..., 1]
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-6
        # frames = floor(own fp32 bins / N) exactly; within 1 LSB of the float64 oracle
        log2n = int(np.log2(n))
        assert np.array_equal(cg.quantize_pack(got.astype(np.complex128), -log2n, 0), out["frames"])
        re, im, mag = g.decode_frame(out["frames"])
        rq, iq_ = g.quantize_bins(ref, -log2n)
        assert np.abs(re - rq).max() <= 1 and np.abs(im - iq_).max() <= 1
        assert ((re != rq) | (im != iq_)).mean() < 1e-3
        # magnitude is bit-identical to the GUI's decode of the same frame; phase to 1e-6
        assert np.array_equal(mag.view(np.uint32), out["mag"].view(np.uint32))
        assert np.abs(out["phase"] - np.arctan2(im, re)).max() < 2e-6
    finally:
        f.close()


@pytest.mark.parametrize("flags", [0, _abi.FRA_K1_SPECULATE])
def test_chain_iir_then_fft_and_host_path_agree(rom, flags):
    n, c = 2048, 9
    x = g.tone_noise(range(c), n=n, seed=3)
    f = EmulFra(c, n, flags)
    try:
        f.command(bytes([0x00]))
        a = f.process(x, want=("filtered", "frames", "iq"))
        b = f.process(x, want=("filtered", "frames", "iq"), host=True)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        y, _ = cg.window_iir(x, rom, 0, g.BANK0_COEFF, B1)
        assert np.array_equal(a["filtered"], y)
        ref = np.fft.fft(y.astype(np.float64), axis=-1)
        got = a["iq"][..., 0] + 1j * a["iq"][..., 1]
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-6
    finally:
        f.close()


def test_scale_and_rounding_modes(rom):
    n, c = 1024, 2
    x = adversarial(np.random.default_rng(5), c, n)
    w = g.window(x, rom)
    ref = np.fft.fft(w.astype(np.float64), axis=-1)
    for flags, rounding in ((0, 0), (_abi.FRA_ROUND_NEAREST, 1)):
        f = EmulFra(c, n, flags)
        try:
            for ls in (-10, -6, 0):                  # 1/N, overflowing, heavily overflowing: saturation
                out = f.process(x, log2_scale=ls, want=("frames", "iq"))
                got = (out["iq"][..., 0] + 1j * out["iq"][..., 1]).astype(np.complex128)
                assert np.array_equal(cg.quantize_pack(got, ls, rounding), out["frames"]), (flags, ls)
        finally:
            f.close()


def test_command_protocol_through_abi_matches_oracle_decoder():
    rng = np.random.default_rng(11)
    interesting = [0x00, 0xA1, 0xB1, 0xFF, 0x55, 0xA5, 0xEF, 0xFE, 0xF1, 0x42]
    L = lib()
    for trial in range(20):
        f = EmulFra(2, 1024)
        dec = g.CommandDecoder()
        try:
            stream = bytes(int(rng.choice(interesting)) if rng.random() < 0.7 else int(rng.integers(0, 256))
                           for _ in range(int(rng.integers(1, 80))))
            pos = 0
            while pos < len(stream):                 # arbitrary fragmentation of the byte stream
                k = int(rng.integers(1, 9))
                rc = f.command(stream[pos:pos + k])
                dec.feed(stream[pos:pos + k])
                assert rc == (_abi.FRA_ERR_BUSY if dec.busy else 0)
                pos += k
            mode, tr = C.c_uint8(), C.c_uint8()
            bank = (C.c_int8 * 12)()
            L.fra_get_mode(f.h, C.byref(mode)); L.fra_get_transport(f.h, C.byref(tr)); L.fra_get_bank(f.h, 1, bank)
            assert mode.value == dec.mode and tr.value == dec.transport
            assert list(bank) == [int(v) for v in dec.bank1]
            cnt = [C.c_uint64() for _ in range(4)]
            L.fra_get_counters(f.h, *[C.byref(v) for v in cnt])
            assert cnt[0].value == dec.events.count("start") and cnt[1].value == dec.events.count("request")
            assert cnt[2].value == dec.events.count("reset") and cnt[3].value == dec.events.count("load")
        finally:
            f.close()


def test_reset_clears_history_and_bank1(rom):
    n, c = 1024, 3
    x = adversarial(np.random.default_rng(2), c, n)
    f = EmulFra(c, n)
    try:
        f.command(bytes([0xF1]) + B1.tobytes() + bytes([0xA1]))
        f.process(x, want=("filtered",))
        assert f.get_state().any()
        f.command(bytes([0xFF]))
        assert not f.get_state().any()
        out = f.process(x, continuous=True, want=("filtered",))      # mode is bypass again after reset
        assert np.array_equal(out["filtered"], g.window(x, rom))
        f.command(bytes([0xA1]))                                     # bank 1 is all zero after reset -> output 0
        out = f.process(x, want=("filtered",))
        assert not out["filtered"].any()
    finally:
        f.close()


def test_stream_exact_and_chunked_modes(rom):
    rng = np.random.default_rng(9)
    x = g.tone_noise([5], n=1 << 16, seed=2)[0]
    yr, st = cg.window_iir(x[None], rom, 0x00, g.BANK0_COEFF, B1)
    f = EmulFra(1, 16384)
    try:
        f.command(bytes([0x00]))
        y, stats = f.iir_stream(x[:4096], exact=True)                # six-lane systolic chain: bit-exact
        assert stats["exact"] == 1 and np.array_equal(y, yr[0, :4096])
        y, stats = f.iir_stream(x, exact=False)                       # chunked scan: dead-band error only
        assert stats["exact"] == 0 and stats["n_chunks"] == 16 and stats["chunk"] == 4096
        assert 0 < stats["max_state_dev"] <= 32           # block-scan prediction: inside the dead band, not exact
        assert np.abs(y.astype(int) - yr[0].astype(int)).max() <= 64
        first = stats["chunk"]
        assert np.array_equal(y[:first], yr[0, :first])               # chunk 0 starts from the true state
        f.command(bytes([0xB1]))
        y, stats = f.iir_stream(x, exact=False)                       # bypass: window only, exact
        assert np.array_equal(y, g.window(x[None], rom)[0]) and stats["n_mismatch"] == 0
    finally:
        f.close()


@pytest.mark.parametrize("flags", [_abi.FRA_K1_FORCE_LANE, _abi.FRA_K1_FORCE_DUO])
def test_all_biased_step_at_the_edge_of_its_condition(flags, rom):
    """biquad_step_biased (every operand carries the PRMT bias, no FADD): coefficient sets on the
    edge of the eligibility condition, random history, full-range input; and the same through
    the general step (FRA_K1_NO_BIASED) for the same bits."""
    rng = np.random.default_rng(77)
    c, n = 7, 1024
    for trial in range(3):
        sec = biased_edge_sections(rng)
        if trial == 2:
            sec[:, 1] = 0
            sec[:, 2] = np.clip(sec[:, 2], -60, 60)
            sec[:, 4] = np.clip(sec[:, 4], -30, 30)
            sec[:, 0] = -sec[:, 2]
        st0 = rng.integers(-32768, 32768, (c, 6, 4)).astype(np.int16)
        x = adversarial(rng, c, n)
        y, st = cg.window_iir_sections(x, rom, sec, st0)
        for fl in (flags, flags | _abi.FRA_K1_NO_BIASED):
            f = EmulFra(c, n, fl)
            try:
                f.load_sections(sec)
                f.command(bytes([0xA1]))
                f.set_state(st0)
                out = f.process(x, continuous=True, want=("filtered",))
                assert np.array_equal(out["filtered"], y), (trial, fl)
                assert np.array_equal(f.get_state(), st)
            finally:
                f.close()


@pytest.mark.parametrize("n", [1024, 2048, 16384])
def test_fixed_point_fft_mode_bit_exact_vs_integer_oracle(n, rom):
    """FRA_FFT_FIXED16: the 16-bit scaled, truncating radix-2^2 pipeline (what xfft_0.xci configures;
    parity with the proprietary core unpinned) equals oracle/fixed_fft.py bit for bit, and stays
    within a few LSB of the float64 FFT / N."""
    from oracle.fixed_fft import fixed_fft
    rng = np.random.default_rng(n)
    c = 3
    x = g.tone_noise(range(c), n=n, seed=3)
    x[2] = rng.integers(-32768, 32768, n)
    f = EmulFra(c, n, _abi.FRA_FFT_FIXED16)
    try:
        f.command(bytes([0x00]))
        out = f.process(x, want=("filtered", "frames", "iq", "mag"))
        y, _ = cg.window_iir(x, rom, 0x00, g.BANK0_COEFF, B1)
        assert np.array_equal(out["filtered"], y)
        re, im = fixed_fft(y)
        gre, gim, gmag = g.decode_frame(out["frames"])
        assert np.array_equal(gre, re) and np.array_equal(gim, im)
        assert np.array_equal(out["mag"].view(np.uint32), gmag.astype(np.float32).view(np.uint32))
        assert np.array_equal(out["iq"][..., 0], re.astype(np.float32) * n)
        ref = np.fft.fft(y.astype(np.float64), axis=-1) / n
        assert np.abs((re + 1j * im) - ref).max() < 8.0
        # bypass: the window fused into the load
        f.command(bytes([0xB1]))
        out = f.process(x, want=("frames",))
        re, im = fixed_fft(g.window(x, rom))
        gre, gim, _ = g.decode_frame(out["frames"])
        assert np.array_equal(gre, re) and np.array_equal(gim, im)
        # only the core's 1/N schedule exists in this mode
        import ctypes as C2
        o = _abi.FraOutputs(None, out["frames"].ctypes.data, None, None, None)
        assert f.L.fra_process(f.h, x.ctypes.data, 0, -3, C2.byref(o), None) == _abi.FRA_ERR_INVALID
    finally:
        f.close()
    h = C.c_void_p()
    assert lib().fra_create(C.byref(h), 0, 1, 65536, _abi.FRA_FFT_FIXED16) == _abi.FRA_ERR_UNSUPPORTED


def test_stream_scan_with_short_chunks_and_aggregate_levels(rom):
    """2^20 samples: 512-sample chunks, 2048 lanes in 64 warps, the warps' aggregates chained by the
    parallel scan.  Still inside the dead band of the truncating sections, chunk 0 exact."""
    x = g.tone_noise([9], n=1 << 20, seed=4)[0]
    yr, st = cg.window_iir(x[None], rom, 0x00, g.BANK0_COEFF, B1)
    f = EmulFra(1, 16384)
    try:
        f.command(bytes([0x00]))
        y, stats = f.iir_stream(x, exact=False)
        assert stats["exact"] == 0 and stats["chunk"] == 512 and stats["n_chunks"] == 2048
        assert 0 < stats["max_state_dev"] <= 32
        assert np.abs(y.astype(int) - yr[0].astype(int)).max() <= 64
        assert np.array_equal(y[:512], yr[0, :512])
        assert np.abs(f.get_state()[0].astype(int) - st[0].astype(int)).max() <= 32
    finally:
        f.close()


@pytest.mark.parametrize("n", [1024, 16384])
def test_half_spectrum_host_transfer_rebuilds_identical_frames(n, rom):
    """FRA_HOST_HALF_SPECTRUM: bins 0..N/2 + one bit per bin cross the link, the host completes the mirror;
    byte-identical to the full transfer (silence, full-scale DC and full-range noise among the channels)."""
    c = 5
    rng = np.random.default_rng(n)
    x = rng.integers(-32768, 32768, (c, n)).astype(np.int16)
    x[1] = 0
    x[2] = -32768
    x[3] = g.tone_noise([3], n=n, seed=1)[0]
    a, b = EmulFra(c, n), EmulFra(c, n, _abi.FRA_HOST_HALF_SPECTRUM)
    try:
        for mode in (0x00, 0xB1):
            a.command(bytes([mode])); b.command(bytes([mode]))
            fa = a.process(x, want=("frames", "mag"), host=True)
            fb = b.process(x, want=("frames", "mag"), host=True)
            assert np.array_equal(fa["frames"], fb["frames"]) and np.array_equal(fa["mag"], fb["mag"])
        # a scale that can saturate falls back to the full transfer: still identical
        fa = a.process(x, log2_scale=-6, want=("frames",), host=True)
        fb = b.process(x, log2_scale=-6, want=("frames",), host=True)
        assert np.array_equal(fa["frames"], fb["frames"])
    finally:
        a.close(); b.close()
