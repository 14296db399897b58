"""The oracle against everything that pins it: the reference-derived fixtures
(tests/golden, made by make_golden.py from /root/reference), a second cycle-accurate
restatement of the VHDL, and the known-answer values of SURVEY Appendix A."""
import hashlib
import os

import numpy as np
import pytest

from oracle import cgolden as cg
from oracle import golden as g
from oracle.rtl_cycle_model import BiquadRTL, Cascade12RTL, WindowRTL

B1 = np.array([32, 10, -33, 119, 35, 0, 52, -16, 11, 84, -10, 0], dtype=np.int8)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype("<i2").tobytes()).hexdigest()


# ------------------------------------------------------------------ ROM (a1)
def test_rom_matches_reference_fixture(rom, reference_facts):
    assert sha(rom) == reference_facts["hann_rom_sha256"] == \
        "c02f53d7f7e3fc7c6eb8787a2e02be9a8516654e635db83ff11cf36e0cbea5a1"
    assert np.array_equal(g.hann_rom(), rom)
    assert np.array_equal(cg.hann_rom(), rom)
    assert rom[0] == -32768 and rom[8177] == 32767 and rom[8178] == -32768 and rom[16383] == -32768
    assert int((rom[8178:8206] == -32768).sum()) == 28          # the wrapped centre entries


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_rom_against_reference_tree(rom):
    import re
    vals = {}
    with open("/root/reference/SDR_v2.srcs/sources_1/new/hann.vhd") as f:
        for line in f:
            m = re.search(r"(\d+)\s*=>\s*to_signed\((-?\d+),16\)", line)
            if m:
                vals[int(m.group(1))] = int(m.group(2))
    assert np.array_equal(np.array([vals[i] for i in range(16384)], dtype=np.int16), rom)


def test_bank0_matches_reference_constants(reference_facts):
    b = reference_facts["bank0"]
    want = [b["ALPHA_B0"], b["ALPHA_B1"], b["ALPHA_B2"], b["ALPHA_A0"], b["ALPHA_A1"], b["ALPHA_A2"],
            b["BETA_B0"], b["BETA_B1"], b["BETA_B2"], b["BETA_A0"], b["BETA_A1"], b["BETA_A2"]]
    assert list(g.BANK0_COEFF) == want


def test_constants_match_gui(reference_facts):
    gui = reference_facts["gui"]
    assert gui["FRAME_SIZE_BYTES"] == g.FRAME_SIZE_BYTES == 4 * g.FFT_SIZE
    assert (gui["UART_REQUEST_CMD"], gui["FPGA_RESET_CMD"], gui["ETHERNET_MODE_CMD"], gui["UART_MODE_CMD"],
            gui["START_COMMAND"], gui["FILTER_UPDATE_CMD"], gui["FILTER_DEFAULT_CMD"], gui["FILTER_CUSTOM_CMD"],
            gui["FILTER_NONE_CMD"]) == (g.CMD_UART_REQUEST, g.CMD_RESET, g.CMD_ETHERNET_MODE, g.CMD_UART_MODE,
                                        g.CMD_START, g.CMD_FILTER_UPDATE, g.MODE_BANK0, g.MODE_BANK1, g.MODE_BYPASS)
    x = reference_facts["xfft"]
    assert x["transform_length"] == "16384" and x["scaling_options"] == "scaled" and x["rounding_modes"] == "truncation"


# --------------------------------------------------------------- window (a2)
def test_window_known_answers(rom):
    table = {0: [-2047, 2048, -1000, 0], 15: [-2047, 2048, -1000, 32767], 4096: [0, 0, 0, -3],
             6000: [1366, -1366, 667, -21860], 8100: [2046, -2047, 999, -32748], 8177: [2047, -2048, 1000, -32767],
             8178: [-2047, 2048, -1000, 0], 12288: [-1, 1, 0, 9], 16383: [-2047, 2048, -1000, 0]}
    for n, want in table.items():
        got = [int(g.window(np.array([v], np.int16), rom, n)[0]) for v in (2047, -2048, 1000, -32768)]
        assert got == want, (n, got)


def test_window_vs_rtl_round_resize_exhaustive_column(rom):
    xs = np.arange(-32768, 32768, dtype=np.int64)
    for c in (-32768, -32767, -1, 0, 1, 3, 21860, 32767):
        want = np.array([WindowRTL.round_resize(int(x) * c) for x in xs[::17]])
        got = g.window(xs[::17].astype(np.int16)[None, :], np.array([c] * 1, np.int16), 0)
        assert np.array_equal(got[0], want)


def test_window_rtl_skew_alignment(rom):
    """The registered RTL lags the coefficient by one sample: out[n] = f(x[n-1], rom[n-2])."""
    rng = np.random.default_rng(0)
    x = rng.integers(-2048, 2048, 64)
    w = WindowRTL(rom)
    outs = []
    for v in x:
        w.clock(1, int(v))
        outs.append(w.sample_out)
    for n in range(3, 64):
        assert outs[n] == WindowRTL.round_resize(int(x[n - 1]) * int(rom[n - 2]))


def test_window_c_equals_numpy(rom):
    rng = np.random.default_rng(1)
    x = rng.integers(-32768, 32768, (3, 40000)).astype(np.int16)
    x[0, :30] = -32768
    y, _ = cg.window_iir(x, rom, 0xB1, g.BANK0_COEFF, B1)
    assert np.array_equal(y, g.window(x, rom))


# --------------------------------------------------------------- biquad (a3)
def test_slice_known_answers():
    assert [int(g.slice_T(v, c)) for v, c in ((-32768, -128), (-32768, 127), (32767, -128), (-1, 1), (1, -1))] == \
        [-32768, -32512, -32767, -1, -1]


@pytest.mark.parametrize("seed", range(6))
def test_biquad_closed_form_equals_cycle_accurate(seed):
    rng = np.random.default_rng(seed)
    coef = [int(v) for v in rng.integers(-128, 128, 5)]
    if seed == 0:
        coef = [-128, 127, -128, 127, -128]
    x = rng.integers(-32768, 32768, 300).astype(np.int16)
    x[:4] = [-32768, 32767, -32768, -32768]
    st = BiquadRTL(*coef)
    outs = []
    for v in x:
        st.clock(1, int(v))
        outs.append(st.comb()[0])
    y, _ = g.biquad(x[None, :], coef)
    assert np.array_equal(y[0], np.array(outs))


@pytest.mark.parametrize("seed", range(4))
def test_cascade_equals_cycle_accurate_and_burst_clears_state(seed):
    rng = np.random.default_rng(100 + seed)
    c12 = rng.integers(-128, 128, 12).astype(np.int8)
    x = rng.integers(-32768, 32768, 200).astype(np.int16)
    rtl = Cascade12RTL(c12)
    # burst, gap of 3 idle clocks, burst: history must restart from zero
    valid = [1] * 100 + [0] * 3 + [1] * 100
    data = list(x[:100]) + [0] * 3 + list(x[100:])
    out = np.array(rtl.run(valid, data, extra_cycles=12))
    y1, _ = g.iir12(x[None, :100], c12)
    y2, _ = g.iir12(x[None, 100:], c12)
    assert np.array_equal(out, np.concatenate([y1[0], y2[0]]))
    # continuous: no gap == state carried across two calls
    rtl = Cascade12RTL(c12)
    out = np.array(rtl.run([1] * 200, list(x), extra_cycles=12))
    ya, sa = g.iir12(x[None, :77], c12)
    yb, _ = g.iir12(x[None, 77:], c12, sa)
    assert np.array_equal(out, np.concatenate([ya[0], yb[0]]))


def test_bank0_known_answers():
    imp = np.zeros(16, np.int16)
    imp[0] = 32767
    assert list(g.iir12(imp[None], g.BANK0_COEFF)[0][0]) == [0, 0, -1, 0, 3, 0, -14, -1, 38, 1, -85, 0, 163, 0, -276, -2]
    assert list(g.iir12(np.full((1, 16), 1000, np.int16), g.BANK0_COEFF)[0][0]) == \
        [0, 0, -1, 0, 1, 1, -1, -1, 1, 2, -1, -3, 2, 4, -4, -5]
    x = g.lcg_stimulus(4096)
    assert list(x[:8]) == [18046, -336, 484, -5221, -13345, 31604, 25270, 31662]
    y, st = cg.iir12(x[None], g.BANK0_COEFF)
    assert sha(y[0]) == "cd9ccbcd411bd2044119e3ed9b3573274bb55f07b805aa91572ac685843afbaf"
    assert list(st[0, 0]) == [-29328, -12607, -5009, -7038] and list(st[0, 5]) == [-2487, 1057, -2853, 2821]
    y1, _ = cg.iir12(x[None], B1)
    assert sha(y1[0]) == "f2e20af4701ebd64d406ce2485a2c8dd5737da69808c1a16e410f0be3bffe8bc"
    ya, sa = cg.iir12(x[None, :2048], g.BANK0_COEFF)
    yb, _ = cg.iir12(x[None, 2048:], B1, sa)                      # mid-stream reload keeps the history (D11)
    assert list(yb[0, :4]) == [-302, 957, 1190, 246]
    assert sha(np.concatenate([ya[0], yb[0]])) == "afd3b6ce1b5fc6d0493ffc7e1c1b9001718dd219aabc95aa3a54d2f738a33c39"


def test_c_golden_equals_numpy_golden_fuzz():
    for seed in range(1, 9):
        rng = np.random.default_rng(seed)
        c12 = rng.integers(-128, 128, 12).astype(np.int8)
        x = rng.integers(-32768, 32768, (5, 700)).astype(np.int16)
        st0 = rng.integers(-32768, 32768, (5, 6, 4)).astype(np.int16)
        yn, sn = g.iir12(x, c12, st0)
        yc, sc = cg.iir12(x, c12, st0)
        assert np.array_equal(yn, yc) and np.array_equal(sn, sc)


# ---------------------------------------------------------- commands (a7)
def test_command_decoder_protocol():
    d = g.CommandDecoder()
    assert d.mode == 0xB1 and not d.bank1.any()
    d.feed(bytes([0xA1]))
    assert d.mode == 0xA1
    # during an upload every byte is data, including bytes that look like commands
    payload = bytes([0xFF, 0x00, 0xA1, 0xB1, 0x55, 0xF1, 1, 2, 3, 4, 5, 6])
    d.feed(bytes([0xF1]) + payload[:5])
    assert d.busy and d.mode == 0xA1
    d.feed(payload[5:] + bytes([0x00]))
    assert not d.busy and d.mode == 0x00
    assert d.bank1.tobytes() == payload
    d.feed(bytes([0x42, 0x55, 0xA5, 0xFE]))
    assert d.events[-2:] == ["start", "request"] and d.transport == 0xFE
    d.feed(bytes([0xFF]))
    assert d.mode == 0xB1 and not d.bank1.any() and d.transport == 0xEF


# ------------------------------------------------------- framing (a10, a11)
def test_pack_and_decode_match_gui_functions(gui_vectors):
    v = gui_vectors["decode"]
    rng = np.random.default_rng(v["seed"])
    frame = rng.integers(0, 256, size=65536, dtype=np.uint8)
    re, im, mag = g.decode_frame(frame)
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert str(mag.dtype) == v["mag_dtype"] == "float32"
    assert h(mag) == v["mag_sha256"] and h(re) == v["re_sha256"] and h(im) == v["im_sha256"]
    assert [float(x) for x in mag[:8]] == v["mag_head"]
    assert np.array_equal(g.pack_frame(re.astype(np.int16), im.astype(np.int16)), frame)     # round trip


def test_quantize_pack_c_equals_numpy():
    rng = np.random.default_rng(3)
    bins = (rng.standard_normal((2, 512)) + 1j * rng.standard_normal((2, 512))) * 3e7
    for rounding in (0, 1):
        re, im = g.quantize_bins(bins, -10, "floor" if rounding == 0 else "nearest")
        assert np.array_equal(g.pack_frame(re, im), cg.quantize_pack(bins, -10, rounding))


def test_float_sos_is_linearisation_of_integer_filter(rom):
    """The float CPU baseline (scipy sosfilt) tracks the integer cascade to a few LSB."""
    from scipy.signal import sosfilt
    x = g.tone_noise([3], n=4096, seed=1)
    w = g.window(x, rom)
    yi, _ = g.iir12(w, g.BANK0_COEFF)
    yf = sosfilt(g.float_sos(g.BANK0_COEFF), w.astype(np.float64), axis=-1)
    assert np.abs(yi - yf).max() < 40
