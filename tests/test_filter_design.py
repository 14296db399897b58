"""filter design -> int8 bytes with an explicit GUI / RTL-compatible switch,
and six independent sections as a superset of the 12-byte bank.  CPU tests: the GUI convention is
pinned by the reference quantiser's own outputs (tests/golden/gui_vectors.json, produced by
tests/golden/make_golden.py from scripts/fft_analyzer_gui.py); the RTL-compatible convention is
checked through the bit-exact oracle."""
import numpy as np
import pytest

from fpga_real_time_fft_analyzer_b200 import filter_design as fd
from oracle import golden as g

signal = pytest.importorskip("scipy.signal")


def design(kind, order, ftype, cutoff, cutoff2, fs):
    wn = cutoff / (fs / 2) if cutoff2 is None else [cutoff / (fs / 2), cutoff2 / (fs / 2)]
    assert kind == "butter"
    return signal.butter(order, wn, btype=ftype, output="sos")


def test_gui_convention_matches_reference_quantiser(gui_vectors):
    checked = 0
    for case in gui_vectors["quantize"]:
        if case["kind"] != "butter":
            continue                      # the GUI's ripple parameters for the other kinds are not in the fixture
        sos = design(case["kind"], case["order"], case["filter_type"], case["cutoff"], case["cutoff2"], case["fs"])
        got = fd.quantize_sections(sos, rtl_compatible=False)
        assert got.tolist() == case["sections"], case
        checked += 1
    assert checked >= 4


def tone_gain(sections6, freq, n=4096, amp=8000.0):
    """Steady-state gain of the BIT-EXACT cascade (oracle) for a tone at `freq` (cycles/sample)."""
    t = np.arange(n)
    x = np.round(amp * np.sin(2 * np.pi * freq * t)).astype(np.int16)[None, :]
    y, _ = g.iir_sections(x, sections6)
    return float(np.std(y[0, n // 2:].astype(np.float64)) / np.std(x[0, n // 2:].astype(np.float64)))


def test_rtl_compatible_design_realises_the_filter():
    # band-pass around fs/4: poles near +-j, |a1| < 1 - the family a >> 7 datapath can hold
    sos = signal.butter(3, [0.42, 0.58], btype="bandpass", output="sos")            # 3 sections
    q = fd.quantize_sections(sos, rtl_compatible=True)
    assert q.dtype == np.int8 and q.shape == (3, 6) and np.all(q[:, 5] == 0)
    assert fd.unrealizable(sos) == []
    stages = fd.expand_to_stages(q)
    # float response of what was realised vs the design (normalised): same shape within quantisation
    realised = fd.realized_sos(q)
    want = fd.normalize_gain(sos)
    hr = np.prod(fd.section_response(realised, 512), axis=0)
    hw = np.prod(fd.section_response(want, 512), axis=0)
    assert np.max(np.abs(hr - hw)) < 0.08
    # and the bit-exact integer cascade passes the pass-band and rejects the stop-band
    pad = (127.0 / 128.0) ** 3                                                       # three pass-through stages
    assert tone_gain(stages, 0.25) == pytest.approx(hw[256] * pad, rel=0.08)
    assert tone_gain(stages, 0.05) < 0.02
    assert tone_gain(stages, 0.45) < 0.02
    # the GUI convention on the same design does NOT realise it: that is the reason for the switch
    gui = fd.quantize_sections(sos, rtl_compatible=False)
    assert abs(tone_gain(fd.expand_to_stages(gui), 0.25) - hw[256] * pad) > 0.2


def test_unrealisable_design_is_reported_not_saturated():
    sos = signal.butter(4, 0.2, btype="lowpass", output="sos")                      # a1 ~ -1.1 ... -1.3
    bad = fd.unrealizable(sos)
    assert bad and all(name == "A1" for _, name, _ in bad)
    with pytest.raises(ValueError, match="not realisable"):
        fd.quantize_sections(sos, rtl_compatible=True)
    q = fd.quantize_sections(sos, rtl_compatible=True, strict=False)
    assert q.min() >= -128 and q.max() <= 127 and (q[:, 4] == -128).any()


def test_bank12_or_sections():
    a = np.array([-14, 0, 14, 107, 21, 127], dtype=np.int8)
    b = np.array([-15, 0, 15, 107, -21, 127], dtype=np.int8)
    assert fd.sections_to_bank12(np.stack([a, b, a, b, a, b])).tolist() == g.BANK0_COEFF.tolist()
    assert fd.sections_to_bank12(np.stack([a, b, a, b, b, a])) is None
    assert fd.sections_to_bank12(np.stack([a, b])) is None                            # padded stages differ from a, b

    class Ctx:                                                                       # records what upload() does
        def __init__(self):
            self.log = []

        def command(self, c):
            self.log.append(("command", bytes([c]) if isinstance(c, int) else bytes(c)))

        def load_sections(self, s):
            self.log.append(("sections", np.asarray(s).tolist()))

    c = Ctx()
    assert fd.upload(c, np.stack([a, b, a, b, a, b])) == "bank12"
    assert c.log == [("command", bytes([0xF1]) + g.BANK0_COEFF.tobytes()), ("command", bytes([0xA1]))]
    c = Ctx()
    assert fd.upload(c, np.stack([a, b])) == "sections"
    assert c.log[0][0] == "sections" and c.log[0][1][2] == [0, 0, 127, 0, 0, 0] and c.log[1] == ("command", bytes([0xA1]))


def test_oracle_sections_equal_alternated_bank12():
    rng = np.random.default_rng(3)
    x = rng.integers(-32768, 32768, (2, 300)).astype(np.int16)
    b = rng.integers(-128, 128, 12).astype(np.int8)
    y12, s12 = g.iir12(x, b)
    y36, s36 = g.iir_sections(x, np.stack([b[:6], b[6:]] * 3))
    assert np.array_equal(y12, y36) and np.array_equal(s12, s36)
