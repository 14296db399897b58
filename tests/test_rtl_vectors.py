"""Vectors obtained from the reference's VHDL TEXT (tests/golden/rtl_vectors.npz, produced in the
build container by tests/golden/make_golden.py --rtl: the window, the biquad and both cascades
parsed from /root/reference and executed with numeric_std semantics by tests/golden/vhdl_eval.py).

They pin (i) the oracle - oracle/golden.py and oracle/golden.c must reproduce them bit for bit -
and (ii), under -m gpu, the CUDA path itself through the C ABI.  The reference holds no vectors of
its own for this path and cannot be simulated here (no ghdl / nvc / xsim): these are as close as the
image gets to running it."""
import os

import numpy as np
import pytest

from oracle import cgolden as cg
from oracle import golden as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHAINS = ["chain_bank0_tone", "chain_bank0_full_2frames", "chain_bank1_full", "chain_bank1_extreme_gap"]


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(ROOT, "tests", "golden", "rtl_vectors.npz"))


def test_biquad_entity_vectors(vec):
    """filter_iir_cust.vhd alone: ALPHA takes registers 0..4, BETA 6..10; i_valid gaps clear the history."""
    for ft, off in (("alpha", 0), ("beta", 6)):
        for case in range(3):
            key = f"biquad_{ft}_{case}"
            x, valid, coef, y = vec[key + "_x"], vec[key + "_valid"], vec[key + "_coef"], vec[key + "_y"]
            exp = np.zeros_like(y)
            i = 0
            while i < len(x):
                if not valid[i]:
                    i += 1
                    continue
                j = i
                while j < len(x) and valid[j]:
                    j += 1
                seg, _ = g.biquad(x[i:j][None], coef[off:off + 5])
                exp[i:j] = seg[0]
                i = j
            assert np.array_equal(y[valid], exp[valid]), key
            assert not y[~valid].any()                    # vs(0) of an idle stage is the sum of zero products


def test_window_vectors_every_rom_address(vec, rom):
    x, y = vec["window_x"], vec["window_y"]
    assert len(x) == 16384
    assert np.array_equal(g.window(x[None], rom)[0], y)
    # the resize quirk is in there: x = c = -32768 -> 0
    hit = (x == -32768) & (rom == -32768)
    assert hit.sum() >= 40 and not y[hit].any()


@pytest.mark.parametrize("key", CHAINS)
def test_chain_vectors_numpy_and_c_oracle(vec, rom, key):
    x, win, y, coef = vec[key + "_x"], vec[key + "_win"], vec[key + "_y"], vec[key + "_coef"]
    gap = bool(vec[key + "_gap"][0])
    mode = 0x00 if "bank0" in key else 0xA1
    st_np = st_c = None
    for f in range(x.shape[0]):
        assert np.array_equal(g.window(x[f][None], rom)[0], win[f]), (key, f)
        if gap:
            st_np = st_c = None                           # the burst after a gap starts from zero history
        c12 = g.BANK0_COEFF if mode == 0x00 else coef
        y_np, st_np = g.iir12(win[f][None], c12, st_np)
        assert np.array_equal(y_np[0], y[f]), (key, f, "numpy oracle")
        y_c, st_c = cg.window_iir(x[f][None], rom, mode, g.BANK0_COEFF, coef, st_c)
        assert np.array_equal(y_c[0], y[f]), (key, f, "C oracle")


def test_bank0_constants_come_from_the_package(vec):
    """the bank-0 chain was elaborated with filter_pkg.vhd's constants, not with ours"""
    assert not vec["chain_bank0_tone_coef"].any()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["auto", "lane", "duo", "split"])
@pytest.mark.parametrize("key", CHAINS)
def test_cuda_path_reproduces_the_vhdl_vectors(vec, key, variant):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; there is no CPU fallback")
    from fpga_real_time_fft_analyzer_b200 import FraContext, _abi
    flags = {"auto": 0, "lane": _abi.FRA_K1_FORCE_LANE, "duo": _abi.FRA_K1_FORCE_DUO,
             "split": _abi.FRA_K1_FORCE_SPLIT}[variant]
    x, win, y, coef = vec[key + "_x"], vec[key + "_win"], vec[key + "_y"], vec[key + "_coef"]
    gap = bool(vec[key + "_gap"][0])
    c = 33                                               # the vector in channel 0, 17 and 32; noise elsewhere
    rng = np.random.default_rng(1)
    with FraContext(c, 16384, flags=flags) as ctx:
        if "bank0" in key:
            ctx.command(0x00)
        else:
            ctx.command(bytes([0xF1]) + coef.tobytes() + bytes([0xA1]))      # the 0xF1 byte protocol
        for f in range(x.shape[0]):
            batch = rng.integers(-32768, 32768, (c, 16384)).astype(np.int16)
            batch[[0, 17, 32]] = x[f]
            out = ctx.process(torch.from_numpy(batch).cuda(), continuous=(f > 0 and not gap), want=("filtered",))
            got = out["filtered"].cpu().numpy()
            for ch in (0, 17, 32):
                assert np.array_equal(got[ch], y[f]), (key, variant, f, ch)
        # the window alone (bypass, 0xB1): the FFT input stream is the windowed frame
        ctx.command(0xB1)
        batch = np.repeat(x[-1][None], c, axis=0)
        got = ctx.process(torch.from_numpy(batch).cuda(), want=("filtered",))["filtered"].cpu().numpy()
        assert np.array_equal(got[5], win[-1])


# ------------------------------------------------------------ the window exactly as the text clocks it
@pytest.fixture(scope="module")
def raw():
    return np.load(os.path.join(ROOT, "tests", "golden", "rtl_window_raw.npz"))


def test_window_register_skew_oracle(raw, rom):
    """hann8192.vhd evaluated with NO stimulus compensation (sample n presented at strobe n, two frames from
    power-up): output n = W(x[n-1], ROM[n-2]), the first two outputs 0."""
    x, y = raw["x"], raw["y"]
    assert np.array_equal(g.window_rtl_skew(x[0][None], rom)[0], y[0])
    assert np.array_equal(g.window_rtl_skew(x[1][None], rom, prev=x[0][-1:])[0], y[1])
    assert not y[0][:2].any() and not np.array_equal(y[0], g.window(x[0][None], rom)[0])


def test_window_register_skew_mode_emulated(raw, rom):
    from fpga_real_time_fft_analyzer_b200 import _abi
    from tests.emul.emul_lib import EmulFra
    x, y = raw["x"], raw["y"]
    f = EmulFra(2, 16384, _abi.FRA_WINDOW_RTL_SKEW)
    try:
        for fr in range(2):                                   # bypass: the FFT input stream is the window's output
            out = f.process(np.stack([x[fr], x[fr]]), continuous=fr > 0, want=("filtered",))["filtered"]
            assert np.array_equal(out[1], y[fr]), fr
        f.command(bytes([0x00]))                              # ... and the filter runs on that stream
        out = f.process(np.stack([x[0], x[0]]), want=("filtered",))["filtered"]
        want, _ = g.iir12(y[0][None], g.BANK0_COEFF)
        assert np.array_equal(out[0], want[0])
    finally:
        f.close()
    # the lane kernel's fast window path must know where the rotated table holds -32768
    f = EmulFra(2, 16384, _abi.FRA_WINDOW_RTL_SKEW | _abi.FRA_K1_FORCE_LANE)
    try:
        f.command(bytes([0x00]))
        out = f.process(np.stack([x[0], x[0]]), want=("filtered",))["filtered"]
        assert np.array_equal(out[1], want[0])
    finally:
        f.close()


@pytest.mark.gpu
def test_window_register_skew_mode(raw, rom):
    """FRA_WINDOW_RTL_SKEW on the GPU against the raw VHDL-evaluated window, two continuous frames, then through
    the cascade (every K1 kernel the dispatcher can pick at this size) and the FFT."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; there is no CPU fallback")
    from fpga_real_time_fft_analyzer_b200 import FraContext, FraError, _abi
    x, y = raw["x"], raw["y"]
    c = 40
    for flags in (0, _abi.FRA_K1_FORCE_LANE, _abi.FRA_K1_FORCE_DUO):
        with FraContext(c, 16384, flags=flags | _abi.FRA_WINDOW_RTL_SKEW) as ctx:
            for fr in range(2):
                batch = torch.from_numpy(np.repeat(x[fr][None], c, axis=0)).cuda()
                out = ctx.process(batch, continuous=fr > 0, want=("filtered", "iq"))
                assert np.array_equal(out["filtered"][c - 1].cpu().numpy(), y[fr]), (flags, fr)
                iq = out["iq"][0].cpu().numpy()
                ref = np.fft.fft(y[fr].astype(np.float64))
                assert np.linalg.norm((iq[..., 0] + 1j * iq[..., 1]) - ref) / np.linalg.norm(ref) < 1e-4
            ctx.command(0xFF)                                 # reset: power-up state again
            ctx.command(0x00)
            st = None
            for fr in range(2):
                batch = torch.from_numpy(np.repeat(x[fr][None], c, axis=0)).cuda()
                got = ctx.process(batch, continuous=fr > 0, want=("filtered",))["filtered"][7].cpu().numpy()
                want, st = g.iir12(y[fr][None], g.BANK0_COEFF, st)
                assert np.array_equal(got, want[0]), (flags, fr)
            with pytest.raises(FraError):
                ctx.iir_stream(torch.zeros(4096, dtype=torch.int16, device="cuda"))
    with pytest.raises(FraError):
        FraContext(4, 16384, flags=_abi.FRA_WINDOW_RTL_SKEW | _abi.FRA_PIPELINE)
