"""CPU golden model of the reference receive chain.  TEST INFRASTRUCTURE ONLY.

This module restates, in numpy integer arithmetic, what the reference's VHDL
computes on the hot path  window -> IIR12 -> FFT -> bin framing.  It is the
checker for the CUDA kernels.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it;
the product (``fpga_real_time_fft_analyzer_b200``) never does.

Pinning status
--------------
* window ROM: PINNED - identical to the reference's checked-in
  ``SDR_v2.srcs/sources_1/new/hann.vhd`` and to the output of the reference's
  ``scripts/hann_coeff.py`` run in the build container (tests/golden/hann_rom.i16,
  sha256 c02f53d7...cbea5a1, made by tests/golden/make_golden.py).
* window / biquad arithmetic: the reference holds NO test vectors for this path
  (sim_1 covers only UART/RMII).  Pinned instead against (a) a second,
  cycle-accurate register-level restatement of ``filter_iir_cust.vhd``
  (oracle/rtl_cycle_model.py) and (b) the known-answer values of SURVEY.md
  Appendix A.  "parity unpinned" at the level of reference-owned vectors.
* FFT: third-party Xilinx xfft v9.1 (source absent) - "parity unpinned" at bit
  level; the contract is <= 1e-4 relative L2 against numpy float64 FFT.
* frame bytes / host decode: PINNED against the reference GUI's own
  ``decode_mag_16iq_le`` / ``decode_iq_components`` run in the build container
  (tests/golden/gui_*.npz).

Reference line citations use the path legend of SURVEY.md:
  NEW/ = SDR_v2.srcs/sources_1/new/, IMP/ = SDR_v2.srcs/sources_1/imports/new/,
  GUI = scripts/fft_analyzer_gui.py.
"""
from __future__ import annotations

import numpy as np

FFT_SIZE = 16384            # GUI:44, IP/xfft_0/xfft_0.xci:12
FRAME_SIZE_BYTES = 65536    # GUI:41-43
FS_HZ = 1_000_000.0         # GUI:45

# command bytes, GUI:28-37
CMD_UART_REQUEST = 0xA5
CMD_RESET = 0xFF
CMD_ETHERNET_MODE = 0xEF
CMD_UART_MODE = 0xFE
CMD_START = 0x55
CMD_FILTER_UPDATE = 0xF1
MODE_BANK0 = 0x00
MODE_BANK1 = 0xA1
MODE_BYPASS = 0xB1

# Fixed bank 0 in the 12-byte register order of NEW/filter_iir12_cust.vhd:83-94
# (B0,B1,B2,A0,A1,A2) x {ALPHA, BETA}; values from IMP/filter_pkg.vhd:54-68.
BANK0_COEFF = np.array([-14, 0, 14, 107, 21, 127,
                        -15, 0, 15, 107, -21, 127], dtype=np.int8)


# --------------------------------------------------------------------------- a1
def hann_rom(n: int = FFT_SIZE) -> np.ndarray:
    """Window ROM, restating scripts/hann_coeff.py:3-5.

    ``round((w-0.5)*2**16).astype(int16)``: offset-coded, and the centre entries
    whose value is +32768 wrap to -32768 in the int16 cast.
    The float->int16 cast of an out-of-range value is done through int64 so the
    wrap is defined behaviour, matching what numpy 2.x does on x86-64.
    """
    k = np.arange(n)
    w = 0.5 * (1 - np.cos(2 * np.pi * k / (n - 1)))
    q = np.round((w - 0.5) * 2 ** 16).astype(np.int64)
    return _wrap16(q).astype(np.int16)


def _wrap16(v):
    """Two's-complement wrap of an integer array to 16 bit (numeric_std slice)."""
    v = np.asarray(v, dtype=np.int64)
    return ((v + 32768) & 0xFFFF) - 32768


# --------------------------------------------------------------------------- a2
def window(x: np.ndarray, rom: np.ndarray, start: int = 0) -> np.ndarray:
    """hann_window arithmetic, NEW/hann8192.vhd:36-39, 'aligned' indexing.

    p = x * rom[n mod 16384] (32 bit);  r17 = p(31 downto 15) + p(14)  (17 bit);
    out = resize(r17, 16): numeric_std keeps the sign bit and the low 15 bits,
    so the single out-of-range value +32768 (x = rom = -32768) becomes 0.
    ``x`` is int16 [..., T]; sample t uses rom[(start + t) mod len(rom)]
    (the 14-bit address counter free-runs, hann8192.vhd:23,41).
    """
    x = np.asarray(x)
    t = x.shape[-1]
    idx = (start + np.arange(t)) % len(rom)
    p = x.astype(np.int64) * rom.astype(np.int64)[idx]
    r17 = (p >> 15) + ((p >> 14) & 1)
    sign = (r17 >> 16) & 1
    low15 = r17 & 0x7FFF
    return (low15 - (sign << 15)).astype(np.int16)


def window_rtl_skew(x: np.ndarray, rom: np.ndarray, prev=None) -> np.ndarray:
    """hann_window as WRITTEN (NEW/hann8192.vhd:36-39): coef_s, product and sample_out are
    registers of the same clocked branch, so output n is the rounding of x[n-1] * ROM[n-2]; from zero
    registers (prev is None: power-up) the first two outputs are 0.  ``prev`` = the sample in front of
    x[..., 0] when the stream continues (then frame lengths must be multiples of the ROM length for the
    free-running address to restart at 0).  x: int16 [..., T]."""
    x = np.asarray(x)
    lead = np.zeros(x.shape[:-1] + (1,), dtype=x.dtype) if prev is None else np.asarray(prev, dtype=x.dtype).reshape(x.shape[:-1] + (1,))
    delayed = np.concatenate([lead, x[..., :-1]], axis=-1)
    if prev is None:
        delayed[..., :2] = 0
    return window(delayed, np.roll(np.asarray(rom), 2))


# --------------------------------------------------------------------------- a3
def slice_T(v, c):
    """One product term of the biquad: mult(22 downto 7) of a 16x8->24-bit
    signed product, NEW/filter_iir_cust.vhd:96-108.  floor(v*c/128) wrapped to
    16 bit (wraps only for v=-32768, c=-128)."""
    p = np.asarray(v, dtype=np.int64) * np.asarray(c, dtype=np.int64)
    return _wrap16(p >> 7)


def biquad(x: np.ndarray, coef5, state: np.ndarray | None = None):
    """One filter_iir[_cust] stage under continuous i_valid.

    coef5 = (B0, B1, B2, A0, A1) in the RTL's naming (NEW/filter_iir_cust.vhd:104-108):
        y[n] = T(x[n],B2) + T(x[n-1],B1) + T(x[n-2],B0) - T(y[n-2],A0) - T(y[n-1],A1)
    all terms and the sum 16-bit two's complement (wrap, :96-100).
    x: int16 [C, T];  state: int16 [C, 4] = (x[n-1], x[n-2], y[n-1], y[n-2]),
    zeros when None (history is cleared whenever i_valid drops, :139-193).
    Returns (y int16 [C, T], new_state int16 [C, 4]).
    """
    x = np.atleast_2d(np.asarray(x)).astype(np.int64)
    c, t = x.shape
    b0, b1, b2, a0, a1 = (int(v) for v in coef5)
    if state is None:
        state = np.zeros((c, 4), dtype=np.int16)
    st = np.asarray(state).astype(np.int64)
    x1, x2, y1, y2 = (st[:, i].copy() for i in range(4))
    y = np.empty((c, t), dtype=np.int64)
    for n in range(t):
        xn = x[:, n]
        s = (slice_T(xn, b2) + slice_T(x1, b1) + slice_T(x2, b0)
             - slice_T(y2, a0) - slice_T(y1, a1))
        yn = _wrap16(s)
        y[:, n] = yn
        x2, x1 = x1, xn
        y2, y1 = y1, yn
    new_state = np.stack([x1, x2, y1, y2], axis=1).astype(np.int16)
    return y.astype(np.int16), new_state


# ----------------------------------------------------------------------- a4, a5
def stage_coeffs(coeff12):
    """The 6 stages' (B0,B1,B2,A0,A1): stages 1,3,5 take bytes 0..4 (set 0,
    ALPHA), stages 2,4,6 bytes 6..10 (set 1, BETA); bytes 5 and 11 (A2) are not
    connected to any product (NEW/filter_iir12_cust.vhd:68-240,
    NEW/filter_iir_cust.vhd:103-118)."""
    c = np.asarray(coeff12, dtype=np.int64).reshape(12)
    alpha = tuple(int(v) for v in c[0:5])
    beta = tuple(int(v) for v in c[6:11])
    return [alpha, beta, alpha, beta, alpha, beta]


def iir12(x: np.ndarray, coeff12, state: np.ndarray | None = None):
    """filter_iir12[_cust]: six biquads in series.
    x int16 [C, T]; state int16 [C, 6, 4] (per stage x1,x2,y1,y2) or None (zeros).
    Returns (y [C, T], new_state [C, 6, 4])."""
    x = np.atleast_2d(np.asarray(x, dtype=np.int16))
    c = x.shape[0]
    if state is None:
        state = np.zeros((c, 6, 4), dtype=np.int16)
    state = np.asarray(state, dtype=np.int16).reshape(c, 6, 4)
    new_state = np.empty_like(state)
    v = x
    for s, coef5 in enumerate(stage_coeffs(coeff12)):
        v, new_state[:, s, :] = biquad(v, coef5, state[:, s, :])
    return v, new_state


def iir_sections(x: np.ndarray, coeff6x6, state: np.ndarray | None = None):
    """Superset of iir12: six INDEPENDENT sections, int8 [6][6] in the
    RTL's register order B0,B1,B2,A0,A1,A2 (A2 unconnected); same per-stage arithmetic
    (NEW/filter_iir_cust.vhd:96-118)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.int16))
    c = x.shape[0]
    k = np.asarray(coeff6x6, dtype=np.int64).reshape(6, 6)
    if state is None:
        state = np.zeros((c, 6, 4), dtype=np.int16)
    state = np.asarray(state, dtype=np.int16).reshape(c, 6, 4)
    new_state = np.empty_like(state)
    v = x
    for s in range(6):
        v, new_state[:, s, :] = biquad(v, tuple(int(t) for t in k[s, :5]), state[:, s, :])
    return v, new_state


# --------------------------------------------------------------------------- a7
class CommandDecoder:
    """Byte protocol of the control plane, restating
    NEW/command_control.vhd:46-78 (mode / reset / start, ignored while busy),
    NEW/rx_filter_coeff.vhd:41-66 (0xF1 then 12 data bytes, busy meanwhile),
    NEW/filter_iir12_cust.vhd:48-60,83-94 (12 registers, cleared on reset),
    IMP/sequ2.vhd:83-96,216 (0xEF/0xFE transport, 0xA5 frame request; sequ_2
    sees uart_rx_valid AND NOT busy, IMP/dsp_system_top.vhd:644).

    Intended (GUI-assumed) semantics per SURVEY D11: a completed 0xF1 upload
    replaces the 12 bank-1 registers atomically and leaves filter state alone.
    """

    def __init__(self):
        self.mode = MODE_BYPASS                      # command_control.vhd:31,50
        self.bank1 = np.zeros(12, dtype=np.int8)     # filter_iir12_cust.vhd:51-52
        self.transport = CMD_ETHERNET_MODE           # sequ2.vhd:86
        self.pending = None                          # bytes collected while busy
        self.events = []                             # ('reset'|'start'|'request'|'load',)

    @property
    def busy(self):
        return self.pending is not None

    def feed(self, data: bytes):
        for b in bytes(data):
            if self.pending is not None:             # ACQUIRE: bytes are data
                self.pending.append(b)
                if len(self.pending) == 12:
                    self.bank1 = np.frombuffer(bytes(self.pending), dtype=np.int8).copy()
                    self.pending = None
                    self.events.append('load')
                continue
            if b == CMD_FILTER_UPDATE:
                self.pending = []
            elif b in (MODE_BANK0, MODE_BANK1, MODE_BYPASS):
                self.mode = b
            elif b == CMD_RESET:
                # rst_n pulse: mode -> bypass, bank-1 registers -> 0, filter
                # history -> 0, transport -> Ethernet
                self.mode = MODE_BYPASS
                self.bank1 = np.zeros(12, dtype=np.int8)
                self.transport = CMD_ETHERNET_MODE
                self.events.append('reset')
            elif b == CMD_START:
                self.events.append('start')
            elif b == CMD_UART_REQUEST:
                self.events.append('request')
            elif b in (CMD_ETHERNET_MODE, CMD_UART_MODE):
                self.transport = b
            # every other byte matches no decoder branch and is dropped
        return self


# -------------------------------------------------------------------- a6, a8
def fft_bins(filtered: np.ndarray) -> np.ndarray:
    """FFT oracle: forward DFT of {re = sample, im = 0} (NEW/command_control.vhd:123),
    natural order, float64 (stands in for the proprietary xfft)."""
    return np.fft.fft(np.asarray(filtered).astype(np.float64), axis=-1)


def quantize_bins(bins: np.ndarray, log2_scale: int = -14, rounding: str = "floor"):
    """Scale bins by 2**log2_scale (xfft default schedule: 1/N),
    truncate (xfft 'truncation' rounding) and saturate to int16.
    Returns (re int16, im int16)."""
    s = 2.0 ** log2_scale
    fn = np.floor if rounding == "floor" else np.rint
    re = np.clip(fn(bins.real * s), -32768, 32767).astype(np.int16)
    im = np.clip(fn(bins.imag * s), -32768, 32767).astype(np.int16)
    return re, im


# --------------------------------------------------------------------- a9, a10
def pack_frame(re: np.ndarray, im: np.ndarray) -> np.ndarray:
    """One GUI frame per row: per bin re_lo, re_hi, im_lo, im_hi
    (FIFO word {im[31:16], re[15:0]} drained LSB first, IMP/sequ2.vhd:153,234;
    consumer GUI:255-257).  re/im int16 [..., N] -> uint8 [..., 4N]."""
    re = np.asarray(re, dtype="<i2")
    im = np.asarray(im, dtype="<i2")
    iq = np.stack([re, im], axis=-1)
    return np.ascontiguousarray(iq).view(np.uint8).reshape(*re.shape[:-1], -1)


# ------------------------------------------------------------------------- a11
def decode_frame(frame: np.ndarray):
    """Host decode, restating GUI:250-270: int16 LE re/im -> float32, and
    mag = sqrt(re^2 + im^2) in float32."""
    arr = np.asarray(frame, dtype=np.uint8)
    iq = arr.reshape(*arr.shape[:-1], -1, 2, 2)
    v = (iq[..., 0].astype(np.uint16) | (iq[..., 1].astype(np.uint16) << 8)).astype(np.int16)
    re = v[..., 0].astype(np.float32)
    im = v[..., 1].astype(np.float32)
    mag = np.sqrt(re ** 2 + im ** 2)
    return re, im, mag


# ------------------------------------------------------------------ whole path
def receive_chain(x, mode, bank1, rom=None, state=None, start=0,
                  log2_scale=-14, rounding="floor"):
    """x int16 [C, N] -> dict(filtered, state, bins, re, im, frame).
    mode: 0x00 bank 0, 0xA1 bank 1, anything else bypass
    (NEW/command_control.vhd:90-116)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.int16))
    if rom is None:
        rom = hann_rom()
    w = window(x, rom, start)
    if mode == MODE_BANK0:
        filt, state = iir12(w, BANK0_COEFF, state)
    elif mode == MODE_BANK1:
        filt, state = iir12(w, bank1, state)
    else:
        filt = w
    bins = fft_bins(filt)
    re, im = quantize_bins(bins, log2_scale, rounding)
    return dict(windowed=w, filtered=filt, state=state, bins=bins, re=re, im=im,
                frame=pack_frame(re, im))


# ---------------------------------------------------------- float CPU baseline
def float_sos(coeff12) -> np.ndarray:
    """Float second-order sections equivalent to the int8 register set
    (SURVEY 8d): b = [B2,B1,B0]/128, a = [1, A1/128, A0/128] per stage."""
    sos = []
    for (b0, b1, b2, a0, a1) in stage_coeffs(coeff12):
        sos.append([b2 / 128.0, b1 / 128.0, b0 / 128.0, 1.0, a1 / 128.0, a0 / 128.0])
    return np.asarray(sos, dtype=np.float64)


def cpu_float_chain(x: np.ndarray, coeff12, rom: np.ndarray, zi=None):
    """The 'repo's CPU path' of BASELINE.md section 3: numpy window, scipy sosfilt,
    np.fft.fft, abs.  float64.  Used only as the timed CPU baseline.
    zi (sosfilt's [sections, channels, 2] delay values): the filter history carried from the
    previous frame (BASELINE config 3, continuous channels); the new one is returned third."""
    from scipy.signal import sosfilt
    xf = x.astype(np.float64) * (rom.astype(np.float64) / 32768.0)
    if zi is None:
        yf = sosfilt(float_sos(coeff12), xf, axis=-1)
        bins = np.fft.fft(yf, axis=-1)
        return bins, np.abs(bins)
    yf, zf = sosfilt(float_sos(coeff12), xf, axis=-1, zi=zi)
    bins = np.fft.fft(yf, axis=-1)
    return bins, np.abs(bins), zf


# ------------------------------------------------------------------- stimulus
def lcg_stimulus(n: int, seed: int = 1) -> np.ndarray:
    """LCG stimulus of SURVEY Appendix A.3."""
    out = np.empty(n, dtype=np.int64)
    s = seed
    for i in range(n):
        s = (1103515245 * s + 12345) % (1 << 31)
        out[i] = ((s >> 8) & 0xFFFF) - 32768
    return out.astype(np.int16)


def tone_noise(channels, n=FFT_SIZE, seed=0, amp=1400.0, sigma=100.0):
    """Synthetic 12-bit tone + noise of SURVEY 8d; ``channels`` = iterable of
    channel numbers.  int16 [len(channels), n]."""
    channels = np.asarray(list(channels), dtype=np.int64)
    rng = np.random.default_rng(seed)
    t = np.arange(n) / FS_HZ
    f = 20e3 + (channels % 4096) * 100.0
    phi = 2 * np.pi * np.mod(channels * 0.6180339887, 1.0)
    sig = amp * np.sin(2 * np.pi * f[:, None] * t[None, :] + phi[:, None])
    sig = sig + sigma * rng.standard_normal(sig.shape)
    return np.clip(np.rint(sig), -2048, 2047).astype(np.int16)
