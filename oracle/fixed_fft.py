"""fixed_fft - integer model of the 16-bit scaled, truncating FFT mode (FRA_FFT_FIXED16).

TEST INFRASTRUCTURE (oracle/): imported by tests/ only; the product never loads it.

PARITY UNPINNED vs xfft.  The reference's FFT is the proprietary Xilinx LogiCORE xfft v9.1
(IP/xfft_0/xfft_0.xci; no source or simulation model under /root/reference).  What the
reference pins is its CONFIGURATION, and this model follows exactly that:

  xfft_0.xci:12    transform_length 16384            -> any power of two 1024..32768 here
  xfft_0.xci:13,15 pipelined_streaming_io            -> radix-2^2 decimation in frequency: pairs of
                                                        radix-2 butterfly stages with the trivial -j
                                                        rotation between them and ONE twiddle
                                                        multiplier behind each pair
  xfft_0.xci:18    input_width 16                    -> int16 re / im between the stage pairs
  xfft_0.xci:19    phase_factor_width 16             -> twiddles round(cos, -sin * 2^15), clipped to 32767
  xfft_0.xci:20    scaling_options scaled            -> every stage pair divides by 4 (the core's default
                                                        schedule [10 10 ... 10], 1/N in total; a lone last
                                                        radix-2 stage divides by 2)
  xfft_0.xci:21    rounding_modes truncation         -> floor (arithmetic shift right), never rounding
  xfft_0.xci:27    output_ordering natural_order     -> bit reversal undone at the end
  no ovflo port (xfft_0.xci: ovflo false)            -> a result that leaves 16 bits WRAPS

The core's internal word growth and the exact place of its truncations are not published; here
the two butterflies of a pair grow to 18 bits, the twiddle product is formed exactly (34 bits) and
ONE truncation (>> 15 + 2) brings the value back to 16 bits.  The un-rotated quarter of a pair
(twiddle W^0) bypasses the multiplier.
"""
from __future__ import annotations

import numpy as np


def twiddle_table(n: int):
    """W_n^t = exp(-2 pi i t / n), t = 0..n-1, as int64 (re, im) in Q1.15, clipped to +-32767."""
    t = np.arange(n, dtype=np.float64)
    wr = np.clip(np.rint(np.cos(2 * np.pi * t / n) * 32768.0), -32767, 32767).astype(np.int64)
    wi = np.clip(np.rint(-np.sin(2 * np.pi * t / n) * 32768.0), -32767, 32767).astype(np.int64)
    return wr, wi


def _wrap16(v):
    return ((v + 32768) & 0xFFFF) - 32768


def fixed_fft(x: np.ndarray):
    """x: int16 real frames [..., N] (imag = 0, NEW/command_control.vhd:123).
    Returns (re, im) int16 [..., N]: the N-point forward DFT scaled by 1/N, natural order."""
    x = np.asarray(x)
    n_total = x.shape[-1]
    log2n = int(np.log2(n_total))
    assert 1 << log2n == n_total
    lead = x.shape[:-1]
    re = x.astype(np.int64).reshape(-1, n_total)
    im = np.zeros_like(re)
    wr_t, wi_t = twiddle_table(n_total)
    n = n_total
    while n >= 4:
        q = n // 4
        blocks = n_total // n
        r = re.reshape(-1, blocks, 4, q)
        i = im.reshape(-1, blocks, 4, q)
        # first butterfly (spacing n/2): sums in quarters 0,1 / differences in quarters 2,3
        sr0, si0 = r[:, :, 0] + r[:, :, 2], i[:, :, 0] + i[:, :, 2]
        sr1, si1 = r[:, :, 1] + r[:, :, 3], i[:, :, 1] + i[:, :, 3]
        dr0, di0 = r[:, :, 0] - r[:, :, 2], i[:, :, 0] - i[:, :, 2]
        dr1, di1 = r[:, :, 1] - r[:, :, 3], i[:, :, 1] - i[:, :, 3]
        dr1, di1 = di1, -dr1                                   # the trivial rotation: * (-j)
        # second butterfly (spacing n/4)
        outs = [(sr0 + sr1, si0 + si1), (sr0 - sr1, si0 - si1), (dr0 + dr1, di0 + di1), (dr0 - dr1, di0 - di1)]
        k = np.arange(q, dtype=np.int64) * (n_total // n)      # W_n^k' = W_N^(k' N / n)
        new_r = np.empty_like(r)
        new_i = np.empty_like(i)
        for quarter, mult in ((0, 0), (1, 2), (2, 1), (3, 3)):  # twiddles W^0 | W^2k | W^k | W^3k
            ar, ai = outs[quarter]
            if mult == 0:
                pr, pi = ar >> 2, ai >> 2                      # no multiplier on this path: scaling only
            else:
                wr, wi = wr_t[(mult * k) % n_total], wi_t[(mult * k) % n_total]
                pr = (ar * wr - ai * wi) >> 17                 # exact product, one truncation: Q15 and the /4
                pi = (ar * wi + ai * wr) >> 17
            new_r[:, :, quarter] = _wrap16(pr)
            new_i[:, :, quarter] = _wrap16(pi)
        re, im = new_r.reshape(-1, n_total), new_i.reshape(-1, n_total)
        n = q
    if n == 2:                                                  # odd log2 N: one radix-2 stage, scaled by 2
        r = re.reshape(-1, n_total // 2, 2)
        i = im.reshape(-1, n_total // 2, 2)
        re = np.stack([_wrap16((r[:, :, 0] + r[:, :, 1]) >> 1), _wrap16((r[:, :, 0] - r[:, :, 1]) >> 1)], axis=-1).reshape(-1, n_total)
        im = np.stack([_wrap16((i[:, :, 0] + i[:, :, 1]) >> 1), _wrap16((i[:, :, 0] - i[:, :, 1]) >> 1)], axis=-1).reshape(-1, n_total)
    # natural order: X[k] sits at the bit-reversed position
    idx = np.arange(n_total)
    rev = np.zeros(n_total, dtype=np.int64)
    for b in range(log2n):
        rev |= ((idx >> b) & 1) << (log2n - 1 - b)
    re, im = re[:, rev], im[:, rev]
    return re.astype(np.int16).reshape(*lead, n_total), im.astype(np.int16).reshape(*lead, n_total)
