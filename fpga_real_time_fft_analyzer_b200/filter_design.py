"""Filter design -> int8 coefficient bytes, with a consistent convention.

The reference's GUI designs a filter with scipy, quantises each second-order section as
``round(64 * [b0, b1, b2, a0, a1, a2])`` (scripts/fft_analyzer_gui.py:159-179, scale 64 at :168)
and uploads two sections as the 12 bytes after 0xF1 (:571-613).  The RTL
(NEW/filter_iir_cust.vhd:96-118) wires those bytes as

    y[n] = T(x[n], B2) + T(x[n-1], B1) + T(x[n-2], B0) - T(y[n-2], A0) - T(y[n-1], A1),
    T(v, c) = (v * c) >> 7,            byte order B0, B1, B2, A0, A1, A2  (A2 unconnected)

so a GUI section lands with half its feed-forward gain, its ``a0`` (= 64) on the y[n-2] tap, its
``a1`` on the y[n-1] tap and its ``a2`` dropped.  ``libfra`` reproduces the RTL, and
therefore that mismatch, bit for bit.  This module is the explicit switch around it:

* ``quantize_sections(sos, rtl_compatible=False)`` gives the GUI's bytes (what the reference
  sends today; pinned against its own quantiser's output in tests/golden/gui_vectors.json);
* ``quantize_sections(sos, rtl_compatible=True)`` gives bytes that make the RTL arithmetic
  realise the designed sections: scale 128 to match the ``>> 7``, taps in the RTL's order.
  A ``>> 7`` datapath only represents coefficients in [-1, 127/128], so sections with
  ``|a1| >= 1`` (poles away from the middle of the band) cannot be realised by this hardware;
  they raise instead of silently saturating (``strict=False`` saturates, ``unrealizable`` lists them);
* the 12-byte protocol carries two sections that the RTL alternates over its six stages;
  ``sections_to_bank12`` says whether a design fits it, ``FraContext.load_sections`` /
  ``GpuReceiver.send_filter_sections`` upload six independent sections otherwise.

numpy only: ``sos`` is any array of rows ``[b0, b1, b2, a0, a1, a2]`` (e.g. from scipy.signal).
"""
from __future__ import annotations

import numpy as np

N_STAGES = 6
RTL_SCALE = 128        # products >> 7
GUI_SCALE = 64         # scripts/fft_analyzer_gui.py:168


def _as_sos(sos) -> np.ndarray:
    sos = np.atleast_2d(np.asarray(sos, dtype=np.float64))
    if sos.shape[1] != 6:
        raise ValueError("sos rows must be [b0, b1, b2, a0, a1, a2]")
    if np.any(sos[:, 3] == 0.0):
        raise ValueError("a0 must not be zero")
    return sos / sos[:, 3:4]                      # a0 = 1


def section_response(sos, n_points: int = 2048) -> np.ndarray:
    """|H_i(e^{jw})| of every section on a grid over [0, pi]: array [sections, n_points]."""
    sos = _as_sos(sos)
    z = np.exp(-1j * np.linspace(0.0, np.pi, n_points))
    num = sos[:, 0:1] + sos[:, 1:2] * z + sos[:, 2:3] * z * z
    den = 1.0 + sos[:, 4:5] * z + sos[:, 5:6] * z * z
    return np.abs(num / den)


def normalize_gain(sos) -> np.ndarray:
    """Redistribute the gain so that every section peaks at 1 (L-infinity scaling) and the
    overall gain of the cascade is kept by the LAST section when that does not exceed 1.

    scipy puts a design's whole gain in the first section (b ~ 1e-4, which quantises to zero -
    the leading ``[0, 0, 0, ...]`` / ``[0, 1, 0, ...]`` rows of the GUI fixtures), and the RTL
    datapath wraps instead of saturating, so sections should neither vanish nor exceed unity."""
    sos = _as_sos(sos).copy()
    total = 1.0
    for i in range(len(sos)):
        peak = float(section_response(sos[i:i + 1]).max())
        if peak > 0.0:
            sos[i, :3] /= peak
            total *= peak
    if total < 1.0:
        sos[-1, :3] *= total                      # an overall attenuation is realisable; a gain > 1 is not kept
    return sos


def quantize_sections(sos, rtl_compatible: bool = True, strict: bool = True, normalize: bool = True):
    """SOS rows -> int8 [sections][6] in the RTL's register order B0,B1,B2,A0,A1,A2.

    rtl_compatible=False reproduces the GUI: ``round(64 * row)`` saturated to int8, the row
    left in ITS order [b0,b1,b2,a0,a1,a2] - these are the bytes the reference uploads, and the RTL
    then mis-assigns them as described in the module docstring.
    rtl_compatible=True: scale 128, ``[b2, b1, b0, a2, a1, 0]`` so that the RTL's wiring realises
    the section; coefficients outside int8 raise (strict) or saturate (``unrealizable`` lists
    them)."""
    if not rtl_compatible:
        raw = np.atleast_2d(np.asarray(sos, dtype=np.float64))
        q = np.clip(np.round(raw * GUI_SCALE), -128, 127).astype(np.int8)
        return q
    want = _rtl_targets(sos, normalize)
    bad = _out_of_range(want)
    if bad and strict:
        raise ValueError("not realisable with >> 7 coefficients (range [-1, 127/128]): "
                         + ", ".join(f"section {i} {n} = {v:+.4f}" for i, n, v in bad))
    return np.clip(np.round(want), -128, 127).astype(np.int8)


def _rtl_targets(sos, normalize):
    s = normalize_gain(sos) if normalize else _as_sos(sos)
    return np.stack([s[:, 2], s[:, 1], s[:, 0], s[:, 5], s[:, 4], np.zeros(len(s))], axis=1) * RTL_SCALE


def _out_of_range(want):
    q = np.round(want)
    names = "B0 B1 B2 A0 A1 A2".split()
    return [(int(i), names[int(j)], float(want[i, j] / RTL_SCALE)) for i, j in zip(*np.nonzero((q < -128) | (q > 127)))]


def unrealizable(sos, normalize: bool = True):
    """[(section, register, wanted value)] for every coefficient a >> 7 datapath cannot hold."""
    return _out_of_range(_rtl_targets(sos, normalize))


def realized_sos(sections) -> np.ndarray:
    """The float sections the RTL arithmetic realises from bytes in its register order
    (ignoring the per-product floor): rows [b0, b1, b2, 1, a1, a2]."""
    k = np.atleast_2d(np.asarray(sections, dtype=np.float64)) / RTL_SCALE
    return np.stack([k[:, 2], k[:, 1], k[:, 0], np.ones(len(k)), k[:, 4], k[:, 3]], axis=1)


def expand_to_stages(sections) -> np.ndarray:
    """int8 [6][6] for the six stages: fewer sections are padded with pass-through stages
    (B2 = 127: the closest a >> 7 product gets to unity, gain 127/128 per padded stage)."""
    k = np.atleast_2d(np.asarray(sections, dtype=np.int8))
    if k.shape[1] != 6 or not 1 <= k.shape[0] <= N_STAGES:
        raise ValueError("1..6 sections of 6 bytes")
    out = np.zeros((N_STAGES, 6), dtype=np.int8)
    out[:, 2] = 127
    out[:k.shape[0]] = k
    return out


def sections_to_bank12(sections):
    """The 12 bytes of the 0xF1 protocol if the six stages are ALPHA, BETA, ALPHA, BETA, ALPHA,
    BETA (what filter_iir12_cust.vhd can hold), else None - use the six-section upload."""
    k = expand_to_stages(sections)
    if np.array_equal(k[0], k[2]) and np.array_equal(k[0], k[4]) and np.array_equal(k[1], k[3]) \
            and np.array_equal(k[1], k[5]):
        return np.concatenate([k[0], k[1]]).astype(np.int8)
    return None


def upload(ctx, sections, select: bool = True):
    """Load a design into a FraContext: through the byte protocol when it fits the 12-byte bank
    (0xF1 + 12 bytes, then 0xA1), as six independent sections otherwise.  Returns "bank12" or
    "sections"."""
    bank = sections_to_bank12(sections)
    if bank is not None:
        ctx.command(bytes([0xF1]) + bank.tobytes())
        how = "bank12"
    else:
        ctx.load_sections(expand_to_stages(sections))
        how = "sections"
    if select:
        ctx.command(0xA1)
    return how
