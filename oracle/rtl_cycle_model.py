"""Cycle-accurate register-level restatement of the reference's biquad, cascade
and window entities.  TEST INFRASTRUCTURE ONLY (see oracle/golden.py header).

Written independently of oracle/golden.py's closed-form difference equation:
this file models every clocked register of the VHDL and evaluates the
concurrent (combinational) assignments each cycle, so that agreement between
the two pins the closed form.  Pure Python integers,
one channel, small sizes only.
"""
from __future__ import annotations


def _s(v: int, bits: int) -> int:
    """Interpret the low ``bits`` bits of v as two's complement."""
    v = int(v) & ((1 << bits) - 1)
    return v - (1 << bits) if v >> (bits - 1) else v


class BiquadRTL:
    """NEW/filter_iir_cust.vhd (== IMP/filter_iir.vhd with package constants)."""

    def __init__(self, b0, b1, b2, a0, a1):
        self.coef = (b0, b1, b2, a0, a1)
        self.ve = [0, 0, 0]     # :45  input history registers
        self.vs12 = [0, 0]      # :46  vs(1), vs(2) registers; vs(0) is combinational
        self.valid = 0          # :55

    def comb(self):
        """Concurrent assignments :96-118 (24-bit products, slice 22..7, 16-bit sum)."""
        b0, b1, b2, a0, a1 = self.coef
        m_b0 = _s(self.ve[2] * b0, 24)
        m_b1 = _s(self.ve[1] * b1, 24)
        m_b2 = _s(self.ve[0] * b2, 24)
        m_a0 = _s(self.vs12[1] * a0, 24)
        m_a1 = _s(self.vs12[0] * a1, 24)
        sl = lambda m: _s(m >> 7, 16)                  # (22 downto 7)
        vs0 = _s(sl(m_b0) + sl(m_b1) + sl(m_b2) - sl(m_a0) - sl(m_a1), 16)
        return vs0, self.valid                          # o_data, o_valid (:90-91)

    def clock(self, i_valid: int, i_data: int):
        """Rising edge (rst_n = '1'): :121-194.  All registers sample the
        pre-edge values, so compute vs(0) first."""
        vs0, _ = self.comb()
        if i_valid:
            self.ve = [_s(i_data, 16), self.ve[0], self.ve[1]]
            self.vs12 = [vs0, self.vs12[0]]
        else:
            self.ve = [0, 0, 0]
            self.vs12 = [0, 0]
        self.valid = 1 if i_valid else 0


class Cascade12RTL:
    """NEW/filter_iir12_cust.vhd:68-240: stage k+1's i_valid/i_data are stage
    k's o_valid/o_data (combinational), coefficients ALPHA,BETA alternating."""

    def __init__(self, coeff12):
        c = [int(v) for v in coeff12]
        alpha, beta = c[0:5], c[6:11]
        self.stages = [BiquadRTL(*(alpha if k % 2 == 0 else beta)) for k in range(6)]

    def run(self, valid_seq, data_seq, extra_cycles=8):
        """Drive (i_valid, i_data) per cycle; collect o_data on cycles with
        o_valid = '1' (what command_control registers into the FFT stream)."""
        out = []
        seq = list(zip(valid_seq, data_seq)) + [(0, 0)] * extra_cycles
        for v, d in seq:
            # combinational outputs of every stage before the edge
            outs = [st.comb() for st in self.stages]
            if outs[5][1]:
                out.append(outs[5][0])
            ins = [(v, d)] + [(outs[k][1], outs[k][0]) for k in range(5)]
            for st, (iv, idat) in zip(self.stages, ins):
                st.clock(iv, idat)
        return out


class WindowRTL:
    """NEW/hann8192.vhd:28-47 with its three same-branch registers
    (coef_s, product, sample_out): the 'rtl_skew' alignment of SURVEY D10."""

    def __init__(self, rom):
        self.rom = [int(v) for v in rom]
        self.addr = 0
        self.coef_s = 0
        self.product = 0
        self.sample_out = 0
        self.valid = 0

    @staticmethod
    def round_resize(product: int) -> int:
        """:39  resize(product(31 downto 15) + product(14), 16)."""
        p = _s(product, 32)
        r17 = _s((p >> 15) + ((p >> 14) & 1), 17)
        sign = (r17 >> 16) & 1
        return (r17 & 0x7FFF) - (sign << 15)

    def clock(self, sample_en: int, sample_in: int):
        if sample_en:
            new_coef = self.rom[self.addr]
            new_product = _s(_s(sample_in, 16) * self.coef_s, 32)
            new_out = self.round_resize(self.product)
            self.coef_s, self.product, self.sample_out = new_coef, new_product, new_out
            self.valid = 1
            self.addr = (self.addr + 1) & 0x3FFF
        else:
            self.sample_out = 0
            self.valid = 0
