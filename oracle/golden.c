/* CPU golden model of the reference's window + IIR12 + framing arithmetic, in
 * plain C.  TEST INFRASTRUCTURE ONLY: the checker for the CUDA kernels at
 * sizes where oracle/golden.py (numpy) is too slow.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product never links or calls it.
 *
 * Pinning: see the header of oracle/golden.py (window ROM pinned to the
 * reference's hann.vhd; arithmetic pinned to a cycle-accurate restatement and
 * SURVEY Appendix A; reference holds no vectors for this path).
 *
 * Path legend: NEW/ = SDR_v2.srcs/sources_1/new/, IMP/ = SDR_v2.srcs/sources_1/imports/new/.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static inline int32_t wrap16(int32_t v) { return (int32_t)(int16_t)(uint16_t)(uint32_t)v; }

/* scripts/hann_coeff.py:3-5 - offset-coded table, int16 wrap of +32768 */
void gold_hann_rom(int16_t *out, int n)
{
    for (int k = 0; k < n; ++k) {
        double w = 0.5 * (1.0 - cos(2.0 * M_PI * (double)k / (double)(n - 1)));
        double q = nearbyint((w - 0.5) * 65536.0);      /* np.round = half-to-even */
        out[k] = (int16_t)wrap16((int32_t)q);
    }
}

/* NEW/hann8192.vhd:36-39 - 32-bit product, round-half-up >>15 in 17 bits,
 * numeric_std resize (sign bit + low 15 bits) */
static inline int16_t window_one(int16_t x, int16_t c)
{
    int32_t p = (int32_t)x * (int32_t)c;
    int32_t r17 = (p >> 15) + ((p >> 14) & 1);
    int32_t sign = (r17 >> 16) & 1;
    return (int16_t)((r17 & 0x7FFF) - (sign << 15));
}

void gold_window(const int16_t *x, int16_t *y, size_t channels, size_t t,
                 const int16_t *rom, size_t rom_len, size_t start)
{
    for (size_t c = 0; c < channels; ++c)
        for (size_t n = 0; n < t; ++n)
            y[c * t + n] = window_one(x[c * t + n], rom[(start + n) % rom_len]);
}

/* NEW/filter_iir_cust.vhd:96-108 - mult(22 downto 7) of the 24-bit product */
static inline int32_t slice_T(int32_t v, int32_t c) { return wrap16((v * c) >> 7); }

/* NEW/filter_iir12_cust.vhd:68-240 - six stages, ALPHA (bytes 0..4) on stages
 * 1,3,5 and BETA (bytes 6..10) on stages 2,4,6; state[c][stage] = x1,x2,y1,y2 */
void gold_iir12(const int16_t *x, int16_t *y, size_t channels, size_t t,
                const int8_t coeff12[12], int16_t *state)
{
    for (size_t c = 0; c < channels; ++c) {
        int16_t *st = state + c * 24;
        for (size_t n = 0; n < t; ++n) {
            int32_t v = x[c * t + n];
            for (int s = 0; s < 6; ++s) {
                const int8_t *k = coeff12 + ((s & 1) ? 6 : 0);
                int32_t b0 = k[0], b1 = k[1], b2 = k[2], a0 = k[3], a1 = k[4];
                int16_t *q = st + 4 * s;                 /* x1 x2 y1 y2 */
                int32_t sum = slice_T(v, b2) + slice_T(q[0], b1) + slice_T(q[1], b0)
                            - slice_T(q[3], a0) - slice_T(q[2], a1);
                int32_t yn = wrap16(sum);
                q[1] = q[0]; q[0] = (int16_t)v;
                q[3] = q[2]; q[2] = (int16_t)yn;
                v = yn;
            }
            y[c * t + n] = (int16_t)v;
        }
    }
}

/* Superset used by SURVEY section 8 row f3: six INDEPENDENT sections, 6 bytes each in the same
 * register order (B0,B1,B2,A0,A1,A2; A2 unconnected).  The per-stage arithmetic is the one of
 * gold_iir12 (NEW/filter_iir_cust.vhd:96-118); only the coefficient routing differs. */
void gold_iir_sections(const int16_t *x, int16_t *y, size_t channels, size_t t,
                       const int8_t coeff36[36], int16_t *state)
{
    for (size_t c = 0; c < channels; ++c) {
        int16_t *st = state + c * 24;
        for (size_t n = 0; n < t; ++n) {
            int32_t v = x[c * t + n];
            for (int s = 0; s < 6; ++s) {
                const int8_t *k = coeff36 + 6 * s;
                int16_t *q = st + 4 * s;                 /* x1 x2 y1 y2 */
                int32_t yn = wrap16(slice_T(v, k[2]) + slice_T(q[0], k[1]) + slice_T(q[1], k[0])
                                    - slice_T(q[3], k[3]) - slice_T(q[2], k[4]));
                q[1] = q[0]; q[0] = (int16_t)v;
                q[3] = q[2]; q[2] = (int16_t)yn;
                v = yn;
            }
            y[c * t + n] = (int16_t)v;
        }
    }
}

/* window followed by the selected path (NEW/command_control.vhd:90-116):
 * mode 0x00 -> bank0, 0xA1 -> bank1, else bypass.  state may be NULL in bypass. */
void gold_window_iir(const int16_t *x, int16_t *y, size_t channels, size_t t,
                     const int16_t *rom, size_t rom_len, size_t start,
                     int mode, const int8_t bank0[12], const int8_t bank1[12],
                     int16_t *state)
{
    gold_window(x, y, channels, t, rom, rom_len, start);
    if (mode == 0x00)      gold_iir12(y, y, channels, t, bank0, state);
    else if (mode == 0xA1) gold_iir12(y, y, channels, t, bank1, state);
}

/* scale by 2^log2_scale, floor (rounding=0) or nearest-even (1), saturate;
 * then IMP/sequ2.vhd:153,234 byte order: re_lo re_hi im_lo im_hi per bin */
void gold_quantize_pack(const double *re, const double *im, size_t n_bins,
                        int log2_scale, int rounding, uint8_t *frame)
{
    double s = ldexp(1.0, log2_scale);
    for (size_t k = 0; k < n_bins; ++k) {
        double r = re[k] * s, i = im[k] * s;
        r = rounding ? nearbyint(r) : floor(r);
        i = rounding ? nearbyint(i) : floor(i);
        r = fmin(fmax(r, -32768.0), 32767.0);
        i = fmin(fmax(i, -32768.0), 32767.0);
        uint16_t ur = (uint16_t)(int16_t)r, ui = (uint16_t)(int16_t)i;
        frame[4 * k + 0] = (uint8_t)(ur & 0xFF);
        frame[4 * k + 1] = (uint8_t)(ur >> 8);
        frame[4 * k + 2] = (uint8_t)(ui & 0xFF);
        frame[4 * k + 3] = (uint8_t)(ui >> 8);
    }
}
