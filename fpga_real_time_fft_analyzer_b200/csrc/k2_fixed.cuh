// k2_fixed.cuh - the FFT as a 16-bit fixed-point, scaled, truncating pipeline (FRA_FFT_FIXED16):
// what the Xilinx core is CONFIGURED to be (IP/xfft_0/xfft_0.xci:12-27: 16-bit data, 16-bit phase
// factors, scaled, truncation, pipelined streaming = radix-2^2, natural order, no overflow flag),
// for users who want the FPGA's quantisation noise in the spectrum instead of an fp32 FFT
// (SURVEY section 8 row f4).  PARITY UNPINNED vs xfft: the core is proprietary, its internal word
// growth and truncation points are not published; the arithmetic here is the one below, restated
// as a numpy integer model for the tests (which are bit-exact against it) and documented in DESIGN.md:
//   * decimation in frequency, pairs of radix-2 butterfly stages with the trivial -j rotation
//     between them and one twiddle multiplier behind each pair;
//   * int16 re / im between the pairs; inside a pair the two butterflies grow to 18 bits, the product
//     with the Q1.15 twiddle (round(cos, -sin * 2^15), clipped to 32767) is exact (int64) and ONE
//     arithmetic shift by 15 + 2 truncates it (floor) and divides the pair by 4 - the core's default
//     schedule, 1/N overall; the un-rotated quarter (W^0) bypasses the multiplier; a lone last radix-2
//     stage (odd log2 N) divides by 2;
//   * a value that leaves 16 bits wraps (two's complement), as a core without ovflo would.
// One frame per CTA, the whole frame in shared memory as packed int16 pairs (4 N bytes), one HBM
// round trip per frame.  Not the hot path: plain indexing, no swizzle.
#pragma once
#include "fra_common.cuh"
#include "k2_fft.cuh"

namespace fra {

constexpr int kFixedThreads = 256;

FRA_DEV uint32_t fx_pack(long long re, long long im) { return ((uint32_t)re & 0xFFFFu) | ((uint32_t)im << 16); }
FRA_DEV int fx_re(uint32_t w) { return lo16(w); }
FRA_DEV int fx_im(uint32_t w) { return hi16(w); }

// (ar + i ai) * (wr + i wi) >> 17, wrapped to 16 bits
FRA_DEV uint32_t fx_twiddle(int ar, int ai, uint32_t w)
{
    const long long wr = fx_re(w), wi = fx_im(w);
    const long long pr = ((long long)ar * wr - (long long)ai * wi) >> 17;
    const long long pi = ((long long)ar * wi + (long long)ai * wr) >> 17;
    return fx_pack(pr, pi);
}

template <bool WIN>
__global__ void __launch_bounds__(kFixedThreads) k2_fixed(K2Args a, const uint32_t *twfx, int log2n)
{
    FRA_DYN_SMEM(smem_raw);
    uint32_t *buf = reinterpret_cast<uint32_t *>(smem_raw);
    const int tid = threadIdx.x;
    const int n_total = 1 << log2n;
    const size_t frame = blockIdx.x;
    const int16_t *src = reinterpret_cast<const int16_t *>(a.in) + frame * n_total;
    for (int i = tid; i < n_total; i += kFixedThreads) {
        int x = src[i];
        if (WIN) x = window_int(x, __ldg(a.rom32 + (i & (kWindowLen - 1))));
        buf[i] = (uint32_t)x & 0xFFFFu;                   // {im = 0, re = sample}: NEW/command_control.vhd:123
    }
    __syncthreads();
    int n = n_total;
    while (n >= 4) {
        const int q = n >> 2;
        const int tstep = n_total / n;                     // W_n^k = W_N^(k N / n)
        for (int b = tid; b < (n_total >> 2); b += kFixedThreads) {
            const int k = b & (q - 1);
            const int base = (b / q) * n + k;
            const uint32_t x0 = buf[base], x1 = buf[base + q], x2 = buf[base + 2 * q], x3 = buf[base + 3 * q];
            // first butterfly (spacing n/2)
            const int sr0 = fx_re(x0) + fx_re(x2), si0 = fx_im(x0) + fx_im(x2);
            const int sr1 = fx_re(x1) + fx_re(x3), si1 = fx_im(x1) + fx_im(x3);
            const int dr0 = fx_re(x0) - fx_re(x2), di0 = fx_im(x0) - fx_im(x2);
            const int er = fx_re(x1) - fx_re(x3), ei = fx_im(x1) - fx_im(x3);
            const int dr1 = ei, di1 = -er;                  // * (-j)
            // second butterfly (spacing n/4) and the pair's twiddles W^0 | W^2k | W^k | W^3k
            const int kt = k * tstep;
            buf[base] = fx_pack((long long)(sr0 + sr1) >> 2, (long long)(si0 + si1) >> 2);
            buf[base + q] = fx_twiddle(sr0 - sr1, si0 - si1, __ldg(twfx + ((2 * kt) & (n_total - 1))));
            buf[base + 2 * q] = fx_twiddle(dr0 + dr1, di0 + di1, __ldg(twfx + kt));
            buf[base + 3 * q] = fx_twiddle(dr0 - dr1, di0 - di1, __ldg(twfx + ((3 * kt) & (n_total - 1))));
        }
        __syncthreads();
        n = q;
    }
    if (n == 2) {
        for (int b = tid; b < (n_total >> 1); b += kFixedThreads) {
            const uint32_t x0 = buf[2 * b], x1 = buf[2 * b + 1];
            buf[2 * b] = fx_pack((long long)(fx_re(x0) + fx_re(x1)) >> 1, (long long)(fx_im(x0) + fx_im(x1)) >> 1);
            buf[2 * b + 1] = fx_pack((long long)(fx_re(x0) - fx_re(x1)) >> 1, (long long)(fx_im(x0) - fx_im(x1)) >> 1);
        }
        __syncthreads();
    }
    // natural order (bit reversal undone), the frame's byte order re_lo re_hi im_lo im_hi (IMP/sequ2.vhd:153,234)
    const size_t out0 = frame * n_total;
    const float unscale = (float)n_total;
    for (int k = tid; k < n_total; k += kFixedThreads) {
        const uint32_t w = buf[__brev((unsigned)k) >> (32 - log2n)];
        if (a.frames != nullptr) a.frames[out0 + k] = w;
        const float fre = (float)fx_re(w), fim = (float)fx_im(w);
        if (a.iq != nullptr) a.iq[out0 + k] = make_float2(fre * unscale, fim * unscale);
        if (a.mag != nullptr) {
            const float m = __fsqrt_rn(__fadd_rn(__fmul_rn(fre, fre), __fmul_rn(fim, fim)));
            float *dst = a.mag + out0 + k;
            *dst = (a.mag_alpha < 1.0f) ? __fmaf_rn(a.mag_alpha, m - *dst, *dst) : m;
        }
        if (a.phase != nullptr) a.phase[out0 + k] = atan2f(fim, fre);
    }
}

}  // namespace fra
