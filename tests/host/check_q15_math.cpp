// Exhaustive host check of the FP32-pipe identities the K1 kernels rely on
// (csrc/fra_common.cuh).  Built and run by tests/test_q15_math_host.py with
//   g++ -std=c++20 -O1 -frounding-math -DFRA_HOST_EMUL -Itests/emul ...
// TEST INFRASTRUCTURE ONLY.
#include "fra_common.cuh"
#include <cstdio>
// a full biquad step against the integer recipe is covered by tests/test_emul_kernels.py

static int floordiv128(int p) { return p >> 7; }

int main()
{
    using namespace fra;
    long bad = 0;
    const float accs[3] = {kMagicB, kMagicB + 131072.0f, kMagicB - 131072.0f};
    // 1. fma.rm / fma.rp accumulate floor(v*c/128) exactly, for every int16 x int8
    for (int c = -128; c <= 127; ++c) {
        float kc = (float)c / 128.0f, nkc = -(float)c / 128.0f;
        for (int v = -32768; v <= 32767; ++v) {
            int f = floordiv128(v * c);
            for (float a0 : accs) {
                float rd = __fmaf_rd((float)v, kc, a0);
                float ru = __fmaf_ru((float)v, nkc, a0);
                if (rd != a0 + (float)f || ru != a0 - (float)f) {
                    if (bad < 5) std::printf("fma mismatch v=%d c=%d a0=%f rd=%f ru=%f f=%d\n", v, c, a0, rd, ru, f);
                    ++bad;
                }
            }
        }
    }
    // 2. low-16-bit wrap of the offset-binary accumulator -> float, for every reachable sum
    for (int s = -5 * 32768 - 8; s <= 5 * 32768 + 8; ++s) {
        float acc = kMagicB + (float)s;
        float y = wrap16_to_float(acc, 0x4B000000u);
        int want = (int)(short)(unsigned short)(s & 0xFFFF);
        if (y != (float)want) { if (bad < 10) std::printf("wrap mismatch s=%d y=%f want=%d\n", s, y, want); ++bad; }
        if ((int)(short)acc_to_u16(acc) != want) { ++bad; }
    }
    // 3. window: every int16 sample x every int16 coefficient vs the VHDL bit recipe
    for (int c = -32768; c <= 32767; ++c) {
        for (int x = -32768; x <= 32767; ++x) {
            int p = x * c;
            int r17 = (p >> 15) + ((p >> 14) & 1);
            int sign = (r17 >> 16) & 1;
            int want = (r17 & 0x7FFF) - (sign << 15);
            if (window_int(x, c) != want) { if (bad < 15) std::printf("window mismatch x=%d c=%d\n", x, c); ++bad; }
        }
    }
    // 4. small int -> float
    for (int v = -32768; v <= 32768; ++v)
        if (small_int_to_float(v) != (float)v) ++bad;
    // 5. pack16 / lo16 / hi16
    for (int a = -32768; a <= 32767; a += 257)
        for (int b = -32768; b <= 32767; b += 263) {
            unsigned w = pack16(__float_as_uint(kMagic + (float)a), __float_as_uint(kMagic + (float)b));
            if (lo16(w) != a || hi16(w) != b) ++bad;
            unsigned w2 = pack16_acc(kMagicB + (float)a, kMagicB + (float)b);
            if (lo16(w2) != a || hi16(w2) != b) ++bad;
        }
    // 6. biquad_step_fast's biased y[n-1] product: for every |A1| <= kFastMaxA1 and every int16 y,
    //    fma.ru(y + kBias16, -A1/128, k0 + s4) == kMagicB + s4 - floor(y*A1/128), with the partial
    //    sum s4 of the four earlier terms at its extremes (|T| <= 2^15 each) and at zero
    for (int a1 = -kFastMaxA1; a1 <= kFastMaxA1; ++a1) {
        const float na1 = -(float)a1 / 128.0f, k0 = kMagicB + (float)a1 * 65792.0f;
        for (int y = -32768; y <= 32767; ++y) {
            const int f = floordiv128(y * a1);
            for (int s4 : {-4 * 32768, -1, 0, 1, 4 * 32768}) {
                const float got = __fmaf_ru((float)y + kBias16, na1, k0 + (float)s4);
                if (got != kMagicB + (float)(s4 - f)) {
                    if (bad < 20) std::printf("fast-step mismatch a1=%d y=%d s4=%d got=%f\n", a1, y, s4, got);
                    ++bad;
                }
            }
        }
    }
    // 7. biquad_step_biased: every operand biased by kBias16.  For every int8 coefficient and every
    //    int16 v, a product added to an accumulator that holds kMagicB + s + 65792 r (s = true partial
    //    sum, r = coefficient sum of the products still to come, BEFORE this one is added:
    //    r_before = r_after + c for fma.rm, r_after - c ... handled below) gives
    //    kMagicB + (s + floor(v c / 128)) + 65792 r_after exactly, at the extremes the host-side
    //    condition biased_order_ok() allows (|r_after| <= kBiasedMaxSuffix, |s| <= 4 * 2^15).
    for (int c = -128; c <= 127; ++c) {
        const float kc = (float)c / 128.0f, nkc = -(float)c / 128.0f;
        for (int v = -32768; v <= 32767; ++v) {
            const int f = floordiv128(v * c);
            const float u = biased_from_int(v);
            if (u != (float)v + kBias16) ++bad;
            for (int r_after : {-kBiasedMaxSuffix, 0, kBiasedMaxSuffix}) {
                for (int s4 : {-4 * 32768, 0, 4 * 32768}) {
                    // '+' product (B taps): the accumulator before it carries -65792 c on top of r_after
                    const double before_p = (double)kMagicB + s4 - 65792.0 * r_after - 65792.0 * c;
                    const float got_p = __fmaf_rd(u, kc, (float)before_p);
                    const double want_p = (double)kMagicB + s4 + f - 65792.0 * r_after;
                    // '-' product (A taps): carries +65792 c
                    const double before_m = (double)kMagicB + s4 - 65792.0 * r_after + 65792.0 * c;
                    const float got_m = __fmaf_ru(u, nkc, (float)before_m);
                    const double want_m = (double)kMagicB + s4 - f - 65792.0 * r_after;
                    if ((double)(float)before_p != before_p || (double)(float)before_m != before_m) ++bad;   // start values are exact
                    if ((double)got_p != want_p || (double)got_m != want_m) {
                        if (bad < 25) std::printf("biased-step mismatch c=%d v=%d r=%d s=%d\n", c, v, r_after, s4);
                        ++bad;
                    }
                }
            }
        }
    }
    //    ... and the fixed bank of the reference passes the condition, extreme sets do not
    if (!biased_order_ok(-14, 0, 14, 107, 21) || !biased_order_ok(-15, 0, 15, 107, -21)) ++bad;
    if (biased_order_ok(-128, 127, -128, -128, 127) || biased_order_ok(0, 0, 0, 0, 61)) ++bad;
    // 8. window_biased == window_int + bias for every int16 sample and every ROM value but -32768
    for (int c = -32767; c <= 32767; ++c)
        for (int x = -32768; x <= 32767; ++x)
            if (window_biased(x, 2 * c, 0x4B000000u) != biased_from_int(window_int(x, c))) {
                if (bad < 30) std::printf("window_biased mismatch x=%d c=%d\n", x, c);
                ++bad;
            }
    std::printf("bad=%ld\n", bad);
    return bad ? 1 : 0;
}
