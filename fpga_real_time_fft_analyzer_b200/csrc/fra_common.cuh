// fra_common.cuh - shared definitions of the libfra kernels (sm_100a).
//
// The arithmetic contract (bit-exact window + biquad cascade) is described in
// DESIGN.md section 3; reference line numbers are given at each primitive.
#pragma once

#ifdef FRA_HOST_EMUL
// tests/emul only: the same sources compiled as C++ on host threads (cusim.h).
#include "cusim.h"
#define FRA_DYN_SMEM(name) unsigned char *name = cusim::g_dyn_smem
#define FRA_LAUNCH(kfn, grid, block, smem, stream, ...) \
    cusim::launch((grid), (block), (smem), [&] { kfn(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#define FRA_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define FRA_LAUNCH(kfn, grid, block, smem, stream, ...) \
    kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#include <stdint.h>

#define FRA_DEV __device__ __forceinline__

namespace fra {

#ifdef FRA_TIMELINE
// Diagnostic build only (tools/timeline_probe.py): first-CTA start and last-CTA end of every
// window+IIR ([0], [1]) and FFT ([2], [3]) launch, by call index, in globaltimer nanoseconds.
__device__ unsigned long long g_timeline[4][4096];
FRA_DEV unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
FRA_DEV void timeline_mark(int row, int step)
{
    if (threadIdx.x == 0 && step >= 0 && step < 4096) {
        if (row & 1) atomicMax(&g_timeline[row][step], globaltimer_ns());
        else atomicMin(&g_timeline[row][step], globaltimer_ns());
    }
}
#endif

constexpr int kWindowLen = 16384;          // NEW/hann.vhd: Hann_ROM(0 to 16383)
constexpr int kStages = 6;                 // NEW/filter_iir12_cust.vhd:68-240
constexpr float kMagic = 12582912.0f;      // 1.5 * 2^23: ulp = 1 on [2^23, 2^24)
constexpr float kMagicB = 12615680.0f;     // kMagic + 2^15 (offset-binary accumulator)
constexpr float kBias16 = 8421376.0f;      // 2^23 + 2^15

// Per-stage coefficients prepared on the host from the 12 int8 registers:
// products on the FP32 pipe, exactly.  |v*c| <= 2^22 fits the 24-bit significand,
// c/128 is exact, and an FFMA rounds once, so
//   fma.rm(v, c/128, acc)  = acc + floor(v*c/128)      (acc integer, ulp 1)
//   fma.rp(v, -c/128, acc) = acc - floor(v*c/128)
// which is mult(22 downto 7) of NEW/filter_iir_cust.vhd:96-100 up to the 16-bit
// wrap; the wrap is applied once to the sum (two's-complement addition commutes
// with it), by taking the low 16 bits of the accumulator's bit pattern.
struct StageCoef {
    float b2, b1, b0;    //  B2/128 -> x[n],  B1/128 -> x[n-1],  B0/128 -> x[n-2]
    float na0, na1;      // -A0/128 -> y[n-2], -A1/128 -> y[n-1]
    unsigned exp23;      // 0x4B000000, passed as data so that it lives in a register and
                         // PRMT's selector can be the immediate (one instruction, no re-materialised selector)
    float k0;            // kMagicB + A1 * 65792: accumulator start for biquad_step_fast
    float kb;            // kMagicB - 65792 (B0 + B1 + B2 - A0 - A1): accumulator start for biquad_step_biased
};

// biquad_step_fast is usable when kMagicB + A1 * 65792 +- (the partial sums, < 2^17.4) stays
// inside [2^23, 2^24)
constexpr int kFastMaxA1 = 60;

struct CascadeCoef {
    StageCoef set[kStages];   // one per stage; the 12-byte protocol fills them ALPHA, BETA, ALPHA, ... (stages 1,3,5 / 2,4,6)
};

struct StageState {      // registers ve(1), ve(2), vs(1), vs(2) as exact floats
    float x1, x2, y1, y2;
};

// The biquad accumulator starts at kMagicB = 1.5*2^23 + 2^15, so the low 16 bits
// of its bit pattern hold (sum + 32768) mod 2^16: the 16-bit wrapped sum in
// offset-binary.  One PRMT splices those 16 bits under the exponent of 2^23 and
// one FADD removes 2^23 + 2^15: the wrapped int16 as an exact float, with no
// integer<->float conversion on the recurrence.
FRA_DEV float wrap16_to_float(float acc, unsigned exp23)
{
    unsigned bits = __float_as_uint(acc);
    return __uint_as_float(__byte_perm(bits, exp23, 0x7610)) - kBias16;
}

// low 16 bits of the accumulator as two's complement (undo the offset-binary)
FRA_DEV unsigned acc_to_u16(float acc) { return (__float_as_uint(acc) ^ 0x8000u) & 0xFFFFu; }

// One biquad evaluation, NEW/filter_iir_cust.vhd:96-118 + :133-194 (state shift).
// Returns the accumulator; *y_out gets the wrapped value as a float.  Every partial
// sum is an exact integer, so the order of the five terms is free: the ones that
// are known early go first.
// B1Z: the x[n-1] coefficient is zero in both sets (true of the reference's fixed bank,
// IMP/filter_pkg.vhd:58,67): T(v, 0) = 0 exactly, so that product is skipped.
template <bool B1Z = false>
FRA_DEV float biquad_step(float x, const StageCoef &k, StageState &s, float *y_out)
{
    float acc = __fmaf_rd(s.x2, k.b0, kMagicB);
    if (!B1Z) acc = __fmaf_rd(s.x1, k.b1, acc);
    acc = __fmaf_ru(s.y2, k.na0, acc);
    acc = __fmaf_rd(x, k.b2, acc);          // x arrives late (previous stage / shuffle): fourth
    acc = __fmaf_ru(s.y1, k.na1, acc);      // y[n-1] is the recurrence: last
    float y = wrap16_to_float(acc, k.exp23);
    s.x2 = s.x1; s.x1 = x;
    s.y2 = s.y1; s.y1 = y;
    *y_out = y;
    return acc;
}

// The same step with a two-instruction recurrence.  PRMT's output u = y + kBias16 (the wrapped
// value still carrying the exponent bias) is fed to the y[n-1] product as it is:
//   u * na1 = y * na1 + kBias16 * na1,   kBias16 * na1 = -A1 * 65792, an integer,
// so with the accumulator started at k0 = kMagicB + A1 * 65792 the FFMA's exact sum is the
// one biquad_step forms and, while every partial result stays inside [2^23, 2^24)
// (|A1| <= kFastMaxA1), its single rounding gives the same integer.  The FADD that removes the
// bias now feeds only the later uses of y (y[n-2] product, next stage, output): the loop-carried
// chain is FFMA -> PRMT (~9.5 cycles) instead of FFMA -> PRMT -> FADD (14.2).
template <bool B1Z = false>
FRA_DEV float biquad_step_fast(float x, const StageCoef &k, StageState &s, float &u1, float *y_out)
{
    float acc = __fmaf_rd(s.x2, k.b0, k.k0);
    if (!B1Z) acc = __fmaf_rd(s.x1, k.b1, acc);
    acc = __fmaf_ru(s.y2, k.na0, acc);
    acc = __fmaf_rd(x, k.b2, acc);
    acc = __fmaf_ru(u1, k.na1, acc);
    const float u = __uint_as_float(__byte_perm(__float_as_uint(acc), k.exp23, 0x7610));
    const float y = u - kBias16;
    s.x2 = s.x1; s.x1 = x;
    s.y2 = s.y1; s.y1 = y;
    u1 = u;
    *y_out = y;
    return acc;
}

// The all-biased step: EVERY operand - x[n], x[n-1], x[n-2], y[n-1], y[n-2] - is the PRMT output
// u = v + kBias16 of the stage (or window) that produced it, so no FADD removes a bias anywhere:
// five FFMAs and one PRMT per stage and sample.  An FFMA forms u * (c/128) + acc exactly before
// its single rounding, and u * (c/128) = v * c/128 + 65792 c with 65792 c an integer, so each
// biased operand shifts the accumulator by a known integer.  The accumulator starts at
// kb = kMagicB - 65792 (B0 + B1 + B2 - A0 - A1); after the k-th product it holds
//   kMagicB + (true partial sum) - 65792 * (signed coefficients of the products still to come)
// and the rounding of every FFMA gives the same integer as in biquad_step as long as that value
// stays inside [2^23, 2^24): the coefficient sum of every SUFFIX of the product order
// (y[n-2], x[n-2], x[n-1], x[n], y[n-1]) must lie within +-kBiasedMaxSuffix.  The start value may be
// anywhere (it is an exact constant; only results are rounded), which is why the large y[n-2]
// coefficient of the reference's fixed bank (A0 = 107) goes first.  biased_order_ok() is the
// host-side test; coefficient sets that fail it use biquad_step / biquad_step_fast.
// tests/host/check_q15_math.cpp proves the identity for every int16 x int8 at the extreme
// accumulator values the condition allows.
constexpr int kBiasedMaxSuffix = 60;

inline bool biased_order_ok(int b0, int b1, int b2, int a0, int a1)
{
    const int term[5] = {-a0, b0, b1, b2, -a1};          // product order of biquad_step_biased
    int suffix = 0;
    for (int k = 4; k >= 1; --k) {
        suffix += term[k];
        if (suffix > kBiasedMaxSuffix || suffix < -kBiasedMaxSuffix) return false;
    }
    return true;
}

typedef StageState StageStateB;     // the same four registers, each biased by kBias16 (bit pattern 0x4B00xxxx)

// int in [-32768, 32767] -> v + kBias16 with one integer add (0x4B008000 = bits of kBias16)
FRA_DEV float biased_from_int(int v) { return __uint_as_float((unsigned)v + 0x4B008000u); }
// biased float -> its int16 as the low 16 bits (two's complement)
FRA_DEV unsigned biased_to_u16(float u) { return (__float_as_uint(u) ^ 0x8000u) & 0xFFFFu; }

template <bool B1Z = false>
FRA_DEV float biquad_step_biased(float ux, const StageCoef &k, StageStateB &s, float *u_out)
{
    float acc = __fmaf_ru(s.y2, k.na0, k.kb);
    acc = __fmaf_rd(s.x2, k.b0, acc);
    if (!B1Z) acc = __fmaf_rd(s.x1, k.b1, acc);
    acc = __fmaf_rd(ux, k.b2, acc);                 // x arrives late (previous stage): fourth
    acc = __fmaf_ru(s.y1, k.na1, acc);              // y[n-1] is the recurrence: last
    const float u = __uint_as_float(__byte_perm(__float_as_uint(acc), k.exp23, 0x7610));
    s.x2 = s.x1; s.x1 = ux;
    s.y2 = s.y1; s.y1 = u;
    *u_out = u;
    return acc;
}

// Speculative form of the same step for the latency-bound systolic kernel: assume the
// five-term sum does not leave the int16 range (true unless the filter overflows), so
// y = acc - kMagicB exactly and the recurrence is FFMA -> FADD instead of
// FFMA -> PRMT -> FADD.  *absmax tracks |y|; the caller re-runs the block with
// biquad_step when any |y| > 32767 was seen, so the result is always the exact one.
FRA_DEV float biquad_step_spec(float x, const StageCoef &k, StageState &s, float *y_out, float *absmax)
{
    float acc = __fmaf_rd(s.x2, k.b0, kMagicB);
    acc = __fmaf_rd(s.x1, k.b1, acc);
    acc = __fmaf_ru(s.y2, k.na0, acc);
    acc = __fmaf_rd(x, k.b2, acc);
    acc = __fmaf_ru(s.y1, k.na1, acc);
    float y = acc - kMagicB;
    *absmax = fmaxf(*absmax, fabsf(y));
    s.x2 = s.x1; s.x1 = x;
    s.y2 = s.y1; s.y1 = y;
    *y_out = y;
    return acc;
}

// hann_window arithmetic, NEW/hann8192.vhd:36-39: 32-bit product, +2^14, >>15 is
// product(31 downto 15) + product(14); numeric_std resize keeps sign + low 15
// bits, so the only out-of-range value (+32768 for x = c = -32768) becomes 0.
FRA_DEV int window_int(int x, int c)
{
    int v = (x * c + 16384) >> 15;
    return v == 32768 ? 0 : v;
}

// the same without the resize quirk: valid whenever c != -32768 (all but 30 ROM entries)
FRA_DEV int window_int_fast(int x, int c) { return (x * c + 16384) >> 15; }

// The window straight to the biased float of biquad_step_biased, two instructions: with c2 = 2 c
// (a second ROM table), (x c + 2^14) >> 15 = (x c2 + 2^15) >> 16, i.e. the HIGH half of
// p = x c2 + 0x8000; adding 0x80000000 as well turns that half into offset-binary, and one PRMT
// splices it under the exponent of 2^23: 2^23 + 2^15 + window(x, c).  One IMAD (the constant
// 0x80008000 is its addend) and one PRMT.  Valid whenever c != -32768 (x c2 then fits 32 bits and the
// result fits int16); the 30 ROM entries equal to -32768 take window_int + biased_from_int.
FRA_DEV float window_biased(int x, int c2, unsigned exp23)
{
    const unsigned p = (unsigned)(x * c2) + 0x80008000u;
    return __uint_as_float(__byte_perm(p, exp23, 0x7632));
}

// int in [-32768, 32767] -> float without the conversion pipe
FRA_DEV float small_int_to_float(int v)
{
    return (float)v;     // 32-bit source: I2FP.F32.S32 (ALU pipe), not the 16-bit I2F of the conversion unit
}

// two packed int16 (little-endian pair) -> ints
FRA_DEV int lo16(unsigned w) { return ((int)(w << 16)) >> 16; }     // shifts, so int->float stays a 32-bit I2FP
FRA_DEV int hi16(unsigned w) { return ((int)w) >> 16; }

// low 16 bits of two words -> one packed pair (first -> low half)
FRA_DEV unsigned pack16(unsigned a_bits, unsigned b_bits) { return __byte_perm(a_bits, b_bits, 0x5410); }
// the same for two offset-binary accumulators -> packed two's-complement int16 pair
FRA_DEV unsigned pack16_acc(float a, float b)
{
    return __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x5410) ^ 0x80008000u;
}

// ---- bulk asynchronous global -> shared copies (cp.async.bulk, the TMA engine's
// 1-D form: SASS UBLKCP) completing on an mbarrier.  One lane issues a whole row.
#ifdef FRA_HOST_EMUL
FRA_DEV void mbar_init(uint64_t *, int) {}
FRA_DEV void mbar_expect_tx(uint64_t *, unsigned) {}
FRA_DEV void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *) { std::memcpy(dst, src, bytes); }
FRA_DEV void mbar_wait(uint64_t *, unsigned) { __syncwarp(); }   // the issuing lanes' memcpy is complete after this
FRA_DEV void fence_proxy_async() {}
#else
FRA_DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
FRA_DEV void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
FRA_DEV void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
FRA_DEV void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// bounded wait: a mis-programmed barrier traps instead of hanging the GPU
FRA_DEV void mbar_wait(uint64_t *bar, unsigned parity)
{
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
// orders this thread's generic-proxy shared-memory accesses before later async-proxy (bulk copy) writes
FRA_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// ---- per-thread asynchronous 16-byte global -> shared copies (cp.async, SASS LDGSTS).
// Unlike a register prefetch they occupy no scoreboard slot, so a later wait on an OLDER
// load can never be held up by them (B200: six counting scoreboards per warp, shared).
#ifdef FRA_HOST_EMUL
FRA_DEV void cp_async16(void *dst, const void *src) { std::memcpy(dst, src, 16); }
FRA_DEV void cp_async_commit() {}
template <int N> FRA_DEV void cp_async_wait() {}
#else
FRA_DEV void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
FRA_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> FRA_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

FRA_DEV uint4 ldg128(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
FRA_DEV void stg128(void *p, uint4 v) { *reinterpret_cast<uint4 *>(p) = v; }

}  // namespace fra
