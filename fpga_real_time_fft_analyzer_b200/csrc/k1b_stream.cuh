// k1b_stream.cuh - K1b: window + IIR12 of ONE long stream, parallel in time.
//
// The biquad of NEW/filter_iir_cust.vhd:96-100 truncates every product, so the
// recurrence is not linear: trajectories started from different histories do not
// merge, they settle a few LSB apart (the dead band of a truncating second-order
// section, the same mechanism as its zero-input limit cycles).  A time-parallel
// evaluation therefore cannot be bit-exact; K1b is the chunked scan with that
// stated error, and the bit-exact answer for one stream is the systolic k1_split
// kernel with one channel (fra_iir_stream(exact = 1)).
//   1. speculate  one lane per chunk: start `warm` samples before the chunk from
//                 a zero history and run the exact cascade up to the chunk start.
//                 The linear part of the missing history is A^warm * s, below half
//                 an LSB once warm >= 18 ln2 / -ln(pole radius): the block scan of
//                 the state-space recurrence truncated to its nearest neighbour,
//                 which is all that survives for a stable cascade.  Then filter
//                 the chunk; record entry and exit states;
//   2. verify     neighbouring chunks compare exit(p-1) with entry(p) - the
//                 neighbour's state travels one lane up by warp shuffle - and
//                 report how many differ and by how many LSB at most.
#pragma once
#include "fra_common.cuh"

namespace fra {

struct K1bArgs {
    const int16_t *in;      // [n]
    int16_t *out;           // [n]
    const int *rom32;       // window ROM widened to int32
    CascadeCoef coef;
    int16_t *entry;         // [P][24] state at each chunk's first sample
    int16_t *exit_;         // [P][24] state after each chunk's last sample
    const int16_t *state0;  // [24] true state before sample 0 (used when continuous)
    int *stats;             // [0] chunks whose entry state differs from the neighbour's exit, [1] max |difference| (LSB)
    unsigned long long n;   // samples
    int chunk;              // samples per chunk (multiple of 8)
    int warm;               // warm-up samples
    int n_chunks;
    int continuous;
    int apply_window;
    int iir;                // 0: bypass (window only)
};

FRA_DEV void state_load(StageState (&st)[kStages], const int16_t *p)
{
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
        st[s].x1 = small_int_to_float((int)p[4 * s + 0]);
        st[s].x2 = small_int_to_float((int)p[4 * s + 1]);
        st[s].y1 = small_int_to_float((int)p[4 * s + 2]);
        st[s].y2 = small_int_to_float((int)p[4 * s + 3]);
    }
}

FRA_DEV void state_store(const StageState (&st)[kStages], int16_t *p)
{
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
        p[4 * s + 0] = (int16_t)(int)st[s].x1;
        p[4 * s + 1] = (int16_t)(int)st[s].x2;
        p[4 * s + 2] = (int16_t)(int)st[s].y1;
        p[4 * s + 3] = (int16_t)(int)st[s].y2;
    }
}

FRA_DEV void state_zero(StageState (&st)[kStages])
{
#pragma unroll
    for (int s = 0; s < kStages; ++s) st[s].x1 = st[s].x2 = st[s].y1 = st[s].y2 = 0.0f;
}

// run samples [begin, end) (multiples of 8) of the stream through window + cascade
template <bool WRITE>
FRA_DEV void stream_run(const K1bArgs &a, StageState (&st)[kStages], unsigned long long begin, unsigned long long end)
{
    for (unsigned long long n0 = begin; n0 < end; n0 += 8) {
        const uint4 xv = ldg128(a.in + n0);
        const unsigned xw[4] = {xv.x, xv.y, xv.z, xv.w};
        const int wbase = (int)(n0 & (unsigned long long)(kWindowLen - 1));
        unsigned ow[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float acc2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int x = e ? hi16(xw[h]) : lo16(xw[h]);
                if (a.apply_window) x = window_int(x, __ldg(a.rom32 + wbase + 2 * h + e));
                float v = small_int_to_float(x);
                float acc = __int_as_float((x + 32768) + 0x4B400000);     // offset-binary, as biquad_step leaves it
                if (a.iir) {
#pragma unroll
                    for (int s = 0; s < kStages; ++s) acc = biquad_step(v, a.coef.set[s & 1], st[s], &v);
                }
                acc2[e] = acc;
            }
            ow[h] = pack16_acc(acc2[0], acc2[1]);
        }
        if (WRITE) stg128(a.out + n0, make_uint4(ow[0], ow[1], ow[2], ow[3]));
    }
}

__global__ void __launch_bounds__(64) k1b_speculate(K1bArgs a)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_chunks) return;
    const unsigned long long start = (unsigned long long)p * (unsigned long long)a.chunk;
    unsigned long long end = start + (unsigned long long)a.chunk;
    if (end > a.n) end = a.n;
    const unsigned long long w0 = (start > (unsigned long long)a.warm) ? start - (unsigned long long)a.warm : 0ull;
    StageState st[kStages];
    if (w0 == 0 && a.continuous) state_load(st, a.state0);
    else state_zero(st);
    stream_run<false>(a, st, w0, start);
    state_store(st, a.entry + (size_t)p * 24);
    stream_run<true>(a, st, start, end);
    state_store(st, a.exit_ + (size_t)p * 24);
}

// verify: lane p compares its entry state with lane p-1's exit state; the exit
// state moves one lane up by shuffle (the warp's first lane reads it from memory)
FRA_DEV int absdiff16(unsigned a, unsigned b)
{
    const int d0 = lo16(a) - lo16(b), d1 = hi16(a) - hi16(b);
    return max(d0 < 0 ? -d0 : d0, d1 < 0 ? -d1 : d1);
}

__global__ void __launch_bounds__(128) k1b_verify(K1bArgs a)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = p < a.n_chunks;
    const int pc = live ? p : a.n_chunks - 1;
    const uint2 *ex = reinterpret_cast<const uint2 *>(a.exit_ + (size_t)pc * 24);
    const uint2 *en = reinterpret_cast<const uint2 *>(a.entry + (size_t)pc * 24);
    int dev = 0;
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
        const uint2 mine = ex[s];
        unsigned px = __shfl_up_sync(0xffffffffu, mine.x, 1);
        unsigned py = __shfl_up_sync(0xffffffffu, mine.y, 1);
        if (lane == 0 && pc > 0) {
            const uint2 prev = reinterpret_cast<const uint2 *>(a.exit_ + (size_t)(pc - 1) * 24)[s];
            px = prev.x; py = prev.y;
        }
        const uint2 e = en[s];
        dev = max(dev, max(absdiff16(e.x, px), absdiff16(e.y, py)));
    }
    if (live && p > 0 && dev != 0) {
        atomicAdd(a.stats, 1);
        atomicMax(a.stats + 1, dev);
    }
}

}  // namespace fra
