// k1b_stream.cuh - K1b: window + IIR12 of ONE long stream, parallel in time.
//
// The biquad of NEW/filter_iir_cust.vhd:96-100 truncates every product, so the
// recurrence is not linear: trajectories started from different histories do not
// merge, they settle a few LSB apart (the dead band of a truncating second-order
// section, the same mechanism as its zero-input limit cycles).  A time-parallel
// evaluation therefore cannot be bit-exact; K1b is the chunked scan with that
// stated error, and the bit-exact answer for one stream is the systolic k1_split
// kernel with one channel (fra_iir_stream(exact = 1)).
//
// The scan is a block scan of the cascade's STATE-SPACE recurrence.  The affine
// model of one sample step,  s' = A s + B u + d  (s = the 24 history values of the
// six stages, u = the windowed sample, d = the mean truncation offset -1/2 per
// non-zero product), gives over a chunk of L samples  s(end) = A^L s(begin) + e,
// with e the zero-state response of the chunk:
//   1. k1b_lin_ends    one lane per chunk: e_p, by running the float model over the chunk
//   2. k1b_scan_warp   inside each warp a Hillis-Steele scan by __shfl_up with the powers
//                      (A^L)^1,2,4,8,16 (24x24, from the host, in shared memory);
//      k1b_aggr_level  the warps' aggregates chained with (A^L)^32 - again a Hillis-Steele scan,
//                      one launch per level, stopped when the matrix power is negligible;
//      k1b_scan_warp   again, with each warp's carry injected: the state at every chunk
//                      boundary
//   3. k1b_speculate   one lane per chunk: start from the rounded predicted state, filter
//                      the chunk with the EXACT integer cascade, record the exit state
//   4. k1b_verify      neighbours compare exit(p-1) with entry(p) - the neighbour's state
//                      travels one lane up by warp shuffle - and report how many differ
//                      and by how many LSB at most.
#pragma once
#include "fra_common.cuh"

namespace fra {

constexpr int kStateDim = 4 * kStages;      // (x1, x2, y1, y2) x 6 stages
constexpr int kScanLevels = 5;              // lane distances 1, 2, 4, 8, 16
constexpr int kAggrLevels = 14;             // scan over the warps' aggregates: distances 1 .. 8192 warps (2^19 chunks)
constexpr int kScanMats = kScanLevels + kAggrLevels;  // M^(1,2,4,8,16) with M = A^L, then Q^(1,2,4,...) with Q = M^32

struct K1bArgs {
    const int16_t *in;      // [n]
    int16_t *out;           // [n]
    const int *rom32;       // window ROM widened to int32
    CascadeCoef coef;
    int16_t *entry;         // [P][24] state at each chunk's first sample
    int16_t *exit_;         // [P][24] state after each chunk's last sample
    const int16_t *state0;  // [24] true state before sample 0 (used when continuous)
    int *stats;             // [0] chunks whose entry state differs from the neighbour's exit, [1] max |difference| (LSB)
    float *ends;            // [P][24] zero-state responses e_p, then predicted states s_p (in place)
    float *aggr;            // [4][ceil(P/32)][24]: per-warp aggregates G_w, carries C_w, and the aggregate scan's two buffers
    const float *mats;      // [kScanMats][24][24] row-major: M^(1,2,4,8,16), then Q^(2^j) with Q = M^32
    int aggr_levels;        // levels of the aggregate scan that matter (Q^(2^j) not yet negligible), <= kAggrLevels
    unsigned long long n;   // samples
    int chunk;              // samples per chunk L (multiple of 8)
    int n_chunks;
    int n_warps;            // ceil(n_chunks / 32)
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

// the affine float model of one biquad step: products are not truncated, each non-zero
// product carries the mean of its truncation (-1/2 LSB; +1/2 for the subtracted terms)
FRA_DEV float biquad_linear(float x, const StageCoef &k, StageState &s, float bias)
{
    float y = bias;
    y = fmaf(s.x2, k.b0, y);
    y = fmaf(s.x1, k.b1, y);
    y = fmaf(s.y2, k.na0, y);
    y = fmaf(x, k.b2, y);
    y = fmaf(s.y1, k.na1, y);
    s.x2 = s.x1; s.x1 = x;
    s.y2 = s.y1; s.y1 = y;
    return y;
}

FRA_DEV float stage_bias(const StageCoef &k)
{
    // floor() lowers each added product by 1/2 on average and raises each subtracted one
    return -0.5f * ((k.b0 != 0.0f) + (k.b1 != 0.0f) + (k.b2 != 0.0f)) + 0.5f * ((k.na0 != 0.0f) + (k.na1 != 0.0f));
}

// samples [begin, end) (multiples of 8) through window + cascade.
// MODE 0: exact, no output; 1: exact, write output; 2: float model, no output
template <int MODE>
FRA_DEV void stream_run(const K1bArgs &a, StageState (&st)[kStages], unsigned long long begin, unsigned long long end)
{
    float bias[kStages];
#pragma unroll
    for (int s = 0; s < kStages; ++s) bias[s] = stage_bias(a.coef.set[s]);
    for (unsigned long long n0 = begin; n0 < end; n0 += 8) {
        const uint4 xv = ldg128(a.in + n0);
        const unsigned xw[4] = {xv.x, xv.y, xv.z, xv.w};
        const int wbase = (int)(n0 & (unsigned long long)(kWindowLen - 1));
        // the chunk's lanes sit at different window phases: eight ROM entries per lane and trip as two
        // 16-byte loads (one scalar load per sample made the L1 the bottleneck of the small-chunk scan)
        const int4 ra = __ldg(reinterpret_cast<const int4 *>(a.rom32 + wbase));
        const int4 rb = __ldg(reinterpret_cast<const int4 *>(a.rom32 + wbase + 4));
        const int rom8[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
        unsigned ow[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float acc2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int x = e ? hi16(xw[h]) : lo16(xw[h]);
                if (a.apply_window) x = window_int(x, rom8[2 * h + e]);
                float v = small_int_to_float(x);
                float acc = __int_as_float((x + 32768) + 0x4B400000);     // offset-binary, as biquad_step leaves it
                if (a.iir) {
                    if (MODE == 2) {
#pragma unroll
                        for (int s = 0; s < kStages; ++s) v = biquad_linear(v, a.coef.set[s], st[s], bias[s]);
                    } else {
#pragma unroll
                        for (int s = 0; s < kStages; ++s) acc = biquad_step(v, a.coef.set[s], st[s], &v);
                    }
                }
                acc2[e] = acc;
            }
            ow[h] = pack16_acc(acc2[0], acc2[1]);
        }
        if (MODE == 1) stg128(a.out + n0, make_uint4(ow[0], ow[1], ow[2], ow[3]));
    }
}

// the scan is over the chunk boundaries b_p = p L: segment p is [b_p, b_{p+1})
FRA_DEV unsigned long long scan_boundary(const K1bArgs &a, int p)
{
    unsigned long long b = (unsigned long long)p * (unsigned long long)a.chunk;
    return b < a.n ? b : a.n;
}

// 1. zero-state response of segment p in the float model
__global__ void __launch_bounds__(64) k1b_lin_ends(K1bArgs a)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_chunks) return;
    StageState st[kStages];
    state_zero(st);
    stream_run<2>(a, st, scan_boundary(a, p), scan_boundary(a, p + 1));
    float *e = a.ends + (size_t)p * kStateDim;
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
        e[4 * s + 0] = st[s].x1; e[4 * s + 1] = st[s].x2; e[4 * s + 2] = st[s].y1; e[4 * s + 3] = st[s].y2;
    }
}

// 2. Hillis-Steele scan of  s_{p+1} = M s_p + e_p  (M = A^L) inside one warp: 32 boundaries,
// the neighbour's partial sum arriving by __shfl_up, M^(2^j) broadcast from shared memory.
// With a_0 = s_0 and a_p = e_{p-1}:  s_p = sum_{q <= p} M^(p-q) a_q.  Lane i of warp w ends up
// with the sum over q = 32 w .. 32 w + i.  PHASE 0 (no carry) keeps only the warp's aggregate
// G_w (lane 31); PHASE 1 first adds the carry C_w = M s_{32 w - 1} at lane 0, so that every
// lane holds the full state s_p, and writes it, rounded, as chunk p's predicted entry state.
template <int PHASE>
__global__ void __launch_bounds__(32) k1b_scan_warp(K1bArgs a)
{
    __shared__ float sm[kScanLevels][kStateDim * kStateDim];
    const int lane = threadIdx.x;
    const int w = blockIdx.x;
    for (int i = lane; i < kScanLevels * kStateDim * kStateDim; i += 32) (&sm[0][0])[i] = __ldg(a.mats + i);
    __syncwarp();
    const int p = 32 * w + lane;
    const bool live = p < a.n_chunks;
    float v[kStateDim];
#pragma unroll
    for (int i = 0; i < kStateDim; ++i) v[i] = 0.0f;
    if (live && p > 0) {
        const float *e = a.ends + (size_t)(p - 1) * kStateDim;
#pragma unroll
        for (int i = 0; i < kStateDim; ++i) v[i] = e[i];
    }
    if (lane == 0) {
        if (w == 0) {
            if (a.continuous) {
#pragma unroll
                for (int i = 0; i < kStateDim; ++i) v[i] = (float)a.state0[i];     // the true state before sample 0
            }
        } else if (PHASE == 1) {
            const float *c = a.aggr + (size_t)(a.n_warps + w) * kStateDim;
#pragma unroll
            for (int i = 0; i < kStateDim; ++i) v[i] += c[i];
        }
    }
#pragma unroll 1
    for (int lvl = 0; lvl < kScanLevels; ++lvl) {
        const int d = 1 << lvl;
        float t[kStateDim];
#pragma unroll
        for (int i = 0; i < kStateDim; ++i) t[i] = __shfl_up_sync(0xffffffffu, v[i], d);
        if (lane >= d) {
            const float *m = sm[lvl];
#pragma unroll
            for (int r = 0; r < kStateDim; ++r) {
                float acc = v[r];
#pragma unroll
                for (int c = 0; c < kStateDim; ++c) acc = fmaf(m[r * kStateDim + c], t[c], acc);
                v[r] = acc;
            }
        }
    }
    if (PHASE == 0) {
        if (lane == 31) {                              // only full warps hand an aggregate on
            float *g = a.aggr + (size_t)w * kStateDim;
#pragma unroll
            for (int i = 0; i < kStateDim; ++i) g[i] = v[i];
        }
    } else if (live && p > 0) {
        int16_t *o = a.entry + (size_t)p * kStateDim;
#pragma unroll
        for (int i = 0; i < kStateDim; ++i) {
            const float r = fminf(fmaxf(rintf(v[i]), -32768.0f), 32767.0f);
            o[i] = (int16_t)(int)r;
        }
    }
}

// The warps' aggregates chained:  S_w = G_w + Q S_{w-1}  (Q = M^32; S_w = state at boundary 32 w + 31).
// Itself a scan - Hillis-Steele over the array of aggregates, one launch per level, one warp per
// element (24 lanes = the 24 rows of the matrix-vector product), ping-pong between two buffers:
//   level j:  X'[w] = X[w] + Q^(2^j) X[w - 2^j]   (w >= 2^j)
// The host stops at the first level whose matrix is negligible (a.aggr_levels): for a stable
// cascade and chunks of a few hundred samples Q = A^(32 L) is already zero to fp32 precision.
__global__ void __launch_bounds__(128) k1b_aggr_level(K1bArgs a, int level, const float *src, float *dst)
{
    const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w + 1 >= a.n_warps || lane >= kStateDim) return;       // the last (possibly partial) warp hands nothing on
    const int d = 1 << level;
    float acc = src[(size_t)w * kStateDim + lane];
    if (w >= d) {
        const float *m = a.mats + (size_t)(kScanLevels + level) * kStateDim * kStateDim + lane * kStateDim;
        const float *x = src + (size_t)(w - d) * kStateDim;
#pragma unroll
        for (int c = 0; c < kStateDim; ++c) acc = fmaf(__ldg(m + c), x[c], acc);
    }
    dst[(size_t)w * kStateDim + lane] = acc;
}

// carries  C_{w+1} = M S_w  (what lane 0 of warp w + 1 adds before its own scan)
__global__ void __launch_bounds__(128) k1b_aggr_carry(K1bArgs a, const float *s_final)
{
    const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w + 1 >= a.n_warps || lane >= kStateDim) return;
    const float *m = a.mats + lane * kStateDim;                // M^1
    const float *x = s_final + (size_t)w * kStateDim;
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < kStateDim; ++c) acc = fmaf(__ldg(m + c), x[c], acc);
    a.aggr[(size_t)(a.n_warps + w + 1) * kStateDim + lane] = acc;
}

// 3. exact filtering of chunk p from the scan's predicted state (chunk 0: the true state)
__global__ void __launch_bounds__(64) k1b_speculate(K1bArgs a)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_chunks) return;
    StageState st[kStages];
    if (p == 0) {
        if (a.continuous) state_load(st, a.state0);
        else state_zero(st);
        state_store(st, a.entry);
    } else {
        state_load(st, a.entry + (size_t)p * kStateDim);
    }
    stream_run<1>(a, st, scan_boundary(a, p), scan_boundary(a, p + 1));
    state_store(st, a.exit_ + (size_t)p * kStateDim);
}

// 4. verify: lane p compares its entry state with lane p-1's exit state; the exit
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
