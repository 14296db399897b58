// k1_window_iir.cuh - K1: fused window ROM multiply + 12th-order IIR (six
// cascaded int16 x int8 biquads), bit-exact to the VHDL.
//
// Replaces hann_window (NEW/hann8192.vhd:28-47), filter_iir12 / filter_iir12_cust
// (IMP/filter_iir12.vhd:38-137, NEW/filter_iir12_cust.vhd:68-240) and the stream
// mux of command_control (NEW/command_control.vhd:90-116).
//
// Two kernels, same arithmetic (fra_common.cuh):
//   k1_lane   one warp lane per channel, all six stages in the lane's registers.
//             ~50 issue slots per sample; needs >= ~19k channels to fill 148 SMs.
//   k1_split  one warp lane per (channel, stage): a warp is five 6-lane systolic
//             chains, stage s works on sample i-s, outputs pass to the next lane
//             by __shfl_up.  6x the parallelism for small channel counts (the
//             4096-channel configuration), ~65 issue slots per sample.
// Both read int16 [C][N] channel-major, write int16 [C][N], and carry the
// per-stage history (x[n-1], x[n-2], y[n-1], y[n-2]) in state[C][6][4].
#pragma once
#include "fra_common.cuh"

namespace fra {

struct K1Args {
    const int16_t *in;      // [C][N]
    int16_t *out;           // [C][N]
    int16_t *state;         // [C][6][4], read when continuous, always written
    const int *rom32;       // [16384] window ROM widened to int32
    CascadeCoef coef;
    int channels;
    int n;                  // samples per frame, multiple of 256
    int continuous;
};

// ------------------------------------------------------------------ k1_lane
constexpr int kLaneBlock = 128;

__global__ void __launch_bounds__(kLaneBlock) k1_lane(K1Args a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.channels) return;

    StageState st[kStages];
    {
        const uint2 *sp = reinterpret_cast<const uint2 *>(a.state + (size_t)c * 24);
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            uint2 v = a.continuous ? __ldg(sp + s) : make_uint2(0u, 0u);
            st[s].x1 = small_int_to_float(lo16(v.x));
            st[s].x2 = small_int_to_float(hi16(v.x));
            st[s].y1 = small_int_to_float(lo16(v.y));
            st[s].y2 = small_int_to_float(hi16(v.y));
        }
    }

    const int16_t *src = a.in + (size_t)c * a.n;
    int16_t *dst = a.out + (size_t)c * a.n;

    for (int n0 = 0; n0 < a.n; n0 += 16) {
        uint4 xa = ldg128(src + n0);
        uint4 xb = ldg128(src + n0 + 8);
        const unsigned xw[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const int4 *rp = reinterpret_cast<const int4 *>(a.rom32 + (n0 & (kWindowLen - 1)));
        unsigned ow[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int4 r = __ldg(rp + q);                       // warp-uniform address: one L1 broadcast
            const int rom[4] = {r.x, r.y, r.z, r.w};
            float acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                unsigned w = xw[2 * q + (j >> 1)];
                int x = (j & 1) ? hi16(w) : lo16(w);
                float v = small_int_to_float(window_int(x, rom[j]));
#pragma unroll
                for (int s = 0; s < kStages; ++s)
                    acc[j] = biquad_step(v, a.coef.set[s & 1], st[s], &v);
            }
            ow[2 * q] = pack16_acc(acc[0], acc[1]);
            ow[2 * q + 1] = pack16_acc(acc[2], acc[3]);
        }
        stg128(dst + n0, make_uint4(ow[0], ow[1], ow[2], ow[3]));
        stg128(dst + n0 + 8, make_uint4(ow[4], ow[5], ow[6], ow[7]));
    }

    {
        uint2 *sp = reinterpret_cast<uint2 *>(a.state + (size_t)c * 24);
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            uint2 v;
            v.x = pack16((unsigned)(int)st[s].x1, (unsigned)(int)st[s].x2);
            v.y = pack16((unsigned)(int)st[s].y1, (unsigned)(int)st[s].y2);
            sp[s] = v;
        }
    }
}

// ----------------------------------------------------------------- k1_split
constexpr int kSplitGroups = 5;      // channels per warp (5 x 6 = 30 lanes, 2 idle)
constexpr int kSplitChunk = 256;     // samples staged per chunk
constexpr int kSplitWarps = 4;       // warps per CTA
constexpr int kSplitRing = 2 * kSplitChunk;
constexpr int kSplitSmemPerWarp =
    kSplitGroups * kSplitChunk * (int)sizeof(float) + kSplitGroups * kSplitRing * (int)sizeof(int16_t);

struct SplitLane {
    StageCoef k;
    StageState st;
    float y;            // this lane's latest output, read by lane+1 next iteration
    unsigned carry[3];  // accumulator bits of the 3 newest outputs not yet stored
    int s;              // stage 0..5
    bool first, last;
};

// Eight systolic iterations i_base .. i_base+7.  Lane (g, s) processes sample
// i - s at iteration i.  GUARD = true for the blocks that contain samples outside
// [0, n): the first block of a frame (stages still empty) and the flush block.
template <bool GUARD>
FRA_DEV void split_block8(SplitLane &L, int i_base, int n, const float *win8, int16_t *ring_row, bool store_ok)
{
    float w[8];
    if (win8 != nullptr) {
        const float4 w0 = *reinterpret_cast<const float4 *>(win8);
        const float4 w1 = *reinterpret_cast<const float4 *>(win8 + 4);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
        w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = 0.0f;
    }
    unsigned ob[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float up = __shfl_up_sync(0xffffffffu, L.y, 1);
        float x = L.first ? w[j] : up;
        if (GUARD) {
            StageState keep = L.st;
            float ykeep = L.y;
            float y;
            float acc = biquad_step(x, L.k, L.st, &y);
            bool active = (unsigned)(i_base + j - L.s) < (unsigned)n;
            if (active) {
                L.y = y;
            } else {
                L.st = keep;
                L.y = ykeep;
            }
            ob[j] = __float_as_uint(acc);
        } else {
            float acc = biquad_step(x, L.k, L.st, &L.y);
            ob[j] = __float_as_uint(acc);
        }
    }
    // the last-stage lane emitted samples i_base-5 .. i_base+2; together with the
    // carry (i_base-8 .. i_base-6) that completes the aligned group [i_base-8, i_base-1]
    uint4 o;
    o.x = __byte_perm(L.carry[0], L.carry[1], 0x5410) ^ 0x80008000u;
    o.y = __byte_perm(L.carry[2], ob[0], 0x5410) ^ 0x80008000u;
    o.z = __byte_perm(ob[1], ob[2], 0x5410) ^ 0x80008000u;
    o.w = __byte_perm(ob[3], ob[4], 0x5410) ^ 0x80008000u;
    L.carry[0] = ob[5]; L.carry[1] = ob[6]; L.carry[2] = ob[7];
    if (store_ok && L.last && i_base >= 8)
        *reinterpret_cast<uint4 *>(ring_row + ((i_base - 8) & (kSplitRing - 1))) = o;
}

__global__ void __launch_bounds__(kSplitWarps * 32) k1_split(K1Args a)
{
    FRA_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    const int warp_global = blockIdx.x * kSplitWarps + warp_in_cta;
    const int c_base = warp_global * kSplitGroups;
    if (c_base >= a.channels) return;                       // warp-uniform; no block-wide barrier is used
    const int n_ch = min(kSplitGroups, a.channels - c_base);

    float *win = reinterpret_cast<float *>(smem_raw + (size_t)warp_in_cta * kSplitSmemPerWarp);   // [5][256]
    int16_t *ring = reinterpret_cast<int16_t *>(win + kSplitGroups * kSplitChunk);                 // [5][512]

    const int g_raw = lane / kStages;
    const int g = min(g_raw, kSplitGroups - 1);
    const bool lane_valid = g_raw < n_ch;
    SplitLane L;
    L.s = lane - g_raw * kStages;
    L.first = (L.s == 0);
    L.last = (L.s == kStages - 1);
    L.k = a.coef.set[L.s & 1];
    L.y = 0.0f;
    L.carry[0] = L.carry[1] = L.carry[2] = 0u;
    {
        uint2 v = make_uint2(0u, 0u);
        if (a.continuous && lane_valid)
            v = __ldg(reinterpret_cast<const uint2 *>(a.state + ((size_t)(c_base + g) * kStages + L.s) * 4));
        L.st.x1 = small_int_to_float(lo16(v.x));
        L.st.x2 = small_int_to_float(hi16(v.x));
        L.st.y1 = small_int_to_float(lo16(v.y));
        L.st.y2 = small_int_to_float(hi16(v.y));
    }

    // staging map: lane handles samples [4*lane, +4) and [128 + 4*lane, +4) of every row
    uint2 pre[kSplitGroups][2];
    auto prefetch = [&](int i0) {
#pragma unroll
        for (int r = 0; r < kSplitGroups; ++r) {
            pre[r][0] = make_uint2(0u, 0u);
            pre[r][1] = make_uint2(0u, 0u);
            if (r < n_ch) {
                const int16_t *row = a.in + (size_t)(c_base + r) * a.n + i0;
                pre[r][0] = __ldg(reinterpret_cast<const uint2 *>(row + 4 * lane));
                pre[r][1] = __ldg(reinterpret_cast<const uint2 *>(row + 128 + 4 * lane));
            }
        }
    };
    auto convert = [&](int i0) {
        const int w0 = (i0 + 4 * lane) & (kWindowLen - 1);
        const int w1 = (i0 + 128 + 4 * lane) & (kWindowLen - 1);
        const int4 ra = __ldg(reinterpret_cast<const int4 *>(a.rom32 + w0));
        const int4 rb = __ldg(reinterpret_cast<const int4 *>(a.rom32 + w1));
#pragma unroll
        for (int r = 0; r < kSplitGroups; ++r) {
            float4 fa, fb;
            fa.x = small_int_to_float(window_int(lo16(pre[r][0].x), ra.x));
            fa.y = small_int_to_float(window_int(hi16(pre[r][0].x), ra.y));
            fa.z = small_int_to_float(window_int(lo16(pre[r][0].y), ra.z));
            fa.w = small_int_to_float(window_int(hi16(pre[r][0].y), ra.w));
            fb.x = small_int_to_float(window_int(lo16(pre[r][1].x), rb.x));
            fb.y = small_int_to_float(window_int(hi16(pre[r][1].x), rb.y));
            fb.z = small_int_to_float(window_int(lo16(pre[r][1].y), rb.z));
            fb.w = small_int_to_float(window_int(hi16(pre[r][1].y), rb.w));
            *reinterpret_cast<float4 *>(win + r * kSplitChunk + 4 * lane) = fa;
            *reinterpret_cast<float4 *>(win + r * kSplitChunk + 128 + 4 * lane) = fb;
        }
    };
    // store the finished output group [m0, m0 + 256) from the ring (m0 may be -8)
    auto flush = [&](int m0, int count8) {
#pragma unroll
        for (int r = 0; r < kSplitGroups; ++r) {
            const int m = m0 + 8 * lane;
            if (r < n_ch && lane < count8 && m >= 0) {
                uint4 v = *reinterpret_cast<const uint4 *>(ring + r * kSplitRing + (m & (kSplitRing - 1)));
                stg128(a.out + (size_t)(c_base + r) * a.n + m, v);
            }
        }
    };

    prefetch(0);
    convert(0);
    __syncwarp();
    const float *win_row = win + g * kSplitChunk;
    int16_t *ring_row = ring + g * kSplitRing;

    for (int i0 = 0; i0 < a.n; i0 += kSplitChunk) {
        const bool more = (i0 + kSplitChunk) < a.n;
        if (more) prefetch(i0 + kSplitChunk);
        if (i0 == 0) {
            split_block8<true>(L, 0, a.n, win_row, ring_row, lane_valid);
            for (int b = 8; b < kSplitChunk; b += 8)
                split_block8<false>(L, b, a.n, win_row + b, ring_row, lane_valid);
        } else {
            for (int b = 0; b < kSplitChunk; b += 8)
                split_block8<false>(L, i0 + b, a.n, win_row + b, ring_row, lane_valid);
        }
        __syncwarp();
        flush(i0 - 8, kSplitChunk / 8);
        if (more) convert(i0 + kSplitChunk);
        __syncwarp();
    }
    // flush block: iterations n .. n+7 drain stages 1..5 and complete group [n-8, n-1]
    split_block8<true>(L, a.n, a.n, nullptr, ring_row, lane_valid);
    __syncwarp();
    flush(a.n - 8, 1);

    if (lane_valid) {
        uint2 v;
        v.x = pack16((unsigned)(int)L.st.x1, (unsigned)(int)L.st.x2);
        v.y = pack16((unsigned)(int)L.st.y1, (unsigned)(int)L.st.y2);
        *reinterpret_cast<uint2 *>(a.state + ((size_t)(c_base + g) * kStages + L.s) * 4) = v;
    }
}

// ------------------------------------------------------------ window only
// Bypass mode with the FFT input stream requested as an output: the window
// alone (in bypass the FFT kernel applies the window itself while loading).
__global__ void __launch_bounds__(256) k1_window_only(const int16_t *in, int16_t *out, const int *rom32,
                                                      size_t total8, int n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total8) return;
    const size_t e = i * 8;
    const int w0 = (int)(e % (size_t)n) & (kWindowLen - 1);
    uint4 x = ldg128(in + e);
    int4 ra = __ldg(reinterpret_cast<const int4 *>(rom32 + w0));
    int4 rb = __ldg(reinterpret_cast<const int4 *>(rom32 + w0 + 4));
    uint4 o;
    o.x = pack16((unsigned)window_int(lo16(x.x), ra.x), (unsigned)window_int(hi16(x.x), ra.y));
    o.y = pack16((unsigned)window_int(lo16(x.y), ra.z), (unsigned)window_int(hi16(x.y), ra.w));
    o.z = pack16((unsigned)window_int(lo16(x.z), rb.x), (unsigned)window_int(hi16(x.z), rb.y));
    o.w = pack16((unsigned)window_int(lo16(x.w), rb.z), (unsigned)window_int(hi16(x.w), rb.w));
    stg128(out + e, o);
}

}  // namespace fra
