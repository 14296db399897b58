// k1_window_iir.cuh - K1: fused window ROM multiply + 12th-order IIR (six
// cascaded int16 x int8 biquads), bit-exact to the VHDL.
//
// Replaces hann_window (NEW/hann8192.vhd:28-47), filter_iir12 / filter_iir12_cust
// (IMP/filter_iir12.vhd:38-137, NEW/filter_iir12_cust.vhd:68-240) and the stream
// mux of command_control (NEW/command_control.vhd:90-116).
//
// Four kernels, same arithmetic (fra_common.cuh):
//   k1_duo          a warp per PAIR of stages, one lane per channel, 64-sample chunks handed from warp to warp
//                   through shared memory (loader, three stage pairs, writer): the kernel for channel counts
//                   that cannot fill the machine with a lane per channel (up to ~24k channels).
//   k1_lane_biased  one lane per channel, all six stages in the lane's registers as a SKEWED cascade (stage s on
//                   sample i - s: six independent chains per warp), global memory in whole 128-byte lines through
//                   a per-warp staging buffer, outputs in place: the kernel from ~24k channels up, for coefficient
//                   sets that allow the all-biased step (the reference's fixed bank does).
//   k1_lane         the general-step lane kernel (any int8 coefficients), sample-major, per-lane 32-byte fetches.
//   k1_split        one lane per (channel, stage): a warp is five 6-lane systolic chains, stage s works on sample
//                   i-4s, outputs pass to the next lane by __shfl_up three iterations ahead of their use; input rows
//                   arrive by cp.async.bulk (TMA 1-D bulk copies on an mbarrier).  The EXACT single-stream path
//                   (fra_iir_stream(exact = 1)) and on request.
// All read int16 [C][N] channel-major, write int16 [C][N], and carry the
// per-stage history (x[n-1], x[n-2], y[n-1], y[n-2]) in state[C][6][4].
#pragma once
#include "fra_common.cuh"

namespace fra {

struct K1Args {
    const int16_t *in;      // [C][N]
    int16_t *out;           // [C][N]
    int16_t *state;         // [C][6][4], read when continuous, always written
    const int *rom32;       // [16384] window ROM widened to int32
    const int *rom2x;       // [16384] 2 * ROM (window_biased)
    CascadeCoef coef;
    int channels;
    int n;                  // samples per frame, multiple of 256
    int continuous;
    int speculate;          // k1_split: try the no-overflow recurrence first (rolled back when it was wrong)
#ifdef FRA_TIMELINE
    int tl_step;
#endif
};

// ------------------------------------------------------------------ k1_lane
constexpr int kLaneBlock = 32;       // one warp per CTA: 2048 warps spread evenly over 148 SMs x 4 schedulers

template <bool B1Z>
__global__ void __launch_bounds__(kLaneBlock) k1_lane(K1Args a)
{
    const int c_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = c_raw < a.channels;
    const int c = live ? c_raw : a.channels - 1;     // inactive lanes shadow the last channel (no stores)

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

    // The next 16 samples (and their 16 window ROM entries) are requested one trip ahead by
    // cp.async into shared memory: a register prefetch would share one of the warp's six
    // scoreboards with the trip's other loads and stall on them (measured).
    __shared__ uint4 stage_x[2][2][kLaneBlock];      // [buffer][half][lane]: 16 samples per lane
    __shared__ int4 stage_rom[2][4];                 // [buffer][4 x 4 ROM entries]
    const int lane = threadIdx.x;
    auto request = [&](int n0, int b) {
        cp_async16(&stage_x[b][0][lane], src + n0);
        cp_async16(&stage_x[b][1][lane], src + n0 + 8);
        if (lane < 4) cp_async16(&stage_rom[b][lane], a.rom32 + ((n0 + 4 * lane) & (kWindowLen - 1)));
        cp_async_commit();
    };
    request(0, 0);
    for (int n0 = 0, it = 0; n0 < a.n; n0 += 16, ++it) {
        const int b = it & 1;
        if (n0 + 16 < a.n) request(n0 + 16, b ^ 1);
        else cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();                                 // the ROM words were copied by lanes 0..3
        const uint4 xa = stage_x[b][0][lane], xb = stage_x[b][1][lane];
        const unsigned xw[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        unsigned ow[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int4 r = stage_rom[b][q];           // same address in every lane: broadcast
            const int rom[4] = {r.x, r.y, r.z, r.w};
            float acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                unsigned w = xw[2 * q + (j >> 1)];
                int x = (j & 1) ? hi16(w) : lo16(w);
                float v = small_int_to_float(window_int(x, rom[j]));
#pragma unroll
                for (int s = 0; s < kStages; ++s)
                    acc[j] = biquad_step<B1Z>(v, a.coef.set[s], st[s], &v);
            }
            ow[2 * q] = pack16_acc(acc[0], acc[1]);
            ow[2 * q + 1] = pack16_acc(acc[2], acc[3]);
        }
        if (live) {
            stg128(dst + n0, make_uint4(ow[0], ow[1], ow[2], ow[3]));
            stg128(dst + n0 + 8, make_uint4(ow[4], ow[5], ow[6], ow[7]));
        }
        __syncwarp();                                 // everyone has read buffer b before it is refilled
    }

    if (live) {
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

// byte offset of 16-byte piece `piece` of 128-byte line `row` in a staged [row][8 pieces] tile: the piece index is
// XOR-swizzled with (row & 7), so line-wise access (eight lanes per row) and lane-per-row access are both conflict-free
FRA_DEV int duo_swz(int row, int piece) { return row * 128 + ((piece ^ (row & 7)) << 4); }

// ------------------------------------------------------------- k1_lane_biased
// The lane-per-channel kernel on the all-biased step (fra_common.cuh: biquad_step_biased): every
// value that moves between the window, the six stages and the history registers is the PRMT
// output v + kBias16, so a stage is five FFMAs (four when the x[n-1] coefficient is zero) and one
// PRMT, the window is IMAD + PRMT, and nothing on the path converts or removes a bias.
// ~35 issue slots per sample with the reference's fixed bank.
FRA_DEV float biased_from_i16(unsigned half_word_lo16)      // int16 in the low half of a word -> v + kBias16
{
    return __uint_as_float(((half_word_lo16 ^ 0x8000u) & 0xFFFFu) | 0x4B000000u);
}

FRA_DEV StageStateB load_state_biased(const int16_t *state, size_t c, int s, bool continuous)
{
    const uint2 v = continuous ? __ldg(reinterpret_cast<const uint2 *>(state + (c * kStages + s) * 4)) : make_uint2(0u, 0u);
    StageStateB st;
    st.x1 = biased_from_i16(v.x);
    st.x2 = biased_from_i16(v.x >> 16);
    st.y1 = biased_from_i16(v.y);
    st.y2 = biased_from_i16(v.y >> 16);
    return st;
}

FRA_DEV void store_state_biased(int16_t *state, size_t c, int s, const StageStateB &st)
{
    uint2 v;
    v.x = pack16(__float_as_uint(st.x1), __float_as_uint(st.x2)) ^ 0x80008000u;
    v.y = pack16(__float_as_uint(st.y1), __float_as_uint(st.y2)) ^ 0x80008000u;
    *reinterpret_cast<uint2 *>(state + (c * kStages + s) * 4) = v;
}

// window + bias for 8 packed samples whose 8 doubled ROM entries are ra, rb; QUIRK: the entries may
// be -32768 (2c = -65536), where the product overflows and the result needs the resize fix-up
template <bool QUIRK>
FRA_DEV void window8_biased(uint4 x, int4 ra, int4 rb, unsigned exp23, float (&u)[8])
{
    auto w = [&](int v, int c2) {
        return QUIRK ? biased_from_int(window_int(v, c2 >> 1)) : window_biased(v, c2, exp23);
    };
    u[0] = w(lo16(x.x), ra.x); u[1] = w(hi16(x.x), ra.y); u[2] = w(lo16(x.y), ra.z); u[3] = w(hi16(x.y), ra.w);
    u[4] = w(lo16(x.z), rb.x); u[5] = w(hi16(x.z), rb.y); u[6] = w(lo16(x.w), rb.z); u[7] = w(hi16(x.w), rb.w);
}

// ROM entries equal to -32768 live in [0, 15), [8178, 8206) and [16369, 16384) - two entries later in the table
// rotated for FRA_WINDOW_RTL_SKEW, so the ranges are taken two wider: does [w0, w0 + len) touch one?
FRA_DEV bool rom_quirk_range(int w0, int len)
{
    return (w0 < 17) || (w0 + len > 8178 && w0 < 8208) || (w0 + len > 16369);
}

// One iteration of the skewed cascade: stage s works on sample i - s, so the six stage steps of an
// iteration are independent of each other (each consumes what its predecessor produced in the
// PREVIOUS iteration) and a single in-order warp has six dependency chains to interleave instead of
// one chain through all six stages (75 cycles per sample measured for the sample-major order; the
// critical path here is one stage step).  Stages are visited last to first so that pipe[s - 1]
// still holds the previous iteration's value.  Returns the last stage's accumulator (sample i - 5).
// ALT: the six stages use two alternating coefficient sets (the RTL's bank layout): a dozen values that
// stay in uniform registers, instead of 36 of which most are re-read from the constant bank every trip
template <bool B1Z, bool ALT>
FRA_DEV float lane_skewed_iteration(float ux, const CascadeCoef &coef, StageStateB (&st)[kStages], float (&pipe)[kStages])
{
    float acc = 0.0f;
#pragma unroll
    for (int s = kStages - 1; s >= 0; --s) {
        const float in = (s == 0) ? ux : pipe[s - 1];
        const float r = biquad_step_biased<B1Z>(in, coef.set[ALT ? (s & 1) : s], st[s], &pipe[s]);
        if (s == kStages - 1) acc = r;
    }
    return acc;
}

constexpr int kLaneLag = kStages - 1;      // the last stage emits sample i - 5 at iteration i

// One trip of 16 iterations i0 .. i0 + 15.  FIRST (i0 = 0): stage s sees its first real sample at
// iteration s; what it computed before that is junk, so its history is (re)loaded right there - no
// guard anywhere.  The accumulators of samples i0 - 5 .. i0 + 10 come out; with the three carried
// from the trip before (i0 - 8 .. i0 - 6) they complete the 16-byte groups [i0 - 8, i0) and
// [i0, i0 + 8); the last three are carried on.
template <bool B1Z, bool ALT, bool FIRST>
FRA_DEV void lane_trip16(const K1Args &a, size_t c, const float (&u)[16], StageStateB (&st)[kStages], float (&pipe)[kStages],
                         float (&carry)[3], uint4 &ga, uint4 &gb)
{
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (FIRST && j < kStages) st[j] = load_state_biased(a.state, c, j, a.continuous != 0);
        acc[j] = lane_skewed_iteration<B1Z, ALT>(u[j], a.coef, st, pipe);
    }
    ga.x = pack16_acc(carry[0], carry[1]);
    ga.y = pack16_acc(carry[2], acc[0]);
    ga.z = pack16_acc(acc[1], acc[2]);
    ga.w = pack16_acc(acc[3], acc[4]);
    gb.x = pack16_acc(acc[5], acc[6]);
    gb.y = pack16_acc(acc[7], acc[8]);
    gb.z = pack16_acc(acc[9], acc[10]);
    gb.w = pack16_acc(acc[11], acc[12]);
    carry[0] = acc[13]; carry[1] = acc[14]; carry[2] = acc[15];
}

// Four independent warps per CTA (no CTA-wide barrier), one lane per channel, and global memory touched in
// whole 128-byte lines only, like k1_duo: a chunk of 64 samples of one channel IS one line, eight lanes move
// it (cp.async 16 B in, STG.128 out), 32 lines = one 4 KiB buffer per warp, piece index XOR-swizzled with
// (row & 7) so that both the line-wise and the lane-per-channel side are conflict-free.  (The first version
// had every lane fetch its own 32 bytes per 16 samples: one sector per lane and request, DRAM pages opened
// for 32 bytes each - the SM's instruction rate stopped rising at ~1.9 per cycle however many warps were
// resident, profiles/r02_k1_lane_16384ch.txt.)  Outputs are written IN PLACE into the buffer the samples came
// from (a lane only ever touches its own row there), and a chunk's lines are stored one 16-sample trip into
// the next chunk, when the cascade's five-sample lag has delivered its last samples; the buffer is then
// refilled with the chunk after next.  Two buffers per warp: 8.5 KiB, 34 KiB per CTA.
constexpr int kLaneBiasedBlock = 128;
constexpr int kLaneChunk = 64;                                  // samples per staged line
constexpr int kLaneBufBytes = 32 * 128;                         // one chunk of 32 channels
constexpr int kLaneWarpBytes = 2 * kLaneBufBytes + 2 * kLaneChunk * 4;      // two line buffers + two ROM slices
constexpr int kLaneBiasedSmem = (kLaneBiasedBlock / 32) * kLaneWarpBytes;

template <bool B1Z, bool ALT>
__global__ void __launch_bounds__(kLaneBiasedBlock) k1_lane_biased(K1Args a)
{
    FRA_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * (kLaneBiasedBlock / 32) + wrp) * 32;
    if (c0 >= a.channels) return;                                  // warp-uniform; no CTA-wide barrier below
    const bool live = c0 + lane < a.channels;
    const size_t c = (size_t)(live ? c0 + lane : a.channels - 1); // inactive lanes shadow the last channel (no stores)
    const unsigned exp23 = a.coef.set[0].exp23;
    unsigned char *lines = smem_raw + wrp * kLaneWarpBytes;        // [2][32 rows][8 pieces of 16 B]
    int *roms = reinterpret_cast<int *>(lines + 2 * kLaneBufBytes);      // [2][64] doubled ROM entries
    const int n_chunks = a.n / kLaneChunk;

    StageStateB st[kStages];
    float pipe[kStages], carry[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
        st[s].x1 = st[s].x2 = st[s].y1 = st[s].y2 = kBias16;      // junk phase of the first trip: any finite value
        pipe[s] = kBias16;
    }

    // chunk t -> buffer t & 1, whole lines: eight lanes per channel row, four rows per instruction
    auto request = [&](int t) {
        unsigned char *dst = lines + (t & 1) * kLaneBufBytes;
        const int p = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + (lane >> 3);
            const int ch = min(c0 + r, a.channels - 1);                       // rows past the end read valid memory
            cp_async16(dst + duo_swz(r, p), a.in + (size_t)ch * a.n + (size_t)t * kLaneChunk + 8 * p);
        }
        if (lane < 16)
            cp_async16(roms + (t & 1) * kLaneChunk + 4 * lane, a.rom2x + (((t * kLaneChunk) & (kWindowLen - 1)) + 4 * lane));
        cp_async_commit();
    };
    // the finished chunk t: buffer -> global memory, whole lines
    auto store_lines = [&](int t) {
        const unsigned char *src = lines + (t & 1) * kLaneBufBytes;
        const int p = lane & 7;
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const uint4 *>(src + duo_swz(4 * i + (lane >> 3), p));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ch = c0 + 4 * i + (lane >> 3);
            if (ch < a.channels) stg128(a.out + (size_t)ch * a.n + (size_t)t * kLaneChunk + 8 * p, v[i]);
        }
    };
    // samples [16 g, 16 g + 16) of this lane's channel (trip g of the frame) -> window -> biased floats
    auto fetch = [&](int g, float (&u)[16]) {
        const int t = g >> 2, j = g & 3;
        const unsigned char *buf = lines + (t & 1) * kLaneBufBytes;
        const uint4 xa = *reinterpret_cast<const uint4 *>(buf + duo_swz(lane, 2 * j));
        const uint4 xb = *reinterpret_cast<const uint4 *>(buf + duo_swz(lane, 2 * j + 1));
        const int4 *rom = reinterpret_cast<const int4 *>(roms + (t & 1) * kLaneChunk + 16 * j);
        const int4 r0 = rom[0], r1 = rom[1], r2 = rom[2], r3 = rom[3];       // same address in every lane: broadcast
        float ua[8], ub[8];
        if (rom_quirk_range((16 * g) & (kWindowLen - 1), 16)) {               // warp-uniform
            window8_biased<true>(xa, r0, r1, exp23, ua);
            window8_biased<true>(xb, r2, r3, exp23, ub);
        } else {
            window8_biased<false>(xa, r0, r1, exp23, ua);
            window8_biased<false>(xb, r2, r3, exp23, ub);
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) { u[jj] = ua[jj]; u[8 + jj] = ub[jj]; }
    };
    // the two finished 8-sample groups of trip g, in place: samples [16 g - 8, 16 g) and [16 g, 16 g + 8)
    auto put = [&](int g, uint4 ga, uint4 gb) {
        const int t = g >> 2, j = g & 3;
        unsigned char *buf = lines + (t & 1) * kLaneBufBytes;
        *reinterpret_cast<uint4 *>(buf + duo_swz(lane, 2 * j)) = gb;
        if (j > 0) *reinterpret_cast<uint4 *>(buf + duo_swz(lane, 2 * j - 1)) = ga;
        else if (t > 0) *reinterpret_cast<uint4 *>(lines + ((t - 1) & 1) * kLaneBufBytes + duo_swz(lane, 7)) = ga;
    };

    request(0);
    cp_async_wait<0>();
    __syncwarp();                                     // every line was copied by eight different lanes
    uint4 ga, gb;
    {
        float u[16];
        fetch(0, u);
        lane_trip16<B1Z, ALT, true>(a, c, u, st, pipe, carry, ga, gb);
        put(0, ga, gb);
        if (n_chunks > 1) request(1);
    }
    const int n_trips = a.n / 16;
#pragma unroll 1
    for (int g = 1; g < n_trips; ++g) {
        const int t = g >> 2;
        if ((g & 3) == 0) {
            cp_async_wait<0>();                       // chunk t (requested three trips ago) has landed
            __syncwarp();
        }
        float u[16];
        fetch(g, u);
        lane_trip16<B1Z, ALT, false>(a, c, u, st, pipe, carry, ga, gb);
        put(g, ga, gb);
        if ((g & 3) == 0) {
            __syncwarp();                             // chunk t - 1 is complete in its buffer
            store_lines(t - 1);
            __syncwarp();                             // ... and read out, before the copy engine refills it
            if (t + 1 < n_chunks) request(t + 1);
        }
    }
    // drain: iterations n .. n + 4 push samples n - 5 .. n - 1 through the remaining stages.  Stage s saw
    // its last real sample at iteration n - 1 + s: its history is final (and stored) right after that;
    // what it computes later is junk and goes nowhere.
    if (live) store_state_biased(a.state, c, 0, st[0]);
    float tail[kLaneLag];
#pragma unroll
    for (int j = 0; j < kLaneLag; ++j) {
        tail[j] = lane_skewed_iteration<B1Z, ALT>(kBias16, a.coef, st, pipe);
        if (live) store_state_biased(a.state, c, j + 1, st[j + 1]);
    }
    ga.x = pack16_acc(carry[0], carry[1]);
    ga.y = pack16_acc(carry[2], tail[0]);
    ga.z = pack16_acc(tail[1], tail[2]);
    ga.w = pack16_acc(tail[3], tail[4]);
    *reinterpret_cast<uint4 *>(lines + ((n_chunks - 1) & 1) * kLaneBufBytes + duo_swz(lane, 7)) = ga;
    __syncwarp();
    store_lines(n_chunks - 1);
}

// ----------------------------------------------------------------- k1_split
constexpr int kSplitGroups = 5;      // channels per warp (5 x 6 = 30 lanes, 2 idle)
constexpr int kSplitChunk = 256;     // samples staged per chunk
constexpr int kSplitWarps = 4;       // warps per CTA
constexpr int kSplitRing = 2 * kSplitChunk;
constexpr int kSplitWinStride = kSplitChunk + 20;   // floats; rows start in different banks, and the look-ahead
                                                    // load of the block after the last one stays inside the row
constexpr int kSplitRingStride = kSplitRing + 8;    // int16;  +16 B, same reason
constexpr int kSplitRawBytes = kSplitGroups * kSplitChunk * 2;
constexpr int kSplitWinBytes = kSplitGroups * kSplitWinStride * 4;
constexpr int kSplitRingBytes = (kSplitGroups + 1) * kSplitRingStride * 2;   // + one dummy row: every lane stores, no divergence
constexpr int kSplitSmemPerWarp = ((kSplitRawBytes + kSplitWinBytes + kSplitRingBytes + 16 + 127) / 128) * 128;

constexpr int kSkew = 4;             // iterations between a stage and the next one (shuffle latency hiding)
constexpr int kSplitDelay = (kStages - 1) * kSkew;   // the last stage emits sample i - 20 at iteration i

struct SplitLane {
    StageCoef k;
    StageState st;
    float y;                // this lane's latest output
    float up[kSkew - 1];    // lane-1's outputs in flight: up[0] is this iteration's input
    unsigned carry[4];      // accumulator bits of the 4 newest outputs not yet stored
    int s;                  // stage 0..5
    bool first, last;
};

FRA_DEV void split_load_w(const float *p, float (&w)[16])
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = *reinterpret_cast<const float4 *>(p + 4 * q);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
}

// Sixteen systolic iterations i_base .. i_base+15.  Lane (g, s) processes sample
// i - kSkew*s at iteration i.  The skew is kSkew = 4 iterations per stage: the
// shuffle that carries stage s-1's output to stage s is issued three iterations
// before its result is consumed, so its ~30-cycle latency (plus the select and the
// two FFMAs behind it) fits in the slack and the per-iteration critical path is
// only the lane's own y[n-1] recurrence (FFMA -> PRMT -> FADD).  GUARD = true for
// the blocks that contain samples outside [0, n): the first two blocks of a frame
// (stages still empty) and the two flush blocks.
// MODE 0: exact, guarded (frame edges); 1: exact; 2: speculative (no-overflow assumption).
template <int MODE>
FRA_DEV float split_iterate16(SplitLane &L, int i_base, int n, const float (&w)[16], unsigned (&ob)[16])
{
    float absmax = 0.0f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float up_new = __shfl_up_sync(0xffffffffu, L.y, 1);   // consumed kSkew - 1 iterations from now
        const float x = L.first ? w[j] : L.up[0];
#pragma unroll
        for (int q = 0; q + 1 < kSkew - 1; ++q) L.up[q] = L.up[q + 1];
        L.up[kSkew - 2] = up_new;
        if (MODE == 0) {
            StageState keep = L.st;
            float ykeep = L.y;
            float y;
            float acc = biquad_step(x, L.k, L.st, &y);
            bool active = (unsigned)(i_base + j - kSkew * L.s) < (unsigned)n;
            if (active) {
                L.y = y;
            } else {
                L.st = keep;
                L.y = ykeep;
            }
            ob[j] = __float_as_uint(acc);
        } else if (MODE == 1) {
            ob[j] = __float_as_uint(biquad_step(x, L.k, L.st, &L.y));
        } else {
            ob[j] = __float_as_uint(biquad_step_spec(x, L.k, L.st, &L.y, &absmax));
        }
    }
    return absmax;
}

// One block of 16 iterations, then the last-stage lane's outputs go to the ring.
// `spec_ok` (warp-uniform) selects the speculative recurrence; a block in which any
// lane saw |y| > 32767 is rolled back and re-run exactly, and the function returns
// true so that the caller can stop speculating on a cascade that keeps overflowing.
template <bool GUARD>
FRA_DEV bool split_block16(SplitLane &L, int i_base, int n, const float (&w)[16], int16_t *ring_row, bool store_ok,
                           bool spec_ok)
{
    unsigned ob[16];
    bool redone = false;
    if (GUARD) {
        split_iterate16<0>(L, i_base, n, w, ob);
    } else if (spec_ok) {
        const StageState st0 = L.st;
        const float y0 = L.y;
        float up0[kSkew - 1];
#pragma unroll
        for (int q = 0; q < kSkew - 1; ++q) up0[q] = L.up[q];
        const float absmax = split_iterate16<2>(L, i_base, n, w, ob);
        if (__any_sync(0xffffffffu, absmax > 32767.0f)) {
            L.st = st0;
            L.y = y0;
#pragma unroll
            for (int q = 0; q < kSkew - 1; ++q) L.up[q] = up0[q];
            split_iterate16<1>(L, i_base, n, w, ob);
            redone = true;
        }
    } else {
        split_iterate16<1>(L, i_base, n, w, ob);
    }
    // the last-stage lane emitted samples i_base-20 .. i_base-5; with the carry
    // (i_base-24 .. i_base-21) that completes the aligned groups [i_base-24, i_base-17]
    // and [i_base-16, i_base-9].  Still offset-binary: the flush flips the sign bits.
    uint4 g1, g2;
    g1.x = __byte_perm(L.carry[0], L.carry[1], 0x5410);
    g1.y = __byte_perm(L.carry[2], L.carry[3], 0x5410);
    g1.z = __byte_perm(ob[0], ob[1], 0x5410);
    g1.w = __byte_perm(ob[2], ob[3], 0x5410);
    g2.x = __byte_perm(ob[4], ob[5], 0x5410);
    g2.y = __byte_perm(ob[6], ob[7], 0x5410);
    g2.z = __byte_perm(ob[8], ob[9], 0x5410);
    g2.w = __byte_perm(ob[10], ob[11], 0x5410);
#pragma unroll
    for (int j = 0; j < 4; ++j) L.carry[j] = ob[12 + j];
    // every lane stores: lanes that are not a valid last stage own the dummy ring row,
    // so there is no divergent branch at the block boundary.  Blocks in the interior of
    // a frame (GUARD = false: 32 <= i_base <= n - 16) always complete both groups.
    (void)store_ok;
    if (!GUARD || (i_base >= 24 && i_base - 16 <= n)) *reinterpret_cast<uint4 *>(ring_row + ((i_base - 24) & (kSplitRing - 1))) = g1;
    if (!GUARD || (i_base >= 16 && i_base - 8 <= n)) *reinterpret_cast<uint4 *>(ring_row + ((i_base - 16) & (kSplitRing - 1))) = g2;
    return redone;
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

    unsigned char *wbase = smem_raw + (size_t)warp_in_cta * kSplitSmemPerWarp;
    int16_t *raw = reinterpret_cast<int16_t *>(wbase);                                   // [5][256] int16, bulk-copy target
    float *win = reinterpret_cast<float *>(wbase + kSplitRawBytes);                      // [5][260] windowed samples
    int16_t *ring = reinterpret_cast<int16_t *>(wbase + kSplitRawBytes + kSplitWinBytes);  // [5][520] output ring
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + kSplitRawBytes + kSplitWinBytes + kSplitRingBytes);

    const int g_raw = lane / kStages;
    const int g = min(g_raw, kSplitGroups - 1);
    const bool lane_valid = g_raw < n_ch;
    SplitLane L;
    L.s = lane - g_raw * kStages;
    L.first = (L.s == 0);
    L.last = (L.s == kStages - 1);
    L.k = a.coef.set[L.s];
    L.y = 0.0f;
#pragma unroll
    for (int j = 0; j < kSkew - 1; ++j) L.up[j] = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) L.carry[j] = 0u;
    {
        uint2 v = make_uint2(0u, 0u);
        if (a.continuous && lane_valid)
            v = __ldg(reinterpret_cast<const uint2 *>(a.state + ((size_t)(c_base + g) * kStages + L.s) * 4));
        L.st.x1 = small_int_to_float(lo16(v.x));
        L.st.x2 = small_int_to_float(hi16(v.x));
        L.st.y1 = small_int_to_float(lo16(v.y));
        L.st.y2 = small_int_to_float(hi16(v.y));
    }

    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    unsigned phase = 0;

    // input staging: one 512-byte bulk copy (cp.async.bulk) per channel row, issued by
    // lanes 0..n_ch-1, landing in `raw` while the previous chunk is being filtered
    int4 rom_a, rom_b;            // window coefficients of the chunk in flight, fetched with it
    auto prefetch = [&](int i0) {
        rom_a = __ldg(reinterpret_cast<const int4 *>(a.rom32 + ((i0 + 4 * lane) & (kWindowLen - 1))));
        rom_b = __ldg(reinterpret_cast<const int4 *>(a.rom32 + ((i0 + 128 + 4 * lane) & (kWindowLen - 1))));
        if (lane == 0) mbar_expect_tx(bar, (unsigned)(n_ch * kSplitChunk * 2));
        if (lane < n_ch)
            bulk_g2s(raw + lane * kSplitChunk, a.in + (size_t)(c_base + lane) * a.n + i0, kSplitChunk * 2, bar);
    };
    // raw int16 -> window -> float; lane handles samples [4*lane, +4) and [128 + 4*lane, +4) of every row
    auto convert = [&]() {
        mbar_wait(bar, phase);
        phase ^= 1u;
        const int4 ra = rom_a, rb = rom_b;
#pragma unroll
        for (int r = 0; r < kSplitGroups; ++r) {
            uint2 pa = make_uint2(0u, 0u), pb = make_uint2(0u, 0u);
            if (r < n_ch) {
                pa = *reinterpret_cast<const uint2 *>(raw + r * kSplitChunk + 4 * lane);
                pb = *reinterpret_cast<const uint2 *>(raw + r * kSplitChunk + 128 + 4 * lane);
            }
            float4 fa, fb;
            fa.x = small_int_to_float(window_int(lo16(pa.x), ra.x));
            fa.y = small_int_to_float(window_int(hi16(pa.x), ra.y));
            fa.z = small_int_to_float(window_int(lo16(pa.y), ra.z));
            fa.w = small_int_to_float(window_int(hi16(pa.y), ra.w));
            fb.x = small_int_to_float(window_int(lo16(pb.x), rb.x));
            fb.y = small_int_to_float(window_int(hi16(pb.x), rb.y));
            fb.z = small_int_to_float(window_int(lo16(pb.y), rb.z));
            fb.w = small_int_to_float(window_int(hi16(pb.y), rb.w));
            *reinterpret_cast<float4 *>(win + r * kSplitWinStride + 4 * lane) = fa;
            *reinterpret_cast<float4 *>(win + r * kSplitWinStride + 128 + 4 * lane) = fb;
        }
        fence_proxy_async();       // my reads of `raw` are done before the next bulk copy may overwrite it
        __syncwarp();
    };
    // store the finished output groups [m0, m0 + 8 * count8) from the ring (m0 may be negative)
    auto flush = [&](int m0, int count8) {
#pragma unroll
        for (int r = 0; r < kSplitGroups; ++r) {
            const int m = m0 + 8 * lane;
            if (r < n_ch && lane < count8 && m >= 0) {
                uint4 v = *reinterpret_cast<const uint4 *>(ring + r * kSplitRingStride + (m & (kSplitRing - 1)));
                v.x ^= 0x80008000u; v.y ^= 0x80008000u; v.z ^= 0x80008000u; v.w ^= 0x80008000u;   // offset-binary -> two's complement
                stg128(a.out + (size_t)(c_base + r) * a.n + m, v);
            }
        }
    };

    prefetch(0);
    convert();
    const float *win_row = win + g * kSplitWinStride;
    int16_t *ring_row = ring + ((L.last && lane_valid) ? g : kSplitGroups) * kSplitRingStride;
    float wa[16], wb[16];
    bool spec_ok = a.speculate != 0;   // warp-uniform

    for (int i0 = 0; i0 < a.n; i0 += kSplitChunk) {
        const bool more = (i0 + kSplitChunk) < a.n;
        if (more) prefetch(i0 + kSplitChunk);
        int b = 0;
        split_load_w(win_row, wa);
        if (i0 == 0) {
            split_load_w(win_row + 16, wb);
            split_block16<true>(L, 0, a.n, wa, ring_row, lane_valid, false);
            split_load_w(win_row + 32, wa);
            split_block16<true>(L, 16, a.n, wb, ring_row, lane_valid, false);
            b = 32;
        }
        // two blocks per trip: the samples of block b+16 are fetched while block b runs
        int redo = 0;
        for (; b < kSplitChunk; b += 32) {
            split_load_w(win_row + b + 16, wb);
            redo += split_block16<false>(L, i0 + b, a.n, wa, ring_row, lane_valid, spec_ok) ? 1 : 0;
            split_load_w(win_row + b + 32, wa);      // past the chunk's end on the last trip: padding, never used
            redo += split_block16<false>(L, i0 + b + 16, a.n, wb, ring_row, lane_valid, spec_ok) ? 1 : 0;
        }
        // a cascade that keeps overflowing (16-bit wrap) gains nothing from speculation:
        // after a chunk with more than two rolled-back blocks this warp runs exactly
        if (redo > 2) spec_ok = false;
        __syncwarp();
        flush(i0 - 24, kSplitChunk / 8);
        __syncwarp();
        if (more) convert();
    }
    // flush blocks: iterations n .. n+31 drain stages 1..5 and complete the groups up to [n-8, n-1]
#pragma unroll
    for (int j = 0; j < 16; ++j) wa[j] = 0.0f;
    split_block16<true>(L, a.n, a.n, wa, ring_row, lane_valid, false);
    split_block16<true>(L, a.n + 16, a.n, wa, ring_row, lane_valid, false);
    __syncwarp();
    flush(a.n - 24, 3);

    if (lane_valid) {
        uint2 v;
        v.x = pack16((unsigned)(int)L.st.x1, (unsigned)(int)L.st.x2);
        v.y = pack16((unsigned)(int)L.st.y1, (unsigned)(int)L.st.y2);
        *reinterpret_cast<uint2 *>(a.state + ((size_t)(c_base + g) * kStages + L.s) * 4) = v;
    }
}

// ------------------------------------------------- shared-memory tiles of the pipelined kernel (k1_duo)
// (k1_stage, round 1's one-warp-per-stage pipeline, was superseded by k1_duo at every channel count
// and has been removed; its tile layout and chunk length live on here.)
constexpr int kStageChunk = 64;                       // samples per pipeline step
constexpr int kStageTileFloats = kStageChunk * 32;    // one chunk of 32 channels

// tile layout: [sample / 4][channel][4] floats - lane c reads / writes 16 bytes at 16 c:
// conflict-free 128-bit accesses, four consecutive samples of its channel per access
FRA_DEV float4 *stage_tile(float *smem, int boundary, int parity)
{
    return reinterpret_cast<float4 *>(smem + (size_t)(boundary * 2 + parity) * kStageTileFloats);
}

// window + int -> float for 8 packed samples; QUIRK = false skips the +32768 -> 0 fix-up,
// which only ROM entries equal to -32768 can trigger
template <bool QUIRK>
FRA_DEV void stage_convert8(uint4 x, int4 ra, int4 rb, float4 &fa, float4 &fb)
{
    auto w = [](int v, int c) { return small_int_to_float(QUIRK ? window_int(v, c) : window_int_fast(v, c)); };
    fa.x = w(lo16(x.x), ra.x); fa.y = w(hi16(x.x), ra.y); fa.z = w(lo16(x.y), ra.z); fa.w = w(hi16(x.y), ra.w);
    fb.x = w(lo16(x.z), rb.x); fb.y = w(hi16(x.z), rb.y); fb.z = w(lo16(x.w), rb.z); fb.w = w(hi16(x.w), rb.w);
}

// ------------------------------------------------------------------ k1_duo
// TWO stages per warp, chained in registers.  A warp that runs one biquad
// recurrence is bound by the latency of its dependency chain (FFMA -> PRMT -> FADD, ~14-16
// cycles per sample for 7.5 issue slots); round 1's k1_stage hid that by putting two single-stage warps on a
// scheduler, and all eight warps then queue on the shared-memory pipe (12 tile accesses per
// sample and CTA, l1tex ~2/3 busy).  Here a stage warp runs stages 2w and 2w+1 on the same
// sample stream: stage 2w+1's recurrence trails stage 2w's by one sample, so the warp carries
// two dependency chains that overlap, the value between them never leaves the register
// file, and the tile traffic drops (4 boundaries instead of 6).  CTA = 5 warps: loader, 3
// stage pairs on a scheduler each, writer; one CTA per SM for up to 148 x 32 channels, two
// per SM beyond that.
constexpr int kDuoPairs = kStages / 2;
constexpr int kDuoWarps = 2 + kDuoPairs;                // loader, three stage pairs, writer
constexpr int kDuoBoundaries = kDuoPairs + 1;           // loader -> pair 0 -> pair 1 -> pair 2 -> writer
constexpr int kDuoTilesBytes = kDuoBoundaries * 2 * kStageTileFloats * (int)sizeof(float);    // 64 KiB
// Global memory is touched in whole 128-byte lines only: a chunk of one channel IS one line
// (64 int16), so eight lanes move a channel's line and one warp instruction covers four
// channels (4 LSU tags instead of the 32 of a lane-per-channel access; every sector is moved
// once and written whole).  Lines are staged in shared memory as [channel][8 pieces of 16 B],
// the piece index XOR-swizzled with (channel & 7) so that the lane-per-channel side (loader
// convert, last stage's packed output) is conflict-free too.
constexpr int kDuoLineTileBytes = 32 * 128;              // one chunk of 32 channels as int16
constexpr int kDuoRomBytes = kStageChunk * 4;             // the chunk's 64 window ROM entries (int32)
constexpr int kDuoInOff = kDuoTilesBytes;                                   // raw input lines, 2 buffers
constexpr int kDuoRomOff = kDuoInOff + 2 * kDuoLineTileBytes;               // ROM slices, 2 buffers
constexpr int kDuoSmemBytes = kDuoRomOff + 2 * kDuoRomBytes;                // 72.5 KiB
// requested size: padded to 96 KiB.  (1) At most TWO CTAs share an SM - a third adds no
// throughput (the three stage-pair schedulers are issue-bound with two) and makes the tail wave
// longer.  (2) In FRA_PIPELINE mode the FFT kernel's CTAs (64 KiB each) run beside this kernel:
// one CTA of this kernel and two of the FFT fit an SM (97 + 65 + 65 KiB), but TWO of this
// kernel beside an FFT CTA do not, so its 128 CTAs cannot double up on SMs the FFT already
// occupies (measured: 0.273 -> 0.260 ms per step in the pipeline's good mode).
#ifdef FRA_DUO_SMEM_REQUEST
constexpr int kDuoSmemRequest = FRA_DUO_SMEM_REQUEST;
#else
constexpr int kDuoSmemRequest = 96 * 1024;
#endif


// Every stage warp runs the SAME branch-free straight-line chunk body: the last pair also
// leaves its output as a float tile, and the writer warp (on the scheduler the pair warps do
// not use) packs it to int16.  Two earlier versions are why: per-role instantiations overflowed the
// 32 KB instruction-cache level ("no_instruction" 2.2 of 4.3 stall slots), and a run-time
// `last` branch per group is a basic-block boundary that drains both dependency chains four
// times per chunk (21.4 instead of 17.9 cycles per sample).
// one stage of a pair: the exact step, or its two-instruction-recurrence form (u = y + kBias16)
// MODE 0: biquad_step, 1: biquad_step_fast, 2: biquad_step_biased (tiles and history hold v + kBias16)
template <bool B1Z, int FAST>
FRA_DEV float duo_step(float x, const StageCoef &k, StageState &s, float &u)
{
    float y;
    if (FAST == 2) (void)biquad_step_biased<B1Z>(x, k, s, &y);
    else if (FAST == 1) (void)biquad_step_fast<B1Z>(x, k, s, u, &y);
    else (void)biquad_step<B1Z>(x, k, s, &y);
    return y;
}

template <bool B1Z, int FAST>
FRA_DEV void duo_group16(const float4 (&in)[4], const StageCoef &ka, const StageCoef &kb, StageState &sa,
                         StageState &sb, float &ua, float &ub, float4 *tout)
{
    float y[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        y[4 * q + 0] = duo_step<B1Z, FAST>(duo_step<B1Z, FAST>(in[q].x, ka, sa, ua), kb, sb, ub);
        y[4 * q + 1] = duo_step<B1Z, FAST>(duo_step<B1Z, FAST>(in[q].y, ka, sa, ua), kb, sb, ub);
        y[4 * q + 2] = duo_step<B1Z, FAST>(duo_step<B1Z, FAST>(in[q].z, ka, sa, ua), kb, sb, ub);
        y[4 * q + 3] = duo_step<B1Z, FAST>(duo_step<B1Z, FAST>(in[q].w, ka, sa, ua), kb, sb, ub);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) tout[q * 32] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
}

// one chunk (64 samples) through a stage pair, straight-line
template <bool B1Z, int FAST>
FRA_DEV void duo_chunk(const float4 *tin, float4 *tout, const StageCoef &ka, const StageCoef &kb, StageState &sa,
                       StageState &sb, float &ua, float &ub)
{
    float4 buf[2][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) buf[0][q] = tin[q * 32];
#pragma unroll
    for (int g = 0; g < kStageChunk / 16; ++g) {
        if (g + 1 < kStageChunk / 16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) buf[(g + 1) & 1][q] = tin[((g + 1) * 4 + q) * 32];
        }
        duo_group16<B1Z, FAST>(buf[g & 1], ka, kb, sa, sb, ua, ub, tout + (g * 4) * 32);
    }
}

// loader, input side: chunk t+1's lines are requested (cp.async, whole lines) before chunk t
// is windowed and converted into the loader->pair-0 tile
template <bool BIASED>
FRA_DEV void duo_loader_request(const K1Args &a, int c0, int t, unsigned char *smem_raw, int lane)
{
    unsigned char *dst = smem_raw + kDuoInOff + (t & 1) * kDuoLineTileBytes;
    const int p = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + (lane >> 3);
        const int ch = min(c0 + r, a.channels - 1);                       // clamped: rows past the end read valid memory
        cp_async16(dst + duo_swz(r, p), a.in + (size_t)ch * a.n + (size_t)t * kStageChunk + 8 * p);
    }
    if (lane < 16)
        cp_async16(smem_raw + kDuoRomOff + (t & 1) * kDuoRomBytes + 16 * lane,
                   (BIASED ? a.rom2x : a.rom32) + (((t * kStageChunk) & (kWindowLen - 1)) + 4 * lane));
}

// (all shared-memory loads of a batch are issued before the first store: the compiler cannot
// move a load above a store to memory it cannot tell apart, and a load -> convert -> store
// sequence per piece costs one full shared-memory latency each)
template <bool QUIRK, bool BIASED>
FRA_DEV void duo_convert_chunk(const unsigned char *lines, const int4 *rom, float4 *tile, int lane, unsigned exp23)
{
    uint4 x[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = *reinterpret_cast<const uint4 *>(lines + duo_swz(lane, q));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int4 r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = rom[8 * h + j];            // same address in every lane: broadcast
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = 4 * h + j;
            float4 fa, fb;
            if (BIASED) {
                float u[8];
                window8_biased<QUIRK>(x[q], r[2 * j], r[2 * j + 1], exp23, u);
                fa = make_float4(u[0], u[1], u[2], u[3]);
                fb = make_float4(u[4], u[5], u[6], u[7]);
            } else {
                stage_convert8<QUIRK>(x[q], r[2 * j], r[2 * j + 1], fa, fb);
            }
            tile[(2 * q) * 32] = fa;
            tile[(2 * q + 1) * 32] = fb;
        }
    }
}

// writer warp: the last pair's float tile of chunk `chunk` -> int16 lines (swizzled, lane per
// channel) -> global memory, whole lines (eight lanes per channel).  The lines are staged in
// the first half of the float tile itself, once every lane has read its part of it: the last
// pair will not write this parity again before the next barrier.
template <bool BIASED>
FRA_DEV void duo_store_chunk(const K1Args &a, int c0, int chunk, unsigned char *smem_raw, int lane)
{
    float4 *tile_base = stage_tile(reinterpret_cast<float *>(smem_raw), kDuoPairs, chunk & 1);
    const float4 *tile = tile_base + lane;
    unsigned char *lines = reinterpret_cast<unsigned char *>(tile_base);
    auto pk = [](float u, float v) {
        // biased: the low halves are the values in offset-binary; else u + 1.5 * 2^23 holds u mod 2^16 in its low mantissa bits
        if (BIASED) return pack16(__float_as_uint(u), __float_as_uint(v)) ^ 0x80008000u;
        return pack16(__float_as_uint(u + kMagic), __float_as_uint(v + kMagic));
    };
    uint4 o[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = tile[(8 * h + j) * 32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o[4 * h + j].x = pk(f[2 * j].x, f[2 * j].y);
            o[4 * h + j].y = pk(f[2 * j].z, f[2 * j].w);
            o[4 * h + j].z = pk(f[2 * j + 1].x, f[2 * j + 1].y);
            o[4 * h + j].w = pk(f[2 * j + 1].z, f[2 * j + 1].w);
        }
    }
    __syncwarp();                                              // every lane has read the tile
#pragma unroll
    for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4 *>(lines + duo_swz(lane, q)) = o[q];
    __syncwarp();
    const int p = lane & 7;
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const uint4 *>(lines + duo_swz(4 * i + (lane >> 3), p));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int ch = c0 + 4 * i + (lane >> 3);
        if (ch < a.channels) stg128(a.out + (size_t)ch * a.n + (size_t)chunk * kStageChunk + 8 * p, v[i]);
    }
}

template <bool BIASED>
FRA_DEV void duo_loader_step(const K1Args &a, int c0, int t, int n_chunks, unsigned char *smem_raw, int lane)
{
    if (t >= n_chunks) return;
    if (t + 1 < n_chunks) duo_loader_request<BIASED>(a, c0, t + 1, smem_raw, lane);
    cp_async_commit();
    cp_async_wait<1>();                                        // everything but the group just committed: chunk t is here
    __syncwarp();                                              // each line was copied by eight different lanes
    const unsigned char *lines = smem_raw + kDuoInOff + (t & 1) * kDuoLineTileBytes;
    const int4 *rom = reinterpret_cast<const int4 *>(smem_raw + kDuoRomOff + (t & 1) * kDuoRomBytes);
    float4 *tile = stage_tile(reinterpret_cast<float *>(smem_raw), 0, t & 1) + lane;
    const int w0 = (t * kStageChunk) & (kWindowLen - 1);
    // ROM entries equal to -32768 live in [0, 15), [8178, 8206) and [16369, 16384)
    const bool quirk = (w0 < 64) || (w0 >= 8128 && w0 < 8256) || (w0 >= kWindowLen - 64);
    const unsigned exp23 = a.coef.set[0].exp23;
    if (quirk) duo_convert_chunk<true, BIASED>(lines, rom, tile, lane, exp23);
    else duo_convert_chunk<false, BIASED>(lines, rom, tile, lane, exp23);
}

FRA_DEV StageState duo_load_state(const K1Args &a, int cc, int s)
{
    StageState st = {0.0f, 0.0f, 0.0f, 0.0f};
    if (a.continuous) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(a.state + ((size_t)cc * kStages + s) * 4));
        st.x1 = small_int_to_float(lo16(v.x));
        st.x2 = small_int_to_float(hi16(v.x));
        st.y1 = small_int_to_float(lo16(v.y));
        st.y2 = small_int_to_float(hi16(v.y));
    }
    return st;
}

FRA_DEV void duo_store_state(const K1Args &a, int c, int s, const StageState &st)
{
    uint2 v;
    v.x = pack16((unsigned)(int)st.x1, (unsigned)(int)st.x2);
    v.y = pack16((unsigned)(int)st.y1, (unsigned)(int)st.y2);
    *reinterpret_cast<uint2 *>(a.state + ((size_t)c * kStages + s) * 4) = v;
}

// ALT: the six stages use two alternating coefficient sets (the RTL's bank layout): the pair's
// coefficients are then compile-time addresses and stay in uniform registers (a run-time set index
// moves them to vector registers and costs 5 %)
#ifdef FRA_DUO_MAXNREG
#define FRA_DUO_BOUNDS __maxnreg__(FRA_DUO_MAXNREG)
#else
#define FRA_DUO_BOUNDS __launch_bounds__(kDuoWarps * 32, 2)
#endif
template <bool B1Z, int FAST, bool ALT>
__global__ void FRA_DUO_BOUNDS k1_duo(K1Args a)
{
    FRA_DYN_SMEM(smem_raw);
    float *smem = reinterpret_cast<float *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32;
    const int c = c0 + lane;
    const bool live = c < a.channels;
    const int cc = live ? c : (a.channels - 1);
    const int n_chunks = a.n / kStageChunk;
    const int n_steps = n_chunks + kDuoPairs + 1;                 // + fill of the three pairs + the writer's last chunk
#ifdef FRA_TIMELINE
    timeline_mark(0, a.tl_step);
#endif

    constexpr bool BIASED = (FAST == 2);
    if (warp == 0) {
        duo_loader_request<BIASED>(a, c0, 0, smem_raw, lane);
        cp_async_commit();
        for (int t = 0; t < n_steps; ++t) {
            duo_loader_step<BIASED>(a, c0, t, n_chunks, smem_raw, lane);
            __syncthreads();
        }
#ifdef FRA_TIMELINE
        timeline_mark(1, a.tl_step);
#endif
    } else if (warp == kDuoWarps - 1) {
        // shares the loader's scheduler: both are short, latency-bound instruction streams
        for (int t = 0; t < n_steps; ++t) {
            const int done = t - (kDuoPairs + 1);                // the chunk the last pair finished in the step before
            if (done >= 0) duo_store_chunk<BIASED>(a, c0, done, smem_raw, lane);
            __syncthreads();
        }
    } else {
        const int p = warp - 1;                                  // stage pair: stages 2p and 2p+1
        const StageCoef ka = a.coef.set[ALT ? 0 : 2 * p], kb = a.coef.set[ALT ? 1 : 2 * p + 1];
        StageState sa, sb;
        if (BIASED) {
            sa = load_state_biased(a.state, (size_t)cc, 2 * p, a.continuous != 0);
            sb = load_state_biased(a.state, (size_t)cc, 2 * p + 1, a.continuous != 0);
        } else {
            sa = duo_load_state(a, cc, 2 * p);
            sb = duo_load_state(a, cc, 2 * p + 1);
        }
        float ua = sa.y1 + kBias16, ub = sb.y1 + kBias16;        // biquad_step_fast only
        for (int t = 0; t < n_steps; ++t) {
            const int chunk = t - 1 - p;
            if (chunk >= 0 && chunk < n_chunks) {
                duo_chunk<B1Z, FAST>(stage_tile(smem, p, chunk & 1) + lane, stage_tile(smem, p + 1, chunk & 1) + lane,
                                     ka, kb, sa, sb, ua, ub);
            }
            __syncthreads();
        }
        if (live) {
            if (BIASED) {
                store_state_biased(a.state, (size_t)c, 2 * p, sa);
                store_state_biased(a.state, (size_t)c, 2 * p + 1, sb);
            } else {
                duo_store_state(a, c, 2 * p, sa);
                duo_store_state(a, c, 2 * p + 1, sb);
            }
        }
    }
}

// ------------------------------------------------- FRA_WINDOW_RTL_SKEW: the window's register skew
// hann8192.vhd:36-39 updates coef_s, product and sample_out in the same clocked branch, so at strobe n the
// output is the rounding of x[n-1] * ROM[n-2]; the first two outputs after power-up come from
// zero registers.  In this mode the chain runs on the stream delayed by one sample (this kernel: the
// previous frame's last sample is carried per channel) with the ROM tables rotated by two entries.
__global__ void __launch_bounds__(256) k0_skew_delay(const int16_t *in, int16_t *out, int16_t *prev, size_t total8, int n,
                                                      int continuous)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread = 8 consecutive samples
    if (i >= total8) return;
    const size_t e = i * 8;
    const size_t c = e / (size_t)n;
    const int n0 = (int)(e % (size_t)n);
    const uint4 x = ldg128(in + e);
    // the sample in front of these eight: the previous group's last one, or the carry of the previous frame
    unsigned before;
    if (n0 == 0) before = continuous ? (unsigned)(uint16_t)prev[c] : 0u;
    else before = (unsigned)(uint16_t)in[e - 1];
    uint4 o;
    o.x = (x.x << 16) | before;
    o.y = (x.y << 16) | (x.x >> 16);
    o.z = (x.z << 16) | (x.y >> 16);
    o.w = (x.w << 16) | (x.z >> 16);
    if (n0 == 0 && !continuous) o.x = 0u;            // power-up: x[0] meets the still-zero coefficient register
    stg128(out + e, o);
}

__global__ void __launch_bounds__(256) k0_skew_carry(const int16_t *in, int16_t *prev, int channels, int n)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < channels) prev[c] = in[(size_t)c * n + n - 1];
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
