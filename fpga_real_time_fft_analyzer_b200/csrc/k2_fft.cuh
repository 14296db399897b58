// k2_fft.cuh - K2+K3: batched N-point FFT of real int16 frames with the bin
// framing fused into the last pass.  One HBM round trip per frame: int16 samples
// in, 4-byte bins out; everything in between lives in shared memory.
//
// Replaces xfft_0 (IP/xfft_0/xfft_0.xci, instance IMP/dsp_system_top.vhd:530-545;
// input word {im = 0, re = sample}, NEW/command_control.vhd:123), the frame
// buffer (IMP/sequencer_dsp.vhd:50-82) and the byte order of sequ_2
// (IMP/sequ2.vhd:153,234), plus the GUI's magnitude decode (GUI:250-260).
// In bypass mode (0xB1) the window multiply (NEW/hann8192.vhd:36-39) is applied
// here while loading, so the whole chain is this one kernel.
//
// Algorithm (tools/fft_plan_prototype.py is the numpy statement of the same
// index arithmetic): the real frame is packed as M = N/2 complex points
// z[m] = x[2m] + i x[2m+1]; z is split into F interleaved sub-sequences of length
// L (4096 for N >= 8192, else 256); each gets a radix-16 Stockham (autosort)
// FFT - 3 or 2 passes, butterflies in registers, exchange through one swizzled
// shared-memory buffer; the last pass does the radix-F combine, the real-input
// untangle, the Hermitian mirror and the int16 quantise/pack in registers and
// writes coalesced 4-byte bins.
#pragma once
#include "fra_common.cuh"

namespace fra {

// pass 0 of the 16K kernel: one 64-bit load per row for the thread's two butterflies (1) or two 32-bit loads (0)
#ifndef FRA_K2_PAIRED
#define FRA_K2_PAIRED 1
#endif
// resident CTAs per SM the FFT kernel is compiled for (register cap 65536 / (256 * this)).  With the
// passes unrolled and both butterflies' input words in flight the kernel wants ~100 registers: at 3 CTAs
// per SM (80 registers) it spills 250 bytes and runs 2.00 ms per 65536 frames, at 2 (no spills) 1.62 ms.
// butterfly order in the passes without a read/write hazard: 0 = both butterflies of a thread in flight,
// 1 = one after the other (unrolled), 2 = one after the other in a rolled loop (fewest registers)
#ifndef FRA_K2_SEQ
#define FRA_K2_SEQ 0
#endif
// (Capping the kernel at 112 registers - so that one of its CTAs fits beside three CTAs of the lane-per-channel
// window+IIR kernel in pipelined mode - compiles without spills and changes nothing: 2.93 ms per step either way.)
#ifndef FRA_K2_EARLY_BARRIER
#define FRA_K2_EARLY_BARRIER 1
#endif
#ifndef FRA_K2_MINBLOCKS
#define FRA_K2_MINBLOCKS 2
#endif

struct K2Args {
    const uint32_t *in;     // [B][N/2] words = int16 pairs
    const int *rom32;       // window ROM (WIN only)
    const float2 *tw1;      // [16][16]   W_256^(r k)
    const float2 *tw2;      // [16][256]  W_4096^(r k)
    const float2 *twn;      // [N/2]      W_N^e
    uint32_t *frames;       // [B][N] {re int16, im int16} little-endian, or null
    float2 *iq;             // [B][N] fp32 bins, or null
    float *mag;             // [B][N] or null
    float *phase;           // [B][N] or null
    float qscale;           // 0.5 * 2^log2_scale (the untangle leaves 2 X[k])
    float mag_alpha;        // 1: mag = |bin|; in (0, 1): mag <- mag + alpha (|bin| - mag), a running average across calls
    int batch;
    unsigned exp23;         // 0x4B000000 as data (keeps PRMT's selector an immediate)
    int frame0;             // host side only: index of the first frame within the context (scratch offset of the 64K path)
    int prefetch;           // k2_fft: L2 prefetch distance in CTAs (0 = off)
#ifdef FRA_TIMELINE
    int tl_step;
#endif
};

template <int LOG2N>
struct FftPlan {
    static constexpr int N = 1 << LOG2N;
    static constexpr int M = N / 2;
    static constexpr int L = (LOG2N >= 13) ? 4096 : 256;
    static constexpr int F = M / L;
    static constexpr int NB = L / 16;
    static constexpr int PASSES = (L == 4096) ? 3 : 2;
    static constexpr int FPC = (N >= 16384) ? 1 : 16384 / N;   // frames per CTA
    static constexpr int ITEMS = FPC * F * NB;                  // butterflies per pass per CTA
    static constexpr int THREADS = ITEMS / 2;                  // two butterflies per thread per pass: 256 (512 at 32K)
    static constexpr int SMEM_BYTES = FPC * M * 8;
    static constexpr int SLOTS = FPC * (L / 2);                 // last-pass work items per CTA
    static constexpr int FLOC = F;                              // sub-sequences held by one CTA (all of them)
    static constexpr int MLOC = M;                              // complex points per frame in one CTA's buffer
};

// N = 32768 / 65536 across a thread-block cluster: the F = M / 4096 sub-sequences of a frame are divided over the R
// CTAs of a cluster (FLOC = F / R each), every CTA runs its sub-FFTs locally, and the last pass reads the other
// CTAs' sub-sequences through distributed shared memory (k2_fft_cluster below).
//   <16, 4>: 65536 points (32768 complex points = 256 KiB of fp32, more than one SM has) as four CTAs of 64 KiB / 256
//            threads: two CTAs per SM, so one CTA's passes run beside the other's loads, barriers and stores (what the
//            16K kernel gains from two CTAs per SM; two CTAs of 128 KiB / 512 threads measured 249 against 274 Gsamples/s)
//   <15, 2>: 32768 points as two CTAs of 64 KiB / 256 threads instead of one CTA of 128 KiB / 512 threads
template <int LOG2N, int R>
struct FftPlanCluster {
    static constexpr int N = 1 << LOG2N;
    static constexpr int M = N / 2;
    static constexpr int L = 4096;
    static constexpr int F = M / L;
    static constexpr int NB = L / 16;
    static constexpr int PASSES = 3;
    static constexpr int FPC = 1;
    static constexpr int RANKS = R;
    static constexpr int FLOC = F / R;
    static constexpr int MLOC = FLOC * L;
    static constexpr int ITEMS = FLOC * NB;
    static constexpr int THREADS = ITEMS / 2;                   // 256 (FLOC = 2)
    static constexpr int SMEM_BYTES = MLOC * 8;                 // 64 KiB per CTA (FLOC = 2)
    static constexpr int MINBLOCKS = (SMEM_BYTES <= 64 * 1024) ? 2 : 1;
    static_assert(F % R == 0 && FLOC >= 1, "cluster plan: whole sub-sequences per CTA");
};

// two packed int16 -> two floats: one I2F.S16 each, reading the register's low / high half
// directly (conversion unit; two issue slots per pair instead of the five of an ALU/FMA-pipe
// unpack - XOR, two PRMT, two FADD - and the FFT kernels are issue-bound)
FRA_DEV float2 int16_pair_to_float2(unsigned w, unsigned exp23)
{
    (void)exp23;
    return make_float2((float)(short)(w & 0xFFFFu), (float)(short)(w >> 16));
}

// complex add / subtract as ONE packed fp32x2 instruction (Blackwell FADD2): same FP32 pipe
// time as two FADDs, half the issue slots - and the FFT kernels are issue-bound
#if defined(FRA_HOST_EMUL) || !defined(FRA_USE_F32X2)
FRA_DEV float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FRA_DEV float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#else
FRA_DEV float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
FRA_DEV float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
#endif
FRA_DEV float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
FRA_DEV float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)
FRA_DEV float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
FRA_DEV float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }     // a * (-i)

// shared-memory index swizzle (complex-element units, 8 B): within every 128-byte row
// of 16 elements the 16-byte chunk index is XORed with the row number mod 8.  The
// stride-16 scatter of pass 0 (as 16-byte vector stores), the stride-16 scatter of
// pass 1 and every unit-stride read are then bank-conflict-free, and because
// 256 r shifts the row by a multiple of 8, swz(b + 256 r) = swz(b) + 256 r: one
// swizzled base per thread and immediate offsets for the 16 butterfly inputs.
FRA_DEV int swz(int idx) { return idx ^ (((idx >> 4) & 7) << 1); }

FRA_DEV void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = make_float2(d02.x + d13.y, d02.y - d13.x);
    a3 = make_float2(d02.x - d13.y, d02.y + d13.x);
}

// forward 16-point DFT, natural order in and out: r = 4a + b, q = q1 + 4 q2
FRA_DEV void dft16(const float2 (&v)[16], float2 (&o)[16])
{
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    float2 u[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        u[b][0] = v[b]; u[b][1] = v[4 + b]; u[b][2] = v[8 + b]; u[b][3] = v[12 + b];
        dft4(u[b][0], u[b][1], u[b][2], u[b][3]);          // u[b][q1]
    }
    // twiddle W16^(b q1)
    u[1][1] = cmul(u[1][1], make_float2(c1, -s1));
    u[1][2] = make_float2((u[1][2].x + u[1][2].y) * h, (u[1][2].y - u[1][2].x) * h);
    u[1][3] = cmul(u[1][3], make_float2(s1, -c1));
    u[2][1] = make_float2((u[2][1].x + u[2][1].y) * h, (u[2][1].y - u[2][1].x) * h);
    u[2][2] = mul_mi(u[2][2]);
    u[2][3] = make_float2((u[2][3].y - u[2][3].x) * h, -(u[2][3].x + u[2][3].y) * h);
    u[3][1] = cmul(u[3][1], make_float2(s1, -c1));
    u[3][2] = make_float2((u[3][2].y - u[3][2].x) * h, -(u[3][2].x + u[3][2].y) * h);
    u[3][3] = cmul(u[3][3], make_float2(-c1, s1));
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) {
        float2 t0 = u[0][q1], t1 = u[1][q1], t2 = u[2][q1], t3 = u[3][q1];
        dft4(t0, t1, t2, t3);
        o[q1] = t0; o[q1 + 4] = t1; o[q1 + 8] = t2; o[q1 + 12] = t3;
    }
}

// forward F-point DFT in place, F in {1, 2, 4, 8}
template <int F>
FRA_DEV void dft_small(float2 (&a)[F])
{
    if (F == 2) {
        float2 t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    } else if (F == 4) {
        dft4(a[0], a[1 % F], a[2 % F], a[3 % F]);
    } else if (F == 8) {
        const float h = 0.70710678118654752f;
        float2 e0 = a[0], e1 = a[2 % F], e2 = a[4 % F], e3 = a[6 % F];
        float2 o0 = a[1 % F], o1 = a[3 % F], o2 = a[5 % F], o3 = a[7 % F];
        dft4(e0, e1, e2, e3);
        dft4(o0, o1, o2, o3);
        o1 = make_float2((o1.x + o1.y) * h, (o1.y - o1.x) * h);      // W8^1
        o2 = mul_mi(o2);                                              // W8^2
        o3 = make_float2((o3.y - o3.x) * h, -(o3.x + o3.y) * h);     // W8^3
        a[0] = cadd(e0, o0); a[4 % F] = csub(e0, o0);
        a[1 % F] = cadd(e1, o1); a[5 % F] = csub(e1, o1);
        a[2 % F] = cadd(e2, o2); a[6 % F] = csub(e2, o2);
        a[3 % F] = cadd(e3, o3); a[7 % F] = csub(e3, o3);
    }
}

// QMODE 0: floor, no saturation (scale <= 1/N cannot overflow); 1: floor + saturate;
// 2: round-to-nearest-even + saturate.  Returns a float whose bit pattern carries
// the two's-complement int16 in its low 16 bits (kMagic = 0x4B400000).
template <int QMODE>
FRA_DEV float quant(float v, float s)
{
    float t = (QMODE == 2) ? __fmaf_rn(v, s, kMagic) : __fmaf_rd(v, s, kMagic);
    if (QMODE >= 1) t = fminf(fmaxf(t, kMagic - 32768.0f), kMagic + 32767.0f);
    return t;
}

struct BinOut {
    uint32_t *frames;
    float2 *iq;
    float *mag;
    float *phase;
    float qscale;
    float mag_alpha;
};

// One bin and its conjugate-symmetric partner: X[up + off] = p/2, X[dn + off_conj] = conj(p)/2.
// `up` / `dn` are element indices (frame base +k and frame base -k) computed once per work item,
// `off` / `off_conj` are compile-time constants, so every store is base + immediate.
// ONLY_FRAMES: the int16 frames are the only output and their pointer is known to be non-null (OUT = 0)
template <int QMODE, bool ONLY_FRAMES = false>
FRA_DEV void emit_pair(const BinOut &o, size_t up, size_t dn, int off, int off_conj, bool write_conj, float2 p)
{
    if (ONLY_FRAMES) {
        const float qre = quant<QMODE>(p.x, o.qscale);
        const float qim = quant<QMODE>(p.y, o.qscale);
        const float qimc = quant<QMODE>(p.y, -o.qscale);
        o.frames[up + off] = __byte_perm(__float_as_uint(qre), __float_as_uint(qim), 0x5410);
        if (write_conj) o.frames[dn + off_conj] = __byte_perm(__float_as_uint(qre), __float_as_uint(qimc), 0x5410);
        return;
    }
    if (o.iq != nullptr) {
        o.iq[up + off] = make_float2(0.5f * p.x, 0.5f * p.y);
        if (write_conj) o.iq[dn + off_conj] = make_float2(0.5f * p.x, -0.5f * p.y);
    }
    if (o.frames != nullptr || o.mag != nullptr || o.phase != nullptr) {
        const float qre = quant<QMODE>(p.x, o.qscale);
        const float qim = quant<QMODE>(p.y, o.qscale);
        const float qimc = quant<QMODE>(p.y, -o.qscale);
        if (o.frames != nullptr) {
            o.frames[up + off] = __byte_perm(__float_as_uint(qre), __float_as_uint(qim), 0x5410);
            if (write_conj)
                o.frames[dn + off_conj] = __byte_perm(__float_as_uint(qre), __float_as_uint(qimc), 0x5410);
        }
        if (o.mag != nullptr || o.phase != nullptr) {
            const float fre = qre - kMagic, fim = qim - kMagic, fimc = qimc - kMagic;
            const float re2 = __fmul_rn(fre, fre);
            if (o.mag != nullptr) {
                // spectrum averaging fused into the pack stage (reference README "Contributing": waterfall /
                // averaging display options): the caller's magnitude buffer holds the running average
                auto put = [&](float *dst, float m) { *dst = (o.mag_alpha < 1.0f) ? __fmaf_rn(o.mag_alpha, m - *dst, *dst) : m; };
                put(o.mag + up + off, __fsqrt_rn(__fadd_rn(re2, __fmul_rn(fim, fim))));
                if (write_conj) put(o.mag + dn + off_conj, __fsqrt_rn(__fadd_rn(re2, __fmul_rn(fimc, fimc))));
            }
            if (o.phase != nullptr) {
                o.phase[up + off] = atan2f(fim, fre);
                if (write_conj) o.phase[dn + off_conj] = atan2f(fimc, fre);
            }
        }
    }
}

// W_16^q, q = 0..7 (the last pass needs W_(2F)^q = W_16^(q * 8 / F))
FRA_DEV float2 w16(int q)
{
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    switch (q) {
    case 0: return make_float2(1.0f, 0.0f);
    case 1: return make_float2(c1, -s1);
    case 2: return make_float2(h, -h);
    case 3: return make_float2(s1, -c1);
    case 4: return make_float2(0.0f, -1.0f);
    case 5: return make_float2(-s1, -c1);
    case 6: return make_float2(-h, -h);
    default: return make_float2(-c1, -s1);
    }
}

// One radix-16 Stockham pass of the whole CTA (two butterflies per thread), PASS a compile-time
// constant.  Pass 0 reads the int16 frame from global memory, the others read shared memory; the
// last pass of each size writes back to the positions it read, only the middle pass of L = 4096
// scatters into other threads' read positions and needs a barrier between read and write.
template <class P, bool WIN, int PASS>
FRA_DEV void fft_pass(const K2Args &a, float2 *buf, int tid, int frame0, const uint2 *staged = nullptr, int f0 = 0)
{
    constexpr int IPT = P::ITEMS / P::THREADS;          // butterflies per thread per pass (2)
    // N = 16384: THREADS = NB and F = 2, so a thread's two butterflies are column j = tid of the
    // sub-sequences f = 0 and f = 1, whose input words are adjacent: one 64-bit load for both
    // (in a cluster with two sub-sequences per CTA, f0 and f0 + 1, the same holds with a stride of F words)
    constexpr bool PAIRED = FRA_K2_PAIRED && (PASS == 0) && (P::FLOC == 2) && (P::THREADS == P::NB) && (IPT == 2) && (P::FPC == 1);
    // the middle pass of L = 4096 scatters into other threads' read positions: all reads, a barrier, all
    // writes (both butterflies' outputs live across it).  Every other pass writes where nobody else reads
    // in that pass (pass 0 reads global memory; the last pass of each size writes back to the positions
    // it read), so a butterfly can be stored before the next one is loaded.
    constexpr bool HAZARD = (PASS == 1 && P::PASSES == 3);
    uint2 pre[PAIRED ? 16 : 1];
    if constexpr (PAIRED) {
        if (staged != nullptr) {
            // the frame was brought into shared memory by a bulk copy (k2_fft_staged): unit-stride 8-byte reads
#pragma unroll
            for (int r = 0; r < 16; ++r) pre[r] = staged[tid + P::NB * r];
        } else {
            const bool live = frame0 < a.batch;
            const uint2 *src = reinterpret_cast<const uint2 *>(a.in + (size_t)frame0 * P::M + f0) + (P::F / 2) * tid;
#pragma unroll
            for (int r = 0; r < 16; ++r) pre[r] = live ? __ldg(src + (P::F / 2) * P::NB * r) : make_uint2(0u, 0u);   // z[F (j + NB r) + f0 + {0, 1}]
        }
    }
    // butterfly q of this thread: inputs (+ twiddles) -> 16-point DFT in o; wbase / jlow describe where it goes
    // inputs (+ twiddles) of butterfly q of this thread -> v; wbase / jlow describe where its outputs go
    auto gather = [&](int q, float2 (&v)[16], int &wbase, int &jlow) {
        const int it = tid + P::THREADS * q;
        const int j = it % P::NB;
        const int f = (it / P::NB) % P::FLOC;               // sub-sequence within this CTA's buffer
        const int fr = it / (P::NB * P::FLOC);
        const int base = fr * P::MLOC + f * P::L;
        if (PASS == 0) {
            const int frame = frame0 + fr;
            const bool live = frame < a.batch;
            const uint32_t *src = a.in + (size_t)frame * P::M + (P::F * j + f0 + f);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                unsigned w;
                if constexpr (PAIRED) w = (q == 0) ? pre[r].x : pre[r].y;
                else w = live ? __ldg(src + P::F * P::NB * r) : 0u;      // z[F (j + NB r) + f]
                if (WIN) {
                    const int e = P::F * (j + P::NB * r) + f0 + f;
                    const int2 c = __ldg(reinterpret_cast<const int2 *>(a.rom32 + ((2 * e) & (kWindowLen - 1))));
                    v[r] = make_float2(small_int_to_float(window_int(lo16(w), c.x)),
                                       small_int_to_float(window_int(hi16(w), c.y)));
                } else {
                    v[r] = int16_pair_to_float2(w, a.exp23);
                }
            }
            wbase = base + 16 * j;                     // out[16 j + r]
            jlow = j & 7;
        } else {
            constexpr int ns = (PASS == 1) ? 16 : 256;
            const int k = j % ns;
            const float2 *tw = (PASS == 1) ? (a.tw1 + k) : (a.tw2 + k);
            float2 t[16];
#pragma unroll
            for (int r = 1; r < 16; ++r) t[r] = __ldg(tw + r * ns);
            // in[j + NB r].  NB = 256: 256 r is a whole number of 8-row groups, one swizzled base.
            // NB = 16 (L = 256): j < 16 is the column, row r: chunk XOR is r & 7.
            const float2 *p = buf + swz(base + j);
            const float2 *prow = buf + base;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const float2 x = (P::NB == 256) ? p[256 * r] : prow[16 * r + (j ^ ((r & 7) << 1))];
                v[r] = (r > 0) ? cmul(x, t[r]) : x;
            }
            wbase = base + (j / ns) * ns * 16 + k;     // out[.. + r ns]
            jlow = k;
        }
    };
    auto compute = [&](int q, float2 (&o)[16], int &wbase, int &jlow) {
        float2 v[16];
        gather(q, v, wbase, jlow);
        dft16(v, o);
    };
    auto store = [&](const float2 (&o)[16], int wbase, int jlow) {
        if (PASS == 0) {
            // 16 consecutive elements = one 128-byte row: eight 16-byte stores, chunk c -> c ^ (row & 7)
            float4 *row = reinterpret_cast<float4 *>(buf + wbase);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                row[c ^ jlow] = make_float4(o[2 * c].x, o[2 * c].y, o[2 * c + 1].x, o[2 * c + 1].y);
        } else if (PASS == 1) {
            // out[w + 16 r], w = (8-row-aligned) + k with k < 16: row advances by r, chunk XOR is r & 7
            float2 *prow = buf + (wbase - jlow);
#pragma unroll
            for (int r = 0; r < 16; ++r) prow[16 * r + (jlow ^ ((r & 7) << 1))] = o[r];
        } else {
            float2 *p = buf + swz(wbase);              // out[j + 256 r]: same positions as read
#pragma unroll
            for (int r = 0; r < 16; ++r) p[256 * r] = o[r];
        }
    };
    if (HAZARD && FRA_K2_EARLY_BARRIER) {
        // the barrier sits between the reads and the butterflies (not between the butterflies and the writes): every
        // thread has its inputs in registers, and the scattered 8-byte stores then leave one butterfly at a time,
        // interleaved with the other one's arithmetic, instead of 32 per thread in one burst behind the barrier
        // (that burst was 7.6 % of the kernel's warp samples, 59 % of them mio_throttle)
        float2 v[IPT][16];
        int wbase[IPT], jlow[IPT];
#pragma unroll
        for (int q = 0; q < IPT; ++q) gather(q, v[q], wbase[q], jlow[q]);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            float2 o[16];
            dft16(v[q], o);
            store(o, wbase[q], jlow[q]);
        }
    } else if (HAZARD || FRA_K2_SEQ == 0) {
        float2 o[IPT][16];
        int wbase[IPT], jlow[IPT];
#pragma unroll
        for (int q = 0; q < IPT; ++q) compute(q, o[q], wbase[q], jlow[q]);
        if (HAZARD) __syncthreads();
#pragma unroll
        for (int q = 0; q < IPT; ++q) store(o[q], wbase[q], jlow[q]);
    } else if (FRA_K2_SEQ == 2 && !PAIRED) {
#pragma unroll 1
        for (int q = 0; q < IPT; ++q) {
            float2 o[16];
            int wbase, jlow;
            compute(q, o, wbase, jlow);
            store(o, wbase, jlow);
        }
    } else {
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            float2 o[16];
            int wbase, jlow;
            compute(q, o, wbase, jlow);
            store(o, wbase, jlow);
        }
    }
    __syncthreads();
}

// ---------------------------------------------------- thread-block cluster helpers (N = 32768 / 65536 kernels below)
#ifndef FRA_HOST_EMUL
FRA_DEV unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
FRA_DEV void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the partner CTA's copy of a shared-memory address of this CTA, as a generic pointer (DSMEM window)
FRA_DEV const float2 *cluster_map(const float2 *p, unsigned rank)
{
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)(uintptr_t)p), "r"(rank));
    return reinterpret_cast<const float2 *>((uintptr_t)out);
}

#endif
template <class P, bool CLUSTER>
struct ClusterRanks {
    static constexpr int value = 1;
};
template <class P>
struct ClusterRanks<P, true> {
    static constexpr int value = P::RANKS;
};

template <bool B>
struct BoolTag {
    static constexpr bool value = B;
};

// OUT = 0: int16 frames only (the hot configuration); OUT = 1: any combination of
// outputs, selected at run time by the null pointers in K2Args.
// the last pass of the CTA's frames [frame0, frame0 + FPC): radix-F combine, untangle, mirror, pack.
// CLUSTER (FftPlanCluster): the F sub-sequences live in the buffers of the cluster's RANKS CTAs - f / FLOC == rank in
// `buf`, the others in the buffer of CTA f / FLOC (`buf` mapped into that CTA's distributed-shared-memory window) -
// and each CTA does 1 / RANKS of the items.
template <class P, int QMODE, int OUT, bool CLUSTER = false>
FRA_DEV void fft_last_pass(const K2Args &a, const float2 *buf, int tid, int frame0, int rank = 0)
{
    constexpr int RANKS = ClusterRanks<P, CLUSTER>::value;
    const float2 *peer[RANKS];
#pragma unroll
    for (int r = 0; r < RANKS; ++r) peer[r] = buf;
#if !defined(FRA_HOST_EMUL)
    if constexpr (CLUSTER) {
#pragma unroll
        for (int r = 0; r < RANKS; ++r) peer[r] = (r == rank) ? buf : cluster_map(buf, (unsigned)r);
    }
#endif

    // ---------------- last pass: radix-F combine + untangle + mirror + pack
    BinOut out;
    out.frames = a.frames; out.qscale = a.qscale;
    out.iq = (OUT == 0) ? nullptr : a.iq;
    out.mag = (OUT == 0) ? nullptr : a.mag;
    out.phase = (OUT == 0) ? nullptr : a.phase;
    out.mag_alpha = a.mag_alpha;

    // one work item: the pair of sub-FFT bins (k, L-k) of frame `fr`, all F sub-sequences.
    // MAIN (a std::true_type-like tag): 0 < k < L/2, so every bin has its mirror partner and is written once; `sk` /
    // `skm` = swz(k) / swz(L - k) then come from the caller (they advance by the item stride, see below)
    auto item = [&](auto main_tag, int fr, int k, float2 wn, int sk_in, int skm_in) {
        constexpr bool MAIN = decltype(main_tag)::value;
        const int frame = frame0 + fr;
        const size_t up = (size_t)frame * P::N + k;
        const size_t dn = (size_t)frame * P::N - k;
        const int km = (P::L - k) & (P::L - 1);
        const int sk = MAIN ? sk_in : swz(k), skm = MAIN ? skm_in : swz(km);   // f L is a multiple of 8 rows
        float2 za[P::F], zb[P::F];
        // W_M^(f k) = W_N^(2 f k) from W_N^k by squaring and multiplying (f = 2 m: the square of m's, f = 2 m + 1: f - 1's
        // times W_M^k) instead of F - 2 gathers of stride 2 f from the table: those touched up to 28 cache lines per warp
        // load at F = 8 and put the last pass of the 32K / 64K kernels behind lg_throttle / long_scoreboard; at most five
        // roundings deep, far inside the FFT's own rounding error
        float2 wp[P::F > 1 ? P::F : 2];
        wp[1] = make_float2(wn.x * wn.x - wn.y * wn.y, 2.0f * wn.x * wn.y);
#pragma unroll
        for (int f = 2; f < P::F; ++f)
            wp[f] = (f & 1) ? cmul(wp[f - 1], wp[1])
                            : make_float2(wp[f / 2].x * wp[f / 2].x - wp[f / 2].y * wp[f / 2].y, 2.0f * wp[f / 2].x * wp[f / 2].y);
#pragma unroll
        for (int f = 0; f < P::F; ++f) {
            if (CLUSTER) {
                const float2 *b = peer[f / P::FLOC];
                za[f] = b[(f % P::FLOC) * P::L + sk];
                zb[f] = b[(f % P::FLOC) * P::L + skm];
            } else {
                za[f] = buf[fr * P::M + f * P::L + sk];
                zb[f] = buf[fr * P::M + f * P::L + skm];
            }
            if (f > 0) {
                const float2 w = wp[f];
                za[f] = cmul(za[f], w);
                // W_M^(f (L-k)) = W_F^f * conj(W_M^(f k))
                float2 t = cmulc(zb[f], w);
                if (P::F == 2) t = make_float2(-t.x, -t.y);                          // W_2^1 = -1
                if (P::F == 4) {
                    if (f == 1) t = mul_mi(t);                                        // -i
                    if (f == 2) t = make_float2(-t.x, -t.y);
                    if (f == 3) t = make_float2(-t.y, t.x);                           // +i
                }
                if (P::F == 8) {
                    const float h = 0.70710678118654752f;
                    if (f == 1) t = make_float2((t.x + t.y) * h, (t.y - t.x) * h);
                    if (f == 2) t = mul_mi(t);
                    if (f == 3) t = make_float2((t.y - t.x) * h, -(t.x + t.y) * h);
                    if (f == 4) t = make_float2(-t.x, -t.y);
                    if (f == 5) t = make_float2(-(t.x + t.y) * h, (t.x - t.y) * h);
                    if (f == 6) t = make_float2(-t.y, t.x);
                    if (f == 7) t = make_float2((t.x - t.y) * h, (t.x + t.y) * h);
                }
                zb[f] = t;
            }
        }
        dft_small<P::F>(za);       // za[q] = Z[k + L q]
        dft_small<P::F>(zb);       // zb[q] = Z[(L - k) + L q]
#pragma unroll
        for (int q = 0; q < P::F; ++q) {
            const float2 A = za[q];
            const float2 B = cconj(zb[P::F - 1 - q]);
            const float2 fe = cadd(A, B);                              // 2 Fe
            const float2 fo = mul_mi(csub(A, B));                      // 2 Fo
            // W_N^(k + L q) = W_N^k * W_(2F)^q
            float2 w = wn;
            if (q > 0) w = cmul(wn, w16(q * (8 / P::F)));
            const float2 t = cmul(w, fo);
            const float2 p = cadd(fe, t);                              // 2 X[j],     j = k + L q
            const float2 m = csub(fe, t);                              // 2 X[M + j]
            // X[j], X[N - j] (j = 0: no partner), X[M + j], X[M - j].  For the self-paired items k = 0
            // and k = L/2 the mirrored indices are also another q's direct ones (N - j = M + j',
            // M - j = j'); both writers run in this thread and the mirrored one keeps X[N - k] =
            // conj(X[k]) bit-exact, so both are kept - except when the magnitude output averages,
            // which is a read-modify-write and must touch every bin exactly once
            const bool once = !MAIN && out.mag_alpha < 1.0f && (k == 0 || k == P::L / 2);
            emit_pair<QMODE, OUT == 0>(out, up, dn, P::L * q, P::N - P::L * q, MAIN || (!once && ((q > 0) || (k != 0))), p);
            emit_pair<QMODE, OUT == 0>(out, up, dn, P::M + P::L * q, P::M - P::L * q, MAIN || !once, m);
        }
    };
    using TagMain = BoolTag<true>;
    using TagEdge = BoolTag<false>;

    // k = 1 .. L/2 - 1: uniform work (in a cluster the two CTAs take alternate blocks of THREADS items).  The trip
    // count is a compile-time constant and the loop is unrolled: the item stride is a multiple of 128 elements, which
    // leaves the swizzle's row bits alone - swz(k + i STEP) = swz(k) + i STEP, swz(L - k - i STEP) = swz(L - k) - i STEP -
    // so shared-memory addresses, the twiddle address and all eight store addresses of an item are one base register
    // plus an immediate (the rolled loop spent 30 of its 100 instructions per item on addresses, bounds and branches).
    {
        constexpr int HALF = P::L / 2;
        constexpr int SLOTS = P::FPC * HALF;
        constexpr int STEP = RANKS * P::THREADS;
        constexpr int TRIPS = SLOTS / STEP;
        static_assert(SLOTS % STEP == 0 && STEP % 128 == 0, "last pass: whole trips, swizzle-preserving stride");
        const int slot0 = tid + (CLUSTER ? rank * P::THREADS : 0);
        // (with the optional outputs an item is several times longer and full of run-time branches: unrolling it only
        // costs instruction-cache space - the complex64 FFT-only sweep lost 8 % at 16K - so that variant stays rolled)
        constexpr int UNROLL = (OUT == 0) ? TRIPS : 1;
        if constexpr (STEP <= HALF) {
            // one frame per CTA (FPC = 1 whenever L = 4096 ... or two frames of 8K): k advances, the frame is fixed per trip
            static_assert(HALF % STEP == 0, "last pass: trips per frame");
            const int k0 = slot0 % HALF;
            const int sk0 = swz(k0), skm0 = swz(P::L - k0);           // (k0 = 0: L itself, so that L - ii follows from it)
            float2 wn_next = __ldg(a.twn + k0);                          // the next item's twiddle, one item ahead
#pragma unroll UNROLL
            for (int i = 0; i < TRIPS; ++i) {
                const int fr = (i * STEP) / HALF;
                const int ii = (i * STEP) % HALF;                        // (unrolled: compile-time) offset of k within the frame
                const int k = k0 + ii;
                const float2 wn = wn_next;
                if (i + 1 < TRIPS) wn_next = __ldg(a.twn + k0 + ((i + 1) * STEP) % HALF);
                const bool live = (P::FPC == 1) || (frame0 + fr < a.batch);
                if ((ii > 0 || k0 != 0) && live) item(TagMain(), fr, k, wn, sk0 + ii, skm0 - ii);
            }
        } else {
            // several frames per trip (L = 256): k is fixed per thread, the frame advances
            static_assert(STEP % HALF == 0, "last pass: frames per trip");
            const int k = slot0 % HALF;
            const int sk = swz(k), skm = swz((P::L - k) & (P::L - 1));
            const float2 wn = __ldg(a.twn + k);
#pragma unroll UNROLL
            for (int i = 0; i < TRIPS; ++i) {
                const int fr = slot0 / HALF + i * (STEP / HALF);
                if (k != 0 && frame0 + fr < a.batch) item(TagMain(), fr, k, wn, sk, skm);
            }
        }
    }
    // the two self-paired items k = 0 and k = L/2 of every frame in the CTA (in a cluster: one each)
    if (tid < 2 * P::FPC && (!CLUSTER || (tid & 1) == rank)) {        // (in a cluster: ranks 0 and 1 take one each)
        const int fr = tid >> 1;
        const int k = (tid & 1) ? (P::L / 2) : 0;
        if (frame0 + fr < a.batch) item(TagEdge(), fr, k, __ldg(a.twn + k), 0, 0);
    }
}

template <int LOG2N, bool WIN, int QMODE, int OUT>
__global__ void __launch_bounds__(FftPlan<LOG2N>::THREADS, (LOG2N == 15) ? 1 : FRA_K2_MINBLOCKS) k2_fft(K2Args a)
{
    using P = FftPlan<LOG2N>;
    FRA_DYN_SMEM(smem_raw);
    float2 *buf = reinterpret_cast<float2 *>(smem_raw);
    const int tid = threadIdx.x;
    const int frame0 = blockIdx.x * P::FPC;
#ifdef FRA_TIMELINE
    timeline_mark(2, a.tl_step);
#endif
#if !defined(FRA_HOST_EMUL)
    // the frames of the CTA that will take this one's place (a.prefetch CTAs ahead = the number resident on the
    // whole GPU): one 128-byte line per thread into L2, so that its pass 0 waits for an L2 hit instead of DRAM -
    // 16 warps per SM do not cover ~800 cycles of DRAM latency (1.627 -> 1.592 ms per 65536 frames)
    if (a.prefetch > 0) {
        const size_t ahead = (size_t)frame0 + (size_t)a.prefetch * P::FPC;
        const size_t byte = (size_t)tid * 128;
        if (ahead + P::FPC <= (size_t)a.batch && byte < (size_t)P::FPC * P::M * 4)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(a.in + ahead * P::M) + byte));
    }
#endif
    // ---------------------------------------------- radix-16 Stockham passes
    // (compile-time pass index: twiddle strides, swizzled offsets and the pass-specific store
    // pattern are all immediates; a run-time pass loop cost ~4 issue slots per sample)
    fft_pass<P, WIN, 0>(a, buf, tid, frame0);
    fft_pass<P, WIN, 1>(a, buf, tid, frame0);
    if (P::PASSES == 3) fft_pass<P, WIN, 2>(a, buf, tid, frame0);
    fft_last_pass<P, QMODE, OUT>(a, buf, tid, frame0);
#ifdef FRA_TIMELINE
    timeline_mark(3, a.tl_step);
#endif
}

// The same for N = 16384 with PERSISTENT CTAs (two per SM) and the int16 frame staged by the TMA
// engine: one thread issues ONE 32 KiB cp.async.bulk (SASS UBLKCP) per frame into a raw buffer
// behind the FFT buffer, completing on an mbarrier, and it does so for frame i + 1 as soon as
// pass 0 of frame i has consumed the buffer - so the DRAM latency of a frame's input is hidden
// behind the three remaining passes of the frame before it, no thread spends an issue slot or a
// register on global loads of samples, and pass 0 reads shared memory at unit stride.
// (With per-thread loads 16 warps per SM could not cover that latency: long_scoreboard was the
// first stall reason of k2_fft, profiles/r02_k2_fft_65536ch.txt.)
constexpr int kStagedLog2N = 14;
constexpr int kStagedRawBytes = FftPlan<kStagedLog2N>::M * 4;                       // the int16 frame: 32 KiB
constexpr int kStagedSmemBytes = FftPlan<kStagedLog2N>::SMEM_BYTES + kStagedRawBytes + 16;

template <bool WIN, int QMODE, int OUT>
__global__ void __launch_bounds__(FftPlan<kStagedLog2N>::THREADS, FRA_K2_MINBLOCKS) k2_fft_staged(K2Args a)
{
    using P = FftPlan<kStagedLog2N>;
    FRA_DYN_SMEM(smem_raw);
    float2 *buf = reinterpret_cast<float2 *>(smem_raw);
    uint2 *raw = reinterpret_cast<uint2 *>(smem_raw + P::SMEM_BYTES);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + P::SMEM_BYTES + kStagedRawBytes);
    const int tid = threadIdx.x;
#ifdef FRA_TIMELINE
    timeline_mark(2, a.tl_step);
#endif
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    unsigned phase = 0;
    int frame = blockIdx.x;
    if (tid == 0 && frame < a.batch) {
        mbar_expect_tx(bar, (unsigned)kStagedRawBytes);
        bulk_g2s(raw, a.in + (size_t)frame * P::M, (unsigned)kStagedRawBytes, bar);
    }
    for (; frame < a.batch; frame += (int)gridDim.x) {
        mbar_wait(bar, phase);                               // this frame's samples have landed
        phase ^= 1u;
        fft_pass<P, WIN, 0>(a, buf, tid, frame, raw);      // ends with a barrier: everyone has read `raw`
        const int next = frame + (int)gridDim.x;
        if (tid == 0 && next < a.batch) {
            fence_proxy_async();                             // the generic-proxy reads above before the bulk copy's writes
            mbar_expect_tx(bar, (unsigned)kStagedRawBytes);
            bulk_g2s(raw, a.in + (size_t)next * P::M, (unsigned)kStagedRawBytes, bar);
        }
        fft_pass<P, WIN, 1>(a, buf, tid, frame);
        fft_pass<P, WIN, 2>(a, buf, tid, frame);
        fft_last_pass<P, QMODE, OUT>(a, buf, tid, frame);
        __syncthreads();                                     // the last pass has read `buf` before the next frame's pass 0 writes it
    }
#ifdef FRA_TIMELINE
    timeline_mark(3, a.tl_step);
#endif
}

// ---------------------------------------------------- N = 32768 / 65536 on chip (thread-block cluster)
// One frame per R-CTA thread-block cluster.  CTA `rank` owns the sub-sequences f = FLOC rank .. FLOC rank + FLOC - 1
// of the F (z[F m + f], 4096 points each): three radix-16 passes in its own shared memory, exactly the code of
// the single-CTA sizes.  After a cluster barrier the last pass - radix-F combine, untangle, mirror, pack - takes
// 1 / R of the items in each CTA and reads the other CTAs' sub-sequences through distributed shared memory
// (the SM-to-SM network); a second cluster barrier keeps all buffers alive until every CTA is done.  One HBM round
// trip per frame (2 B in, 4 B out per sample) - for 64K frames instead of the 28 B of the three-kernel path below.
// a.twn = W_N^e.
#ifndef FRA_HOST_EMUL
template <int LOG2N, int R, bool WIN, int QMODE, int OUT>
__global__ void __cluster_dims__(R, 1, 1) __launch_bounds__(FftPlanCluster<LOG2N, R>::THREADS, FftPlanCluster<LOG2N, R>::MINBLOCKS)
    k2_fft_cluster(K2Args a)
{
    using P = FftPlanCluster<LOG2N, R>;
    FRA_DYN_SMEM(smem_raw);
    float2 *buf = reinterpret_cast<float2 *>(smem_raw);
    const int tid = threadIdx.x;
    const int rank = (int)cluster_ctarank();
    const int frame = blockIdx.x / R;                        // all CTAs of a cluster: the same frame, never past the batch
    fft_pass<P, WIN, 0>(a, buf, tid, frame, nullptr, P::FLOC * rank);
    fft_pass<P, WIN, 1>(a, buf, tid, frame, nullptr, P::FLOC * rank);
    fft_pass<P, WIN, 2>(a, buf, tid, frame, nullptr, P::FLOC * rank);
    cluster_barrier();                                       // every part of the frame is transformed
    fft_last_pass<P, QMODE, OUT, true>(a, buf, tid, frame, rank);
    cluster_barrier();                                       // nobody leaves while another CTA still reads its buffer
}
#endif

// ------------------------------------------------------------------ N = 65536
// A 64K frame is 32768 complex points = 256 KiB of fp32, more than an SM's shared memory, so it
// is done as one decimation-in-time step around the 32K kernel:
//   k2_split64k    even / odd samples of every frame -> two 32K frames (window applied here in
//                  bypass mode: the ROM index is the sample's index in the 64K frame)
//   k2_fft<15,...> on 2 x batch frames, fp32 bins to scratch
//   k2_join64k     X[k] = E[k] + W_65536^k O[k],  X[k + 32768] = E[k] - W_65536^k O[k],
//                  then the same quantise / pack / mag / phase as the single-kernel sizes
// 28 B of HBM traffic per sample instead of 6: the price of not fitting on chip.
constexpr int kHalf64k = 32768;

template <bool WIN>
__global__ void __launch_bounds__(256) k2_split64k(const int16_t *in, int16_t *out, const int *rom32, size_t total8)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread = 8 consecutive samples
    if (i >= total8) return;
    const size_t e = i * 8;
    const size_t frame = e / (2 * kHalf64k);
    const int n0 = (int)(e % (2 * kHalf64k));
    const uint4 x = ldg128(in + e);
    int v[8] = {lo16(x.x), hi16(x.x), lo16(x.y), hi16(x.y), lo16(x.z), hi16(x.z), lo16(x.w), hi16(x.w)};
    if (WIN) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = window_int(v[j], __ldg(rom32 + ((n0 + j) & (kWindowLen - 1))));
    }
    uint2 ev, od;
    ev.x = pack16((unsigned)v[0], (unsigned)v[2]); ev.y = pack16((unsigned)v[4], (unsigned)v[6]);
    od.x = pack16((unsigned)v[1], (unsigned)v[3]); od.y = pack16((unsigned)v[5], (unsigned)v[7]);
    int16_t *even = out + (2 * frame) * kHalf64k + n0 / 2;
    *reinterpret_cast<uint2 *>(even) = ev;
    *reinterpret_cast<uint2 *>(even + kHalf64k) = od;
}

template <int QMODE>
__global__ void __launch_bounds__(256) k2_join64k(const float2 *halves, const float2 *twc, BinOut out, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread = bins k and k + 32768 of a frame
    if (i >= total) return;
    const size_t frame = i / kHalf64k;
    const int k = (int)(i % kHalf64k);
    const float2 e = halves[(2 * frame) * kHalf64k + k];
    const float2 o = cmul(halves[(2 * frame + 1) * kHalf64k + k], __ldg(twc + k));
    const size_t base = frame * (2 * (size_t)kHalf64k) + k;
    // emit_pair takes 2 X (the untangle's convention) and halves it
    emit_pair<QMODE>(out, base, 0, 0, 0, false, make_float2(2.0f * (e.x + o.x), 2.0f * (e.y + o.y)));
    emit_pair<QMODE>(out, base, 0, kHalf64k, 0, false, make_float2(2.0f * (e.x - o.x), 2.0f * (e.y - o.y)));
}

// ------------------------------------------------- half-spectrum host transfer (FRA_HOST_HALF_SPECTRUM)
// The input is real, so X[N-j] = conj(X[j]) - up to the truncation: the frame holds floor(Im s) at j and
// floor(-Im s) at N - j, which is -floor(Im s) - 1 unless Im s is an integer.  One bit per bin says which:
// bit j = (im[j] + im[N-j] != 0 mod 2^16), taken from the finished frame itself, so the host can rebuild the
// upper half exactly from bins 0..N/2 and N/2 bits per frame - 33 KiB over PCIe instead of 64 KiB.
// One thread per lower-half bin, one word per warp (ballot).
__global__ void __launch_bounds__(256) k3_mirror_bits(const uint32_t *frames, uint32_t *bits, size_t total, int log2m)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // total is a multiple of 32
    if (t >= total) return;
    const size_t frame = t >> log2m;
    const unsigned j = (unsigned)(t & (((size_t)1 << log2m) - 1));
    const uint32_t *fr = frames + (frame << (log2m + 1));
    unsigned nz = 0;
    if (j != 0) {
        const unsigned a = fr[j] >> 16, b = fr[((size_t)2 << log2m) - j] >> 16;
        nz = ((a + b) & 0xFFFFu) != 0;
    }
    const unsigned word = __ballot_sync(0xffffffffu, nz);
    if ((threadIdx.x & 31) == 0) bits[t >> 5] = word;
}

}  // namespace fra
