// fra_api.cu - the C ABI of libfra.so (include/fra.h): context, control-plane
// byte decoder, kernel dispatch.  Host code only; kernels are in k1_*.cuh / k2_fft.cuh.
//
// Control plane restates NEW/command_control.vhd:46-78 (mode / reset / start),
// NEW/rx_filter_coeff.vhd:41-66 (0xF1 + 12 bytes, busy), NEW/filter_iir12_cust.vhd:
// 48-60,83-94 (register map) and IMP/sequ2.vhd:83-96,216 (transport, request).
#include "../../include/fra.h"

#include "k1_window_iir.cuh"
#include "k1b_stream.cuh"
#include "k2_fft.cuh"
#include "k2_fixed.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <new>
#include <thread>
#include <utility>
#include <vector>
#if !defined(FRA_HOST_EMUL) && (defined(__x86_64__) || defined(_M_X64))
#include <immintrin.h>
#define FRA_HAVE_AVX2_PATH 1
#endif

namespace {

using namespace fra;

const int16_t kHannRom[kWindowLen] = {
#include "hann_rom_q15.inc"
};

// IMP/filter_pkg.vhd:54-68 in the register order of NEW/filter_iir12_cust.vhd:83-94
const int8_t kBank0[12] = {-14, 0, 14, 107, 21, 127, -15, 0, 15, 107, -21, 127};

constexpr int kStreamMaxDeadband = 32;            // LSB; larger boundary differences mean the scan is invalid
// k1_lane wins only in a window of channel counts: below it k1_duo is faster outright (k1_lane is flat at
// 1.0 ms up to ~19k channels), above it k1_duo's asymptote is higher (561 vs 535 Gsamples/s at 65536 channels);
// in between k1_duo's wave quantisation (296 CTAs of 32 channels per wave) costs more than the difference
// (32768 channels: 1.09 vs 1.03 ms, 49152: 1.63 vs 1.53 ms)
constexpr int kLaneMinChannels = 28672;
constexpr int kLaneMaxChannels = 57344;
constexpr int kLaneBiasedMinChannels = 24576;     // k1_lane_biased from here up (measured crossover between 16384 and 32768)
// ... and in FRA_PIPELINE mode from here up: beside the FFT of the previous call the lane kernel's lone, latency-bound warps
// leave the FFT most of the issue slots, whereas two 96 KiB k1_duo CTAs leave an SM no room for an FFT CTA at all
// (whole steps, ms: 12288 channels 0.659 vs 0.612 for k1_duo; 14336: 0.689 vs 0.711; 16384: 0.724 vs 0.814;
// 20480: 0.951 vs 1.011 - profiles/r02_session4_k1_choice_pipelined.txt)
constexpr int kLaneBiasedMinChannelsPipelined = 13312;

}  // namespace

struct fra_ctx {
    int device = 0;
    int channels = 0;
    int n = 0;
    int log2n = 0;
    unsigned flags = 0;
    int sm_count = 0;

    // control plane (what the RTL keeps in registers)
    uint8_t mode = FRA_MODE_BYPASS;
    uint8_t transport = FRA_CMD_ETHERNET_MODE;
    float mag_alpha = 1.0f;          // fra_set_mag_average
    int8_t bank1[12] = {0};
    int8_t sections[6][6] = {{0}};  // fra_load_sections: six independent sections (superset of the 12-byte bank)
    bool sections_loaded = false;   // bank 1 is `sections` instead of `bank1` alternated
    int upload_pos = -1;            // >= 0: inside a 0xF1 upload, bytes received so far
    int8_t upload_buf[12] = {0};
    uint64_t n_start = 0, n_request = 0, n_reset = 0, n_upload = 0;

    // device resources
    cudaStream_t stream = nullptr;
    cudaStream_t copy_streams[3] = {nullptr, nullptr, nullptr};
    int *d_rom32 = nullptr;
    int *d_rom2x = nullptr;           // 2 * ROM: the window straight to the biased float (window_biased)
    int16_t *d_state = nullptr;       // [C][6][4]
    int16_t *d_scratch = nullptr;     // [C][N] filter output when the caller does not ask for it
    float2 *d_tw1 = nullptr, *d_tw2 = nullptr, *d_twn = nullptr;
    uint32_t *d_twfx = nullptr;       // FRA_FFT_FIXED16: W_N^t as packed Q1.15 pairs
    // 64K frames: W_65536^k and the scratch of the even/odd decomposition (k2_fft.cuh)
    float2 *d_twc = nullptr, *d_halves = nullptr;
    int16_t *d_split = nullptr;
    size_t split_frames = 0;
    // staging for fra_process_host
    int16_t *d_in = nullptr;
    uint8_t *d_frames = nullptr;
    int16_t *d_filtered_out = nullptr;
    float *d_iq = nullptr, *d_mag = nullptr, *d_phase = nullptr;
    // K1b work space
    int16_t *d_entry = nullptr, *d_exit = nullptr;
    float *d_ends = nullptr, *d_aggr = nullptr, *d_mats = nullptr;
    int *d_counts = nullptr;
    int k1b_capacity = 0;
    // the scan matrices on the device belong to these coefficients / this chunk length (recomputed only when they change)
    int8_t k1b_key[36] = {0};
    int k1b_key_chunk = 0;
    double k1b_norm[32] = {0};        // max |entry| of each matrix of the table

    // FRA_PIPELINE: window+IIR of call i+1 beside the FFT of call i
    cudaStream_t pipe_k1 = nullptr, pipe_k2 = nullptr;
    cudaEvent_t pipe_in = nullptr, pipe_k1_done[2] = {nullptr, nullptr}, pipe_k2_done[2] = {nullptr, nullptr};
    unsigned long long pipe_calls = 0;
    cudaEvent_t pipe_go = nullptr;    // recorded on pipe_k1 right before a window+IIR launch
    bool fft_pending = false;         // the FFT of the previous call, launched behind the next call's window+IIR
    K2Args fft_args;
    bool fft_win = false;
    int fft_qmode = 0, fft_buf = 0;
    // fra_process_host_async: completion events per call slot (two calls in flight) and copy stream
    cudaEvent_t host_done[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    unsigned long long host_calls = 0;
    size_t scratch_elems = 0;         // int16 elements allocated at d_scratch
    // FRA_HOST_HALF_SPECTRUM: mirror bits on the device, their pinned host copies per call slot, and the frames
    // (host pointer) of each slot that still wait for their upper half
    int16_t *d_skew_in = nullptr;     // FRA_WINDOW_RTL_SKEW: the input delayed by one sample, [C][N]
    int16_t *d_skew_prev = nullptr;   // ... and every channel's last sample of the previous frame
    uint32_t *d_mbits = nullptr;
    uint32_t *h_mbits[2] = {nullptr, nullptr};
    // the frames of each call slot that still wait for their upper half: in every channel slice [s per, s per + per)
    // the first nh[s] frames crossed the link as half spectra
    struct MirrorPlan {
        uint8_t *frames = nullptr;
        size_t per = 0, nh[8] = {0};
        int n_slices = 0;
    } mirror_plan[2];
    // Share of the frames that travel as half spectra; the rest go whole.  More half spectra = less device-to-host
    // link traffic (2.06 instead of 4 B per sample) but more host memory traffic (the mirror reads and writes what the
    // DMA would only have written: 8.06 instead of 6 B per sample in total), and which of the two binds depends on the
    // host.  Adaptive mode measures it: a short search over the cadence of the caller's fra_process_host_async calls -
    // three calls per trial share, the interval before the next trial's first call is the trial's step time
    // (1, 3/4, 1/2, then the two neighbours of the best at 1/8) - and then keeps the best share.
    double half_frac = 1.0;
    bool half_adaptive = true;
    struct HalfTuner {
        int calls = 0;                     // half-spectrum calls seen since the search started
        double trial_share[5] = {1.0, 0.75, 0.5, 0.0, 0.0};
        double trial_time[5] = {0, 0, 0, 0, 0};
        int n_trials = 3;
        bool done = false;
        std::chrono::steady_clock::time_point last_entry;
    } tuner;
    uint64_t last_h2d_bytes = 0, last_d2h_bytes = 0;
    cudaEvent_t slice_done[2][8] = {{nullptr}};      // per call slot and channel slice: its device-to-host copies
    double last_wait_s = 0.0, last_mirror_s = 0.0;   // of the newest finished call: blocked on its copies / mirroring

    bool profiling = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // K1 begin/end, K2 begin/end
    bool ev_k1 = false, ev_k2 = false;

    // kernels whose dynamic shared-memory limit has been raised (once per context, not per launch)
    std::vector<std::pair<const void *, int>> smem_set;

    int last_kernels = 0;
    char err[256] = {0};
};

namespace {

int fail_cuda(fra_ctx *ctx, cudaError_t e, const char *what)
{
    if (ctx) std::snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
    return FRA_ERR_CUDA;
}

#define FRA_TRY(ctx, expr)                                         \
    do {                                                           \
        cudaError_t e_ = (expr);                                   \
        if (e_ != cudaSuccess) return fail_cuda((ctx), e_, #expr); \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel and context
template <class K>
int ensure_smem(fra_ctx *ctx, K kfn, int bytes)
{
    const void *key = reinterpret_cast<const void *>(kfn);
    for (auto &e : ctx->smem_set)
        if (e.first == key && e.second >= bytes) return FRA_OK;
    FRA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    ctx->smem_set.emplace_back(key, bytes);
    return FRA_OK;
}

#define FRA_SMEM(ctx, kfn, bytes)                       \
    do {                                                \
        int rc_ = ensure_smem((ctx), (kfn), (bytes));   \
        if (rc_ != FRA_OK) return rc_;                  \
    } while (0)

StageCoef make_stage(const int8_t *k)       // k = B0,B1,B2,A0,A1 (NEW/filter_iir_cust.vhd:104-108)
{
    StageCoef c;
    c.b0 = (float)k[0] / 128.0f;
    c.b1 = (float)k[1] / 128.0f;
    c.b2 = (float)k[2] / 128.0f;
    c.na0 = -(float)k[3] / 128.0f;
    c.na1 = -(float)k[4] / 128.0f;
    c.exp23 = 0x4B000000u;
    c.k0 = kMagicB + (float)k[4] * 65792.0f;
    c.kb = (float)(12615680 - 65792 * ((int)k[0] + (int)k[1] + (int)k[2] - (int)k[3] - (int)k[4]));   // exact: a multiple of 256 below 2^26
    return c;
}

struct Sections {
    int8_t c[kStages][6];    // B0,B1,B2,A0,A1,A2 per stage
};

// the 12-byte bank as the RTL wires it: ALPHA (bytes 0..5) -> stages 1,3,5, BETA (bytes 6..11) -> stages 2,4,6
// (NEW/filter_iir12_cust.vhd:68-240; bytes 5 and 11 = A2 are unconnected)
Sections alternate(const int8_t coeff12[12])
{
    Sections s;
    for (int i = 0; i < kStages; ++i) std::memcpy(s.c[i], coeff12 + 6 * (i & 1), 6);
    return s;
}

// the coefficients the current mode filters with
Sections current_sections(const fra_ctx *ctx)
{
    if (ctx->mode == FRA_MODE_BANK1) {
        if (!ctx->sections_loaded) return alternate(ctx->bank1);
        Sections s;
        std::memcpy(s.c, ctx->sections, sizeof(s.c));
        return s;
    }
    return alternate(kBank0);
}

CascadeCoef make_cascade(const Sections &s)
{
    CascadeCoef c;
    for (int i = 0; i < kStages; ++i) c.set[i] = make_stage(s.c[i]);
    return c;
}

void do_reset(fra_ctx *ctx)
{
    // rst_n pulse: NEW/command_control.vhd:50, NEW/filter_iir12_cust.vhd:51-52, IMP/sequ2.vhd:86
    ctx->mode = FRA_MODE_BYPASS;
    std::memset(ctx->bank1, 0, sizeof(ctx->bank1));
    std::memset(ctx->sections, 0, sizeof(ctx->sections));
    ctx->sections_loaded = false;
    ctx->transport = FRA_CMD_ETHERNET_MODE;
    ctx->n_reset++;
}

template <int LOG2N, bool WIN, int QMODE>
int launch_k2_inst(fra_ctx *ctx, const K2Args &args, cudaStream_t st)
{
    using P = FftPlan<LOG2N>;
    // frames-only is the hot configuration and gets its own instantiation
    const bool frames_only = args.frames && !args.iq && !args.mag && !args.phase;
    if constexpr (LOG2N == kStagedLog2N) {
        // 16K frames, on request (FRA_K2_STAGED): persistent CTAs, each frame staged by one bulk copy (k2_fft.cuh).
        // Measured slower than the one-frame-per-CTA kernel (1.79 vs 1.62 ms per 65536 frames), so
        // not the default.  The bulk copy needs a 16-byte aligned source.
        const bool staged = (ctx->flags & FRA_K2_STAGED) && !(ctx->flags & FRA_PIPELINE) && (reinterpret_cast<uintptr_t>(args.in) % 16) == 0;
        if (staged && args.batch > 0) {
            auto sfn = frames_only ? k2_fft_staged<WIN, QMODE, 0> : k2_fft_staged<WIN, QMODE, 1>;
            FRA_SMEM(ctx, sfn, kStagedSmemBytes);
            const int grid = std::min(args.batch, FRA_K2_MINBLOCKS * std::max(1, ctx->sm_count));
            FRA_LAUNCH(sfn, dim3(grid), dim3(P::THREADS), (size_t)kStagedSmemBytes, st, args);
            FRA_TRY(ctx, cudaGetLastError());
            ctx->last_kernels++;
            return FRA_OK;
        }
    }
    auto kfn = frames_only ? k2_fft<LOG2N, WIN, QMODE, 0> : k2_fft<LOG2N, WIN, QMODE, 1>;
    FRA_SMEM(ctx, kfn, P::SMEM_BYTES);
    const int grid = (args.batch + P::FPC - 1) / P::FPC;
    if (grid > 0) {
        FRA_LAUNCH(kfn, dim3(grid), dim3(P::THREADS), (size_t)P::SMEM_BYTES, st, args);
        FRA_TRY(ctx, cudaGetLastError());
        ctx->last_kernels++;
    }
    return FRA_OK;
}

template <int LOG2N>
int launch_k2_n(fra_ctx *ctx, const K2Args &args, bool win, int qmode, cudaStream_t st)
{
    if (win) {
        if (qmode == 0) return launch_k2_inst<LOG2N, true, 0>(ctx, args, st);
        if (qmode == 1) return launch_k2_inst<LOG2N, true, 1>(ctx, args, st);
        return launch_k2_inst<LOG2N, true, 2>(ctx, args, st);
    }
    if (qmode == 0) return launch_k2_inst<LOG2N, false, 0>(ctx, args, st);
    if (qmode == 1) return launch_k2_inst<LOG2N, false, 1>(ctx, args, st);
    return launch_k2_inst<LOG2N, false, 2>(ctx, args, st);
}

#ifndef FRA_HOST_EMUL
// N = 32768 / 65536 on chip: one frame per cluster of R CTAs (k2_fft.cuh: k2_fft_cluster)
template <int LOG2N, int R, bool WIN, int QMODE>
int launch_k2_cluster_inst(fra_ctx *ctx, K2Args args, cudaStream_t st)
{
    using P = FftPlanCluster<LOG2N, R>;
    const bool frames_only = args.frames && !args.iq && !args.mag && !args.phase;
    auto kfn = frames_only ? k2_fft_cluster<LOG2N, R, WIN, QMODE, 0> : k2_fft_cluster<LOG2N, R, WIN, QMODE, 1>;
    FRA_SMEM(ctx, kfn, P::SMEM_BYTES);
    if (LOG2N == 16) args.twn = ctx->d_twc;                  // W_65536^e
    if (args.batch <= 0) return FRA_OK;
    kfn<<<dim3((unsigned)R * (unsigned)args.batch), dim3(P::THREADS), (size_t)P::SMEM_BYTES, st>>>(args);    // __cluster_dims__(R, 1, 1)
    FRA_TRY(ctx, cudaGetLastError());
    ctx->last_kernels++;
    return FRA_OK;
}

template <int LOG2N, int R>
int launch_k2_cluster(fra_ctx *ctx, const K2Args &args, bool win, int qmode, cudaStream_t st)
{
    if (win) {
        if (qmode == 0) return launch_k2_cluster_inst<LOG2N, R, true, 0>(ctx, args, st);
        if (qmode == 1) return launch_k2_cluster_inst<LOG2N, R, true, 1>(ctx, args, st);
        return launch_k2_cluster_inst<LOG2N, R, true, 2>(ctx, args, st);
    }
    if (qmode == 0) return launch_k2_cluster_inst<LOG2N, R, false, 0>(ctx, args, st);
    if (qmode == 1) return launch_k2_cluster_inst<LOG2N, R, false, 1>(ctx, args, st);
    return launch_k2_cluster_inst<LOG2N, R, false, 2>(ctx, args, st);
}
#endif

// N = 65536 through HBM (FRA_K2_64K_SPLIT, and the host-emulated build): split into even / odd 32K frames,
// the 32K kernel on both, radix-2 join
int launch_k2_64k(fra_ctx *ctx, const K2Args &args, bool win, int qmode, cudaStream_t st)
{
#ifndef FRA_HOST_EMUL
    if (!(ctx->flags & FRA_K2_64K_SPLIT)) return launch_k2_cluster<16, 4>(ctx, args, win, qmode, st);
#endif
    // scratch is indexed by the frame's position in the context, so channel slices on different
    // streams (fra_process_host) do not share it
    const size_t frames = (size_t)args.batch;
    const size_t need = std::max<size_t>((size_t)ctx->channels, (size_t)args.frame0 + frames);
    if (ctx->split_frames < need) {
        FRA_TRY(ctx, cudaDeviceSynchronize());                 // other slices may still be using the old scratch
        if (ctx->d_split) FRA_TRY(ctx, cudaFree(ctx->d_split));
        if (ctx->d_halves) FRA_TRY(ctx, cudaFree(ctx->d_halves));
        ctx->d_split = nullptr;
        ctx->d_halves = nullptr;
        ctx->split_frames = 0;
        if (cudaMalloc((void **)&ctx->d_split, need * 2 * kHalf64k * sizeof(int16_t)) != cudaSuccess ||
            cudaMalloc((void **)&ctx->d_halves, need * 2 * kHalf64k * sizeof(float2)) != cudaSuccess)
            return FRA_ERR_NOMEM;
        ctx->split_frames = need;
    }
    int16_t *split = ctx->d_split + (size_t)args.frame0 * 2 * kHalf64k;
    float2 *halves = ctx->d_halves + (size_t)args.frame0 * 2 * kHalf64k;
    const size_t total8 = frames * (2 * kHalf64k / 8);
    const int16_t *in16 = reinterpret_cast<const int16_t *>(args.in);
    if (win) {
        auto kfn = k2_split64k<true>;
        FRA_LAUNCH(kfn, dim3((unsigned)((total8 + 255) / 256)), dim3(256), (size_t)0, st, in16, split,
                   (const int *)ctx->d_rom32, total8);
    } else {
        auto kfn = k2_split64k<false>;
        FRA_LAUNCH(kfn, dim3((unsigned)((total8 + 255) / 256)), dim3(256), (size_t)0, st, in16, split,
                   (const int *)ctx->d_rom32, total8);
    }
    FRA_TRY(ctx, cudaGetLastError());
    ctx->last_kernels++;

    K2Args half = args;
    half.in = reinterpret_cast<const uint32_t *>(split);
    half.frames = nullptr;
    half.iq = halves;
    half.mag = nullptr;
    half.phase = nullptr;
    half.batch = 2 * args.batch;
    int rc = launch_k2_inst<15, false, 0>(ctx, half, st);
    if (rc != FRA_OK) return rc;

    BinOut o;
    o.frames = args.frames;
    o.iq = args.iq;
    o.mag = args.mag;
    o.phase = args.phase;
    o.qscale = args.qscale;
    o.mag_alpha = args.mag_alpha;
    const size_t total = frames * kHalf64k;
    const dim3 grid((unsigned)((total + 255) / 256));
    if (qmode == 0) {
        auto kfn = k2_join64k<0>;
        FRA_LAUNCH(kfn, grid, dim3(256), (size_t)0, st, (const float2 *)halves, (const float2 *)ctx->d_twc, o, total);
    } else if (qmode == 1) {
        auto kfn = k2_join64k<1>;
        FRA_LAUNCH(kfn, grid, dim3(256), (size_t)0, st, (const float2 *)halves, (const float2 *)ctx->d_twc, o, total);
    } else {
        auto kfn = k2_join64k<2>;
        FRA_LAUNCH(kfn, grid, dim3(256), (size_t)0, st, (const float2 *)halves, (const float2 *)ctx->d_twc, o, total);
    }
    FRA_TRY(ctx, cudaGetLastError());
    ctx->last_kernels++;
    return FRA_OK;
}

// FRA_FFT_FIXED16: the 16-bit scaled, truncating radix-2^2 pipeline (k2_fixed.cuh), one frame per CTA
int launch_k2_fixed(fra_ctx *ctx, const K2Args &args, bool win, cudaStream_t st)
{
    const int smem = 4 << ctx->log2n;
    if (args.batch <= 0) return FRA_OK;
    if (win) {
        auto kfn = k2_fixed<true>;
        FRA_SMEM(ctx, kfn, smem);
        FRA_LAUNCH(kfn, dim3(args.batch), dim3(kFixedThreads), (size_t)smem, st, args, (const uint32_t *)ctx->d_twfx, ctx->log2n);
    } else {
        auto kfn = k2_fixed<false>;
        FRA_SMEM(ctx, kfn, smem);
        FRA_LAUNCH(kfn, dim3(args.batch), dim3(kFixedThreads), (size_t)smem, st, args, (const uint32_t *)ctx->d_twfx, ctx->log2n);
    }
    FRA_TRY(ctx, cudaGetLastError());
    ctx->last_kernels++;
    return FRA_OK;
}

int launch_k2(fra_ctx *ctx, const K2Args &args, bool win, int qmode, cudaStream_t st)
{
    if (ctx->flags & FRA_FFT_FIXED16) return launch_k2_fixed(ctx, args, win, st);
    switch (ctx->log2n) {
    case 10: return launch_k2_n<10>(ctx, args, win, qmode, st);
    case 11: return launch_k2_n<11>(ctx, args, win, qmode, st);
    case 12: return launch_k2_n<12>(ctx, args, win, qmode, st);
    case 13: return launch_k2_n<13>(ctx, args, win, qmode, st);
    case 14: return launch_k2_n<14>(ctx, args, win, qmode, st);
    case 15:
#ifndef FRA_HOST_EMUL
        if (!(ctx->flags & FRA_K2_WIDE_CTA)) return launch_k2_cluster<15, 2>(ctx, args, win, qmode, st);
#endif
        return launch_k2_n<15>(ctx, args, win, qmode, st);
    case 16: return launch_k2_64k(ctx, args, win, qmode, st);
    default: return FRA_ERR_UNSUPPORTED;
    }
}

// One step over channels [c0, c0 + nch): pointers in `o` and d_in are already
// offset to channel c0.
// scratch: filter output buffer for the whole context when the caller does not ask for it
int launch_fft(fra_ctx *ctx, const K2Args &k2, bool win, int qmode, cudaStream_t st)
{
    if (ctx->profiling) FRA_TRY(ctx, cudaEventRecord(ctx->ev[2], st));
    int rc = launch_k2(ctx, k2, win, qmode, st);
    if (rc != FRA_OK) return rc;
    if (ctx->profiling) {
        FRA_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
        ctx->ev_k2 = true;
    }
    return FRA_OK;
}

// defer_fft (FRA_PIPELINE): the FFT launch is not enqueued but left in ctx->fft_args for the caller
int process_range(fra_ctx *ctx, const int16_t *d_in, int c0, int nch, int continuous, int log2_scale,
                  const fra_outputs &o, cudaStream_t st, bool defer_fft = false, int16_t *scratch = nullptr)
{
    if (!scratch) scratch = ctx->d_scratch;
    const int n = ctx->n;
    if (ctx->flags & FRA_WINDOW_RTL_SKEW) {
        // the stream delayed by one sample (the ROM tables are rotated by two): out[n] = W(x[n-1], ROM[n-2])
        int16_t *delayed = ctx->d_skew_in + (size_t)c0 * n;
        int16_t *prev = ctx->d_skew_prev + c0;
        const size_t total8 = (size_t)nch * n / 8;
        auto dk = k0_skew_delay;
        FRA_LAUNCH(dk, dim3((unsigned)((total8 + 255) / 256)), dim3(256), (size_t)0, st, d_in, delayed, prev, total8, n, continuous);
        FRA_TRY(ctx, cudaGetLastError());
        auto ck = k0_skew_carry;
        FRA_LAUNCH(ck, dim3((unsigned)((nch + 255) / 256)), dim3(256), (size_t)0, st, d_in, prev, nch, n);
        FRA_TRY(ctx, cudaGetLastError());
        d_in = delayed;
    }
    const bool iir = (ctx->mode == FRA_MODE_BANK0 || ctx->mode == FRA_MODE_BANK1);
    const bool want_fft = o.d_frames || o.d_iq || o.d_mag || o.d_phase;
    const int16_t *fft_in = d_in;

    if (iir) {
        int16_t *filt = o.d_filtered ? o.d_filtered : (scratch + (size_t)c0 * n);
        K1Args k1;
        k1.in = d_in;
        k1.out = filt;
        k1.state = ctx->d_state + (size_t)c0 * 24;
        k1.rom32 = ctx->d_rom32;
        k1.rom2x = ctx->d_rom2x;
        const Sections sec = current_sections(ctx);
        k1.coef = make_cascade(sec);
        k1.channels = nch;
        k1.n = n;
        k1.continuous = continuous;
        k1.speculate = (ctx->flags & FRA_K1_SPECULATE) ? 1 : 0;
#ifdef FRA_TIMELINE
        k1.tl_step = (int)ctx->pipe_calls;
#endif
        if (ctx->profiling) FRA_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
        // k1_duo (stage pairs per warp) by default, k1_lane (a lane per channel) in the window of channel
        // counts where it wins; the stage-per-lane systolic k1_split only on request (and for the exact single stream)
        int variant = (nch >= kLaneMinChannels && nch < kLaneMaxChannels) ? 0 : 3;     // 0 lane, 1 split, 3 duo
        // (the all-biased lane kernel wins from ~24k channels up, decided below once the coefficients are classified)
        if (ctx->flags & FRA_K1_FORCE_LANE) variant = 0;
        if (ctx->flags & FRA_K1_FORCE_SPLIT) variant = 1;
        if (ctx->flags & FRA_K1_FORCE_DUO) variant = 3;
        bool b1z = true, fast = true, biased = !(ctx->flags & FRA_K1_NO_BIASED);
        for (int i = 0; i < kStages; ++i) {
            b1z = b1z && sec.c[i][1] == 0;                       // x[n-1] coefficient zero in every stage: skip that product
            fast = fast && std::abs((int)sec.c[i][4]) <= kFastMaxA1;   // two-instruction recurrence (fra_common.cuh)
            // every operand biased, no FADD at all (fra_common.cuh: biquad_step_biased)
            biased = biased && biased_order_ok(sec.c[i][0], sec.c[i][1], sec.c[i][2], sec.c[i][3], sec.c[i][4]);
        }
        // k1_lane_biased (whole-line staging, skewed cascade, every scheduler equally loaded) against k1_duo:
        // 32768 channels 0.82 vs 0.97 ms, 65536 1.57 vs 1.70 ms; 16384 0.56 vs 0.49 ms (too few warps per scheduler)
        if (biased && !(ctx->flags & (FRA_K1_FORCE_LANE | FRA_K1_FORCE_SPLIT | FRA_K1_FORCE_DUO)))
            variant = (nch >= ((ctx->flags & FRA_PIPELINE) ? kLaneBiasedMinChannelsPipelined : kLaneBiasedMinChannels)) ? 0 : 3;
        bool alt = true;                                           // ALPHA, BETA, ALPHA, BETA, ALPHA, BETA
        for (int i = 2; i < kStages; ++i) alt = alt && std::memcmp(sec.c[i], sec.c[i & 1], 5) == 0;
        if (variant == 3) {
            const int grid = (nch + 31) / 32;
            void (*const table[12])(K1Args) = {
                k1_duo<false, 0, false>, k1_duo<true, 0, false>, k1_duo<false, 1, false>, k1_duo<true, 1, false>,
                k1_duo<false, 2, false>, k1_duo<true, 2, false>,
                k1_duo<false, 0, true>,  k1_duo<true, 0, true>,  k1_duo<false, 1, true>,  k1_duo<true, 1, true>,
                k1_duo<false, 2, true>,  k1_duo<true, 2, true>};
            auto kfn = table[(b1z ? 1 : 0) + 2 * (biased ? 2 : fast ? 1 : 0) + (alt ? 6 : 0)];
            static_assert(kDuoSmemRequest >= kDuoSmemBytes, "k1_duo shared memory");
            FRA_SMEM(ctx, kfn, kDuoSmemRequest);
            FRA_LAUNCH(kfn, dim3(grid), dim3(kDuoWarps * 32), (size_t)kDuoSmemRequest, st, k1);
        } else if (variant == 1) {
            const int per_cta = kSplitWarps * kSplitGroups;
            const int grid = (nch + per_cta - 1) / per_cta;
            const size_t smem = (size_t)kSplitWarps * kSplitSmemPerWarp;
            auto kfn = k1_split;
            FRA_SMEM(ctx, kfn, (int)smem);
            FRA_LAUNCH(kfn, dim3(grid), dim3(kSplitWarps * 32), smem, st, k1);
        } else {
            const int block = biased ? kLaneBiasedBlock : kLaneBlock;
            const int grid = (nch + block - 1) / block;
            void (*const lane_table[4])(K1Args) = {k1_lane_biased<false, false>, k1_lane_biased<true, false>,
                                                   k1_lane_biased<false, true>, k1_lane_biased<true, true>};
            auto kfn = biased ? lane_table[(b1z ? 1 : 0) | (alt ? 2 : 0)] : (b1z ? k1_lane<true> : k1_lane<false>);
            if (biased) FRA_SMEM(ctx, kfn, kLaneBiasedSmem);
            FRA_LAUNCH(kfn, dim3(grid), dim3(block), (size_t)(biased ? kLaneBiasedSmem : 0), st, k1);
        }
        FRA_TRY(ctx, cudaGetLastError());
        if (ctx->profiling) {
            FRA_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
            ctx->ev_k1 = true;
        }
        ctx->last_kernels++;
        fft_in = filt;
    } else if (o.d_filtered) {
        const size_t total8 = (size_t)nch * n / 8;
        auto kfn = k1_window_only;
        FRA_LAUNCH(kfn, dim3((unsigned)((total8 + 255) / 256)), dim3(256), (size_t)0, st, d_in, o.d_filtered,
                   (const int *)ctx->d_rom32, total8, n);
        FRA_TRY(ctx, cudaGetLastError());
        ctx->last_kernels++;
    }

    if (want_fft) {
        K2Args k2;
        k2.in = reinterpret_cast<const uint32_t *>(fft_in);
        k2.rom32 = ctx->d_rom32;
        k2.tw1 = ctx->d_tw1;
        k2.tw2 = ctx->d_tw2;
        k2.twn = ctx->d_twn;
        k2.frames = reinterpret_cast<uint32_t *>(o.d_frames);
        k2.iq = reinterpret_cast<float2 *>(o.d_iq);
        k2.mag = o.d_mag;
        k2.phase = o.d_phase;
        k2.qscale = std::ldexp(0.5f, log2_scale);
        k2.mag_alpha = ctx->mag_alpha;
        k2.batch = nch;
        k2.frame0 = c0;
        k2.prefetch = ctx->sm_count * FRA_K2_MINBLOCKS;
#ifdef FRA_TIMELINE
        k2.tl_step = (int)ctx->pipe_calls;
#endif
        k2.exp23 = 0x4B000000u;
        const bool nearest = (ctx->flags & FRA_ROUND_NEAREST) != 0;
        const int qmode = nearest ? 2 : (log2_scale <= -ctx->log2n ? 0 : 1);
        if (defer_fft) {
            ctx->fft_args = k2;
            ctx->fft_win = !iir;
            ctx->fft_qmode = qmode;
            ctx->fft_pending = true;
            return FRA_OK;
        }
        return launch_fft(ctx, k2, /*win=*/!iir, qmode, st);
    }
    return FRA_OK;
}

int pipe_init(fra_ctx *ctx)
{
    if (ctx->pipe_k1) return FRA_OK;
    int lo = 0, hi = 0;                                        // hi is the numerically smaller value
    FRA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // window+IIR first: measured 0.2725 ms/step against 0.276 with equal priorities and 0.326 with the FFT first
    FRA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->pipe_k1, cudaStreamNonBlocking, hi));
    FRA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->pipe_k2, cudaStreamNonBlocking, lo));
    FRA_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_in, cudaEventDisableTiming));
    FRA_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_go, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        FRA_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_k1_done[i], cudaEventDisableTiming));
        FRA_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_k2_done[i], cudaEventDisableTiming));
    }
    return FRA_OK;
}

// The FFT of the last call, if still held back: behind that call's window+IIR kernel.
int pipe_flush_fft(fra_ctx *ctx)
{
    if (!ctx->fft_pending) return FRA_OK;
    ctx->fft_pending = false;
    FRA_TRY(ctx, cudaStreamWaitEvent(ctx->pipe_k2, ctx->pipe_k1_done[ctx->fft_buf], 0));
    int rc = launch_fft(ctx, ctx->fft_args, ctx->fft_win, ctx->fft_qmode, ctx->pipe_k2);
    if (rc != FRA_OK) return rc;
    FRA_TRY(ctx, cudaEventRecord(ctx->pipe_k2_done[ctx->fft_buf], ctx->pipe_k2));
    return FRA_OK;
}

// host waits for the two pipeline streams (before anything that touches state or scratch
// from another stream)
cudaError_t pipe_host_join(fra_ctx *ctx)
{
    if (pipe_flush_fft(ctx) != FRA_OK) return cudaErrorUnknown;
    cudaError_t e = ctx->pipe_k1 ? cudaStreamSynchronize(ctx->pipe_k1) : cudaSuccess;
    if (e == cudaSuccess && ctx->pipe_k2) e = cudaStreamSynchronize(ctx->pipe_k2);
    return e;
}

// host waits for the copy streams (calls of fra_process_host_async still in flight)
size_t host_mirror(fra_ctx *ctx, int slot, double *wait_s);

cudaError_t host_streams_join(fra_ctx *ctx)
{
    cudaError_t e = cudaSuccess;
    for (auto st : ctx->copy_streams)
        if (st && e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess)
        for (int slot = 0; slot < 2; ++slot) host_mirror(ctx, slot, nullptr);   // frames of half-spectrum calls nobody waited for
    return e;
}

// ---- FRA_HOST_HALF_SPECTRUM, host side: the upper half of every frame from bins 0 .. N/2 and the mirror bits.
// word j = {re, im} (int16 pair);  X[N-j] = {re, -im - bit_j} mod 2^16  ((~im) + 1 - bit in the upper half-word).
inline uint32_t mirror_word(uint32_t w, uint32_t bit) { return (w & 0xFFFFu) | ((((~w) >> 16) + 1u - bit) << 16); }

void mirror_frame_scalar(uint32_t *fr, const uint32_t *bits, int n, int j0, int j1)
{
    for (int j = j0; j < j1; ++j) fr[n - j] = mirror_word(fr[j], (bits[j >> 5] >> (j & 31)) & 1u);
}

#ifdef FRA_HAVE_AVX2_PATH
// Eight words per step.  The blocks are chosen so that the STORES are 32-byte aligned (j = 1 mod 8: words
// N-j-7 .. N-j start at a multiple of 8) and go out as non-temporal stores: the upper half is written once and
// not read again here, and a regular store would first read every line it overwrites (the mirror is bound by
// host memory bandwidth, which it shares with the DMA engines).  `bits` has one word of slack behind the last frame.
__attribute__((target("avx2"))) void mirror_frame_avx2(uint32_t *fr, const uint32_t *bits, int n)
{
    const int m = n / 2;
    mirror_frame_scalar(fr, bits, n, 1, 9);
    const __m256i lo = _mm256_set1_epi32(0xFFFF), one = _mm256_set1_epi32(1);
    const __m256i sh = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7), rev = _mm256_setr_epi32(7, 6, 5, 4, 3, 2, 1, 0);
    const bool aligned = (reinterpret_cast<uintptr_t>(fr) & 31u) == 0;
    int j = 9;
    for (; j + 8 <= m; j += 8) {                                       // words j .. j+7 -> N-j-7 .. N-j
        const __m256i w = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(fr + j));
        uint64_t two;
        std::memcpy(&two, bits + (j >> 5), sizeof(two));
        const __m256i b = _mm256_and_si256(_mm256_srlv_epi32(_mm256_set1_epi32((int)((two >> (j & 31)) & 0xFFu)), sh), one);
        const __m256i hi = _mm256_add_epi32(_mm256_srli_epi32(_mm256_xor_si256(w, _mm256_set1_epi32(-1)), 16), _mm256_sub_epi32(one, b));
        const __m256i r = _mm256_permutevar8x32_epi32(_mm256_or_si256(_mm256_and_si256(w, lo), _mm256_slli_epi32(hi, 16)), rev);
        if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i *>(fr + n - j - 7), r);
        else _mm256_storeu_si256(reinterpret_cast<__m256i *>(fr + n - j - 7), r);
    }
    mirror_frame_scalar(fr, bits, n, j, m);
}
#endif

// All half-spectrum frames of one call slot.  The call's channel slices land one after the other (their copies are
// queued in slice order), so the mirror is pipelined with the link: the calling thread waits for slice s's copies
// (its event) and releases it to the worker threads, which each complete their part of every slice - the first
// slices are mirrored while the last ones still cross the link.  Returns the number of frames mirrored;
// *wait_s = how long the calling thread was blocked on copies.
size_t host_mirror(fra_ctx *ctx, int slot, double *wait_s = nullptr)
{
    using clk = std::chrono::steady_clock;
    fra_ctx::MirrorPlan plan = ctx->mirror_plan[slot];
    if (wait_s) *wait_s = 0.0;
    if (!plan.frames) return 0;
    ctx->mirror_plan[slot].frames = nullptr;
    uint8_t *base = plan.frames;
    const int n = ctx->n, words = n / 64;                              // N/2 bits per frame
    const uint32_t *bits = ctx->h_mbits[slot];
#ifdef FRA_HAVE_AVX2_PATH
    const bool avx2 = __builtin_cpu_supports("avx2");
#endif
    size_t total = 0;
    for (int s = 0; s < plan.n_slices; ++s) total += plan.nh[s];
    if (total == 0) return 0;
    // threads: the host's hardware threads (at most 32), or FRA_HOST_THREADS when several processes share the host
    size_t hw = std::max(1u, std::thread::hardware_concurrency());
    if (const char *env = std::getenv("FRA_HOST_THREADS")) {
        const long v = std::strtol(env, nullptr, 10);
        if (v >= 1) hw = (size_t)v;
    }
    const size_t n_thr = std::min<size_t>(std::min<size_t>(hw, 32), std::max<size_t>(1, total * (size_t)n >> 20));
    std::atomic<int> ready{0};                                         // slices whose copies have landed
    // worker t of n_thr: its part of every slice, in slice order
    auto work = [&](size_t t) {
        for (int s = 0; s < plan.n_slices; ++s) {
            const size_t nh = plan.nh[s];
            const size_t i0 = nh * t / n_thr, i1 = nh * (t + 1) / n_thr;
            if (i0 >= i1) continue;
            while (ready.load(std::memory_order_acquire) <= s) std::this_thread::yield();
            for (size_t i = i0; i < i1; ++i) {
                const size_t f = (size_t)s * plan.per + i;
                uint32_t *fr = reinterpret_cast<uint32_t *>(base + f * (size_t)n * 4);
#ifdef FRA_HAVE_AVX2_PATH
                if (avx2) { mirror_frame_avx2(fr, bits + f * words, n); continue; }
#endif
                mirror_frame_scalar(fr, bits + f * words, n, 1, n / 2);
            }
        }
#ifdef FRA_HAVE_AVX2_PATH
        _mm_sfence();                                                   // the non-temporal stores are visible before the join
#endif
    };
    std::vector<std::thread> pool;
    if (n_thr > 1)
        for (size_t t = 0; t < n_thr; ++t) pool.emplace_back(work, t);
    double waited = 0.0;
    for (int s = 0; s < plan.n_slices; ++s) {
        const auto t0 = clk::now();
        if (cudaEvent_t e = ctx->slice_done[slot][s]) (void)cudaEventSynchronize(e);   // errors surface at the call-level events
        waited += std::chrono::duration<double>(clk::now() - t0).count();
        ready.store(s + 1, std::memory_order_release);
    }
    if (n_thr == 1) work(0);                                            // small calls: no threads
    for (auto &th : pool) th.join();
    if (wait_s) *wait_s = waited;
    return total;
}

// everything an earlier call left in slot `slot`: its copies have landed, its frames are whole
int finish_host_slot(fra_ctx *ctx, int slot)
{
    using clk = std::chrono::steady_clock;
    const bool pending = ctx->mirror_plan[slot].frames != nullptr;
    const auto t0 = clk::now();
    double slice_wait = 0.0;
    (void)host_mirror(ctx, slot, &slice_wait);                         // slice by slice, as the copies land
    const auto t1 = clk::now();
    for (cudaEvent_t e : ctx->host_done[slot])
        if (e) FRA_TRY(ctx, cudaEventSynchronize(e));
    const auto t2 = clk::now();
    // (reported by fra_get_host_transfer: how long this thread was blocked on copies, how long the mirror ran on)
    const double wait = slice_wait + std::chrono::duration<double>(t2 - t1).count();
    const double mir = std::chrono::duration<double>(t1 - t0).count() - slice_wait;
    if (pending) {
        ctx->last_wait_s = wait;
        ctx->last_mirror_s = mir;
    }
    return FRA_OK;
}

fra_outputs offset_outputs(const fra_outputs &o, size_t c0, int n)
{
    fra_outputs r = o;
    if (r.d_filtered) r.d_filtered += c0 * n;
    if (r.d_frames) r.d_frames += c0 * n * 4;
    if (r.d_iq) r.d_iq += c0 * n * 2;
    if (r.d_mag) r.d_mag += c0 * n;
    if (r.d_phase) r.d_phase += c0 * n;
    return r;
}

}  // namespace

extern "C" {

int fra_abi_version(void) { return FRA_ABI_VERSION; }

const char *fra_strerror(int status)
{
    switch (status) {
    case FRA_OK: return "ok";
    case FRA_ERR_INVALID: return "invalid argument";
    case FRA_ERR_NO_DEVICE: return "no CUDA device (libfra has no CPU fallback)";
    case FRA_ERR_CUDA: return "CUDA error";
    case FRA_ERR_NOMEM: return "out of memory";
    case FRA_ERR_UNSUPPORTED: return "unsupported configuration";
    case FRA_ERR_BUSY: return "byte stream ended inside a 0xF1 coefficient upload";
    default: return "unknown status";
    }
}

int fra_window_rom(int16_t out[FRA_WINDOW_LEN])
{
    if (!out) return FRA_ERR_INVALID;
    std::memcpy(out, kHannRom, sizeof(kHannRom));
    return FRA_OK;
}

int fra_create(fra_ctx **out, int device, int n_channels, int fft_size, unsigned flags)
{
    if (!out) return FRA_ERR_INVALID;
    *out = nullptr;
    if (n_channels <= 0) return FRA_ERR_INVALID;
    int log2n = 0;
    while ((1 << log2n) < fft_size) ++log2n;
    if ((1 << log2n) != fft_size) return FRA_ERR_INVALID;
    if (log2n < 10 || log2n > 16) return FRA_ERR_UNSUPPORTED;
    if ((flags & FRA_FFT_FIXED16) && log2n > 15) return FRA_ERR_UNSUPPORTED;      // the frame must fit one SM's shared memory
    if ((flags & FRA_WINDOW_RTL_SKEW) && (flags & FRA_PIPELINE)) return FRA_ERR_UNSUPPORTED;   // one delayed-input buffer
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return FRA_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return FRA_ERR_INVALID;

    fra_ctx *ctx = new (std::nothrow) fra_ctx();
    if (!ctx) return FRA_ERR_NOMEM;
    ctx->device = device;
    ctx->channels = n_channels;
    ctx->n = fft_size;
    ctx->log2n = log2n;
    ctx->flags = flags;

    auto bail = [&](int rc) { fra_destroy(ctx); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(FRA_ERR_NO_DEVICE);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(FRA_ERR_NO_DEVICE);
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(FRA_ERR_CUDA);
    // the three copy streams are created by the first host-buffer call

    const size_t n = (size_t)fft_size;
    if (cudaMalloc((void **)&ctx->d_rom32, kWindowLen * sizeof(int)) != cudaSuccess) return bail(FRA_ERR_NOMEM);
    if (cudaMalloc((void **)&ctx->d_rom2x, kWindowLen * sizeof(int)) != cudaSuccess) return bail(FRA_ERR_NOMEM);
    if (cudaMalloc((void **)&ctx->d_state, (size_t)n_channels * 24 * sizeof(int16_t)) != cudaSuccess) return bail(FRA_ERR_NOMEM);
    if (cudaMalloc((void **)&ctx->d_tw1, 256 * sizeof(float2)) != cudaSuccess) return bail(FRA_ERR_NOMEM);
    if (cudaMalloc((void **)&ctx->d_tw2, 4096 * sizeof(float2)) != cudaSuccess) return bail(FRA_ERR_NOMEM);
    // the in-kernel FFT is the frame itself up to 32K, and each half of a 64K frame
    const size_t n_kernel = std::min<size_t>(n, (size_t)kHalf64k);
    if (cudaMalloc((void **)&ctx->d_twn, (n_kernel / 2) * sizeof(float2)) != cudaSuccess) return bail(FRA_ERR_NOMEM);

    std::vector<int> rom32(kWindowLen), rom2x(kWindowLen);
    const int rot = (flags & FRA_WINDOW_RTL_SKEW) ? 2 : 0;            // sample n-1 meets ROM[n-2] (hann8192.vhd:36-39)
    for (int i = 0; i < kWindowLen; ++i) {
        rom32[i] = kHannRom[(i - rot) & (kWindowLen - 1)];
        rom2x[i] = 2 * rom32[i];
    }
    std::vector<float2> tw1(256), tw2(4096), twn(n_kernel / 2);
    const double two_pi = 6.283185307179586476925286766559;
    for (int r = 0; r < 16; ++r)
        for (int k = 0; k < 16; ++k) {
            double a = -two_pi * (double)(r * k) / 256.0;
            tw1[r * 16 + k] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int r = 0; r < 16; ++r)
        for (int k = 0; k < 256; ++k) {
            double a = -two_pi * (double)(r * k) / 4096.0;
            tw2[r * 256 + k] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (size_t e = 0; e < n_kernel / 2; ++e) {
        double a = -two_pi * (double)e / (double)n_kernel;
        twn[e] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    if (n > n_kernel) {
        std::vector<float2> twc(kHalf64k);
        for (int k = 0; k < kHalf64k; ++k) {
            double a = -two_pi * (double)k / (double)n;
            twc[k] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
        if (cudaMalloc((void **)&ctx->d_twc, twc.size() * sizeof(float2)) != cudaSuccess) return bail(FRA_ERR_NOMEM);
        if (cudaMemcpy(ctx->d_twc, twc.data(), twc.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(FRA_ERR_CUDA);
    }
    if (cudaMemcpy(ctx->d_rom32, rom32.data(), rom32.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(ctx->d_rom2x, rom2x.data(), rom2x.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(ctx->d_tw1, tw1.data(), tw1.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(ctx->d_tw2, tw2.data(), tw2.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(ctx->d_twn, twn.data(), twn.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(ctx->d_state, 0, (size_t)n_channels * 24 * sizeof(int16_t)) != cudaSuccess)
        return bail(FRA_ERR_CUDA);
    if (flags & FRA_WINDOW_RTL_SKEW) {
        if (cudaMalloc((void **)&ctx->d_skew_in, (size_t)n_channels * n * sizeof(int16_t)) != cudaSuccess ||
            cudaMalloc((void **)&ctx->d_skew_prev, (size_t)n_channels * sizeof(int16_t)) != cudaSuccess)
            return bail(FRA_ERR_NOMEM);
        if (cudaMemset(ctx->d_skew_prev, 0, (size_t)n_channels * sizeof(int16_t)) != cudaSuccess) return bail(FRA_ERR_CUDA);
    }
    if (flags & FRA_FFT_FIXED16) {
        // phase factors of the fixed-point mode: round(cos, -sin * 2^15) clipped to 32767 (xfft_0.xci:19, 16 bits)
        std::vector<uint32_t> twfx(n);
        for (size_t t = 0; t < n; ++t) {
            const double a = two_pi * (double)t / (double)n;
            const long wr = std::min(32767L, std::max(-32767L, std::lrint(std::cos(a) * 32768.0)));
            const long wi = std::min(32767L, std::max(-32767L, std::lrint(-std::sin(a) * 32768.0)));
            twfx[t] = ((uint32_t)wr & 0xFFFFu) | ((uint32_t)wi << 16);
        }
        if (cudaMalloc((void **)&ctx->d_twfx, n * sizeof(uint32_t)) != cudaSuccess) return bail(FRA_ERR_NOMEM);
        if (cudaMemcpy(ctx->d_twfx, twfx.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(FRA_ERR_CUDA);
    }
    // Work buffers are sized here, not in the first fra_process call: the filter-output scratch
    // (two of them in pipelined mode) and, for 64K frames, the even/odd split and its fp32 halves.
    {
        const size_t frame_elems = (size_t)n_channels * n;
        const size_t want_elems = frame_elems * ((flags & FRA_PIPELINE) ? 2 : 1);
        if (cudaMalloc((void **)&ctx->d_scratch, want_elems * sizeof(int16_t)) != cudaSuccess) return bail(FRA_ERR_NOMEM);
        ctx->scratch_elems = want_elems;
#ifdef FRA_HOST_EMUL
        const bool split64k = true;
#else
        const bool split64k = (flags & FRA_K2_64K_SPLIT) != 0;
#endif
        if (n > n_kernel && split64k) {
            if (cudaMalloc((void **)&ctx->d_split, (size_t)n_channels * 2 * kHalf64k * sizeof(int16_t)) != cudaSuccess ||
                cudaMalloc((void **)&ctx->d_halves, (size_t)n_channels * 2 * kHalf64k * sizeof(float2)) != cudaSuccess)
                return bail(FRA_ERR_NOMEM);
            ctx->split_frames = (size_t)n_channels;
        }
    }
    *out = ctx;
    return FRA_OK;
}

int fra_destroy(fra_ctx *ctx)
{
    if (!ctx) return FRA_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    host_streams_join(ctx);
    for (auto &slot : ctx->host_done)
        for (cudaEvent_t e : slot)
            if (e) cudaEventDestroy(e);
    for (cudaStream_t ps : {ctx->pipe_k1, ctx->pipe_k2})
        if (ps) { cudaStreamSynchronize(ps); cudaStreamDestroy(ps); }
    for (cudaEvent_t pe : {ctx->pipe_in, ctx->pipe_go, ctx->pipe_k1_done[0], ctx->pipe_k1_done[1], ctx->pipe_k2_done[0], ctx->pipe_k2_done[1]})
        if (pe) cudaEventDestroy(pe);
    for (auto &slot_events : ctx->slice_done)
        for (cudaEvent_t se : slot_events)
            if (se) cudaEventDestroy(se);
    for (auto hb : ctx->h_mbits)
        if (hb) cudaFreeHost(hb);
    void *bufs[] = {ctx->d_rom32, ctx->d_rom2x, ctx->d_twfx, ctx->d_mbits, ctx->d_skew_in, ctx->d_skew_prev, ctx->d_state, ctx->d_scratch, ctx->d_tw1, ctx->d_tw2, ctx->d_twn, ctx->d_in, ctx->d_twc, ctx->d_halves, ctx->d_split,
                    ctx->d_frames, ctx->d_filtered_out, ctx->d_iq, ctx->d_mag, ctx->d_phase, ctx->d_entry,
                    ctx->d_exit, ctx->d_counts, ctx->d_ends, ctx->d_aggr, ctx->d_mats};
    for (void *p : bufs)
        if (p) cudaFree(p);
    for (auto e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto s : ctx->copy_streams)
        if (s) cudaStreamDestroy(s);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return FRA_OK;
}

int fra_command(fra_ctx *ctx, const uint8_t *bytes, size_t n)
{
    if (!ctx || (!bytes && n)) return FRA_ERR_INVALID;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t b = bytes[i];
        if (ctx->upload_pos >= 0) {                       // ACQUIRE: every byte is a coefficient
            ctx->upload_buf[ctx->upload_pos++] = (int8_t)b;
            if (ctx->upload_pos == 12) {
                std::memcpy(ctx->bank1, ctx->upload_buf, 12);
                ctx->sections_loaded = false;
                ctx->upload_pos = -1;
                ctx->n_upload++;
            }
            continue;
        }
        switch (b) {
        case FRA_CMD_FILTER_UPDATE: ctx->upload_pos = 0; break;
        case FRA_MODE_BANK0:
        case FRA_MODE_BANK1:
        case FRA_MODE_BYPASS: ctx->mode = b; break;
        case FRA_CMD_RESET: {
            int rc = fra_reset(ctx);
            if (rc != FRA_OK) return rc;
            break;
        }
        case FRA_CMD_START: ctx->n_start++; break;
        case FRA_CMD_UART_REQUEST: ctx->n_request++; break;
        case FRA_CMD_ETHERNET_MODE:
        case FRA_CMD_UART_MODE: ctx->transport = b; break;
        default: break;                                   // matches no decoder branch: dropped
        }
    }
    return ctx->upload_pos >= 0 ? FRA_ERR_BUSY : FRA_OK;
}

int fra_load_bank1(fra_ctx *ctx, const int8_t coeff[12])
{
    if (!ctx || !coeff) return FRA_ERR_INVALID;
    std::memcpy(ctx->bank1, coeff, 12);
    ctx->sections_loaded = false;
    ctx->n_upload++;
    return FRA_OK;
}

int fra_load_sections(fra_ctx *ctx, const int8_t coeff[36])
{
    if (!ctx || !coeff) return FRA_ERR_INVALID;
    std::memcpy(ctx->sections, coeff, 36);
    ctx->sections_loaded = true;
    ctx->n_upload++;
    return FRA_OK;
}

int fra_set_mag_average(fra_ctx *ctx, float alpha)
{
    if (!ctx || !(alpha > 0.0f) || alpha > 1.0f) return FRA_ERR_INVALID;
    ctx->mag_alpha = alpha;
    return FRA_OK;
}

int fra_set_host_half_share(fra_ctx *ctx, double share)
{
    if (!ctx || share > 1.0) return FRA_ERR_INVALID;
    ctx->half_adaptive = share < 0.0;
    ctx->half_frac = share < 0.0 ? 1.0 : share;
    ctx->tuner = fra_ctx::HalfTuner();
    return FRA_OK;
}

int fra_get_host_transfer(const fra_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes, double *half_share, double *wait_s,
                          double *mirror_s)
{
    if (!ctx) return FRA_ERR_INVALID;
    if (h2d_bytes) *h2d_bytes = ctx->last_h2d_bytes;
    if (d2h_bytes) *d2h_bytes = ctx->last_d2h_bytes;
    if (half_share) *half_share = ctx->half_frac;
    if (wait_s) *wait_s = ctx->last_wait_s;
    if (mirror_s) *mirror_s = ctx->last_mirror_s;
    return FRA_OK;
}

int fra_get_sections(const fra_ctx *ctx, int8_t coeff[36])
{
    if (!ctx || !coeff) return FRA_ERR_INVALID;
    const Sections s = current_sections(ctx);
    std::memcpy(coeff, s.c, 36);
    return FRA_OK;
}

int fra_set_mode(fra_ctx *ctx, uint8_t mode)
{
    if (!ctx) return FRA_ERR_INVALID;
    if (mode != FRA_MODE_BANK0 && mode != FRA_MODE_BANK1 && mode != FRA_MODE_BYPASS) return FRA_ERR_INVALID;
    ctx->mode = mode;
    return FRA_OK;
}

int fra_reset(fra_ctx *ctx)
{
    if (!ctx) return FRA_ERR_INVALID;
    do_reset(ctx);
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    FRA_TRY(ctx, pipe_host_join(ctx));
    FRA_TRY(ctx, host_streams_join(ctx));
    // synchronous: a reset must be ordered before work on ANY stream the caller uses next
    FRA_TRY(ctx, cudaMemsetAsync(ctx->d_state, 0, (size_t)ctx->channels * 24 * sizeof(int16_t), ctx->stream));
    if (ctx->d_skew_prev) FRA_TRY(ctx, cudaMemsetAsync(ctx->d_skew_prev, 0, (size_t)ctx->channels * sizeof(int16_t), ctx->stream));
    FRA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FRA_OK;
}

int fra_get_mode(const fra_ctx *ctx, uint8_t *mode)
{
    if (!ctx || !mode) return FRA_ERR_INVALID;
    *mode = ctx->mode;
    return FRA_OK;
}

int fra_get_bank(const fra_ctx *ctx, int bank, int8_t coeff[12])
{
    if (!ctx || !coeff || (bank != 0 && bank != 1)) return FRA_ERR_INVALID;
    std::memcpy(coeff, bank == 0 ? kBank0 : ctx->bank1, 12);
    return FRA_OK;
}

int fra_get_transport(const fra_ctx *ctx, uint8_t *transport)
{
    if (!ctx || !transport) return FRA_ERR_INVALID;
    *transport = ctx->transport;
    return FRA_OK;
}

int fra_get_counters(const fra_ctx *ctx, uint64_t *n_start, uint64_t *n_request, uint64_t *n_reset, uint64_t *n_upload)
{
    if (!ctx) return FRA_ERR_INVALID;
    if (n_start) *n_start = ctx->n_start;
    if (n_request) *n_request = ctx->n_request;
    if (n_reset) *n_reset = ctx->n_reset;
    if (n_upload) *n_upload = ctx->n_upload;
    return FRA_OK;
}

int fra_process(fra_ctx *ctx, const int16_t *d_in, int continuous, int log2_scale, const fra_outputs *out,
                void *cuda_stream)
{
    if (!ctx || !d_in || !out) return FRA_ERR_INVALID;
    if (log2_scale == FRA_SCALE_DEFAULT) log2_scale = -ctx->log2n;
    if (log2_scale < -40 || log2_scale > 16) return FRA_ERR_INVALID;
    if ((ctx->flags & FRA_FFT_FIXED16) && log2_scale != -ctx->log2n) return FRA_ERR_INVALID;   // the fixed pipeline's schedule is 1/N
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;      // NULL = the legacy default stream
    const bool iir = (ctx->mode == FRA_MODE_BANK0 || ctx->mode == FRA_MODE_BANK1);
    const bool pipe = (ctx->flags & FRA_PIPELINE) != 0;
    const size_t frame_elems = (size_t)ctx->channels * ctx->n;
    const size_t want_elems = frame_elems * (pipe ? 2 : 1);
    if (iir && !out->d_filtered && ctx->scratch_elems < want_elems) {
        if (pipe) FRA_TRY(ctx, pipe_host_join(ctx));
        if (ctx->d_scratch) FRA_TRY(ctx, cudaFree(ctx->d_scratch));
        ctx->d_scratch = nullptr;
        ctx->scratch_elems = 0;
        if (cudaMalloc((void **)&ctx->d_scratch, want_elems * sizeof(int16_t)) != cudaSuccess) return FRA_ERR_NOMEM;
        ctx->scratch_elems = want_elems;
    }
    ctx->last_kernels = 0;
    ctx->ev_k1 = ctx->ev_k2 = false;
    if (!pipe) return process_range(ctx, d_in, 0, ctx->channels, continuous, log2_scale, *out, st);

    // ---- FRA_PIPELINE: K1(i) on pipe_k1 (high priority: its few long-lived CTAs are placed
    // first), K2(i-1) on pipe_k2 beside it.  Two filter scratch buffers alternate; K1(i) waits
    // for K2(i-2), the last reader of its buffer.
    //
    // The FFT of call i-1 is enqueued HERE, behind an event recorded right before K1(i)'s
    // launch, not in call i-1.  Released by K1(i-1)'s completion alone it started ~13 us before
    // K1(i) - whose stream still had that completion's event record and two event waits to work
    // through - and filled every SM three CTAs deep; K1(i)'s CTAs then trickled in as FFT CTAs
    // retired (tools/timeline_probe.py: K1 span 316 us instead of 264, step 0.300 instead of
    // 0.268 ms; which of the two a run got depended on its first steps).
    int rc = pipe_init(ctx);
    if (rc != FRA_OK) return rc;
    const int buf = (int)(ctx->pipe_calls & 1);
    FRA_TRY(ctx, cudaEventRecord(ctx->pipe_in, st));
    FRA_TRY(ctx, cudaStreamWaitEvent(ctx->pipe_k1, ctx->pipe_in, 0));
    if (ctx->pipe_calls >= 2) FRA_TRY(ctx, cudaStreamWaitEvent(ctx->pipe_k1, ctx->pipe_k2_done[buf], 0));
    // a caller-owned filter output is rewritten every call: the previous FFT must have read it
    if (out->d_filtered && ctx->pipe_calls >= 1) {
        rc = pipe_flush_fft(ctx);
        if (rc != FRA_OK) return rc;
        FRA_TRY(ctx, cudaStreamWaitEvent(ctx->pipe_k1, ctx->pipe_k2_done[buf ^ 1], 0));
    }
    FRA_TRY(ctx, cudaEventRecord(ctx->pipe_go, ctx->pipe_k1));
    const bool had_pending = ctx->fft_pending;
    const K2Args prev_args = ctx->fft_args;
    const bool prev_win = ctx->fft_win;
    const int prev_qmode = ctx->fft_qmode, prev_buf = ctx->fft_buf;
    ctx->fft_pending = false;
    rc = process_range(ctx, d_in, 0, ctx->channels, continuous, log2_scale, *out, ctx->pipe_k1, /*defer_fft=*/true,
                       ctx->d_scratch ? ctx->d_scratch + (size_t)buf * frame_elems : nullptr);
    if (rc != FRA_OK) return rc;
    FRA_TRY(ctx, cudaEventRecord(ctx->pipe_k1_done[buf], ctx->pipe_k1));
    ctx->fft_buf = buf;
    if (had_pending) {
        FRA_TRY(ctx, cudaStreamWaitEvent(ctx->pipe_k2, ctx->pipe_go, 0));
        FRA_TRY(ctx, cudaStreamWaitEvent(ctx->pipe_k2, ctx->pipe_k1_done[prev_buf], 0));
        rc = launch_fft(ctx, prev_args, prev_win, prev_qmode, ctx->pipe_k2);
        if (rc != FRA_OK) return rc;
        FRA_TRY(ctx, cudaEventRecord(ctx->pipe_k2_done[prev_buf], ctx->pipe_k2));
    }
    ctx->pipe_calls++;
    return FRA_OK;
}

int fra_join(fra_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return FRA_ERR_INVALID;
    if (!(ctx->flags & FRA_PIPELINE) || ctx->pipe_calls == 0) return FRA_OK;
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = pipe_flush_fft(ctx);
    if (rc != FRA_OK) return rc;
    // Both parities of both event pairs: the last call may have asked for no FFT output, in which
    // case the newest FFT event is the other parity's (recorded during the last call for the call
    // before it).  Waiting on an event that was never recorded is a no-op.
    for (int i = 0; i < 2; ++i) {
        FRA_TRY(ctx, cudaStreamWaitEvent(st, ctx->pipe_k1_done[i], 0));
        FRA_TRY(ctx, cudaStreamWaitEvent(st, ctx->pipe_k2_done[i], 0));
    }
    return FRA_OK;
}

int fra_process_host(fra_ctx *ctx, const int16_t *h_in, int continuous, int log2_scale, const fra_outputs *h_out)
{
    uint64_t ticket = 0;
    int rc = fra_process_host_async(ctx, h_in, continuous, log2_scale, h_out, &ticket);
    if (rc != FRA_OK) return rc;
    return fra_host_wait(ctx, ticket);
}

int fra_host_wait(fra_ctx *ctx, uint64_t ticket)
{
    if (!ctx || ticket == 0 || ticket > ctx->host_calls) return FRA_ERR_INVALID;
    if (ticket + 2 <= ctx->host_calls) return FRA_OK;          // its slot was waited for when it was reused
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    return finish_host_slot(ctx, (int)(ticket & 1));
}

int fra_process_host_async(fra_ctx *ctx, const int16_t *h_in, int continuous, int log2_scale,
                           const fra_outputs *h_out, uint64_t *ticket)
{
    if (!ctx || !h_in || !h_out || !ticket) return FRA_ERR_INVALID;
    if (log2_scale == FRA_SCALE_DEFAULT) log2_scale = -ctx->log2n;
    if (log2_scale < -40 || log2_scale > 16) return FRA_ERR_INVALID;
    if ((ctx->flags & FRA_FFT_FIXED16) && log2_scale != -ctx->log2n) return FRA_ERR_INVALID;
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    FRA_TRY(ctx, pipe_host_join(ctx));                 // may launch the held-back FFT: after the device is current
    const size_t C = (size_t)ctx->channels, n = (size_t)ctx->n;
    auto need = [&](void **p, size_t bytes) -> bool { return *p || cudaMalloc(p, bytes) == cudaSuccess; };
    if (!need((void **)&ctx->d_in, C * n * 2)) return FRA_ERR_NOMEM;
    if (h_out->d_frames && !need((void **)&ctx->d_frames, C * n * 4)) return FRA_ERR_NOMEM;
    if (h_out->d_filtered && !need((void **)&ctx->d_filtered_out, C * n * 2)) return FRA_ERR_NOMEM;
    if (h_out->d_iq && !need((void **)&ctx->d_iq, C * n * 8)) return FRA_ERR_NOMEM;
    if (h_out->d_mag && !need((void **)&ctx->d_mag, C * n * 4)) return FRA_ERR_NOMEM;
    if (h_out->d_phase && !need((void **)&ctx->d_phase, C * n * 4)) return FRA_ERR_NOMEM;
    const bool iir = (ctx->mode == FRA_MODE_BANK0 || ctx->mode == FRA_MODE_BANK1);
    if (iir && !h_out->d_filtered && !ctx->d_scratch) {
        if (cudaMalloc((void **)&ctx->d_scratch, C * n * 2) != cudaSuccess) return FRA_ERR_NOMEM;
        ctx->scratch_elems = C * n;
    }
    // the slot this call takes was last used two calls ago: that call is finished first (its events, its mirror)
    const unsigned long long id = ctx->host_calls + 1;
    {
        int rc0 = finish_host_slot(ctx, (int)(id & 1));
        if (rc0 != FRA_OK) return rc0;
    }
    // half-spectrum transfer of the frames: only with the truncating 1/N-or-smaller scale, where the mirrored
    // imaginary part is -im or -im - 1 (a saturating or rounding quantiser breaks that relation)
    const bool half = (ctx->flags & FRA_HOST_HALF_SPECTRUM) && h_out->d_frames && !(ctx->flags & (FRA_ROUND_NEAREST | FRA_FFT_FIXED16)) &&
                      log2_scale <= -ctx->log2n;
    const size_t mwords = n / 64;                                     // mirror bits per frame, in words
    if (half) {
        if (!need((void **)&ctx->d_mbits, C * mwords * 4)) return FRA_ERR_NOMEM;
        for (auto &hb : ctx->h_mbits)
            if (!hb && cudaMallocHost((void **)&hb, C * mwords * 4 + 8) != cudaSuccess) return FRA_ERR_NOMEM;
    }

    fra_outputs dev;
    dev.d_filtered = h_out->d_filtered ? ctx->d_filtered_out : nullptr;
    dev.d_frames = h_out->d_frames ? ctx->d_frames : nullptr;
    dev.d_iq = h_out->d_iq ? ctx->d_iq : nullptr;
    dev.d_mag = h_out->d_mag ? ctx->d_mag : nullptr;
    dev.d_phase = h_out->d_phase ? ctx->d_phase : nullptr;

    // pending control-plane work (a reset's memset) is ordered on ctx->stream
    FRA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &cs : ctx->copy_streams)
        if (!cs) FRA_TRY(ctx, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    ctx->last_kernels = 0;
    // channel slices round-robin over three streams: H2D(i+1) and D2H(i-1) overlap compute(i)
    const int n_slices = (int)std::min<size_t>(C, C * n >= ((size_t)1 << 24) ? 8 : 1);
    const size_t per = (C + n_slices - 1) / n_slices;
    int rc = FRA_OK;
    fra_ctx::MirrorPlan plan;
    plan.frames = h_out->d_frames;
    plan.per = per;
    plan.n_slices = n_slices;
    if (half && ctx->half_adaptive && !ctx->tuner.done) {
        // the search (see HalfTuner): trial i covers calls 3 i + 1 .. 3 i + 3 (call 0 is warm-up: allocations)
        auto &tu = ctx->tuner;
        const auto now = std::chrono::steady_clock::now();
        const int k = tu.calls - 1;                                   // calls of the search proper made so far
        if (k >= 3 && k % 3 == 0) {                                   // first call after a trial: the trial's step time
            tu.trial_time[k / 3 - 1] = std::chrono::duration<double>(now - tu.last_entry).count();
            if (k / 3 == 3) {
                // coarse result: refine around the best of 1, 3/4, 1/2
                int b = 0;
                for (int i = 1; i < 3; ++i)
                    if (tu.trial_time[i] < tu.trial_time[b]) b = i;
                const double c = tu.trial_share[b];
                tu.n_trials = 3;
                if (c + 0.125 <= 1.0) tu.trial_share[tu.n_trials++] = c + 0.125;
                tu.trial_share[tu.n_trials++] = c - 0.125;
            }
            if (k / 3 == tu.n_trials) {
                int b = 0;
                for (int i = 1; i < tu.n_trials; ++i)
                    if (tu.trial_time[i] < tu.trial_time[b]) b = i;
                ctx->half_frac = tu.trial_share[b];
                tu.done = true;
            }
        }
        if (!tu.done && k >= 0) ctx->half_frac = tu.trial_share[std::min(k / 3, tu.n_trials - 1)];
        tu.last_entry = now;
        tu.calls++;
    }
    const double half_frac = ctx->half_frac;
    uint64_t d2h = 0;
    for (int s = 0; s < n_slices && rc == FRA_OK; ++s) {
        const size_t c0 = (size_t)s * per;
        if (c0 >= C) break;
        const size_t nch = std::min(per, C - c0);
        cudaStream_t st = ctx->copy_streams[s % 3];
        FRA_TRY(ctx, cudaMemcpyAsync(ctx->d_in + c0 * n, h_in + c0 * n, nch * n * 2, cudaMemcpyHostToDevice, st));
        fra_outputs o = offset_outputs(dev, c0, (int)n);
        rc = process_range(ctx, ctx->d_in + c0 * n, (int)c0, (int)nch, continuous, log2_scale, o, st);
        if (rc != FRA_OK) break;
        if (h_out->d_filtered) {
            FRA_TRY(ctx, cudaMemcpyAsync(h_out->d_filtered + c0 * n, o.d_filtered, nch * n * 2, cudaMemcpyDeviceToHost, st));
            d2h += nch * n * 2;
        }
        if (h_out->d_frames && half) {
            // the first nh frames of the slice: bins 0 .. N/2 straight into the caller's frames (a pitched copy), the
            // mirror bits into the slot's staging; fra_host_wait completes the upper halves on the host's cores.
            // The other frames of the slice cross the link whole.
            const size_t nh = std::min(nch, (size_t)std::llround((double)nch * half_frac));
            plan.nh[s] = nh;
            if (nh > 0) {
                const size_t total = nh * (n / 2);
                auto kfn = k3_mirror_bits;
                int log2m = ctx->log2n - 1;
                FRA_LAUNCH(kfn, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)0, st,
                           (const uint32_t *)o.d_frames, ctx->d_mbits + c0 * mwords, total, log2m);
                FRA_TRY(ctx, cudaGetLastError());
                ctx->last_kernels++;
                FRA_TRY(ctx, cudaMemcpy2DAsync(h_out->d_frames + c0 * n * 4, n * 4, o.d_frames, n * 4, (n / 2 + 1) * 4, nh,
                                               cudaMemcpyDeviceToHost, st));
                FRA_TRY(ctx, cudaMemcpyAsync(ctx->h_mbits[id & 1] + c0 * mwords, ctx->d_mbits + c0 * mwords, nh * mwords * 4,
                                             cudaMemcpyDeviceToHost, st));
                d2h += nh * ((n / 2 + 1) * 4 + mwords * 4);
            }
            if (nh < nch) {
                FRA_TRY(ctx, cudaMemcpyAsync(h_out->d_frames + (c0 + nh) * n * 4, o.d_frames + nh * n * 4, (nch - nh) * n * 4,
                                             cudaMemcpyDeviceToHost, st));
                d2h += (nch - nh) * n * 4;
            }
        } else if (h_out->d_frames) {
            FRA_TRY(ctx, cudaMemcpyAsync(h_out->d_frames + c0 * n * 4, o.d_frames, nch * n * 4, cudaMemcpyDeviceToHost, st));
            d2h += nch * n * 4;
        }
        if (half) {
            cudaEvent_t &se = ctx->slice_done[id & 1][s];
            if (!se) FRA_TRY(ctx, cudaEventCreateWithFlags(&se, cudaEventDisableTiming));
            FRA_TRY(ctx, cudaEventRecord(se, st));
        }
        if (h_out->d_iq) {
            FRA_TRY(ctx, cudaMemcpyAsync(h_out->d_iq + c0 * n * 2, o.d_iq, nch * n * 8, cudaMemcpyDeviceToHost, st));
            d2h += nch * n * 8;
        }
        if (h_out->d_mag) {
            FRA_TRY(ctx, cudaMemcpyAsync(h_out->d_mag + c0 * n, o.d_mag, nch * n * 4, cudaMemcpyDeviceToHost, st));
            d2h += nch * n * 4;
        }
        if (h_out->d_phase) {
            FRA_TRY(ctx, cudaMemcpyAsync(h_out->d_phase + c0 * n, o.d_phase, nch * n * 4, cudaMemcpyDeviceToHost, st));
            d2h += nch * n * 4;
        }
    }
    if (rc != FRA_OK) {
        host_streams_join(ctx);
        return rc;
    }
    // completion of THIS call on every copy stream; at most two calls are in flight
    for (int s = 0; s < 3; ++s) {
        cudaEvent_t &e = ctx->host_done[id & 1][s];
        if (!e) FRA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        FRA_TRY(ctx, cudaEventRecord(e, ctx->copy_streams[s]));
    }
    if (half) ctx->mirror_plan[id & 1] = plan;
    ctx->last_h2d_bytes = (uint64_t)(C * n * 2);
    ctx->last_d2h_bytes = d2h;
    ctx->host_calls = id;
    *ticket = id;
    return FRA_OK;
}

int fra_get_state(fra_ctx *ctx, int16_t *d_state, void *cuda_stream)
{
    if (!ctx || !d_state) return FRA_ERR_INVALID;
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    FRA_TRY(ctx, pipe_host_join(ctx));                 // FRA_PIPELINE: the state belongs to pipe_k1 until it drains
    FRA_TRY(ctx, host_streams_join(ctx));              // ... or to an asynchronous host call
    cudaStream_t st = (cudaStream_t)cuda_stream;      // NULL = the legacy default stream
    FRA_TRY(ctx, cudaMemcpyAsync(d_state, ctx->d_state, (size_t)ctx->channels * 24 * sizeof(int16_t),
                                 cudaMemcpyDeviceToDevice, st));
    return FRA_OK;
}

int fra_set_state(fra_ctx *ctx, const int16_t *d_state, void *cuda_stream)
{
    if (!ctx || !d_state) return FRA_ERR_INVALID;
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    FRA_TRY(ctx, pipe_host_join(ctx));                 // FRA_PIPELINE: the state belongs to pipe_k1 until it drains
    FRA_TRY(ctx, host_streams_join(ctx));              // ... or to an asynchronous host call
    cudaStream_t st = (cudaStream_t)cuda_stream;      // NULL = the legacy default stream
    FRA_TRY(ctx, cudaMemcpyAsync(ctx->d_state, d_state, (size_t)ctx->channels * 24 * sizeof(int16_t),
                                 cudaMemcpyDeviceToDevice, st));
    return FRA_OK;
}

int fra_iir_stream(fra_ctx *ctx, const int16_t *d_in, int16_t *d_out, size_t n, int continuous, int exact,
                   fra_stream_stats *stats)
{
    if (!ctx || !d_in || !d_out || n == 0 || (n % 8) != 0) return FRA_ERR_INVALID;
    if (ctx->flags & FRA_WINDOW_RTL_SKEW) return FRA_ERR_UNSUPPORTED;      // frames only
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    FRA_TRY(ctx, pipe_host_join(ctx));
    cudaStream_t st = ctx->stream;
    const bool iir = (ctx->mode == FRA_MODE_BANK0 || ctx->mode == FRA_MODE_BANK1);
    const Sections sec = current_sections(ctx);
    fra_stream_stats s = {0, 0, 0, 0, 0, 0};

    // pole radius of z^2 + (A1/128) z + A0/128 (both sets): the block scan needs a stable cascade
    double r = 0.0;
    if (iir && !exact) {
        for (int set = 0; set < kStages; ++set) {
            const double a0 = sec.c[set][3] / 128.0, a1 = sec.c[set][4] / 128.0;
            const double disc = a1 * a1 - 4.0 * a0;
            const double rs = disc < 0.0 ? std::sqrt(a0)
                                         : std::max(std::fabs((-a1 + std::sqrt(disc)) / 2.0), std::fabs((-a1 - std::sqrt(disc)) / 2.0));
            r = std::max(r, rs);
        }
        if (r >= 0.9995) {                                // (A^L) does not decay: use the exact chain
            if ((n % kSplitChunk) != 0) return FRA_ERR_UNSUPPORTED;
            exact = 1;
        }
    }

    ctx->last_kernels = 0;
    // bit-exact path: the six stages as a systolic chain on six lanes of one warp
    auto run_exact = [&]() -> int {
        if ((n % kSplitChunk) != 0 || n > 0x7fffff00u) return FRA_ERR_INVALID;
        K1Args k1;
        k1.in = d_in;
        k1.out = d_out;
        k1.state = ctx->d_state;
        k1.rom32 = ctx->d_rom32;
        k1.rom2x = ctx->d_rom2x;
        k1.coef = make_cascade(sec);
        k1.channels = 1;
        k1.n = (int)n;
        k1.continuous = continuous;
        k1.speculate = (ctx->flags & FRA_K1_SPECULATE) ? 1 : 0;
        const size_t smem = (size_t)kSplitWarps * kSplitSmemPerWarp;
        auto kfn = k1_split;
        FRA_SMEM(ctx, kfn, (int)smem);
        FRA_LAUNCH(kfn, dim3(1), dim3(kSplitWarps * 32), smem, st, k1);
        FRA_TRY(ctx, cudaGetLastError());
        ctx->last_kernels++;
        FRA_TRY(ctx, cudaStreamSynchronize(st));
        s.exact = 1;
        return FRA_OK;
    };
    if (exact && iir) {
        int rc = run_exact();
        s.n_chunks = 1;
        s.chunk = (int)n;
        if (stats) *stats = s;
        return rc;
    }

    // Chunk length = samples per lane.  Long streams get short chunks so that there are lanes for the whole
    // machine (2^26 samples / 512 = 131072 lanes); the scan cost per chunk (a few 24 x 24 products) is small
    // against 512 samples of filtering.  Short streams keep 4096 (fewer boundaries, fewer LSB of dead band).
    const int chunk = (n >= ((size_t)1 << 20)) ? 512 : 4096;
    const int n_chunks = (int)((n + chunk - 1) / chunk);
    const int n_warps = (n_chunks + 31) / 32;
    if (n_chunks > ctx->k1b_capacity) {
        void *old[] = {ctx->d_entry, ctx->d_exit, ctx->d_ends, ctx->d_aggr};
        for (void *p : old)
            if (p) cudaFree(p);
        ctx->d_entry = ctx->d_exit = nullptr;
        ctx->d_ends = ctx->d_aggr = nullptr;
        ctx->k1b_capacity = 0;
        if (cudaMalloc((void **)&ctx->d_entry, (size_t)n_chunks * 24 * sizeof(int16_t)) != cudaSuccess ||
            cudaMalloc((void **)&ctx->d_exit, (size_t)n_chunks * 24 * sizeof(int16_t)) != cudaSuccess ||
            cudaMalloc((void **)&ctx->d_ends, (size_t)n_chunks * 24 * sizeof(float)) != cudaSuccess ||
            cudaMalloc((void **)&ctx->d_aggr, (size_t)4 * n_warps * 24 * sizeof(float)) != cudaSuccess)
            return FRA_ERR_NOMEM;
        ctx->k1b_capacity = n_chunks;
    }
    if (!ctx->d_counts && cudaMalloc((void **)&ctx->d_counts, 2 * sizeof(int)) != cudaSuccess) return FRA_ERR_NOMEM;
    if (!ctx->d_mats && cudaMalloc((void **)&ctx->d_mats, (size_t)kScanMats * 576 * sizeof(float)) != cudaSuccess)
        return FRA_ERR_NOMEM;

    K1bArgs a;
    a.in = d_in;
    a.out = d_out;
    a.rom32 = ctx->d_rom32;
    a.coef = make_cascade(sec);
    a.entry = ctx->d_entry;
    a.exit_ = ctx->d_exit;
    a.state0 = ctx->d_state;
    a.stats = ctx->d_counts;
    a.ends = ctx->d_ends;
    a.aggr = ctx->d_aggr;
    a.mats = ctx->d_mats;
    a.n = n;
    a.chunk = chunk;
    a.n_chunks = n_chunks;
    a.n_warps = n_warps;
    a.continuous = continuous;
    a.apply_window = 1;
    a.iir = iir ? 1 : 0;
    a.aggr_levels = 0;
    FRA_TRY(ctx, cudaMemsetAsync(ctx->d_counts, 0, 2 * sizeof(int), st));

    if (iir && n_chunks > 1) {
        // state-space matrices of the float model: one sample step A (24 x 24), then
        // M = A^L and M^2 .. M^16, Q = M^32 and Q^2, Q^4, ... by squaring, in double.  Kept on the device
        // until the coefficients or the chunk length change (a 2^26-sample call is 0.6 ms: the host-side
        // powering and the upload would be a sixth of it).
        constexpr int D = kStateDim;
        static_assert(kScanMats <= 32, "k1b_norm");
        if (ctx->k1b_key_chunk != chunk || std::memcmp(ctx->k1b_key, sec.c, 36) != 0) {
            const CascadeCoef cc = a.coef;
            auto step = [&](const double *in, double *out) {           // u = 0
                double v = 0.0;
                for (int sg = 0; sg < kStages; ++sg) {
                    const StageCoef &k = cc.set[sg];
                    const double x1 = in[4 * sg], x2 = in[4 * sg + 1], y1 = in[4 * sg + 2], y2 = in[4 * sg + 3];
                    const double y = (double)k.b2 * v + (double)k.b1 * x1 + (double)k.b0 * x2 + (double)k.na0 * y2 + (double)k.na1 * y1;
                    out[4 * sg] = v; out[4 * sg + 1] = x1; out[4 * sg + 2] = y; out[4 * sg + 3] = y1;
                    v = y;
                }
            };
            std::vector<double> A(D * D), M(D * D), T(D * D), R(D * D);
            for (int j = 0; j < D; ++j) {
                double e[D] = {0}, o[D];
                e[j] = 1.0;
                step(e, o);
                for (int i = 0; i < D; ++i) A[i * D + j] = o[i];
            }
            auto matmul = [&](const std::vector<double> &x, const std::vector<double> &y, std::vector<double> &z) {
                for (int i = 0; i < D; ++i)
                    for (int j = 0; j < D; ++j) {
                        double acc = 0.0;
                        for (int k = 0; k < D; ++k) acc += x[i * D + k] * y[k * D + j];
                        z[i * D + j] = acc;
                    }
            };
            // R = A^chunk by binary powering
            for (int i = 0; i < D * D; ++i) R[i] = (i / D == i % D) ? 1.0 : 0.0;
            M = A;
            for (int e = chunk; e > 0; e >>= 1) {
                if (e & 1) { matmul(R, M, T); R = T; }
                matmul(M, M, T);
                M = T;
            }
            std::vector<float> mats((size_t)kScanMats * D * D, 0.0f);
            M = R;                                                       // (A^L)^1
            for (int lvl = 0; lvl < kScanMats; ++lvl) {                  // M^(1..16), then Q^(2^j) with Q = M^32
                double mx = 0.0;
                for (int i = 0; i < D * D; ++i) {
                    mats[(size_t)lvl * D * D + i] = (float)M[i];
                    mx = std::max(mx, std::fabs(M[i]));
                }
                ctx->k1b_norm[lvl] = mx;
                matmul(M, M, T);
                M = T;
                for (int i = 0; i < D * D; ++i)
                    if (!std::isfinite(M[i])) M[i] = 0.0;                // (cannot happen for r < 0.9995; keeps the table finite)
            }
            FRA_TRY(ctx, cudaMemcpyAsync(ctx->d_mats, mats.data(), mats.size() * sizeof(float), cudaMemcpyHostToDevice, st));
            FRA_TRY(ctx, cudaStreamSynchronize(st));                    // `mats` lives on this stack frame
            std::memcpy(ctx->k1b_key, sec.c, 36);
            ctx->k1b_key_chunk = chunk;
        }
        // a level of the aggregate scan matters while its matrix can move a state value (|s| <= 2^15, 24 terms)
        // by more than a hundredth of an LSB, and while there are warps that far apart
        int aggr_levels = 0;
        for (int j = 0; j < kAggrLevels; ++j)
            if (ctx->k1b_norm[kScanLevels + j] * 24.0 * 32768.0 > 0.01 && (1 << j) < n_warps) aggr_levels = j + 1;
        if (aggr_levels == kAggrLevels && (1 << kAggrLevels) < n_warps) return FRA_ERR_UNSUPPORTED;   // stream too long for the table
        // (two levels always run when there are warps to chain: with int8 coefficients and 512-sample chunks
        // Q = A^16384 is negligible for every cascade this path accepts, and the scan code should not depend on that)
        a.aggr_levels = std::max(aggr_levels, std::min(2, n_warps > 2 ? 2 : 0));

        auto l0 = k1b_lin_ends;
        FRA_LAUNCH(l0, dim3((n_chunks + 63) / 64), dim3(64), (size_t)0, st, a);
        FRA_TRY(ctx, cudaGetLastError());
        auto s0 = k1b_scan_warp<0>;
        FRA_LAUNCH(s0, dim3(n_warps), dim3(32), (size_t)0, st, a);
        FRA_TRY(ctx, cudaGetLastError());
        {
            // S_w = G_w + Q S_{w-1}: Hillis-Steele over the aggregates, ping-pong in aggr[2], aggr[3]
            float *bufs[2] = {ctx->d_aggr + (size_t)2 * n_warps * 24, ctx->d_aggr + (size_t)3 * n_warps * 24};
            const float *cur = ctx->d_aggr;                          // G_w
            const dim3 g((unsigned)((n_warps + 3) / 4));
            auto lv = k1b_aggr_level;
            for (int j = 0; j < a.aggr_levels; ++j) {
                FRA_LAUNCH(lv, g, dim3(128), (size_t)0, st, a, j, cur, bufs[j & 1]);
                FRA_TRY(ctx, cudaGetLastError());
                cur = bufs[j & 1];
                ctx->last_kernels++;
            }
            auto cy = k1b_aggr_carry;
            FRA_LAUNCH(cy, g, dim3(128), (size_t)0, st, a, cur);
            FRA_TRY(ctx, cudaGetLastError());
        }
        auto s1 = k1b_scan_warp<1>;
        FRA_LAUNCH(s1, dim3(n_warps), dim3(32), (size_t)0, st, a);
        FRA_TRY(ctx, cudaGetLastError());
        ctx->last_kernels += 4;
    } else if (n_chunks > 1) {
        FRA_TRY(ctx, cudaMemsetAsync(ctx->d_entry, 0, (size_t)n_chunks * 24 * sizeof(int16_t), st));   // bypass: no history
    }
    auto kfn = k1b_speculate;
    FRA_LAUNCH(kfn, dim3((n_chunks + 63) / 64), dim3(64), (size_t)0, st, a);
    FRA_TRY(ctx, cudaGetLastError());
    auto vfn = k1b_verify;
    FRA_LAUNCH(vfn, dim3((n_chunks + 127) / 128), dim3(128), (size_t)0, st, a);
    FRA_TRY(ctx, cudaGetLastError());
    ctx->last_kernels += 2;
    int counts[2] = {0, 0};
    FRA_TRY(ctx, cudaMemcpyAsync(counts, ctx->d_counts, sizeof(counts), cudaMemcpyDeviceToHost, st));
    FRA_TRY(ctx, cudaStreamSynchronize(st));
    s.n_chunks = n_chunks;
    s.chunk = chunk;
    s.warmup = 0;
    s.n_mismatch = iir ? counts[0] : 0;
    s.max_state_dev = iir ? counts[1] : 0;
    int rc = FRA_OK;
    // the dead band of a truncating section grows like 1 / (1 - r^2) (measured: 6-7 LSB at r^2 = 0.84,
    // 65 LSB at r^2 = 0.984); far beyond it the trajectories have separated
    const int dead_band = std::max(kStreamMaxDeadband, (int)(6.0 / std::max(1e-3, 1.0 - r * r)));
    if (iir && counts[1] > dead_band) {
        // beyond the dead band: the cascade is overflowing (16-bit wrap) or barely
        // stable, trajectories do not stay together - recompute exactly.  The exact chain
        // needs n % 256 == 0; otherwise the approximate output is NOT handed out as valid
        // and channel 0's history is left as it was.
        if ((n % kSplitChunk) != 0) {
            if (stats) *stats = s;
            return FRA_ERR_UNSUPPORTED;
        }
        rc = run_exact();
    } else if (iir) {
        // the stream's end state becomes channel 0's history (continuous operation)
        FRA_TRY(ctx, cudaMemcpyAsync(ctx->d_state, ctx->d_exit + (size_t)(n_chunks - 1) * 24, 24 * sizeof(int16_t),
                                     cudaMemcpyDeviceToDevice, st));
        FRA_TRY(ctx, cudaStreamSynchronize(st));
    }
    if (stats) *stats = s;
    return rc;
}

int fra_fft_only(fra_ctx *ctx, const int16_t *d_in, int batch, float *d_iq, void *cuda_stream)
{
    if (!ctx || !d_in || !d_iq || batch <= 0) return FRA_ERR_INVALID;
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;      // NULL = the legacy default stream
    K2Args k2;
    k2.in = reinterpret_cast<const uint32_t *>(d_in);
    k2.rom32 = ctx->d_rom32;
    k2.tw1 = ctx->d_tw1;
    k2.tw2 = ctx->d_tw2;
    k2.twn = ctx->d_twn;
    k2.frames = nullptr;
    k2.iq = reinterpret_cast<float2 *>(d_iq);
    k2.mag = nullptr;
    k2.phase = nullptr;
    k2.qscale = std::ldexp(0.5f, -ctx->log2n);
    k2.mag_alpha = 1.0f;
    k2.batch = batch;
    k2.frame0 = 0;
    k2.prefetch = ctx->sm_count * FRA_K2_MINBLOCKS;
#ifdef FRA_TIMELINE
    k2.tl_step = -1;
#endif
    k2.exp23 = 0x4B000000u;
    ctx->last_kernels = 0;
    return launch_k2(ctx, k2, /*win=*/false, /*qmode=*/0, st);
}

int fra_profile_enable(fra_ctx *ctx, int on)
{
    if (!ctx) return FRA_ERR_INVALID;
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    if (on)
        for (auto &e : ctx->ev)
            if (!e) FRA_TRY(ctx, cudaEventCreateWithFlags(&e, 0));
    ctx->profiling = on != 0;
    ctx->ev_k1 = ctx->ev_k2 = false;
    return FRA_OK;
}

int fra_profile_last(fra_ctx *ctx, float *ms_window_iir, float *ms_fft_pack)
{
    if (!ctx || !ctx->profiling) return FRA_ERR_INVALID;
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    float a = 0.0f, b = 0.0f;
    if (ctx->ev_k1) {
        FRA_TRY(ctx, cudaEventSynchronize(ctx->ev[1]));
        FRA_TRY(ctx, cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]));
    }
    if (ctx->ev_k2) {
        FRA_TRY(ctx, cudaEventSynchronize(ctx->ev[3]));
        FRA_TRY(ctx, cudaEventElapsedTime(&b, ctx->ev[2], ctx->ev[3]));
    }
    if (ms_window_iir) *ms_window_iir = a;
    if (ms_fft_pack) *ms_fft_pack = b;
    return FRA_OK;
}

int fra_sync(fra_ctx *ctx)
{
    if (!ctx) return FRA_ERR_INVALID;
    FRA_TRY(ctx, cudaSetDevice(ctx->device));
    FRA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    FRA_TRY(ctx, pipe_host_join(ctx));
    FRA_TRY(ctx, host_streams_join(ctx));
    return FRA_OK;
}

#ifdef FRA_TIMELINE
// diagnostic build only: reset (out == NULL) or read the [4][4096] timeline
int fra_debug_timeline(unsigned long long *out)
{
    if (!out) {
        static unsigned long long init[4][4096];
        for (int s = 0; s < 4096; ++s) { init[0][s] = init[2][s] = ~0ULL; init[1][s] = init[3][s] = 0; }
        return cudaMemcpyToSymbol(g_timeline, init, sizeof(init)) == cudaSuccess ? FRA_OK : FRA_ERR_CUDA;
    }
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, g_timeline, sizeof(unsigned long long) * 4 * 4096) == cudaSuccess ? FRA_OK : FRA_ERR_CUDA;
}
#endif

int fra_last_kernel_count(const fra_ctx *ctx) { return ctx ? ctx->last_kernels : FRA_ERR_INVALID; }

const char *fra_last_cuda_error(const fra_ctx *ctx) { return ctx ? ctx->err : "null context"; }

}  // extern "C"
