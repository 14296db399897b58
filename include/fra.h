/* fra.h - C ABI of libfra.so, the B200-native receive chain
 *   window ROM multiply -> 12th-order IIR (6 int16 x int8 biquads, 2 banks)
 *   -> N-point FFT -> int16 I/Q bin framing (+ optional fp32 I/Q, magnitude, phase)
 *
 * The reference (mfkiwl/fpga-real-time-fft-analyzer) is an FPGA design with a
 * Python GUI; it has no FFI.  The boundary this library sits behind is the GUI's
 * receiver-backend contract (scripts/fft_analyzer_gui.py, "GUI" below): 65536-byte
 * frames out, 1-byte commands (+ 0xF1 and 12 coefficient bytes) in.  Every entry
 * point cites the reference interface it replaces.  Paths: NEW/ =
 * SDR_v2.srcs/sources_1/new/, IMP/ = SDR_v2.srcs/sources_1/imports/new/.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 (FRA_OK)
 * or a negative fra_status; nothing throws; no global mutable state.  A context
 * is thread-compatible (one thread at a time), matching the GUI's Qt-main-thread
 * discipline (GUI:1009-1053).  Pointers named d_* are device pointers on the
 * context's device, h_* are host pointers.  Calls that take a `cuda_stream` enqueue
 * on exactly that stream (a cudaStream_t; NULL = the CUDA legacy default stream) and
 * return without waiting; the caller orders them against its own work as with any
 * CUDA library.  Calls without a stream argument are synchronous.  Launch errors
 * surface from the call that made them.
 * There is NO CPU fallback: without a CUDA device fra_create fails with
 * FRA_ERR_NO_DEVICE.
 */
#ifndef FRA_H
#define FRA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRA_ABI_VERSION 1

typedef enum fra_status {
    FRA_OK = 0,
    FRA_ERR_INVALID = -1,      /* bad argument (null pointer, size, mode byte) */
    FRA_ERR_NO_DEVICE = -2,    /* no CUDA device / driver */
    FRA_ERR_CUDA = -3,         /* a CUDA call or kernel failed; see fra_last_cuda_error */
    FRA_ERR_NOMEM = -4,
    FRA_ERR_UNSUPPORTED = -5,  /* e.g. fft_size not in 1024..65536 */
    FRA_ERR_BUSY = -6          /* byte stream ended inside a 0xF1 upload (informational) */
} fra_status;

/* Command bytes, GUI:28-37; decoded by NEW/command_control.vhd:46-74,
 * NEW/rx_filter_coeff.vhd:41-66 and IMP/sequ2.vhd:83-96,216. */
#define FRA_CMD_UART_REQUEST  0xA5
#define FRA_CMD_RESET         0xFF
#define FRA_CMD_ETHERNET_MODE 0xEF
#define FRA_CMD_UART_MODE     0xFE
#define FRA_CMD_START         0x55
#define FRA_CMD_FILTER_UPDATE 0xF1
#define FRA_MODE_BANK0        0x00   /* fixed filter, IMP/filter_iir12.vhd + IMP/filter_pkg.vhd:54-68 */
#define FRA_MODE_BANK1        0xA1   /* loadable filter, NEW/filter_iir12_cust.vhd */
#define FRA_MODE_BYPASS       0xB1   /* reset default, NEW/command_control.vhd:31,50 */

#define FRA_WINDOW_LEN   16384       /* NEW/hann.vhd: Hann_ROM(0 to 16383) */
#define FRA_FRAME_BYTES(n) ((size_t)4 * (size_t)(n))   /* GUI:41-44: 65536 at n = 16384 */

/* fra_create flags */
#define FRA_ROUND_NEAREST   0x1u     /* int16 bins rounded to nearest-even instead of truncated (floor) */
#define FRA_K1_FORCE_LANE   0x2u     /* always use the lane-per-channel window+IIR kernel */
/* 0x10u was FRA_K1_FORCE_STAGE (round 1's warp-per-stage kernel, superseded by the stage-pair kernel and removed) */
#define FRA_K1_FORCE_DUO    0x20u    /* always use the two-stages-per-warp pipeline window+IIR kernel */
#define FRA_PIPELINE        0x40u    /* fra_process runs the window+IIR of call i+1 beside the FFT of call i (which is
                                        enqueued by call i+1, or by fra_join / fra_sync after the last call) on two
                                        internal streams (the FPGA does the same: the filter streams the next frame
                                        while xfft_0 unloads the previous one, dsp_system_top.vhd:530-567).  The call
                                        only waits for `cuda_stream` (inputs ready); inputs must stay untouched and
                                        outputs are complete after fra_join() / fra_sync(). */
#define FRA_K1_FORCE_SPLIT  0x4u     /* always use the stage-per-lane (systolic) window+IIR kernel */
#define FRA_K1_SPECULATE    0x8u     /* systolic kernel: try the no-overflow recurrence (FFMA->FADD) first and roll a block
                                        back when a sum left the int16 range; same results, measured no faster on B200 */

#define FRA_K1_NO_BIASED    0x80u    /* never use the all-biased biquad step (five FFMAs + one PRMT per stage);
                                        same results - for A/B timing and to test the general step with eligible coefficients */

#define FRA_K2_STAGED       0x100u   /* 16K frames: persistent FFT CTAs whose frames arrive by one bulk copy each (cp.async.bulk on an
                                        mbarrier) instead of one frame per CTA with per-thread loads; same results, measured
                                        slower on B200, kept for A/B timing */
#define FRA_WINDOW_RTL_SKEW 0x1000u  /* the window with the register skew of the RTL as written (NEW/hann8192.vhd:36-39):
                                        coef_s, product and sample_out update in one clocked branch, so output n is the rounding
                                        of x[n-1] * ROM[n-2], and the first two outputs after a gap come from zero registers.
                                        Default (flag clear) is the aligned window, y[n] = f(x[n], ROM[n]).  Frames only
                                        (fra_iir_stream and FRA_PIPELINE return FRA_ERR_UNSUPPORTED). */
#define FRA_HOST_HALF_SPECTRUM 0x800u /* fra_process_host[_async]: the int16 frames cross PCIe as bins 0..N/2 plus one bit per bin
                                        (33 KiB instead of 64 KiB per 16K frame) and fra_host_wait completes the Hermitian upper
                                        half on the host's cores - X[N-j] = {re, -im - bit}, the bit saying whether floor(-Im s)
                                        is -floor(Im s) or one less.  Byte-identical frames; only with the default truncating
                                        scale (otherwise, and for the other outputs, the full transfer is used).  Pays when the
                                        device-to-host link is the bottleneck and the host has cores and memory bandwidth to spare. */
#define FRA_K2_64K_SPLIT     0x400u   /* 64K frames through HBM (even / odd split, two 32K transforms, radix-2 join: 28 B of traffic per
                                        sample) instead of on chip in a cluster of two CTAs exchanging through distributed shared
                                        memory; same results within fp32 rounding, for A/B timing */
#define FRA_K2_WIDE_CTA      0x2000u  /* 32K frames in one 128 KiB / 512-thread CTA (one per SM) instead of a cluster of two 64 KiB /
                                        256-thread CTAs (two per SM) exchanging through distributed shared memory; same results
                                        within fp32 rounding, for A/B timing */
#define FRA_FFT_FIXED16     0x200u   /* the FFT as the 16-bit fixed-point, scaled, truncating radix-2^2 pipeline the Xilinx core is
                                        configured to be (IP/xfft_0/xfft_0.xci:12-27: 16-bit data and phase factors, scaled 1/N,
                                        truncation, natural order) instead of fp32: the spectrum carries FPGA-like quantisation
                                        noise.  Arithmetic defined in csrc/k2_fixed.cuh / oracle/fixed_fft.py; bit-level parity with
                                        the proprietary core is UNPINNED.  fft_size <= 32768, log2_scale must be the default (1/N);
                                        d_iq receives the int16 bins times N as floats. */

typedef struct fra_ctx fra_ctx;      /* opaque: ROM, two coefficient banks, IIR state, twiddles, one stream */

/* Lifetime.  n_channels independent channels, fft_size in {1024,...,65536}
 * (the reference is fixed at 16384: IP/xfft_0/xfft_0.xci:12, IMP/dsp_system_top.vhd:438-449).
 * After create the control state is the RTL's reset state: mode 0xB1, bank 1 all
 * zero, IIR history zero, window address 0. */
int fra_create(fra_ctx **out, int device, int n_channels, int fft_size, unsigned flags);
int fra_destroy(fra_ctx *ctx);

/* Byte protocol, exactly as it arrives on the reference's UART
 * (UartReceiver.send_command / send_filter_coefficients, GUI:567-613):
 * 0x00/0xA1/0xB1 select the stream fed to the FFT; 0xFF = reset; 0xF1 starts a
 * 12-byte upload into bank 1 during which every byte is data, not a command
 * (rx_filter_coeff 'busy', NEW/command_control.vhd:51); 0x55/0xA5/0xEF/0xFE are
 * accepted and recorded (start / frame request / transport) but do not change the
 * arithmetic.  Takes effect for the next fra_process call; filter history is NOT
 * cleared by an upload or a mode change.  Returns FRA_ERR_BUSY (not a
 * failure) if the stream ends inside an upload; the rest may follow in a later call. */
int fra_command(fra_ctx *ctx, const uint8_t *bytes, size_t n);

/* Same effects without the byte stream. */
int fra_load_bank1(fra_ctx *ctx, const int8_t coeff[12]);

/* Superset of the 12-byte protocol: six INDEPENDENT sections, 6 bytes
 * each in the RTL's register order B0,B1,B2,A0,A1,A2 (B2 -> x[n], B1 -> x[n-1], B0 -> x[n-2],
 * A1 -> y[n-1], A0 -> y[n-2], A2 unconnected; products >> 7), stage 1 first.  Becomes bank 1
 * (mode 0xA1) until the next 0xF1 upload / fra_load_bank1, which restores the RTL's two
 * alternating sets (filter_iir12_cust.vhd:68-240).  Takes effect at the next frame boundary;
 * the IIR history is kept.  fra_get_sections returns the 36 bytes the CURRENT mode filters with
 * (bank 0 and 12-byte uploads expanded ALPHA, BETA, ALPHA, ...). */
int fra_load_sections(fra_ctx *ctx, const int8_t coeff[36]);
int fra_get_sections(const fra_ctx *ctx, int8_t coeff[36]);  /* 0xF1 payload: B0,B1,B2,A0,A1,A2 x {set0,set1}, NEW/filter_iir12_cust.vhd:83-94 */
int fra_set_mode(fra_ctx *ctx, uint8_t mode);              /* 0x00 | 0xA1 | 0xB1, NEW/command_control.vhd:53-58 */
int fra_reset(fra_ctx *ctx);                               /* 0xFF: history 0, bank 1 = 0, mode 0xB1, window address 0 */

/* Introspection of the control state (what the RTL holds in registers). */
int fra_get_mode(const fra_ctx *ctx, uint8_t *mode);
int fra_get_bank(const fra_ctx *ctx, int bank, int8_t coeff[12]);   /* bank 0 = IMP/filter_pkg.vhd constants */
int fra_get_transport(const fra_ctx *ctx, uint8_t *transport);      /* 0xEF | 0xFE, IMP/sequ2.vhd:83-96 */
int fra_get_counters(const fra_ctx *ctx, uint64_t *n_start, uint64_t *n_request, uint64_t *n_reset, uint64_t *n_upload);
int fra_window_rom(int16_t out[FRA_WINDOW_LEN]);                    /* the ROM of NEW/hann.vhd:5-16390 */

/* Optional outputs of one processing step; any pointer may be NULL. */
typedef struct fra_outputs {
    int16_t *d_filtered;  /* [C][N] int16: the FFT's input stream (window, then the selected filter) */
    uint8_t *d_frames;    /* [C][4N] bytes: per bin re_lo,re_hi,im_lo,im_hi (IMP/sequ2.vhd:153,234; GUI:255-257) */
    float   *d_iq;        /* [C][N][2] fp32 unscaled DFT bins (what np.fft.fft returns) */
    float   *d_mag;       /* [C][N] fp32 sqrt(re^2+im^2) of the int16 bins, bit-identical to GUI decode_mag_16iq_le (GUI:250-260) */
    float   *d_phase;     /* [C][N] fp32 atan2(im, re) of the int16 bins (the reference computes no phase) */
} fra_outputs;

/* One step: every channel's next N samples.
 *   d_in        [C][N] int16, channel-major (sample n of channel c at c*N + n)
 *   continuous  != 0: i_valid held high across the frame boundary - IIR history
 *               carries over from the previous step (NEW/filter_iir_cust.vhd:139-193);
 *               == 0: a burst after a gap - history starts from zero.
 *   The window address always restarts at 0 for a frame (N-sample frames aligned to
 *   the 14-bit address counter, NEW/hann8192.vhd:23,41); for N < 16384 the first N
 *   ROM entries are used, for N = 32768 the address wraps.
 * int16 bins = floor(X[k] * 2^log2_scale) (xfft 'scaled' + 'truncation',
 * IP/xfft_0/xfft_0.xci; default schedule = 1/N), saturated.  Pass
 * FRA_SCALE_DEFAULT for 1/N. */
#define FRA_SCALE_DEFAULT 0x7fffffff
int fra_process(fra_ctx *ctx, const int16_t *d_in, int continuous, int log2_scale,
                const fra_outputs *out, void *cuda_stream);

/* Spectrum averaging fused into the pack stage (the reference README lists waterfall / averaging
 * display options under "Contributing").  alpha = 1 (default): d_mag
 * receives |bin| as decode_mag_16iq_le computes it.  0 < alpha < 1: the buffer passed as d_mag is
 * read and updated in place, mag <- mag + alpha (|bin| - mag), i.e. it holds an exponential moving
 * average across calls and must be the same (initially zeroed or primed) buffer every call. */
int fra_set_mag_average(fra_ctx *ctx, float alpha);

/* The same step through host buffers: H2D of h_in, fra_process, D2H of the requested
 * outputs, synchronised on return.  The copies are cudaMemcpyAsync straight from / into the
 * caller's buffers (no staging copy inside the library): page-locked buffers (cudaHostAlloc,
 * cudaHostRegister, torch pin_memory) overlap with the kernels; pageable ones still work but
 * every copy then blocks the calling thread.  This is the call a GpuReceiver backend makes
 * per batch of frames; h_out fields are HOST pointers. */
int fra_process_host(fra_ctx *ctx, const int16_t *h_in, int continuous, int log2_scale,
                     const fra_outputs *h_out);

/* The same without the final wait: everything (H2D, kernels, D2H) is enqueued on the context's
 * copy streams and *ticket identifies the call; fra_host_wait(ctx, ticket) blocks until its
 * outputs are in host memory.  Two calls may be in flight (a third waits for the first), so a
 * receiver loop uploads frame i+1 while frame i is still downloading - PCIe is full duplex and
 * the D2H side (4 B per sample) is the longer one.  The caller keeps h_in and the buffers in
 * *h_out untouched until the wait returns, i.e. alternates two sets of host buffers.
 * (UdpReceiver / UartReceiver do the same with their byte buffers while a frame is decoded,
 * scripts/fft_analyzer_gui.py:373-455.) */
int fra_process_host_async(fra_ctx *ctx, const int16_t *h_in, int continuous, int log2_scale,
                           const struct fra_outputs *h_out, uint64_t *ticket);
int fra_host_wait(fra_ctx *ctx, uint64_t ticket);

/* FRA_HOST_HALF_SPECTRUM contexts: the share of a call's frames that cross the link as half spectra (the rest go
 * whole).  Half spectra halve the device-to-host link traffic but cost host memory traffic (the mirror reads and
 * writes what the DMA would only have written); which of the two binds depends on the host.  share in [0, 1] fixes
 * it; a negative value (the default) lets the library measure: over the first 14-17 calls it tries 1, 3/4, 1/2 and
 * the neighbours of the best (three calls each, timing the cadence of the caller's fra_process_host_async calls) and
 * then keeps the fastest.  fra_get_host_transfer reports the bytes the last fra_process_host[_async] call moved each
 * way, the share the next call will use, and for the newest half-spectrum call that fra_host_wait finished how long
 * it was blocked on the copies and how long the mirror ran on after them (seconds; any pointer may be NULL).  The
 * frames are byte-identical whatever the share. */
int fra_set_host_half_share(fra_ctx *ctx, double share);
int fra_get_host_transfer(const fra_ctx *ctx, uint64_t *h2d_bytes, uint64_t *d2h_bytes, double *half_share, double *wait_s,
                          double *mirror_s);

/* IIR history, [C][6][4] int16 = per stage (x[n-1], x[n-2], y[n-1], y[n-2])
 * (registers ve(1..2), vs(1..2) of NEW/filter_iir_cust.vhd:45-46). Device pointers. */
int fra_get_state(fra_ctx *ctx, int16_t *d_state, void *cuda_stream);
int fra_set_state(fra_ctx *ctx, const int16_t *d_state, void *cuda_stream);

/* Single long stream (BASELINE config 5): window + IIR12 of n samples of ONE
 * channel (channel 0's history; window address free-running from 0) with the
 * currently selected mode.
 *   exact != 0  bit-exact: the six stages run as a systolic chain on six warp lanes
 *               (the stage-per-lane kernel with one channel); n must be a multiple of 256.
 *   exact == 0  time-parallel chunked scan: one lane per 4096-sample chunk; each chunk's
 *               entry state comes from a block scan of the cascade's state-space
 *               recurrence (float model, 24 x 24 chunk-transition powers combined by warp
 *               shuffles); the chunks are then filtered exactly from those states and
 *               neighbours verified by warp shuffles.  Because every product is truncated
 *               (NEW/filter_iir_cust.vhd:96-100) the cascade is not linear and this path
 *               is NOT bit-exact: it stays within the dead band of the truncating
 *               sections, a few LSB from the serial result; `stats` reports how many
 *               chunk boundaries differ and by how many LSB.  n multiple of 8.  Falls
 *               back to the exact path when the poles are too close to the unit circle
 *               or the deviation shows an overflowing (wrapping) cascade (stats->exact);
 *               that fallback needs n % 256 == 0 like the exact path itself - otherwise
 *               the call returns FRA_ERR_UNSUPPORTED, d_out is not valid and the history
 *               is left untouched.
 * Synchronous. */
typedef struct fra_stream_stats {
    int exact;          /* 1 if the exact path produced the output */
    int n_chunks;
    int chunk;          /* samples per chunk */
    int warmup;         /* exact warm-up samples per chunk (0: the block scan predicts the entry states) */
    int n_mismatch;     /* chunk boundaries whose predicted entry state != neighbour's exit state */
    int max_state_dev;  /* largest such difference, LSB */
} fra_stream_stats;
int fra_iir_stream(fra_ctx *ctx, const int16_t *d_in, int16_t *d_out, size_t n,
                   int continuous, int exact, fra_stream_stats *stats);

/* FFT alone (BASELINE config 5 size sweep): batch frames of fft_size int16
 * samples -> fp32 bins, no window, no filter. */
int fra_fft_only(fra_ctx *ctx, const int16_t *d_in, int batch, float *d_iq, void *cuda_stream);

/* Per-kernel device timing of fra_process (bench.py's roofline): when enabled,
 * CUDA events are recorded on the launching stream around the window+IIR kernel
 * and around the FFT+pack kernel.  fra_profile_last waits for the last step and
 * returns the two durations in milliseconds (0 for a kernel that did not run). */
int fra_profile_enable(fra_ctx *ctx, int on);
int fra_profile_last(fra_ctx *ctx, float *ms_window_iir, float *ms_fft_pack);

int fra_sync(fra_ctx *ctx);                       /* host waits for everything the context has enqueued */
/* Device-side join for FRA_PIPELINE contexts: `cuda_stream` waits for all work enqueued by
 * earlier fra_process calls (no host synchronisation).  A no-op without FRA_PIPELINE, where
 * fra_process already runs on the caller's stream. */
int fra_join(fra_ctx *ctx, void *cuda_stream);
int fra_last_kernel_count(const fra_ctx *ctx);    /* kernels launched by the last fra_process */
const char *fra_last_cuda_error(const fra_ctx *ctx);
const char *fra_strerror(int status);
int fra_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FRA_H */
