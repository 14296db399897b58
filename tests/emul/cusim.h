// cusim.h - a minimal "CUDA on host threads" shim.  TEST INFRASTRUCTURE ONLY.
//
// Purpose: there is no GPU in the build container, and a round trip to a B200
// box takes minutes.  This header lets tests/emul build the UNCHANGED kernel
// sources (csrc/*.cu, compiled as C++ with -DFRA_HOST_EMUL) into
// tests/emul/libfra_emul.so, where every CUDA thread of a block is an OS thread,
// __syncthreads() is a barrier, warp shuffles go through a per-warp mailbox, and
// the CUDA runtime calls used by the host side are mapped to malloc/memcpy.
// It exists to catch index/sign/barrier bugs in the kernel source before GPU
// time is spent.  It is never built into, loaded by, or reachable from the
// product library libfra.so or the Python package.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

// ---------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct short2 { short x, y; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }

namespace cusim {

struct Warp {
    std::barrier<> bar;
    uint64_t mail[32];
    explicit Warp(int n) : bar(n) {}
};

struct Block {
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
    explicit Block(int nthreads) : bar(nthreads) {
        for (int w = 0; w * 32 < nthreads; ++w)
            warps.emplace_back(new Warp(std::min(32, nthreads - 32 * w)));
    }
};

inline thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local Block *t_block = nullptr;
inline thread_local Warp *t_warp = nullptr;
inline thread_local int t_lane = 0;
inline unsigned char *g_dyn_smem = nullptr;   // one block runs at a time

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body)
{
    const int nthreads = (int)(block.x * block.y * block.z);
    std::vector<unsigned char> smem(smem_bytes + 128);
    g_dyn_smem = (unsigned char *)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block blk(nthreads);
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t) {
            th.emplace_back([&, t] {
                t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                t_blockIdx = dim3(bx, by, bz);
                t_blockDim = block;
                t_gridDim = grid;
                t_block = &blk;
                t_warp = blk.warps[t / 32].get();
                t_lane = t % 32;
                std::fesetround(FE_TONEAREST);
                body();
            });
        }
        for (auto &x : th) x.join();
    }
    g_dyn_smem = nullptr;
}

template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

template <class T> inline T shfl_from(T v, int src)
{
    Warp *w = t_warp;
    w->mail[t_lane] = to_bits(v);
    w->bar.arrive_and_wait();
    T r = from_bits<T>(w->mail[src & 31]);
    w->bar.arrive_and_wait();
    return r;
}

}  // namespace cusim

#define threadIdx (cusim::t_threadIdx)
#define blockIdx (cusim::t_blockIdx)
#define blockDim (cusim::t_blockDim)
#define gridDim (cusim::t_gridDim)

static inline void __syncthreads() { cusim::t_block->bar.arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cusim::t_warp->bar.arrive_and_wait(); }

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return cusim::shfl_from(v, src); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32)
{
    int src = cusim::t_lane - (int)d;
    return cusim::shfl_from(v, src < 0 ? cusim::t_lane : src);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32)
{
    int src = cusim::t_lane + (int)d;
    return cusim::shfl_from(v, src > 31 ? cusim::t_lane : src);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return cusim::shfl_from(v, cusim::t_lane ^ m); }
static inline unsigned __ballot_sync(unsigned, int pred)
{
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (cusim::shfl_from(pred ? 1u : 0u, l) << l);
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

// ------------------------------------------------------------ math intrinsics
static inline float cusim_fma_mode(float a, float b, float c, int mode)
{
    volatile float va = a, vb = b, vc = c;
    std::fesetround(mode);
    volatile float r = std::fmaf(va, vb, vc);
    std::fesetround(FE_TONEAREST);
    return r;
}
static inline float cusim_add_mode(float a, float b, int mode)
{
    volatile float va = a, vb = b;
    std::fesetround(mode);
    volatile float r = va + vb;
    std::fesetround(FE_TONEAREST);
    return r;
}
static inline float __fmaf_rd(float a, float b, float c) { return cusim_fma_mode(a, b, c, FE_DOWNWARD); }
static inline float __fmaf_ru(float a, float b, float c) { return cusim_fma_mode(a, b, c, FE_UPWARD); }
static inline float __fmaf_rn(float a, float b, float c) { return cusim_fma_mode(a, b, c, FE_TONEAREST); }
static inline float __fadd_rd(float a, float b) { return cusim_add_mode(a, b, FE_DOWNWARD); }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned __float_as_uint(float f) { return cusim::to_bits(f) & 0xffffffffu; }
static inline int __float_as_int(float f) { return (int)__float_as_uint(f); }
static inline float __uint_as_float(unsigned u) { return cusim::from_bits<float>(u); }
static inline float __int_as_float(int u) { return cusim::from_bits<float>((unsigned)u); }
static inline float __int2float_rn(int v) { return (float)v; }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    uint64_t src = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)((src >> (8 * (sel & 7))) & 0xff);
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicMax(int *p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline float fminf_(float a, float b) { return std::fmin(a, b); }
using std::max;
using std::min;

// ----------------------------------------------------------- runtime subset
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; int major, minor; int l2CacheSize; char name[64]; };
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    std::memset(p, 0, sizeof(*p));
    p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024; p->major = 10; p->minor = 0;
    p->l2CacheSize = 126 << 20; std::snprintf(p->name, sizeof(p->name), "cusim host emulation");
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t = nullptr)
{
    for (size_t r = 0; r < height; ++r) std::memmove((char *)d + r * dpitch, (const char *)s + r * spitch, width);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cusim error"; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
