// ubench_duo.cu - stage_chunk / duo_chunk of k1_window_iir.cuh in isolation (not product code):
// cycles per 64-sample chunk for one stage per warp and for two stages per warp chained in registers,
// with 1..3 warps of the CTA active, with and without the per-chunk __syncthreads().
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fpga_real_time_fft_analyzer_b200/csrc/fra_common.cuh"
#include "../../fpga_real_time_fft_analyzer_b200/csrc/k1_window_iir.cuh"
using namespace fra;

// MODE 0: stage_chunk (one chain), 1: duo_chunk (two chains), 2: two stage_chunk calls back to back
template <int MODE, bool SYNC>
__global__ void __launch_bounds__(128, 1) loop(float *out, StageCoef k, int chunks, long long *cycles, int active_warps,
                                               int16_t *gdst, int n)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *smem = reinterpret_cast<float *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kStageTilesBytes / 4; i += blockDim.x) smem[i] = (float)((i * 37) % 2001 - 1000);
    StageState sa = {0.f, 0.f, 0.f, 0.f}, sb = sa;
    float ua = kBias16, ub = kBias16;
    __syncthreads();
    const bool active = warp >= 1 && warp <= active_warps;
    const int s = 2 * (warp > 0 ? warp - 1 : 0);
    long long t0 = clock64();
    for (int c = 0; c < chunks; ++c) {
        if (active) {
            const float4 *tina = stage_tile(smem, s, c & 1) + lane;
            float4 *touta = stage_tile(smem, s + 1, c & 1) + lane;
            const float4 *tinb = stage_tile(smem, s + 1, (c + 1) & 1) + lane;
            float4 *toutb = stage_tile(smem, (s + 2) % 6, (c + 1) & 1) + lane;
            if (MODE == 0) stage_chunk<false, true>(tina, touta, nullptr, k, sa, true);
            if (MODE == 1) duo_chunk<true, false>(tina, touta, k, k, sa, sb, ua, ub);
            if (MODE == 3) duo_chunk<true, true>(tina, touta, k, k, sa, sb, ua, ub);

            if (MODE == 2) {
                stage_chunk<false, true>(tina, touta, nullptr, k, sa, true);
                stage_chunk<false, true>(tinb, toutb, nullptr, k, sb, true);
            }
        }
        if (SYNC) __syncthreads();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sa.y1 + sb.y1;
    if (threadIdx.x == 32) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, bool SYNC>
void run(const char *name)
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 128 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
    int16_t *gdst; const int n = 16384; cudaMalloc(&gdst, (size_t)148 * 32 * n * 2);
    StageCoef k = {14 / 128.f, 0.0f, -14 / 128.f, -107 / 128.f, -21 / 128.f, 0x4B000000u, kMagicB + 21 * 65792.0f};
    const int chunks = 2048;
    cudaFuncSetAttribute(loop<MODE, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageSmemBytes);
    for (int aw = 1; aw <= 3; aw += 2) {
        for (int rep = 0; rep < 2; ++rep) loop<MODE, SYNC><<<148, 128, kStageSmemBytes>>>(out, k, chunks, cyc, aw, gdst, n);
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-44s sync=%d active warps=%d: %8.1f cycles per chunk = %.2f per sample\n", name, (int)SYNC, aw,
               (double)h / chunks, (double)h / chunks / 64);
    }
    cudaFree(out); cudaFree(cyc); cudaFree(gdst);
}

int main()
{
    run<0, false>("stage_chunk (one chain per warp)");
    run<0, true>("stage_chunk (one chain per warp)");
    run<1, false>("duo_chunk (two chained stages per warp)");
    run<1, true>("duo_chunk (two chained stages per warp)");
    run<2, false>("two stage_chunk calls back to back");
    run<3, false>("duo_chunk, two-instruction recurrence");
    run<3, true>("duo_chunk, two-instruction recurrence");
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
