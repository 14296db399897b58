// ubench_split.cu - the inner loop of k1_split in isolation (not product code):
// how many cycles does ONE warp need per systolic iteration, and which ingredient costs what?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fpga_real_time_fft_analyzer_b200/csrc/fra_common.cuh"
using namespace fra;

// VAR bit0: SHFL, bit1: FSEL, bit2: speculative (FADD) instead of PRMT+FADD wrap
template <int VAR, int UNROLL>
__global__ void loop(float *out, StageCoef k, int blocks, long long *cycles, int first_mask)
{
    StageState st = {0.f, 0.f, 0.f, 0.f};
    float y = threadIdx.x * 0.25f, up0 = 1.f, up1 = 2.f, up2 = 3.f, absmax = 0.f;
    const bool first = (first_mask >> (threadIdx.x & 31)) & 1;
    float w = 100.0f + threadIdx.x;
    unsigned sink = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int b = 0; b < blocks; ++b) {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            float up_new = up2;
            if (VAR & 1) up_new = __shfl_up_sync(0xffffffffu, y, 1);
            float x = up0;
            if (VAR & 2) x = first ? w : up0;
            up0 = up1; up1 = up2; up2 = up_new;
            float acc;
            if (VAR & 4) acc = biquad_step_spec(x, k, st, &y, &absmax);
            else acc = biquad_step(x, k, st, &y);
            sink ^= __float_as_uint(acc);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = y + absmax + __uint_as_float(sink);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int VAR, int UNROLL>
void run(const char *name)
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, 4096 * sizeof(float)); cudaMalloc(&cyc, 8 * sizeof(long long));
    StageCoef k = {14 / 128.f, 0.0f, -14 / 128.f, -107 / 128.f, -21 / 128.f, 0x4B000000u};
    const int iters = 1 << 16;
    for (int wps = 1; wps <= 2; ++wps) {
        loop<VAR, UNROLL><<<1, 128 * wps>>>(out, k, iters / UNROLL, cyc, 0x01041041);
        loop<VAR, UNROLL><<<1, 128 * wps>>>(out, k, iters / UNROLL, cyc, 0x01041041);
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-52s unroll %2d warps/sched=%d: %.2f cycles per iteration per warp\n", name, UNROLL, wps, (double)h / iters);
    }
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0, 16>("biquad only (PRMT wrap)");
    run<1, 16>("+ SHFL");
    run<3, 16>("+ SHFL + FSEL (the k1_split iteration)");
    run<3, 32>("+ SHFL + FSEL");
    run<3, 64>("+ SHFL + FSEL");
    run<4, 16>("biquad only, speculative (FADD)");
    run<7, 16>("speculative + SHFL + FSEL");
    run<7, 64>("speculative + SHFL + FSEL");
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
