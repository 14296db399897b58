// ubench_f32x2.cu - issue rate of Blackwell's packed fp32x2 instructions (FFMA2 / FADD2)
// next to scalar FFMA, alone and mixed with ALU work (not product code).
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

template <int KIND>
__global__ void probe(float *out, float a, float b, unsigned sel, long long *cycles)
{
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 ka = make_float2(a, a * 0.5f), kb = make_float2(b, -b);
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = threadIdx.x + i;
    float2 dep = make_float2(threadIdx.x * 0.5f, 1.0f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        if (KIND == 0) {          // 8 independent FFMA2.RM
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rd(acc[i], ka, kb);
        } else if (KIND == 1) {   // 16 independent scalar FFMA.RM (same flops as KIND 0)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i].x = __fmaf_rd(acc[i].x, ka.x, kb.x);
                acc[i].y = __fmaf_rd(acc[i].y, ka.y, kb.y);
            }
        } else if (KIND == 2) {   // 8 independent FADD2
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fadd2_rn(acc[i], ka);
        } else if (KIND == 3) {   // 4 FFMA2 + 4 PRMT interleaved
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i] = __ffma2_rd(acc[i], ka, kb);
                u[i] = __byte_perm(u[i], sel, 0x7610);
            }
        } else if (KIND == 4) {   // dependent FFMA2.RP -> FADD2 chain
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dep = __ffma2_ru(dep, ka, kb);
                dep = __fadd2_rn(dep, make_float2(-12615680.0f, -12615680.0f));
            }
        } else if (KIND == 5) {   // 4 FFMA2 + 8 PRMT (the biquad's mix: 5 FMA2 per 2x(PRMT) + FADD2)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i] = __ffma2_rd(acc[i], ka, kb);
                u[2 * i] = __byte_perm(u[2 * i], sel, 0x7610);
                u[2 * i + 1] = __byte_perm(u[2 * i + 1], sel, 0x7610);
            }
        }
    }
    long long t1 = clock64();
    float s = dep.x + dep.y;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char *name, int per_iter)
{
    float *out;
    long long *cyc, h;
    cudaMalloc(&out, 1024 * sizeof(float));
    cudaMalloc(&cyc, sizeof(long long));
    for (int wps = 1; wps <= 4; wps *= 2) {
        probe<KIND><<<1, 128 * wps>>>(out, 1.0001f, 0.5f, 0x4B000000u, cyc);
        probe<KIND><<<1, 128 * wps>>>(out, 1.0001f, 0.5f, 0x4B000000u, cyc);
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-40s warps/sched=%d  cycles per warp-instr per scheduler = %.2f\n", name, wps,
               (double)h / ((double)ITER * per_iter * wps));
    }
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    run<0>("FFMA2.RM independent", 8);
    run<1>("FFMA.RM independent (2x count)", 16);
    run<2>("FADD2 independent", 8);
    run<3>("FFMA2 + PRMT interleaved", 8);
    run<5>("FFMA2 + 2 PRMT interleaved", 12);
    run<4>("FFMA2.RP -> FADD2 dependent (per pair)", 4);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
