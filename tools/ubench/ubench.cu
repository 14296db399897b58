// ubench.cu - tiny issue-rate / latency probes on the B200 SM (not product code).
// Each probe runs on ONE SM with W warps per scheduler and reports cycles per
// warp-instruction, so DESIGN.md's instruction-bound ceilings rest on measured rates.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

template <int KIND>
__global__ void probe(float *out, float a, float b, unsigned sel, long long *cycles)
{
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 0.001f + i;
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = threadIdx.x + i;
    float dep = threadIdx.x * 0.5f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        if (KIND == 0) {          // 8 independent FFMA, 3 register operands (a, b live in registers)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rd(acc[(i + 1) & 7], a, acc[i]);
        } else if (KIND == 1) {   // 8 independent FFMA, register * register + immediate-like self (2 distinct regs)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rd(acc[i], a, acc[i]);
        } else if (KIND == 2) {   // 8 independent FADD
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = acc[i] + a;
        } else if (KIND == 3) {   // 4 FFMA + 4 PRMT interleaved (two pipes)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i] = __fmaf_rd(acc[i], a, acc[i]);
                u[i] = __byte_perm(u[i], sel, 0x7610);
            }
        } else if (KIND == 4) {   // dependent FFMA -> FADD chain (recurrence latency)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dep = __fmaf_ru(dep, a, b);
                dep = dep - 12615680.0f;
            }
        } else if (KIND == 5) {   // dependent FFMA -> PRMT -> FADD chain
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dep = __fmaf_ru(dep, a, b);
                dep = __uint_as_float(__byte_perm(__float_as_uint(dep), sel, 0x7610)) - 8421376.0f;
            }
        } else if (KIND == 6) {   // dependent SHFL chain (shuffle latency)
#pragma unroll
            for (int i = 0; i < 8; ++i) dep = __shfl_up_sync(0xffffffffu, dep, 1);
        } else if (KIND == 7) {   // 8 independent PRMT
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = __byte_perm(u[i], sel, 0x7610) + 1;
        }
    }
    long long t1 = clock64();
    float s = dep;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i] + __uint_as_float(u[i]);
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
    for (int warps_per_sched = 1; warps_per_sched <= 4; warps_per_sched *= 2) {
        int threads = 128 * warps_per_sched;
        probe<KIND><<<1, threads>>>(out, 1.0001f, 0.5f, 0x4B000000u, cyc);
        probe<KIND><<<1, threads>>>(out, 1.0001f, 0.5f, 0x4B000000u, cyc);
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-44s warps/sched=%d  cycles per warp-instr (per scheduler) = %.2f   per-warp = %.2f\n", name,
               warps_per_sched, (double)h / ((double)ITER * per_iter * warps_per_sched),
               (double)h / ((double)ITER * per_iter));
    }
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    run<0>("FFMA 3 distinct regs, independent", 8);
    run<1>("FFMA acc*a+acc (2 distinct regs), independent", 8);
    run<2>("FADD independent", 8);
    run<3>("FFMA + PRMT interleaved", 8);
    run<7>("PRMT+IADD independent", 16);
    run<4>("FFMA.RP -> FADD dependent chain (per pair)", 4);
    run<5>("FFMA.RP -> PRMT -> FADD dependent (per triple)", 4);
    run<6>("SHFL.UP dependent chain", 8);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
