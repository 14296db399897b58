// ubench_ffma.cu - does the FFMA issue rate on the B200 SM depend on where the operands live?
// (three vector registers / one uniform register or constant-bank operand / an immediate), with the
// directed rounding modes the bit-exact biquad uses (FFMA.RM / .RP).  One SM, W warps per scheduler,
// cycles per warp-instruction per scheduler.  Not product code.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048

template <int KIND>
__global__ void probe(float *out, const float *table, float pa, float pb, unsigned sel, long long *cycles)
{
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 0.001f + i;
    // lane-varying coefficients: cannot live in uniform registers
    const float va = table[threadIdx.x & 7], vb = table[8 + (threadIdx.x & 7)];
    unsigned u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        if (KIND == 0) {          // FFMA.RM  R, R, R   (coefficient in a vector register)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rd(acc[(i + 1) & 7], va, acc[i]);
        } else if (KIND == 1) {   // FFMA.RM  R, UR/c[], R  (coefficient = kernel parameter)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rd(acc[(i + 1) & 7], pa, acc[i]);
        } else if (KIND == 2) {   // FFMA.RM  R, imm, R
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rd(acc[(i + 1) & 7], 0.8359375f, acc[i]);
        } else if (KIND == 3) {   // FFMA (RN) R, R, R
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[(i + 1) & 7], va, acc[i]);
        } else if (KIND == 4) {   // FFMA (RN) R, UR, R
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[(i + 1) & 7], pa, acc[i]);
        } else if (KIND == 5) {   // FFMA.RM R, R, UR  (addend = parameter, multiplier in registers)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rd(acc[(i + 1) & 7], acc[i], pb);
        } else if (KIND == 6) {   // 8 FFMA.RM (R,UR,R) + 2 PRMT : the biased biquad's mix (4 : 1)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rd(acc[(i + 1) & 7], pa, acc[i]);
            u[0] = __byte_perm(u[0], sel, 0x7610);
            u[1] = __byte_perm(u[1], sel, 0x7610);
        } else if (KIND == 7) {   // 8 FFMA.RM (R,R,R) + 2 PRMT
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rd(acc[(i + 1) & 7], va, acc[i]);
            u[0] = __byte_perm(u[0], sel, 0x7610);
            u[1] = __byte_perm(u[1], sel, 0x7610);
        } else if (KIND == 8) {   // 8 independent PRMT
#pragma unroll
            for (int i = 0; i < 4; ++i) { u[i] = __byte_perm(u[i], sel, 0x7610); }
#pragma unroll
            for (int i = 0; i < 4; ++i) { u[i] = __byte_perm(u[i], sel, 0x3254); }
        } else if (KIND == 9) {   // FFMA.RM R,R,R with two DIFFERENT vector coefficients (bank spread)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rd(acc[(i + 1) & 7], (i & 1) ? va : vb, acc[i]);
        }
    }
    long long t1 = clock64();
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char *name, int per_iter)
{
    float *out, *table;
    long long *cyc, h;
    cudaMalloc(&out, 1024 * sizeof(float));
    cudaMalloc(&table, 16 * sizeof(float));
    float ht[16];
    for (int i = 0; i < 16; ++i) ht[i] = 0.5f + 0.01f * i;
    cudaMemcpy(table, ht, sizeof(ht), cudaMemcpyHostToDevice);
    cudaMalloc(&cyc, sizeof(long long));
    for (int warps_per_sched = 1; warps_per_sched <= 4; warps_per_sched *= 2) {
        int threads = 128 * warps_per_sched;
        probe<KIND><<<1, threads>>>(out, table, 0.7001f, 12615680.0f, 0x4B000000u, cyc);
        probe<KIND><<<1, threads>>>(out, table, 0.7001f, 12615680.0f, 0x4B000000u, cyc);
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-52s warps/sched=%d  cycles per warp-instr per scheduler = %.2f\n", name, warps_per_sched,
               (double)h / ((double)ITER * per_iter * warps_per_sched));
    }
    cudaFree(out); cudaFree(table); cudaFree(cyc);
}

int main()
{
    run<0>("FFMA.RM R,R,R (vector-register coefficient)", 8);
    run<9>("FFMA.RM R,R,R two vector coefficients", 8);
    run<1>("FFMA.RM R,param,R (uniform / constant operand)", 8);
    run<2>("FFMA.RM R,imm,R", 8);
    run<3>("FFMA.RN R,R,R", 8);
    run<4>("FFMA.RN R,param,R", 8);
    run<5>("FFMA.RM R,R,param (parameter as addend)", 8);
    run<6>("8 FFMA.RM R,param,R + 2 PRMT", 10);
    run<7>("8 FFMA.RM R,R,R + 2 PRMT", 10);
    run<8>("PRMT independent", 8);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
