// pipe_microbench.cu -- measures per-SM issue rates of the integer instructions the DP uses (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_microbench tools/pipe_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
#define CH 8
template <int OP> __global__ void k(unsigned *out, unsigned a0, unsigned b0)
{
    unsigned v[CH], w[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) { v[c] = a0 + threadIdx.x * (c + 1); w[c] = b0 ^ (threadIdx.x << c); }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (OP == 0) v[c] = __vimax3_s32(v[c], w[c], b0);
            if (OP == 1) v[c] = __vimax3_s16x2(v[c], w[c], b0);
            if (OP == 2) v[c] = __viaddmax_s32(v[c], a0, w[c]);
            if (OP == 3) v[c] = __viaddmax_s16x2(v[c], a0, w[c]);
            if (OP == 4) v[c] = (v[c] & 0x00030003u) | w[c];               // LOP3
            if (OP == 5) v[c] = v[c] * 4u + w[c];                          // IMAD
            if (OP == 6) v[c] = __vadd2(v[c], w[c]);                       // VIADD.16x2
            if (OP == 7) v[c] = __funnelshift_r(v[c], w[c], 2);            // SHF
            if (OP == 8) { v[c] = __vimax3_s16x2(v[c], w[c], b0); w[c] = w[c] * 4u + v[c]; }     // ALU + FMA pair
            if (OP == 9) v[c] = __vmaxs2(v[c], w[c]);
            if (OP == 10) v[c] = v[c] + w[c] + a0;                         // IADD3
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s ^= v[c] ^ w[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char *name, int nsm)
{
    unsigned *d; cudaMalloc(&d, nsm * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<nsm * 2, 1024>>>(d, 12345u, 777u);
    cudaEventRecord(e0);
    k<OP><<<nsm * 2, 1024>>>(d, 12345u, 777u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double ops = (double)nsm * 2 * 1024 * ITER * CH * (OP == 8 ? 2 : 1);
    printf("%-22s %8.3f ms  %7.1f thread-ops/clk/SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3) / (clk * 1e3) / nsm, clk / 1000);
    cudaFree(d);
}
int main()
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    run<0>("VIMNMX3.s32", nsm); run<1>("VIMNMX3.s16x2", nsm); run<2>("VIADDMNMX.s32", nsm); run<3>("VIADDMNMX.s16x2", nsm);
    run<4>("LOP3", nsm); run<5>("IMAD", nsm); run<6>("VIADD.16x2", nsm); run<7>("SHF funnel", nsm);
    run<8>("VIMNMX3.16x2+IMAD", nsm); run<9>("VIMNMX.s16x2", nsm); run<10>("IADD3", nsm);
    return 0;
}
