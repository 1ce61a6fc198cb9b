// bench_issue.cu — does an FP64 instruction take one or two issue slots of its scheduler on B200?
// (development tool.)  Every warp runs a loop of 8 independent DFMAs, each followed by K independent integer
// instructions (K = 0..4; LOP3/IADD3 on the ALU pipe or IMAD on the FMA pipe).  The FP64 pipe accepts one warp
// instruction every 2 cycles per scheduler (16 lanes): if the DFMA occupied ONE issue slot, K = 1 would be free
// (2 cycles per DFMA either way); if the integer instructions cannot use the second cycle, the time grows with K
// from K = 1 on.  Reported: cycles per (DFMA + K integer) group and scheduler, at 2 and 4 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_issue bench_issue.cu && ./bench_issue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int K, int KIND>
__global__ void __launch_bounds__(512, 1) k_issue(double *out, uint32_t *iout, int iters, double a, double b, uint32_t c)
{
    double f[8];
    uint32_t x[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        f[j] = threadIdx.x + j;
#pragma unroll
        for (int k = 0; k < 4; k++) x[j][k] = threadIdx.x * 8 + j + k;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[j]) : "d"(a), "d"(b));
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j][k]) : "r"(c), "r"(x[j][(k + 1) & 3]));
                else asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[j][k]) : "r"(c));
            }
        }
    }
    double s = 0;
    uint32_t u = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        s += f[j];
#pragma unroll
        for (int k = 0; k < 4; k++) u ^= x[j][k];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = u;
}

template <int K, int KIND>
void run(double *d, uint32_t *di, int threads, double ghz)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_issue<K, KIND><<<148, threads>>>(d, di, 100, 1.0000001, 1e-9, 12345u);
    cudaEventRecord(e0);
    k_issue<K, KIND><<<148, threads>>>(d, di, iters, 1.0000001, 1e-9, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * ghz * 1e9;
    const double groups_per_sched = (double)iters * 8 * (threads / 32) / 4;
    printf("%s K=%d warps/scheduler=%d: %.3f ms, %.2f cycles per (DFMA + %d int) per scheduler\n", KIND ? "IMAD" : "LOP3", K, threads / 128, ms,
           cycles / groups_per_sched, K);
}

int main()
{
    double *d;
    uint32_t *di;
    cudaMalloc(&d, 148 * 512 * 8);
    cudaMalloc(&di, 148 * 512 * 4);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("clock %.3f GHz (nominal max; boost clocks may differ)\n", ghz);
    for (int threads : {256, 512}) {
        run<0, 0>(d, di, threads, ghz);
        run<1, 0>(d, di, threads, ghz);
        run<2, 0>(d, di, threads, ghz);
        run<3, 0>(d, di, threads, ghz);
        run<4, 0>(d, di, threads, ghz);
        run<1, 1>(d, di, threads, ghz);
        run<2, 1>(d, di, threads, ghz);
        run<4, 1>(d, di, threads, ghz);
    }
    return 0;
}
