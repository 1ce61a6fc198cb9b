// bench_fp64ops.cu — issue rate of DADD, DMUL, DFMA and of their mixes on B200 (development tool).
// 16 independent chains per thread, 2 / 3 / 4 warps per scheduler.  Reported: cycles per warp instruction and scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_fp64ops bench_fp64ops.cu && ./bench_fp64ops
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(512, 1) k_ops(double *out, int iters, double a, double b)
{
    double f[16];
#pragma unroll
    for (int j = 0; j < 16; j++) f[j] = threadIdx.x + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (KIND == 0) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[j]) : "d"(a));
            if (KIND == 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(f[j]) : "d"(a));
            if (KIND == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[j]) : "d"(a), "d"(b));
            if (KIND == 3) {  // the butterfly mix: 2 DADD : 1 DFMA
                if (j % 3 == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[j]) : "d"(a), "d"(b));
                else asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[j]) : "d"(a));
            }
            if (KIND == 4) {  // dependent pairs (a +- b): two chains share inputs
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[j]) : "d"(f[(j + 1) & 15]));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s += f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char *name, double *d, int threads, double ghz)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_ops<KIND><<<148, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k_ops<KIND><<<148, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * ghz * 1e9;
    const double per_sched = (double)iters * 16 * (threads / 32) / 4;
    printf("%-28s warps/scheduler=%d: %.3f ms, %.2f cycles per warp instruction per scheduler\n", name, threads / 128, ms, cycles / per_sched);
}

int main()
{
    double *d;
    cudaMalloc(&d, 148 * 512 * 8);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    for (int threads : {128, 256, 384, 512}) {
        run<0>("DADD", d, threads, ghz);
        run<1>("DMUL", d, threads, ghz);
        run<2>("DFMA", d, threads, ghz);
        run<3>("2 DADD : 1 DFMA", d, threads, ghz);
        run<4>("DADD, register operands", d, threads, ghz);
    }
    return 0;
}
