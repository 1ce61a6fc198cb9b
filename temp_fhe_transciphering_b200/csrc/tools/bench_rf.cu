// bench_rf.cu — is the issue rate of the blind rotation's instruction mix bounded by REGISTER-FILE READ BANDWIDTH?
// (development tool.)  ncu shows every blind-rotation variant at ~0.55 instructions per cycle and scheduler whatever the
// occupancy (2 or 3 warps per scheduler), and adding instructions of any kind anywhere costs time in proportion.  Here:
// instructions whose source operands are all DISTINCT, CHANGING registers (no reuse cache, no immediates), 16-24
// independent chains, 2 and 4 warps per scheduler.  Reported: cycles per warp instruction and scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_rf bench_rf.cu && ./bench_rf
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(512, 1) k_rf(double *out, uint32_t *iout, int iters)
{
    double f[24];
    uint32_t x[24];
#pragma unroll
    for (int j = 0; j < 24; j++) f[j] = 1.0 + 1e-9 * (threadIdx.x + j), x[j] = threadIdx.x * 24 + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 24; j++) {
            if (KIND == 0) asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(f[j]) : "d"(f[(j + 5) % 24]), "d"(f[(j + 11) % 24]));
            if (KIND == 1) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[j]) : "d"(f[(j + 5) % 24]), "d"(f[(j + 11) % 24]));
            if (KIND == 2) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[j]) : "d"(f[(j + 5) % 24]), "d"(f[(j / 2 * 2 + 11) % 24]));  // operand B shared by pairs
            if (KIND == 3) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[j]) : "r"(x[(j + 3) % 24]), "r"(x[(j + 7) % 24]), "r"(x[(j + 13) % 24]));
            if (KIND == 4) asm volatile("add.u32 %0, %1, %2;" : "=r"(x[j]) : "r"(x[(j + 3) % 24]), "r"(x[(j + 7) % 24]));
            if (KIND == 5) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(x[j]) : "r"(x[(j + 3) % 24]), "r"(x[(j + 7) % 24]), "r"(x[(j + 13) % 24]));
            if (KIND == 6) {  // DADD + LOP3 pairs, all operands distinct registers
                asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(f[j]) : "d"(f[(j + 5) % 24]), "d"(f[(j + 11) % 24]));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[j]) : "r"(x[(j + 3) % 24]), "r"(x[(j + 7) % 24]), "r"(x[(j + 13) % 24]));
            }
            if (KIND == 7) {  // DFMA + LOP3 pairs, all operands distinct registers
                asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[j]) : "d"(f[(j + 5) % 24]), "d"(f[(j + 11) % 24]));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[j]) : "r"(x[(j + 3) % 24]), "r"(x[(j + 7) % 24]), "r"(x[(j + 13) % 24]));
            }
        }
    }
    double s = 0;
    uint32_t u = 0;
#pragma unroll
    for (int j = 0; j < 24; j++) s += f[j], u ^= x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = u;
}

template <int KIND>
void run(const char *name, int per_iter, double *d, uint32_t *di, int threads, double ghz)
{
    const int iters = 10000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_rf<KIND><<<148, threads>>>(d, di, 100);
    cudaEventRecord(e0);
    k_rf<KIND><<<148, threads>>>(d, di, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * ghz * 1e9;
    const double per_sched = (double)iters * per_iter * (threads / 32) / 4;
    printf("%-44s warps/scheduler=%d: %.2f cycles per warp instruction per scheduler\n", name, threads / 128, cycles / per_sched);
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
    for (int threads : {256, 512}) {
        run<0>("DADD  d = a + b (3 distinct registers)", 24, d, di, threads, ghz);
        run<1>("DFMA  d = a * b + d (3 distinct sources)", 24, d, di, threads, ghz);
        run<2>("DFMA  operand B shared by neighbours", 24, d, di, threads, ghz);
        run<3>("LOP3  3 distinct sources", 24, d, di, threads, ghz);
        run<4>("IADD  2 distinct sources", 24, d, di, threads, ghz);
        run<5>("IMAD  3 distinct sources", 24, d, di, threads, ghz);
        run<6>("DADD + LOP3 pairs (per instruction)", 48, d, di, threads, ghz);
        run<7>("DFMA + LOP3 pairs (per instruction)", 48, d, di, threads, ghz);
    }
    return 0;
}
