// bench_dmma.cu — evidence for DESIGN.md section 4 "tensor cores are not used" (development tool, not product).
//
// The north star allows tensor cores "only if an ncu capture shows a dense-contraction formulation beats the
// FFT path".  The only tensor-core path with the precision the B = 2^23 blind rotation needs is FP64 DMMA
// (mma.sync.m8n8k4.f64).  This tool times, per SM:
//   1. the DFMA peak (the roofline denominator bench.py also probes),
//   2. the DMMA m8n8k4 peak,
//   3. both interleaved in one warp (do the two share a pipe?),
//   4. ONE radix-8 pass of the 512-point transform (no twiddles) over 8-point complex vectors
//        a. as the scalar dft8<false>() butterfly network of fft512.cuh (what the kernels run),
//        b. as a dense contraction: the 16 x 16 real DFT-8 matrix times a 16 x 8 panel = 8 DMMA per 64 points.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o bench_dmma bench_dmma.cu && ./bench_dmma
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../fft512.cuh"
using namespace cbs;

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

constexpr int kWarps = 8;

__global__ void __launch_bounds__(32 * kWarps, 1) k_dfma(int iters, double *out, long long *cyc)
{
    double g[16];
#pragma unroll
    for (int k = 0; k < 16; k++) g[k] = 1.0 + k + threadIdx.x;
    const double a = 0.999999, b = 1e-9;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int k = 0; k < 16; k++) g[k] = fma(g[k], a, b);
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s += g[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

// MIX = 0: 16 DMMA per iteration; MIX = 1: 16 DMMA + 64 DFMA per iteration
template <int MIX>
__global__ void __launch_bounds__(32 * kWarps, 1) k_dmma(int iters, double *out, long long *cyc)
{
    double d[8][2], g[16];
#pragma unroll
    for (int k = 0; k < 8; k++) d[k][0] = d[k][1] = 0.0;
#pragma unroll
    for (int k = 0; k < 16; k++) g[k] = 1.0 + k + threadIdx.x;
    const double a = 1e-3 * (threadIdx.x & 31), b = 1e-3;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int k = 0; k < 8; k++) dmma(d[k][0], d[k][1], a, b, d[k][0], d[k][1]);
        if (MIX) {
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int k = 0; k < 16; k++) g[k] = fma(g[k], 0.999999, 1e-9);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += d[k][0] + d[k][1];
#pragma unroll
    for (int k = 0; k < 16; k++) s += g[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

// radix-8 pass, scalar butterflies: every thread transforms one 8-point vector per iteration (256 points per warp)
__global__ void __launch_bounds__(32 * kWarps, 1) k_pass_scalar(int iters, double *out, long long *cyc)
{
    cplx v[8];
#pragma unroll
    for (int m = 0; m < 8; m++) v[m] = cplx{1.0 + threadIdx.x + m, 0.5 * m};
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        dft8<false>(v);
#pragma unroll
        for (int m = 0; m < 8; m++) v[m].x *= 0.35355339059327373;  // keep the values bounded (1 DMUL per real: counted)
#pragma unroll
        for (int m = 0; m < 8; m++) v[m].y *= 0.35355339059327373;
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int m = 0; m < 8; m++) s += v[m].x + v[m].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

// radix-8 pass as a dense contraction: Y(16 x 8) = W(16 x 16) X(16 x 8), W = [[C, S], [-S, C]] / sqrt(8) of the DFT-8,
// X = 8 complex vectors (columns) as [Re; Im].  2 row tiles x 4 k-steps = 8 DMMA per 64 complex points.
// A fragment (m8 x k4, row): lane holds A[lane / 4][lane % 4]; B fragment (k4 x n8, col): lane holds B[lane % 4][lane / 4];
// C/D: lane holds D[lane / 4][2 * (lane % 4) + {0, 1}].  The output panel is fed back as the next input without the
// layout exchange a real transform would need (2 more shuffles per value), which favours the DMMA arm.
__global__ void __launch_bounds__(32 * kWarps, 1) k_pass_dmma(int iters, double *out, long long *cyc)
{
    const int lane = threadIdx.x & 31, row = lane >> 2, col = lane & 3;
    double A[2][4];  // [row tile][k step]
    for (int mt = 0; mt < 2; mt++)
        for (int ks = 0; ks < 4; ks++) {
            const int r = mt * 8 + row, c = ks * 4 + col;  // element W[r][c]
            const int k = r & 7, j = c & 7;
            const double ang = -2.0 * 3.14159265358979323846 * k * j / 8.0, s8 = 0.35355339059327373;
            const double cr = cos(ang) * s8, ci = sin(ang) * s8;
            // [Re y; Im y] = [[cr, -ci], [ci, cr]] [Re x; Im x]
            A[mt][ks] = (r < 8) ? ((c < 8) ? cr : -ci) : ((c < 8) ? ci : cr);
        }
    double X[4];
#pragma unroll
    for (int ks = 0; ks < 4; ks++) X[ks] = 1.0 + lane + ks;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        double D[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int ks = 0; ks < 4; ks++)
#pragma unroll
            for (int mt = 0; mt < 2; mt++) dmma(D[mt][0], D[mt][1], A[mt][ks], X[ks], D[mt][0], D[mt][1]);
        X[0] = D[0][0];
        X[1] = D[0][1];
        X[2] = D[1][0];
        X[3] = D[1][1];
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = X[0] + X[1] + X[2] + X[3];
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, 148 * 32 * kWarps * 8);
    cudaMalloc(&d_cyc, 8);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    auto report = [&](const char *name, double flop_per_iter_per_warp, double points_per_iter_per_warp, float ms, long long cyc) {
        const double tf = flop_per_iter_per_warp * kWarps * 148.0 * iters / (ms * 1e-3) / 1e12;
        printf("%-44s %8.2f cyc/iter/warp-set  %7.2f TFLOP/s chip", name, (double)cyc / iters, tf);
        if (points_per_iter_per_warp > 0)
            printf("  %8.3f cyc per 64 complex points per SM", (double)cyc / iters / (points_per_iter_per_warp * kWarps / 64.0));
        printf("  (%.3f ms)\n", ms);
    };
    auto timed = [&](auto launch, float &ms, long long &cyc) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        launch(10);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        launch(iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    };
    float ms;
    long long cyc;
    timed([&](int n) { k_dfma<<<148, 32 * kWarps>>>(n, d_out, d_cyc); }, ms, cyc);
    report("1. DFMA peak (64 per iter per thread)", 64.0 * 32 * 2, 0, ms, cyc);
    timed([&](int n) { k_dmma<0><<<148, 32 * kWarps>>>(n, d_out, d_cyc); }, ms, cyc);
    report("2. DMMA m8n8k4 peak (16 per iter per warp)", 16.0 * 512, 0, ms, cyc);
    timed([&](int n) { k_dmma<1><<<148, 32 * kWarps>>>(n, d_out, d_cyc); }, ms, cyc);
    report("3. 16 DMMA + 64 DFMA interleaved", 16.0 * 512 + 64.0 * 32 * 2, 0, ms, cyc);
    timed([&](int n) { k_pass_scalar<<<148, 32 * kWarps>>>(n, d_out, d_cyc); }, ms, cyc);
    report("4a. radix-8 pass, scalar dft8 (256 pts/warp)", 0, 256, ms, cyc);
    timed([&](int n) { k_pass_dmma<<<148, 32 * kWarps>>>(n, d_out, d_cyc); }, ms, cyc);
    report("4b. radix-8 pass, 8 DMMA (64 pts/warp)", 0, 64, ms, cyc);
    if (cudaGetLastError() != cudaSuccess) return 1;
    printf("SM clock attribute %d kHz\n", clk_khz);
    return 0;
}
