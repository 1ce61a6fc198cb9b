// bench_fft.cu — microbenchmark of the fft512.cuh group transform (development tool, not product).
// Measures forward+inverse FFT pairs per microsecond for several groups-per-CTA settings, to find the
// ceiling of the 64-thread radix-8 design independent of the blind-rotation bookkeeping.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../fft512.cuh"
#include "../fft_tables.h"
using namespace cbs;

__device__ __forceinline__ void gsync(int bar) { asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory"); }

template <int G, bool TW_SMEM>
__global__ void __launch_bounds__(64 * G, 1) k_fft_loop(const double *twtab, double *out, int iters)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int gi = threadIdx.x >> 6, t = threadIdx.x & 63;
    cplx *scr0 = reinterpret_cast<cplx *>(smem) + gi * 1024;
    cplx *scr1 = scr0 + 512;
    Twiddles tw;
    load_twiddles(tw, twtab, t);
    cplx v[8];
#pragma unroll
    for (int m = 0; m < 8; m++) v[m] = cplx{(double)(t + m), (double)(t - m)};
    const int bar = 1 + gi;
    for (int it = 0; it < iters; it++) {
        fwd_p1(v, scr0, tw, t);
        gsync(bar);
        fwd_p2(v, scr0, tw, t);
        gsync(bar);
        fwd_p3(v, scr0, t);
#pragma unroll
        for (int m = 0; m < 8; m++) v[m].x *= 1.0 / 512.0, v[m].y *= 1.0 / 512.0;
        inv_p3(v, scr1, t);
        gsync(bar);
        inv_p2(v, scr1, tw, t);
        gsync(bar);
        inv_p1(v, scr1, tw, t);
    }
    double s = 0;
#pragma unroll
    for (int m = 0; m < 8; m++) s += v[m].x + v[m].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shuffle-exchange variant: passes 2<->3 exchanged with width-8 shuffles (2 barriers per transform instead of 4)
template <int G>
__global__ void __launch_bounds__(64 * G, 1) k_fft_loop_x(const double *twtab, double *out, int iters)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int gi = threadIdx.x >> 6, t = threadIdx.x & 63;
    cplx *scr0 = reinterpret_cast<cplx *>(smem) + gi * 1024;
    cplx *scr1 = scr0 + 512;
    Twiddles tw;
    load_twiddles_x(tw, twtab, t);
    cplx v[8];
#pragma unroll
    for (int m = 0; m < 8; m++) v[m] = cplx{(double)(t + m), (double)(t - m)};
    const int bar = 1 + gi;
    for (int it = 0; it < iters; it++) {
        fwd_p1(v, scr0, tw, t);
        gsync(bar);
        fwd_p2x(v, scr0, tw, t);
        exchange8<-1>(v, t & 7);
        fwd_p3x(v);
#pragma unroll
        for (int m = 0; m < 8; m++) v[m].x *= 1.0 / 512.0, v[m].y *= 1.0 / 512.0;
        inv_p3x(v);
        exchange8<1>(v, t & 7);
        inv_p2x(v, scr1, tw, t);
        gsync(bar);
        inv_p1(v, scr1, tw, t);
    }
    double s = 0;
#pragma unroll
    for (int m = 0; m < 8; m++) s += v[m].x + v[m].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int G>
void run_x(const double *tw, double *out, int iters, size_t extra_smem)
{
    size_t smem = (size_t)G * 16384 + extra_smem;
    cudaFuncSetAttribute(k_fft_loop_x<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_fft_loop_x<G><<<148, 64 * G, smem>>>(tw, out, 10);
    cudaEventRecord(e0);
    k_fft_loop_x<G><<<148, 64 * G, smem>>>(tw, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ffts = 2.0 * iters * G * 148;
    cudaError_t err = cudaGetLastError();
    double h[4];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("[shuffle] groups/CTA=%d extra_smem=%zuKB: %.3f ms, %.1f FFT/us (%s) check=%.6g\n", G, extra_smem / 1024, ms,
           ffts / (ms * 1e3), cudaGetErrorString(err), h[1]);
}

template <int G>
void run(const double *tw, double *out, int iters, size_t extra_smem)
{
    size_t smem = (size_t)G * 16384 + extra_smem;
    cudaFuncSetAttribute(k_fft_loop<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_fft_loop<G, false><<<148, 64 * G, smem>>>(tw, out, 10);
    cudaEventRecord(e0);
    k_fft_loop<G, false><<<148, 64 * G, smem>>>(tw, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ffts = 2.0 * iters * G * 148;
    cudaError_t err = cudaGetLastError();
    double h[4];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("check=%.6g ", h[1]);
    printf("groups/CTA=%d extra_smem=%zuKB: %.3f ms, %.1f FFT/us, FP64 lane-ops/s=%.2f T (%s)\n", G, extra_smem / 1024, ms,
           ffts / (ms * 1e3), ffts * 16384.0 / (ms * 1e-3) * 1e-12, cudaGetErrorString(err));
}

int main()
{
    std::vector<double> tab = make_twiddle_table();
    double *d_tw, *d_out;
    cudaMalloc(&d_tw, tab.size() * 8);
    cudaMalloc(&d_out, 148 * 1024 * 8);
    cudaMemcpy(d_tw, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice);
    const int iters = 3000;
    run<2>(d_tw, d_out, iters, 0);
    run<4>(d_tw, d_out, iters, 0);
    run<4>(d_tw, d_out, iters, 140 * 1024);  // same occupancy as the blind rotation (1 CTA/SM)
    run_x<2>(d_tw, d_out, iters, 0);
    run_x<4>(d_tw, d_out, iters, 0);
    run_x<4>(d_tw, d_out, iters, 140 * 1024);
    run_x<8>(d_tw, d_out, iters, 0);
    run<6>(d_tw, d_out, iters, 0);
    run<8>(d_tw, d_out, iters, 0);
    run<12>(d_tw, d_out, iters, 0);
    run<16>(d_tw, d_out, iters, 0);
    return 0;
}
