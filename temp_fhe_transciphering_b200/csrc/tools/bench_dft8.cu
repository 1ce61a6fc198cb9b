// bench_dft8.cu — how fast does the register-only part of the transform (the radix-8 butterfly network of fft512.cuh plus
// a twiddle multiplication per value) issue on B200?  (development tool.)  No memory, no barriers: whatever stays below one
// FP64 warp instruction per 2 cycles and scheduler here is dependency latency / register-file bandwidth of the butterfly
// code itself.  Reported per configuration: cycles per FP64 warp instruction and scheduler (2.0 = the pipe's peak).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I.. -o bench_dft8 bench_dft8.cu && ./bench_dft8
#include <cstdio>
#include <cuda_runtime.h>
#include "../fft512.cuh"
using namespace cbs;

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_dft8(double *out, int iters, const double *tw)
{
    cplx v[8], w[8];
#pragma unroll
    for (int m = 0; m < 8; m++) {
        v[m] = cplx{(double)(threadIdx.x + m), (double)(threadIdx.x - m)};
        w[m] = cplx{tw[2 * m], tw[2 * m + 1]};
    }
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {  // butterflies only
            dft8<false>(v);
            dft8<true>(v);
        } else if (MODE == 1) {  // butterflies + twiddles (one pass of the transform)
            dft8<false>(v);
#pragma unroll
            for (int m = 0; m < 8; m++) v[m] = cmul(v[m], w[m]);
            dft8<true>(v);
#pragma unroll
            for (int m = 0; m < 8; m++) v[m] = cmul_conj(v[m], w[m]);
        } else {  // two independent polynomials interleaved (2 x the instruction-level parallelism)
            cplx u[8];
#pragma unroll
            for (int m = 0; m < 8; m++) u[m] = cplx{v[m].y, v[m].x};
            dft8<false>(v);
            dft8<false>(u);
#pragma unroll
            for (int m = 0; m < 8; m++) v[m] = cmul(v[m], w[m]);
#pragma unroll
            for (int m = 0; m < 8; m++) u[m] = cmul(u[m], w[m]);
#pragma unroll
            for (int m = 0; m < 8; m++) v[m] = cplx{v[m].x + u[m].y, v[m].y - u[m].x};
        }
    }
    double s = 0;
#pragma unroll
    for (int m = 0; m < 8; m++) s += v[m].x + v[m].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int fp64_per_iter, double *d, const double *tw, int threads, double ghz)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_dft8<MODE><<<148, threads>>>(d, 100, tw);
    cudaEventRecord(e0);
    k_dft8<MODE><<<148, threads>>>(d, iters, tw);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * ghz * 1e9;
    const double per_sched = (double)iters * fp64_per_iter * (threads / 32) / 4;
    printf("%-34s warps/scheduler=%d: %.3f ms, %.2f cycles per FP64 warp instruction per scheduler\n", name, threads / 128, ms, cycles / per_sched);
}

int main()
{
    double *d, *tw, h[16];
    for (int m = 0; m < 8; m++) h[2 * m] = 0.8 + 0.01 * m, h[2 * m + 1] = 0.6 - 0.01 * m;
    cudaMalloc(&d, 148 * 512 * 8);
    cudaMalloc(&tw, sizeof(h));
    cudaMemcpy(tw, h, sizeof(h), cudaMemcpyHostToDevice);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    // FP64 instruction counts per loop iteration are taken from the SASS (cuobjdump -sass | grep -c 'D(ADD|MUL|FMA)' per loop body)
    for (int threads : {128, 256, 384, 512}) {
        run<0>("dft8 fwd + inv", FP64_MODE0, d, tw, threads, ghz);
        run<1>("dft8 + twiddles, fwd + inv", FP64_MODE1, d, tw, threads, ghz);
        run<2>("two polynomials interleaved", FP64_MODE2, d, tw, threads, ghz);
    }
    return 0;
}
