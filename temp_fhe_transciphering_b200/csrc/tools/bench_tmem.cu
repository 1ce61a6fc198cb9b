// bench_tmem.cu — microbenchmark of tensor memory used as a thread-private table space (development tool).
//
// Question it answers: can the per-thread FFT twiddles (64 registers in
// k_blind_rotate_v3) live in TMEM and be fetched with tcgen05.ld right before use, so that 6 groups
// (12 warps) fit in the register file?  Needed numbers: tcgen05.ld throughput per SM for small shapes,
// its latency, and whether it competes with the shared-memory data pipe / the FP64 pipe.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_tmem bench_tmem.cu && ./bench_tmem
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, int cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, int cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE bit 0: tcgen05.ld of SHAPE words per iteration; bit 1: 8 conflict-free LDS.128 per iteration;
// bit 2: 32 dependent-free DFMA per iteration; bit 3: tcgen05.st x16 per iteration
template <int MODE, int SHAPE>
__global__ void __launch_bounds__(384, 1) k_tmem(int iters, double *out, long long *cycles, int *check)
{
    __shared__ uint32_t slot;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&slot, 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    // warp w owns TMEM lanes 32*(w%4).. and columns 64*(w/4).. (3 warps share a lane quarter at 12 warps)
    const uint32_t taddr = base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(64 * (warp >> 2));
    // fill: word c of this thread = tid * 1000 + c
    {
        uint32_t w[16];
        for (int c0 = 0; c0 < 64; c0 += 16) {
#pragma unroll
            for (int c = 0; c < 16; c++) w[c] = threadIdx.x * 1000u + c0 + c;
            tmem_st16(taddr + c0, w);
        }
        tmem_wait_st();
    }
    double4 *sm = reinterpret_cast<double4 *>(dyn);  // 32 KB region, LDS.128 = 16 B ... use uint4
    uint4 *smq = reinterpret_cast<uint4 *>(dyn);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) smq[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    (void)sm;
    uint32_t acc = 0;
    double f0 = 1.0 + lane, f1 = 0.5, f2 = 0.25, f3 = 2.0, g[8];
#pragma unroll
    for (int k = 0; k < 8; k++) g[k] = 1.0 + k;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE & 1) {
            if (SHAPE == 8) {  // four loads in flight, one wait
                uint32_t r[4][8];
#pragma unroll
                for (int q = 0; q < 4; q++) tmem_ld8(taddr + 8 * ((it + q) & 7), r[q]);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc += r[q][c];
            } else if (SHAPE == 9) {  // wait after every load (latency chain at 1 warp)
                uint32_t r[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    tmem_ld8(taddr + 8 * ((it + q) & 7), r);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 8; c++) acc += r[c];
                }
            } else if (SHAPE == 16) {
                uint32_t r[16];
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    tmem_ld16(taddr + 16 * ((it + q) & 3), r);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 16; c++) acc += r[c];
                }
            } else {
                uint32_t r[32];
                tmem_ld32(taddr + 32 * (it & 1), r);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 32; c++) acc += r[c];
            }
        }
        if (MODE & 2) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                uint4 x = smq[((it + q) & 7) * 256 + (threadIdx.x & 255)];
                acc += x.x ^ x.y ^ x.z ^ x.w;
            }
        }
        if (MODE & 4) {
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int k = 0; k < 8; k++) g[k] = fma(g[k], f1, f2);
        }
        if (MODE & 8) {
            uint32_t w[16];
#pragma unroll
            for (int c = 0; c < 16; c++) w[c] = acc + c;
            tmem_st16(taddr + 16 * (it & 3), w);
            tmem_wait_st();
        }
    }
    const long long t1 = clock64();
    double s = f0 + f3;
#pragma unroll
    for (int k = 0; k < 8; k++) s += g[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    // correctness probe of the lane/column mapping: re-read word 5 and 37
    if (blockIdx.x == 0 && MODE == 1 && SHAPE == 8) {
        // refill (mode 8 not active here so contents are intact)
        uint32_t r[8];
        tmem_ld8(taddr + 32, r);
        tmem_wait_ld();
        if (r[5] != threadIdx.x * 1000u + 37) atomicAdd(check, 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(base, 512);
}

template <int MODE, int SHAPE>
static int run(const char *name, int warps, int iters, double *d_out, long long *d_cyc, int *d_chk)
{
    cudaFuncSetAttribute(k_tmem<MODE, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_tmem<MODE, SHAPE><<<148, warps * 32, 32768>>>(10, d_out, d_cyc, d_chk);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k_tmem<MODE, SHAPE><<<148, warps * 32, 32768>>>(iters, d_out, d_cyc, d_chk);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc;
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double per_it = (double)cyc / iters;
    const double tm_bytes = (MODE & 1) ? 128.0 * 32 * warps : 0;  // per iteration per SM (32 words x 32 lanes x 4 B per warp)
    const double st_bytes = (MODE & 8) ? 128.0 * 16 * warps : 0;
    const double ls_bytes = (MODE & 2) ? 8.0 * 512 * warps : 0;
    const double dfma = (MODE & 4) ? 32.0 * warps : 0;  // warp instructions
    printf("%-28s warps %2d  %8.1f cyc/iter  tmem-ld %6.1f B/clk/SM  tmem-st %6.1f B/clk/SM  lds %6.1f B/clk/SM  dfma %5.2f warp-instr/clk/SM  (%.3f ms)\n",
           name, warps, per_it, tm_bytes / per_it, st_bytes / per_it, ls_bytes / per_it, dfma / per_it, ms);
    return 0;
}

int main()
{
    double *d_out;
    long long *d_cyc;
    int *d_chk;
    CK(cudaMalloc(&d_out, 148 * 384 * 8));
    CK(cudaMalloc(&d_cyc, 8));
    CK(cudaMalloc(&d_chk, 4));
    CK(cudaMemset(d_chk, 0, 4));
    const int it = 20000;
    for (int warps : {1, 4, 8, 12}) {
        run<1, 8>("tmem.ld x8 (4 per iter)", warps, it, d_out, d_cyc, d_chk);
        run<1, 9>("tmem.ld x8 wait each", warps, it, d_out, d_cyc, d_chk);
        run<1, 16>("tmem.ld x16 (2 per iter)", warps, it, d_out, d_cyc, d_chk);
        run<1, 32>("tmem.ld x32 (1 per iter)", warps, it, d_out, d_cyc, d_chk);
        run<2, 8>("lds.128 x8", warps, it, d_out, d_cyc, d_chk);
        run<3, 8>("tmem.ld x8 + lds", warps, it, d_out, d_cyc, d_chk);
        run<3, 32>("tmem.ld x32 + lds", warps, it, d_out, d_cyc, d_chk);
        run<4, 8>("dfma x32", warps, it, d_out, d_cyc, d_chk);
        run<5, 8>("tmem.ld x8 + dfma", warps, it, d_out, d_cyc, d_chk);
        run<6, 8>("lds + dfma", warps, it, d_out, d_cyc, d_chk);
        run<7, 8>("tmem.ld x8 + lds + dfma", warps, it, d_out, d_cyc, d_chk);
        run<8, 8>("tmem.st x16", warps, it, d_out, d_cyc, d_chk);
        run<9, 16>("tmem.ld x16x2 + st x16", warps, it, d_out, d_cyc, d_chk);
    }
    int chk;
    cudaMemcpy(&chk, d_chk, 4, cudaMemcpyDeviceToHost);
    printf("mapping check mismatches: %d\n", chk);
    return chk != 0;
}
