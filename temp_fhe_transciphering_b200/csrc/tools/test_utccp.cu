// test_utccp.cu — which shared-memory matrix descriptor and which lane mapping does
//   tcgen05.cp.cta_group::1.64x128b.warpx2::02_13
// use?  (development tool: the blind rotation wants BSK tiles copied smem -> tensor memory by the async proxy so that the
// key reads leave the LSU shared-memory pipe.)  64 rows x 16 B in shared memory, row r = words {100 r + 0..3};
// expectation: lane l of quarter q receives row 32 (q & 1) + l.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o test_utccp test_utccp.cu && ./test_utccp
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_test(int variant, uint32_t *out /*[128][4]*/)
{
    __shared__ __align__(1024) uint32_t mat[64 * 4 * 4];  // room for variants with other strides
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 4 * 4; i += blockDim.x) mat[i] = 0xdead0000u + i;
    __syncthreads();
    // rows of 16 B at stride 16 B: core matrices (8 rows x 16 B = 128 B) back to back
    for (int r = threadIdx.x; r < 64; r += blockDim.x)
        for (int w = 0; w < 4; w++) mat[r * 4 + w] = 100u * r + w;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = slot;
    if (threadIdx.x == 0) {
        const uint64_t addr = (uint64_t)((smem_u32(mat) & 0x3FFFF) >> 4);
        uint64_t lbo = 0, sbo = 0;
        if (variant == 0) { lbo = 0; sbo = 128 >> 4; }
        if (variant == 1) { lbo = 128 >> 4; sbo = 0; }
        if (variant == 2) { lbo = 16 >> 4; sbo = 128 >> 4; }
        if (variant == 3) { lbo = 128 >> 4; sbo = 128 >> 4; }
        const uint64_t desc = addr | (lbo << 16) | (sbo << 32) | (1ull << 46);  // version 1, SWIZZLE_NONE
        asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(tbase), "l"(desc) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the copy
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r0, r1, r2, r3;
    const uint32_t taddr = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[threadIdx.x * 4 + 0] = r0;
    out[threadIdx.x * 4 + 1] = r1;
    out[threadIdx.x * 4 + 2] = r2;
    out[threadIdx.x * 4 + 3] = r3;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tbase) : "memory");
}

int main()
{
    uint32_t *d, h[512];
    cudaMalloc(&d, sizeof(h));
    for (int variant = 0; variant < 4; variant++) {
        cudaMemset(d, 0xff, sizeof(h));
        k_test<<<1, 128>>>(variant, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e));
            return 1;
        }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int ok = 0;
        for (int t = 0; t < 128; t++) {
            const int q = t >> 5, l = t & 31, row = 32 * (q & 1) + l;
            bool good = true;
            for (int w = 0; w < 4; w++) good &= h[t * 4 + w] == 100u * row + w;
            ok += good;
        }
        printf("variant %d: %d / 128 threads hold the expected row;  thread 0: %u %u %u %u  thread 1: %u  thread 8: %u  thread 32: %u  thread 64: %u  thread 96: %u\n",
               variant, ok, h[0], h[1], h[2], h[3], h[4], h[32], h[128], h[256], h[384]);
    }
    return 0;
}
