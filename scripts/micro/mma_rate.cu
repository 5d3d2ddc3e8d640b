// Microbenchmark: cycles per tcgen05.mma.kind::tf32 (M = 128, SS operands in SWIZZLE_128B K-major shared memory) when
// `n` accumulating MMAs are issued back to back by one thread, for N = 64 / 128 / 256.  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_tf32(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) k_rate(int N, int n_mma, int same_operands, unsigned long long *out, int mode) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar, bar2, bar3;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 * 8 + 32768 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3f800000u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar3)));     // phase 0 of bar3 is complete
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = slot;
    if (warp == 0 && (mode < 10 ? threadIdx.x == 0 : true)) {
        const uint32_t idesc = idesc_tf32(N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 16384 * 8);
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            // A: 8 K blocks of [128 x 128 B]; B: 2 K blocks of [256 x 128 B]; K step of 8 tf32 = 32 bytes inside a block
            const int kb = same_operands ? 0 : (i >> 2);
            const uint64_t da = make_desc(a0 + (kb & 7) * 16384 + (i & 3) * 32), db = make_desc(b0 + (kb & 1) * 32768 + (i & 3) * 32);
            const uint32_t accum = i > 0 ? 1u : 0u;
            uint32_t elected = 1;
            if (mode >= 10) asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
            if (elected)
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(accum)
                : "memory");
            if ((i & 3) == 3 && (mode % 10) >= 1 && (mode < 10 || (threadIdx.x & 31) == 0)) {      // end of a K block: what the kernel's MMA loop does there
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
                if ((mode % 10) >= 3) {
                    uint32_t ok;
                    do {
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(ok) : "r"(smem_u32(&bar3)), "r"(0u) : "memory");
                    } while (!ok);
                }
                if ((mode % 10) >= 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                
            }
        }
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        } while (!ok);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            out[0] = (unsigned long long)(t1 - t0);
            out[1] = (unsigned long long)(t2 - t0);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

__global__ void __launch_bounds__(128, 1) k_clean(int N, int n_mma, unsigned long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_tf32(N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 16384 * 8);
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const uint64_t da = make_desc(a0 + ((i >> 2) & 7) * 16384 + (i & 3) * 32), db = make_desc(b0 + ((i >> 2) & 1) * 32768 + (i & 3) * 32);
            const uint32_t accum = i > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(accum)
                : "memory");
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        } while (!ok);
        const long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[0] = (unsigned long long)(t1 - t0);
            out[1] = (unsigned long long)(t2 - t0);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

int main() {
    unsigned long long *d, h[2];
    cudaMalloc(&d, 16);
    const size_t smem = 16384 * 8 + 32768 * 2 + 1024;
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_clean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int N : {64, 256})
        for (int n : {32, 256}) {
            k_clean<<<sms, 128, smem>>>(N, n, d);
            k_clean<<<sms, 128, smem>>>(N, n, d);
            cudaDeviceSynchronize();
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("clean N=%3d %3d MMAs: issue %6llu complete %6llu -> %6.1f clk/MMA\n", N, n, h[0], h[1], (double)h[1] / n);
        }
    for (int grid : {sms})
        for (int same : {0})
          for (int mode : {0, 10, 1, 11})
            for (int N : {64, 256})
                for (int n : {32, 256}) {
                    k_rate<<<grid, 128, smem>>>(N, n, same, d, mode);
                    k_rate<<<grid, 128, smem>>>(N, n, same, d, mode);
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    printf("mode %d grid %3d  %s  N=%3d  %3d MMAs: issue %6llu clk  complete %6llu clk  -> %6.1f clk/MMA  (floor %d)  %s\n", mode, grid,
                           same ? "same operands " : "walking K blks", N, n, h[0], h[1], (double)h[1] / n, 128 * N / 256, cudaGetErrorString(e));
                }
    return 0;
}
