// Microbenchmark: one-way latency of a small message between two GPUs over NVLink (peer-mapped memory), for the
// write / poll instruction choices available to the per-attempt shared-step exchange (csrc/b2ode_fused.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nvlink_pingpong nvlink_pingpong.cu && ./nvlink_pingpong
// GPU 0 writes sequence number s into GPU 1's memory, GPU 1 polls its LOCAL memory, answers into GPU 0's memory, GPU 0
// polls its local memory: time per round trip / 2 = one hop (write latency + poll detection).
//   WM 0 st.relaxed.sys   1 st.volatile   2 red.relaxed.sys.max (remote atomic)   3 st.relaxed.sys + fence.sys
//   RM 0 ld.relaxed.sys   1 ld.relaxed.gpu 2 ld.volatile       3 ld.global.cg (weak)   4 ld.acquire.sys
//   words: how many 8-byte words the message has (each polled word is one more strong load per poll)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void wr(int wm, unsigned long long *p, unsigned long long v) {
    if (wm == 1) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    else if (wm == 2) asm volatile("red.relaxed.sys.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    else {
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
        if (wm == 3) asm volatile("fence.acq_rel.sys;" ::: "memory");
    }
}
__device__ __forceinline__ unsigned long long rd(int rm, const unsigned long long *p) {
    unsigned long long v;
    if (rm == 1) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else if (rm == 2) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else if (rm == 3) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else if (rm == 4) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// who = 0 starts; `mine` is local memory (polled), `theirs` is the peer's memory (written)
__global__ void k(int who, int wm, int rm, int words, int iters, unsigned long long *mine, unsigned long long *theirs,
                  unsigned long long *cycles) {
    unsigned long long t0 = 0;
    for (int it = 1; it <= iters; ++it) {
        if (it == 33) t0 = clock64();
        const unsigned long long s = (unsigned long long)it;
        if (who == 0)
            for (int w = 0; w < words; ++w) wr(wm, theirs + w, s);
        for (int w = words - 1; w >= 0; --w)
            while (rd(rm, mine + w) < s) {}
        if (who == 1)
            for (int w = 0; w < words; ++w) wr(wm, theirs + w, s);
    }
    if (who == 0) *cycles = clock64() - t0;
}

int main() {
    int n = 0;
    cudaGetDeviceCount(&n);
    if (n < 2) {
        printf("needs 2 GPUs\n");
        return 0;
    }
    unsigned long long *buf[2], *cyc;
    cudaStream_t st[2];
    for (int d = 0; d < 2; ++d) {
        cudaSetDevice(d);
        cudaDeviceEnablePeerAccess(1 - d, 0);
        cudaMalloc(&buf[d], 4096);
        cudaStreamCreate(&st[d]);
    }
    cudaSetDevice(0);
    cudaMalloc(&cyc, 64);
    const int iters = 4032;
    const int combos[][3] = {{0, 0, 1}, {0, 1, 1}, {0, 2, 1}, {0, 3, 1}, {0, 4, 1}, {1, 2, 1}, {2, 0, 1}, {2, 1, 1}, {2, 3, 1}, {3, 0, 1},
                             {0, 0, 2}, {0, 0, 4}, {0, 1, 4}, {0, 3, 4}};
    for (auto &c : combos) {
        for (int d = 0; d < 2; ++d) {
            cudaSetDevice(d);
            cudaMemset(buf[d], 0, 4096);
            cudaDeviceSynchronize();
        }
        for (int d = 0; d < 2; ++d) {
            cudaSetDevice(d);
            k<<<1, 1, 0, st[d]>>>(d, c[0], c[1], c[2], iters, buf[d], buf[1 - d], cyc);
        }
        cudaError_t e = cudaSuccess;
        for (int d = 0; d < 2; ++d) {
            cudaSetDevice(d);
            cudaError_t e2 = cudaDeviceSynchronize();
            if (e2 != cudaSuccess) e = e2;
        }
        if (e != cudaSuccess) {
            printf("WM %d RM %d words %d: %s\n", c[0], c[1], c[2], cudaGetErrorString(e));
            return 1;
        }
        unsigned long long v;
        cudaSetDevice(0);
        cudaMemcpy(&v, cyc, 8, cudaMemcpyDeviceToHost);
        printf("WM %d RM %d words %d: %.0f cycles per one-way hop\n", c[0], c[1], c[2], (double)v / (iters - 32) / 2.0);
        fflush(stdout);
    }
    return 0;
}
