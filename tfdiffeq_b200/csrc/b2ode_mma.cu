// b2ode_mma.cu -- the GEMM-shaped part of the path: a dense layer of an ODENet-style func on tcgen05.
//
// SURVEY.md 8(f)-3 / north_star: "tensor cores used only when func is the ODENet linear block (a genuine dense
// GEMM)".  The reference's ODEFunc is fc1 -> relu -> fc2 -> relu -> fc3 (tfdiffeq/models/dense_odenet.py:85-92).
// One kernel computes   out[M, N] = act( A[M, K] . W[N, K]^T + bias[N] )   with fp32 storage and TF32 tensor-core
// math (TensorFlow's own default for fp32 matmuls on Ampere and later), where A is either a plain activation
// matrix or -- for the first layer -- the Runge-Kutta stage input produced on the fly,
//     A = y0 + sum_j (dt * beta_j) * k_j          (tfdiffeq/rk_common.py:51, dt read from the device state),
// i.e. the stage combine is the A-operand producer and the stage input never round-trips HBM for the GEMM.
//
// Blackwell mapping: one CTA (128 threads) per 128-row tile; operands are written by the threads themselves into
// the canonical K-major SWIZZLE_128B shared-memory layout (they are computed, not copied, so there is nothing
// for TMA to fetch); tcgen05.mma.kind::tf32 (M = 128, N <= 256, K = 8 per instruction) is issued by one
// thread with the accumulator in TMEM; tcgen05.commit -> mbarrier tells the CTA when shared memory may be
// refilled / the accumulator read; the epilogue reads TMEM with tcgen05.ld (one accumulator row per thread), adds
// the bias, applies the activation and stores fp32 rows.

#include "b2ode_dev.cuh"

#include <stdint.h>
#include <stdlib.h>

constexpr int kMmaThreads = 128;
constexpr int kTileM = 128;
constexpr int kKChunk = 64;              // K elements staged per round: 2 swizzle blocks of 32 tf32 (128 bytes)
constexpr int kMaxNK = 8;

struct DenseParams {
    const float *x;                      // A (nk == 0) or y0 (nk > 0), row-major [M, K]
    const float *k[kMaxNK];              // stage derivatives, row-major [M, K]
    double coef[kMaxNK];                 // beta_j of the stage row (zeros already dropped)
    int nk;
    const b2ode_state *st;               // dt lives here when nk > 0
    float *ystage;                       // optional: also materialise the stage input (needed for the last stage)
    const float *W;                      // [N, K] row-major (torch nn.Linear.weight layout), values already rounded to TF32
    const float *W_lo;                   // null: plain TF32.  Else tf32(W_fp32 - W): the 3xTF32 split (fp32-accurate products)
    const float *bias;                   // [N] or null
    float *out;                          // [M, N]
    int M, K, N;
    int act;                             // 0 none, 1 relu, 2 tanh, 3 softplus
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// low half of the 3xTF32 split: tf32(x - tf32(x)); the subtraction is exact in fp32
__device__ __forceinline__ uint32_t tf32_lo(float x) {
    return to_tf32(__fsub_rn(x, __uint_as_float(to_tf32(x))));
}

// Round to TF32's 11 significant bits with three FP32 operations (Veltkamp's split, C = 2^13 + 1): round-to-nearest
// (ties to even), low 13 mantissa bits come out zero, NaN and Inf stay non-finite.  cvt.rna.tf32.f32 has no fast
// hardware path (ptxas emulates it in ~7 integer/predicate instructions); this is what the activation epilogues use.
// |x| > 4e34 overflows the intermediate product and yields NaN.
__device__ __forceinline__ uint32_t to_tf32_fast(float x) {
    const float g = __fmul_rn(x, 8193.0f);
    const float d = __fsub_rn(x, g);
    return __float_as_uint(__fadd_rn(g, d));
}

// byte offset of (row, 16-byte chunk c in 0..7) inside one [rows x 128 B] K-major SWIZZLE_128B block:
// 8-row atoms of 1024 B, the chunk index XOR-ed with the row inside the atom (Swizzle<3,4,3>)
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B, dense 8-row atoms (SBO = 1024 B), sm_100 version bit
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address  [0,14)
    d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)  [46,48)
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B  [61,64)
    return d;
}

// The same descriptor as two 32-bit words.  The high word is a constant; the low word is (address >> 4) | LBO, so a
// K step of 8 TF32 (32 bytes) is `+ 2` and the next 16 KB block is `+ 1024`.  The MMA-issuing lane is a single thread:
// every integer instruction it spends on descriptor arithmetic is ~4-5 cycles of MMA issue latency, and the unrolled
// make_desc() form cost ~300 cycles per MMA against a tensor-pipe floor of 128 (scripts/micro/mma_rate.cu).
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFF) >> 4) | (1u << 16); }

// Four accumulating MMAs over one 32-column K block (cta_group 1 or 2).  Executed by ALL 32 lanes of the MMA warp,
// convergently; `elect.sync` inside the statement picks the one lane that issues.  Written as `if (lane == 0) { mma }`
// the instruction sits in a divergent region and nvcc feeds every operand through ELECT + R2UR.BROADCAST in a
// per-thread waterfall loop (~380 cycles per MMA measured, tensor-pipe floor 128); in this form the operands are
// computed in the uniform datapath and the four UTCHMMA issue back to back.
template <int CG>
__device__ __forceinline__ void mma_kblock_tf32(uint32_t tacc, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum_first) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const uint64_t da = ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + 2u * ks);
        const uint64_t db = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + 2u * ks);
        const uint32_t accum = ks > 0 ? 1u : accum_first;
        if (CG == 1)
            asm volatile(
                "{\n\t.reg .pred p, q;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "elect.sync _|q, 0xffffffff;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tacc),
                "l"(da), "l"(db), "r"(idesc), "r"(accum)
                : "memory");
        else
            asm volatile(
                "{\n\t.reg .pred p, q;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "elect.sync _|q, 0xffffffff;\n\t"
                "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tacc),
                "l"(da), "l"(db), "r"(idesc), "r"(accum)
                : "memory");
    }
}

// tcgen05.commit by one elected lane of a converged warp (same reasoning)
__device__ __forceinline__ void commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
        : "memory");
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int N) {
    uint32_t d = 0;
    d |= 1u << 4;                                         // D format F32
    d |= 2u << 7;                                         // A format TF32
    d |= 2u << 10;                                        // B format TF32
    d |= (uint32_t)(N >> 3) << 17;                        // N
    d |= (uint32_t)(kTileM >> 4) << 24;                   // M = 128
    return d;                                             // A, B K-major; no negate; dense
}

// The spin loop lives INSIDE the asm statement: to the compiler the wait is one convergent instruction, so the control
// flow of the calling warp stays provably uniform and the descriptor / barrier-address arithmetic of the MMA-issuing
// warp is done in the uniform datapath.  With the loop written in C++ (exit condition = a per-thread asm output) every
// tcgen05.mma operand went through ELECT + R2UR.BROADCAST, ~380 cycles per MMA against a 128-cycle tensor-pipe floor.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// compile-time activation: a run-time switch inlined per element bloats the unrolled epilogues past the
// instruction cache (tanhf + log1pf/expf bodies 32 times over) even when only the relu branch ever executes
template <int ACT>
__device__ __forceinline__ float act_t(float v) {
    if (ACT == 1) return v > 0.f ? v : 0.f;
    if (ACT == 2) return tanhf(v);
    if (ACT == 3) return (v > 20.f) ? v : log1pf(expf(v));
    return v;
}

// one output column of a transposed 32 x 32 block: bias + activation, one 128-byte line per warp store
template <int ACT>
__device__ __forceinline__ void store_column_t(float *dst, const float *tile, int lane, float bv, int ld, int nrows) {
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr)
        if (rr < nrows) dst[(size_t)rr * ld] = act_t<ACT>(tile[rr * 33 + lane] + bv);
}

// ---- CTA-pair plumbing (k_mlp3_tf32_pair) ----
// arrive on the mbarrier at the same shared-memory offset in CTA 0 of the cluster (the MMA-issuing CTA).  The plain
// (release.cta) form, as CUTLASS's ClusterBarrier::arrive(cta_id) uses for cross-CTA hand-offs: the operand bytes were
// already pushed to the async proxy by the fence.proxy.async that precedes every call; a cluster-scope release here
// costs ~1 us per call (measured: it doubled the activation epilogues).
__device__ __forceinline__ void mbar_arrive_cta0(uint32_t local_bar) {
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (rank == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(local_bar) : "memory");
    } else {
        uint32_t r;
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(local_bar));
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
    }
}

// wait on a barrier that threads of the OTHER CTA arrive on
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// for waits that last microseconds (keeps the spinning warps off the issue ports the working warps need)
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "nanosleep.u32 64;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return v > 0.f ? v : 0.f;
    if (act == 2) return tanhf(v);
    if (act == 3) return (v > 20.f) ? v : log1pf(expf(v));
    return v;
}

// ---- operand producers (shared by the simple kernel and the warp-specialised one) ----------------------------------
// Fill one K chunk (64 columns) of the B operand (NT rows of W, host-rounded TF32) and of the A operand (128 rows:
// a plain activation tile, or the Runge-Kutta stage input formed on the fly) in the K-major SWIZZLE_128B layout.
// Called by NTHR threads with ids 0..NTHR-1.  NK (the number of k's in the stage combine) is a template parameter
// so that every global load of a batch -- BQ positions x (1 + NK) streams -- is issued before the first one is
// consumed: the producers are latency-bound, memory-level parallelism is what feeds the tensor core.
// 3xTF32 (p.W_lo != null): every fp32 operand is split as x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and the
// product is accumulated in fp32 (TMEM) as A_lo.W_hi + A_hi.W_lo + A_hi.W_hi -- the K loop simply runs three times
// over the same columns with `mode` selecting which halves are staged (1: A_lo/W_hi, 2: A_hi/W_lo, 0: A_hi/W_hi).
// The dropped A_lo.W_lo term is 2^-22 relative: the result is as accurate as an fp32 FMA chain.
template <int NTHR, int NK>
__device__ __forceinline__ void produce_chunk_t(const DenseParams &p, uint8_t *sA, uint8_t *sB, int m0, int n0, int NT, int kc,
                                                int tid, const float (&cf)[kMaxNK], int mode) {
    const bool vec_ok = (p.K & 3) == 0;
    const float *Wsrc = (mode == 2) ? p.W_lo : p.W;
    // ---- B chunk first: NT rows (output features) x 64 columns of W[N, K], already TF32-rounded by the host:
    //      raw 16-byte async copies straight into the swizzled layout (no register staging), zero-filled past K
    if (vec_ok) {
        for (int f = tid; f < NT * (kKChunk / 4); f += NTHR) {
            const int row = f / (kKChunk / 4), c4 = f % (kKChunk / 4);
            const int gk = kc + c4 * 4;
            const float *src = Wsrc + (size_t)(n0 + row) * p.K + (gk < p.K ? gk : 0);
            const uint32_t dst = smem_u32(sB + (c4 >> 3) * (256 * 128) + sw128_offset(row, c4 & 7));
            const int nbytes = gk < p.K ? 16 : 0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
        }
    } else {
        for (int f = tid; f < NT * (kKChunk / 4); f += NTHR) {
            const int row = f / (kKChunk / 4), c4 = f % (kKChunk / 4);
            const int gk = kc + c4 * 4;
            uint32_t e[4] = {0u, 0u, 0u, 0u};
            for (int q = 0; q < 4 && gk + q < p.K; ++q) e[q] = __float_as_uint(Wsrc[(size_t)(n0 + row) * p.K + gk + q]);
            *reinterpret_cast<uint4 *>(sB + (c4 >> 3) * (256 * 128) + sw128_offset(row, c4 & 7)) = make_uint4(e[0], e[1], e[2], e[3]);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // ---- A chunk: 128 rows x 64 columns = PP float4 per thread
    constexpr int PP = kTileM * (kKChunk / 4) / NTHR;          // 16 (128 threads) or 8 (256 threads)
    constexpr int BQ = (NK == 0) ? PP : (PP < 4 ? PP : 4);     // positions per batch
    if (vec_ok) {
#pragma unroll 1
        for (int b0 = 0; b0 < PP; b0 += BQ) {
            float4 v[BQ];
            float4 kv[NK > 0 ? NK : 1][BQ];
            size_t off[BQ];
            bool in[BQ];
#pragma unroll
            for (int q = 0; q < BQ; ++q) {
                const int f = tid + (b0 + q) * NTHR;
                const int row = f / (kKChunk / 4), c4 = f % (kKChunk / 4);
                const int gm = m0 + row, gk = kc + c4 * 4;
                in[q] = gm < p.M && gk < p.K;
                off[q] = in[q] ? (size_t)gm * p.K + gk : 0;
            }
            // all loads of the batch first ...
#pragma unroll
            for (int q = 0; q < BQ; ++q)
                v[q] = in[q] ? *reinterpret_cast<const float4 *>(p.x + off[q]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < NK; ++j)
#pragma unroll
                for (int q = 0; q < BQ; ++q)
                    kv[j][q] = in[q] ? *reinterpret_cast<const float4 *>(p.k[j] + off[q]) : make_float4(0.f, 0.f, 0.f, 0.f);
            // ... then the combine (rk_common.py:51: (dt*beta_j)*k_j summed left to right, then y0 + sum), in fp32
            if (NK > 0) {
#pragma unroll
                for (int q = 0; q < BQ; ++q) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < NK; ++j) {
                        const float c = cf[j];
                        const float tx = __fmul_rn(c, kv[j][q].x), ty = __fmul_rn(c, kv[j][q].y);
                        const float tz = __fmul_rn(c, kv[j][q].z), tw = __fmul_rn(c, kv[j][q].w);
                        acc.x = j ? __fadd_rn(acc.x, tx) : tx;
                        acc.y = j ? __fadd_rn(acc.y, ty) : ty;
                        acc.z = j ? __fadd_rn(acc.z, tz) : tz;
                        acc.w = j ? __fadd_rn(acc.w, tw) : tw;
                    }
                    v[q].x = __fadd_rn(v[q].x, acc.x);
                    v[q].y = __fadd_rn(v[q].y, acc.y);
                    v[q].z = __fadd_rn(v[q].z, acc.z);
                    v[q].w = __fadd_rn(v[q].w, acc.w);
                    if (p.ystage && n0 == 0 && mode == 0 && in[q]) *reinterpret_cast<float4 *>(p.ystage + off[q]) = v[q];
                }
            }
#pragma unroll
            for (int q = 0; q < BQ; ++q) {
                const int f = tid + (b0 + q) * NTHR;
                const int row = f / (kKChunk / 4), c4 = f % (kKChunk / 4);
                const uint4 t = (mode == 1) ? make_uint4(tf32_lo(v[q].x), tf32_lo(v[q].y), tf32_lo(v[q].z), tf32_lo(v[q].w))
                                            : make_uint4(to_tf32(v[q].x), to_tf32(v[q].y), to_tf32(v[q].z), to_tf32(v[q].w));
                *reinterpret_cast<uint4 *>(sA + (c4 >> 3) * (kTileM * 128) + sw128_offset(row, c4 & 7)) = t;
            }
        }
    } else {
        for (int f = tid; f < kTileM * (kKChunk / 4); f += NTHR) {
            const int row = f / (kKChunk / 4), c4 = f % (kKChunk / 4);
            const int gm = m0 + row, gk = kc + c4 * 4;
            float e[4] = {0.f, 0.f, 0.f, 0.f};
            if (gm < p.M) {
                for (int q = 0; q < 4 && gk + q < p.K; ++q) {
                    const size_t off = (size_t)gm * p.K + gk + q;
                    float a = p.x[off];
                    if (NK > 0) {
                        float acc = 0.f;
#pragma unroll
                        for (int j = 0; j < NK; ++j) {
                            const float t = __fmul_rn(cf[j], p.k[j][off]);
                            acc = j ? __fadd_rn(acc, t) : t;
                        }
                        a = __fadd_rn(a, acc);
                        if (p.ystage && n0 == 0 && mode == 0) p.ystage[off] = a;
                    }
                    e[q] = a;
                }
            }
            const uint4 t = (mode == 1) ? make_uint4(tf32_lo(e[0]), tf32_lo(e[1]), tf32_lo(e[2]), tf32_lo(e[3]))
                                        : make_uint4(to_tf32(e[0]), to_tf32(e[1]), to_tf32(e[2]), to_tf32(e[3]));
            *reinterpret_cast<uint4 *>(sA + (c4 >> 3) * (kTileM * 128) + sw128_offset(row, c4 & 7)) = t;
        }
    }
}

template <int NTHR>
__device__ __forceinline__ void produce_chunk(const DenseParams &p, uint8_t *sA, uint8_t *sB, int m0, int n0, int NT, int kc,
                                              int tid, const float (&cf)[kMaxNK], int mode = 0) {
    switch (p.nk) {
        case 0: produce_chunk_t<NTHR, 0>(p, sA, sB, m0, n0, NT, kc, tid, cf, mode); break;
        case 1: produce_chunk_t<NTHR, 1>(p, sA, sB, m0, n0, NT, kc, tid, cf, mode); break;
        case 2: produce_chunk_t<NTHR, 2>(p, sA, sB, m0, n0, NT, kc, tid, cf, mode); break;
        case 3: produce_chunk_t<NTHR, 3>(p, sA, sB, m0, n0, NT, kc, tid, cf, mode); break;
        case 4: produce_chunk_t<NTHR, 4>(p, sA, sB, m0, n0, NT, kc, tid, cf, mode); break;
        case 5: produce_chunk_t<NTHR, 5>(p, sA, sB, m0, n0, NT, kc, tid, cf, mode); break;
        case 6: produce_chunk_t<NTHR, 6>(p, sA, sB, m0, n0, NT, kc, tid, cf, mode); break;
        case 7: produce_chunk_t<NTHR, 7>(p, sA, sB, m0, n0, NT, kc, tid, cf, mode); break;
        default: produce_chunk_t<NTHR, 8>(p, sA, sB, m0, n0, NT, kc, tid, cf, mode); break;
    }
}

__global__ void __launch_bounds__(kMmaThreads, 2) k_dense_layer_tf32(const __grid_constant__ DenseParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the swizzle atoms
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                   // 2 blocks x [128 rows x 128 B]          = 32 KB
    uint8_t *sB = smem + 2 * kTileM * 128;                // 2 blocks x [256 rows x 128 B] (max)    = 64 KB
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * kTileM;
    const int NT_max = p.N < 256 ? p.N : 256;
    uint32_t ncols = 32;
    while ((int)ncols < NT_max) ncols <<= 1;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_slot;
    uint32_t phase = 0;

    float cf[kMaxNK];
    if (p.nk > 0) {
        const float dt = (float)p.st->dt;
#pragma unroll
        for (int j = 0; j < kMaxNK; ++j) cf[j] = (j < p.nk) ? __fmul_rn(dt, (float)p.coef[j]) : 0.f;
    }

    for (int n0 = 0; n0 < p.N; n0 += 256) {
        const int NT = (p.N - n0) < 256 ? (p.N - n0) : 256;
        const uint32_t idesc = make_idesc_tf32(NT);
        const int passes = p.W_lo ? 3 : 1;
        for (int pass = 0; pass < passes; ++pass)
        for (int kc = 0; kc < p.K; kc += kKChunk) {
            const int mode = p.W_lo ? (pass == 0 ? 1 : pass == 1 ? 2 : 0) : 0;
            produce_chunk<kMmaThreads>(p, sA, sB, m0, n0, NT, kc, tid, cf, mode);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            // make the generic-proxy writes visible to the tensor core (async proxy), then hand over
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t da = make_desc(smem_u32(sA + kb * (kTileM * 128)) + ks * 32);
                        const uint64_t db = make_desc(smem_u32(sB + kb * (256 * 128)) + ks * 32);
                        const uint32_t accum = (pass > 0 || kc > 0 || kb > 0 || ks > 0) ? 1u : 0u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "setp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(accum)
                            : "memory");
                    }
                }
                // arrives on the mbarrier once every MMA issued so far has finished reading shared memory / writing TMEM
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                             : "memory");
            }
            mbar_wait(smem_u32(&mbar), phase);
            phase ^= 1u;
        }
        // ---- epilogue: one accumulator row per thread (TMEM lane = row), 32 columns at a time -----------------------
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int gm = m0 + tid;
        for (int c0 = 0; c0 < NT; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (gm < p.M) {
                float *dst = p.out + (size_t)gm * p.N + n0 + c0;
                const bool full = (c0 + 32 <= NT) && ((p.N & 3) == 0);
#pragma unroll
                for (int q = 0; q < 32; q += 4) {
                    float o[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float v = __uint_as_float(r[q + u]);
                        if (p.bias && c0 + q + u < NT) v += p.bias[n0 + c0 + q + u];
                        o[u] = apply_act(v, p.act);
                    }
                    if (full) {
                        *reinterpret_cast<float4 *>(dst + q) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (c0 + q + u < NT) dst[q + u] = o[u];
                    }
                }
            }
        }
        // the accumulator columns are reused by the next N tile
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols));
}

// ================================================================================================
// Warp-specialised, persistent version (the one that is launched by default).
//   warps 0-7  producers : fill a 2-stage ring of (A chunk, B chunk) shared-memory buffers
//   warp  8    MMA       : one elected lane issues tcgen05.mma into one of two 256-column TMEM accumulators,
//                          tcgen05.commit releases the stage back to the producers / hands the accumulator over
//   warps 9-12 epilogue  : tcgen05.ld (each warp its own 32-lane quarter), bias + activation, fp32 row stores
// so operand production, tensor-core math and the epilogue of the previous tile overlap.  mbarriers: full[s]
// (128 producer arrivals: the group that owns the stage), empty[s] (1 commit), tfull[b] (1 commit), tempty[b] (128 epilogue arrivals).
// ================================================================================================
constexpr int kProdWarps = 8;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kWsThreads = (kProdWarps + 1 + 4) * 32;      // producers + MMA warp + 4 epilogue warps
constexpr int kStages = 2;
constexpr int kStageBytes = kTileM * 128 * 2 + 256 * 128 * 2;     // A 32 KB + B 64 KB

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(kWsThreads, 1) k_dense_layer_tf32_ws(const __grid_constant__ DenseParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[kStages], bar_empty[kStages], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_m = (p.M + kTileM - 1) / kTileM;
    const int tiles_n = (p.N + 255) / 256;
    const int items = tiles_m * tiles_n;
    const int kchunks = (p.K + kKChunk - 1) / kKChunk;
    const int chunks = p.W_lo ? 3 * kchunks : kchunks;      // 3xTF32: three passes over K (A_lo.W_hi, A_hi.W_lo, A_hi.W_hi)

    if (warp == kProdWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_full[i])), "r"((unsigned)(kProdThreads / kStages)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_empty[i])), "r"(1u));
        }
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_tfull[i])), "r"(1u));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_tempty[i])), "r"(128u));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_slot;

    if (warp < kProdWarps) {
        // ===== producers =====
        float cf[kMaxNK];
        if (p.nk > 0) {
            const float dt = (float)p.st->dt;
#pragma unroll
            for (int j = 0; j < kMaxNK; ++j) cf[j] = (j < p.nk) ? __fmul_rn(dt, (float)p.coef[j]) : 0.f;
        } else {
#pragma unroll
            for (int j = 0; j < kMaxNK; ++j) cf[j] = 0.f;
        }
        // two producer groups of 4 warps, each owning one stage of the ring: group g fills every chunk with
        // (chunk counter % 2) == g, so the global loads of two consecutive chunks are in flight at the same time
        const uint32_t grp = (uint32_t)(warp / (kProdWarps / kStages));
        const int ltid = tid - (int)grp * (kProdThreads / kStages);
        uint32_t it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int m0 = (item / tiles_n) * kTileM, n0 = (item % tiles_n) * 256;
            const int NT = (p.N - n0) < 256 ? (p.N - n0) : 256;
            for (int c = 0; c < chunks; ++c, ++it) {
                const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                if (s != grp) continue;
                mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);                       // the MMAs that read this stage are done
                uint8_t *sA = smem + s * kStageBytes, *sB = sA + kTileM * 128 * 2;
                const int pass = c / kchunks;
                const int mode = p.W_lo ? (pass == 0 ? 1 : pass == 1 ? 2 : 0) : 0;
                produce_chunk<kProdThreads / kStages>(p, sA, sB, m0, n0, NT, (c - pass * kchunks) * kKChunk, ltid, cf, mode);
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy (tensor core)
                mbar_arrive(smem_u32(&bar_full[s]));
            }
        }
    } else if (warp == kProdWarps) {
        // ===== MMA issuer =====
        uint32_t it = 0, acc = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++acc) {
            const int n0 = (item % tiles_n) * 256;
            const int NT = (p.N - n0) < 256 ? (p.N - n0) : 256;
            const uint32_t idesc = make_idesc_tf32(NT);
            const uint32_t ab = acc & 1u, aph = (acc >> 1) & 1u;
            mbar_wait(smem_u32(&bar_tempty[ab]), aph ^ 1u);                        // the epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t tacc = tmem_base + ab * 256u;
            for (int c = 0; c < chunks; ++c, ++it) {
                const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                mbar_wait(smem_u32(&bar_full[s]), ph);
                asm volatile("tcgen05.fence::after_thread_sync;");
                {
                    const uint32_t a_lo = desc_lo(smem_u32(smem)) + s * (uint32_t)(kStageBytes >> 4);
                    const uint32_t b_lo = a_lo + (uint32_t)((kTileM * 128 * 2) >> 4);
                    mma_kblock_tf32<1>(tacc, a_lo, b_lo, idesc, c > 0 ? 1u : 0u);
                    mma_kblock_tf32<1>(tacc, a_lo + (uint32_t)((kTileM * 128) >> 4), b_lo + (uint32_t)((256 * 128) >> 4), idesc, 1u);
                    commit_elect(smem_u32(&bar_empty[s]));
                    if (c == chunks - 1) commit_elect(smem_u32(&bar_tfull[ab]));
                }
            }
        }
    } else {
        // ===== epilogue =====
        // TMEM gives each lane one accumulator ROW (32 consecutive columns per load); storing that directly would make
        // every warp store touch 32 different rows, 16 bytes each.  The 32 x 32 block is transposed through a padded
        // shared-memory tile instead, so that each warp store writes 32 consecutive floats (one 128-byte line) of one row.
        const int q = warp & 3;                                                    // TMEM lane quarter this warp may read
        float *tile = reinterpret_cast<float *>(smem + kStages * kStageBytes) + q * (32 * 33);
        uint32_t acc = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++acc) {
            const int m0 = (item / tiles_n) * kTileM, n0 = (item % tiles_n) * 256;
            const int NT = (p.N - n0) < 256 ? (p.N - n0) : 256;
            const uint32_t ab = acc & 1u, aph = (acc >> 1) & 1u;
            mbar_wait(smem_u32(&bar_tfull[ab]), aph);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const int row0 = m0 + q * 32;
            for (int c0 = 0; c0 < NT; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ab * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                __syncwarp();                                                       // previous block fully read back
#pragma unroll
                for (int w = 0; w < 32; ++w) tile[lane * 33 + w] = __uint_as_float(r[w]);   // row = lane (conflict-free: stride 33)
                __syncwarp();
                const int col = c0 + lane;                                          // after the transpose a lane owns a column
                const bool col_ok = col < NT;
                const float bv = (p.bias && col_ok) ? p.bias[n0 + col] : 0.f;
                float *dst = p.out + (size_t)row0 * p.N + n0 + col;
                if (col_ok) {
                    const int nrows = p.M - row0 < 32 ? p.M - row0 : 32;
                    switch (p.act) {
                        case 1: store_column_t<1>(dst, tile, lane, bv, p.N, nrows); break;
                        case 2: store_column_t<2>(dst, tile, lane, bv, p.N, nrows); break;
                        case 3: store_column_t<3>(dst, tile, lane, bv, p.N, nrows); break;
                        default: store_column_t<0>(dst, tile, lane, bv, p.N, nrows); break;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            mbar_arrive(smem_u32(&bar_tempty[ab]));                                // accumulator may be overwritten
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == kProdWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// ================================================================================================
// The whole ODENet-style func in ONE kernel: fc1 -> act -> fc2 -> act -> fc3 chained per 128-row tile, the hidden
// activations never leave the SM.
//
//   ACT   : 128 rows x up to 256 TF32 columns of shared memory (<= 128 KB), the A operand of whichever GEMM is running:
//           first the (optionally stage-combined) input tile, then act(h1), then act(h2) -- each written in the K-major
//           SWIZZLE_128B layout by the warps that produced it
//   ring  : S stages x [N rows x 32 columns] of the current layer's weights.  The weights are packed ONCE per weight
//           version (k_mlp3_pack) into exactly this shared-memory image -- TF32-rounded, swizzled, zero-padded -- so one
//           thread streams them with cp.async.bulk (the TMA engine's linear mode) and an mbarrier transaction count
//   TMEM  : acc0 (256 columns: GEMM1, later GEMM3), acc1 (256 columns: GEMM2)
//   warps : 0-7 input producers (the next tile is loaded and stage-combined into REGISTERS while the current tile is in
//           its GEMMs, then dropped into ACT the moment GEMM3 releases it), 8 MMA issuer, 9-12 activation epilogues
//           (TMEM -> bias/act -> ACT), 13 weight loader
// HBM traffic per evaluation: the input tile(s) and the output tile -- (1 + nk) x 4D + 4D bytes per row instead of
// 4(D + 4H + D) bytes per row through three separate layers; the weights stream from L2.
// ================================================================================================
constexpr int kSub = 32;                                   // K columns per weight sub-chunk (one 128-byte swizzle row)
constexpr int kEpiWarps = 4;                               // 4, or 8 (two per TMEM lane quarter: even / odd K blocks); measured equal
constexpr int kChainThreads = (kProdWarps + 1 + kEpiWarps + 1) * 32;
constexpr int kLoaderWarp = kProdWarps + 1 + kEpiWarps;
constexpr int kMaxRing = 8;
constexpr int kMlp3ActSlots = 8;           // ACT ring depth of k_mlp3_tf32: 8 = hold the whole activation tile (measured best); 2..7 = ring
constexpr bool kMlp3PairDefault = false;   // k_mlp3_tf32_pair (cta_group::2) instead of k_mlp3_tf32
#ifndef B2_BULK_PIECE
#define B2_BULK_PIECE 16384
#endif
constexpr uint32_t kBulkPiece = B2_BULK_PIECE;   // bytes per cp.async.bulk request
constexpr int kTileScratchBytes = kProdWarps * 32 * 17 * 4;  // the output epilogue's transposition scratch, one per producer warp

struct Mlp3Params {
    DenseParams in;                    // x / k / coef / nk / st / ystage / M / K(= D) describe the input tile
    const uint8_t *packed;             // k_mlp3_pack's image: layer 1 sub-chunks, layer 2, layer 3
    const float *b1, *b2, *b3;
    float *out;                        // [M, D]
    int M, D, H, act;
    int act_bytes, stage_bytes, stages;
    int act_slots;                     // R: ACT holds R K blocks (16 KB each); K block kb lives in slot kb % R
};

__host__ __device__ inline int mlp3_subs(int K) { return (K + kSub - 1) / kSub; }
__host__ __device__ inline size_t mlp3_layer_bytes(int N, int K) { return (size_t)mlp3_subs(K) * N * 128; }

// TF32-round, swizzle and zero-pad W[N, K] into sub-chunk images of [N rows x 128 B]
__global__ void k_mlp3_pack(const float *W1, const float *W2, const float *W3, int D, int H, uint8_t *packed) {
    const size_t n1 = mlp3_layer_bytes(H, D) / 16, n2 = mlp3_layer_bytes(H, H) / 16, n3 = mlp3_layer_bytes(D, H) / 16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + n3; i += (size_t)gridDim.x * blockDim.x) {
        const float *W;
        int N, K;
        size_t j = i;
        if (j < n1) {
            W = W1, N = H, K = D;
        } else if (j < n1 + n2) {
            W = W2, N = H, K = H, j -= n1;
        } else {
            W = W3, N = D, K = H, j -= n1 + n2;
        }
        const int q = (int)(j & 7);                       // 16-byte slot inside the 128-byte row of the image
        const int row = (int)((j >> 3) % N);
        const int c = (int)((j >> 3) / N);
        const int piece = q ^ (row & 7);                  // which 4 source columns live in that slot (Swizzle<3,4,3>)
        uint32_t e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int gk = c * kSub + piece * 4 + u;
            e[u] = gk < K ? to_tf32(W[(size_t)row * K + gk]) : 0u;
        }
        reinterpret_cast<uint4 *>(packed)[i] = make_uint4(e[0], e[1], e[2], e[3]);
    }
}

#ifdef B2ODE_TRACE
__device__ unsigned long long g_mlp3_trace[16 * 16];
__device__ __forceinline__ void trace(int tile_no, int slot) {
    if (blockIdx.x == 0 && tile_no < 16) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_mlp3_trace[tile_no * 16 + slot] = t;
    }
}
#define TRACE(tile_no, slot) trace((int)(tile_no), slot)
__device__ unsigned long long g_mlp3_clk[80];
__device__ int g_mlp3_clk_n;
#define CLK(cond)                                                            \
    if ((cond) && blockIdx.x == 0 && g_mlp3_clk_n < 80) g_mlp3_clk[g_mlp3_clk_n++] = clock64()
#else
#define TRACE(tile_no, slot)
#define CLK(cond)
#endif

// ---- input tile (D <= 64: one 64-column chunk) loaded and stage-combined into registers ----
template <int NK>
__device__ __forceinline__ void load_tile_regs_t(const DenseParams &p, int m0, int tid, const float (&cf)[kMaxNK], float4 (&v)[8]) {
    constexpr int BQ = NK >= 4 ? 2 : 4;
#pragma unroll
    for (int b0 = 0; b0 < 8; b0 += BQ) {
        float4 kv[NK > 0 ? NK : 1][BQ];
        size_t off[BQ];
        bool in[BQ];
#pragma unroll
        for (int q = 0; q < BQ; ++q) {
            const int f = tid + (b0 + q) * kProdThreads;
            const int row = f >> 4, c4 = f & 15;
            const int gm = m0 + row, gk = c4 * 4;
            in[q] = gm < p.M && gk < p.K;
            off[q] = in[q] ? (size_t)gm * p.K + gk : 0;
        }
#pragma unroll
        for (int q = 0; q < BQ; ++q)
            v[b0 + q] = in[q] ? *reinterpret_cast<const float4 *>(p.x + off[q]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NK; ++j)
#pragma unroll
            for (int q = 0; q < BQ; ++q)
                kv[j][q] = in[q] ? *reinterpret_cast<const float4 *>(p.k[j] + off[q]) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (NK > 0) {
#pragma unroll
            for (int q = 0; q < BQ; ++q) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < NK; ++j) {
                    const float c = cf[j];
                    const float tx = __fmul_rn(c, kv[j][q].x), ty = __fmul_rn(c, kv[j][q].y);
                    const float tz = __fmul_rn(c, kv[j][q].z), tw = __fmul_rn(c, kv[j][q].w);
                    acc.x = j ? __fadd_rn(acc.x, tx) : tx;
                    acc.y = j ? __fadd_rn(acc.y, ty) : ty;
                    acc.z = j ? __fadd_rn(acc.z, tz) : tz;
                    acc.w = j ? __fadd_rn(acc.w, tw) : tw;
                }
                float4 &r = v[b0 + q];
                r.x = __fadd_rn(r.x, acc.x);
                r.y = __fadd_rn(r.y, acc.y);
                r.z = __fadd_rn(r.z, acc.z);
                r.w = __fadd_rn(r.w, acc.w);
                if (p.ystage && in[q]) *reinterpret_cast<float4 *>(p.ystage + off[q]) = r;
            }
        }
    }
}

__device__ __forceinline__ void load_tile_regs(const DenseParams &p, int m0, int tid, const float (&cf)[kMaxNK], float4 (&v)[8]) {
    switch (p.nk) {
        case 0: load_tile_regs_t<0>(p, m0, tid, cf, v); break;
        case 1: load_tile_regs_t<1>(p, m0, tid, cf, v); break;
        case 2: load_tile_regs_t<2>(p, m0, tid, cf, v); break;
        case 3: load_tile_regs_t<3>(p, m0, tid, cf, v); break;
        case 4: load_tile_regs_t<4>(p, m0, tid, cf, v); break;
        case 5: load_tile_regs_t<5>(p, m0, tid, cf, v); break;
        case 6: load_tile_regs_t<6>(p, m0, tid, cf, v); break;
        case 7: load_tile_regs_t<7>(p, m0, tid, cf, v); break;
        default: load_tile_regs_t<8>(p, m0, tid, cf, v); break;
    }
}

__device__ __forceinline__ void store_tile_regs(uint8_t *act_buf, int tid, const float4 (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int f = tid + i * kProdThreads;
        const int row = f >> 4, c4 = f & 15;
        const uint4 t = make_uint4(to_tf32(v[i].x), to_tf32(v[i].y), to_tf32(v[i].z), to_tf32(v[i].w));
        *reinterpret_cast<uint4 *>(act_buf + (c4 >> 3) * (kTileM * 128) + sw128_offset(row, c4 & 7)) = t;
    }
}

// bias + activation + TF32 rounding of one accumulator block, written as the next GEMM's A operand, one 32-column
// K block at a time; each finished K block is handed to the MMA warp through its own mbarrier (`bar_kb`, one arrival
// per epilogue warp), so the next GEMM runs one K block behind this epilogue instead of after it.
// `bias` is the shared-memory copy (zero-filled when the layer has none, padded).  ncols is a multiple of 16.
//
// ACT is a ring of R slots: K block kb is written to slot kb % R.  For kb >= R the slot still holds K block kb - R of the
// same activation matrix until the GEMM that trails this epilogue has consumed it; the MMA warp commits `bar_free + slot`
// after every K block it consumes, and `gemm_no` (2 * tile + {0: h1, 1: h2}) locates the commit to wait for:
// every GEMM over K = ncols commits ceil((nblk - slot) / R) times on a slot.
template <int ACT>
__device__ __noinline__ void epilogue_to_act_t(uint32_t act_s, uint32_t tmem_acc, int q, int half, int lane, int ncols,
                                               const float *bias, uint32_t bar_kb, bool pair, int R, uint32_t bar_free,
                                               uint32_t gemm_no) {
    const int row = q * 32 + lane;
    const uint32_t row_s = act_s + (uint32_t)row * 128u;                  // (row >> 3) * 1024 + (row & 7) * 128
    const uint32_t rx = (uint32_t)(row & 7);
    const int ncols_pad = (ncols + kSub - 1) / kSub * kSub;
    for (int c0 = 32 * half; c0 < ncols_pad; c0 += 32 * (kEpiWarps / 4)) {   // with 8 warps: K blocks half, half + 2, ...
        uint32_t r[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        CLK(q == 0 && half == 0 && lane == 0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        CLK(q == 0 && half == 0 && lane == 0);
        const int kb = c0 / kSub, slot = kb % R;
        if (kb >= R) {
            const int nblk = ncols_pad / kSub;
            const uint32_t cnt = (uint32_t)((nblk - slot + R - 1) / R);          // commits per GEMM on this slot
            mbar_wait(bar_free + (uint32_t)slot * 8u, (gemm_no * cnt + (uint32_t)(kb / R) - 1u) & 1u);
        }
        const uint32_t blk_s = row_s + (uint32_t)slot * (kTileM * 128);            // K block of 32 columns: [128 rows x 128 B]
        const bool second_half = c0 + 16 < ncols;                                  // warp-uniform: ncols is a multiple of 16
        // all 32 values first (independent 5-op chains the scheduler can interleave), then the eight 16-byte stores;
        // the store asm carries no memory clobber so that nothing pins the bias reads or the math between stores --
        // the fence below is a volatile asm WITH a clobber and stays behind every one of them
        uint32_t t[32];
#pragma unroll
        for (int w = 0; w < 32; w += 4) {
            const float4 bv = *reinterpret_cast<const float4 *>(bias + c0 + w);       // broadcast read
            const bool in = w < 16 || second_half;
            // two elements per instruction (Blackwell's packed fp32x2 add / mul / fma: same IEEE results as the scalar
            // forms): bias add, activation, then the Veltkamp split g = 8193 v, d = v - g, r = g + d
            float2 v01 = __fadd2_rn(make_float2(__uint_as_float(r[w + 0]), __uint_as_float(r[w + 1])), make_float2(bv.x, bv.y));
            float2 v23 = __fadd2_rn(make_float2(__uint_as_float(r[w + 2]), __uint_as_float(r[w + 3])), make_float2(bv.z, bv.w));
            v01 = make_float2(act_t<ACT>(v01.x), act_t<ACT>(v01.y));
            v23 = make_float2(act_t<ACT>(v23.x), act_t<ACT>(v23.y));
            const float2 c8193 = make_float2(8193.0f, 8193.0f), m1 = make_float2(-1.0f, -1.0f);
            const float2 g01 = __fmul2_rn(v01, c8193), g23 = __fmul2_rn(v23, c8193);
            const float2 r01 = __fadd2_rn(g01, __ffma2_rn(g01, m1, v01)), r23 = __fadd2_rn(g23, __ffma2_rn(g23, m1, v23));
            t[w + 0] = in ? __float_as_uint(r01.x) : 0u;
            t[w + 1] = in ? __float_as_uint(r01.y) : 0u;
            t[w + 2] = in ? __float_as_uint(r23.x) : 0u;
            t[w + 3] = in ? __float_as_uint(r23.y) : 0u;
        }
#pragma unroll
        for (int w = 0; w < 32; w += 4)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk_s + ((((uint32_t)w >> 2) ^ rx) << 4)), "r"(t[w]),
                         "r"(t[w + 1]), "r"(t[w + 2]), "r"(t[w + 3]));
        CLK(q == 0 && half == 0 && lane == 0);
        asm volatile("tcgen05.fence::before_thread_sync;");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> the tensor core's async proxy
        CLK(q == 0 && half == 0 && lane == 0);
        __syncwarp();
        if (lane == 0) {
            if (pair)
                mbar_arrive_cta0(bar_kb + (uint32_t)(c0 / kSub) * 8u);     // the leader CTA issues the MMAs of both CTAs
            else
                mbar_arrive(bar_kb + (uint32_t)(c0 / kSub) * 8u);
        }
    }
}

__device__ __forceinline__ void epilogue_to_act(uint32_t act_s, uint32_t tmem_acc, int q, int half, int lane, int ncols,
                                                const float *bias, int act, uint32_t bar_kb, bool pair, int R, uint32_t bar_free,
                                                uint32_t gemm_no) {
    switch (act) {
        case 1: epilogue_to_act_t<1>(act_s, tmem_acc, q, half, lane, ncols, bias, bar_kb, pair, R, bar_free, gemm_no); break;
        case 2: epilogue_to_act_t<2>(act_s, tmem_acc, q, half, lane, ncols, bias, bar_kb, pair, R, bar_free, gemm_no); break;
        case 3: epilogue_to_act_t<3>(act_s, tmem_acc, q, half, lane, ncols, bias, bar_kb, pair, R, bar_free, gemm_no); break;
        default: epilogue_to_act_t<0>(act_s, tmem_acc, q, half, lane, ncols, bias, bar_kb, pair, R, bar_free, gemm_no); break;
    }
}

// Output tile: accumulator (D columns) + bias -> global rows.  Run by the eight producer warps after they have
// dropped the next tile's input into ACT, i.e. off the GEMM / activation-epilogue critical path.  Warp w reads TMEM
// lane quarter w & 3 (the hardware's per-warp access window) and the 32-column chunks (w >> 2), (w >> 2) + 2, ...
// TMEM hands each lane one ROW; 16 columns at a time are transposed through a [32][17] shared-memory scratch so that
// a warp store writes two 64-byte row segments instead of 32 scattered 4-byte words.
__device__ __forceinline__ void output_tile(const Mlp3Params &P, uint32_t tacc, int warp, int lane, int m0, float *scratch,
                                            const float *bias) {
    const int q = warp & 3;
    const int row0 = m0 + q * 32;
    const int D = P.D;
    for (int c0 = 32 * (warp >> 2); c0 < D; c0 += 64) {
        uint32_t r[32];
        const uint32_t taddr = tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (c0 + half * 16 >= D) break;                               // D is a multiple of 16: warp-uniform
            __syncwarp();
#pragma unroll
            for (int w = 0; w < 16; ++w) scratch[lane * 17 + w] = __uint_as_float(r[half * 16 + w]);
            __syncwarp();
            const int col = c0 + half * 16 + (lane & 15);
            const float bv = bias[col];
            float *dst = P.out + (size_t)row0 * D + col;
#pragma unroll 8
            for (int i = 0; i < 16; ++i) {
                const int rr = 2 * i + (lane >> 4);
                if (row0 + rr < P.M) dst[(size_t)rr * D] = scratch[rr * 17 + (lane & 15)] + bv;
            }
        }
    }
}

__global__ void __launch_bounds__(kChainThreads, 1) k_mlp3_tf32(const __grid_constant__ Mlp3Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *act_buf = smem;
    uint8_t *ring = smem + P.act_bytes;
    float *tiles = reinterpret_cast<float *>(ring + (size_t)P.stages * P.stage_bytes);
    __shared__ __align__(8) uint64_t bar_full[kMaxRing], bar_empty[kMaxRing];
    __shared__ __align__(8) uint64_t bar_a1, bar_actfree, bar_t1, bar_t2, bar_outdone;
    __shared__ __align__(8) uint64_t bar_a2[8], bar_a3[8];             // act(h1) / act(h2), one per 32-column K block
    __shared__ __align__(8) uint64_t bar_hfree[8];                     // ACT slot consumed by the trailing GEMM
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[3][256];

    const DenseParams &p = P.in;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = P.D, H = P.H;
    for (int i = tid; i < 3 * 256; i += kChainThreads) {
        const int l = i >> 8, c = i & 255;
        const float *b = l == 0 ? P.b1 : (l == 1 ? P.b2 : P.b3);
        sbias[l][c] = (b && c < (l == 2 ? D : H)) ? b[c] : 0.f;
    }
    const int tiles_m = (P.M + kTileM - 1) / kTileM;
    const int sub1 = mlp3_subs(D), sub2 = mlp3_subs(H);
    const uint32_t S = (uint32_t)P.stages;

    if (warp == kProdWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        auto init = [](uint64_t *b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); };
        for (int i = 0; i < kMaxRing; ++i) {
            init(&bar_full[i], 1u);              // the loader's expect_tx arrive; the bytes complete the phase
            init(&bar_empty[i], 1u);             // tcgen05.commit
        }
        init(&bar_a1, (unsigned)kProdThreads);   // input tile written by the producers
        init(&bar_actfree, 1u);                  // GEMM3 finished reading ACT (commit)
        init(&bar_t1, 1u);
        init(&bar_t2, 1u);
        init(&bar_outdone, (unsigned)kProdWarps); // output tile read out of its accumulator (one arrival per producer warp)
        for (int i = 0; i < 8; ++i) {
            init(&bar_a2[i], 4u);                // one arrival per epilogue warp
            init(&bar_a3[i], 4u);
            init(&bar_hfree[i], 1u);             // tcgen05.commit after the K block in that slot
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_slot;
    const uint32_t acc0 = tmem_base, acc1 = tmem_base + 256u;

    if (warp < kProdWarps) {
        // ===== input producers =====
        float cf[kMaxNK];
#pragma unroll
        for (int j = 0; j < kMaxNK; ++j) cf[j] = 0.f;
        if (p.nk > 0) {
            const float dt = (float)p.st->dt;
#pragma unroll
            for (int j = 0; j < kMaxNK; ++j)
                if (j < p.nk) cf[j] = __fmul_rn(dt, (float)p.coef[j]);
        }
        const bool prefetch = D <= kKChunk && (D & 3) == 0;
        float *scratch = tiles + warp * (32 * 17);
        uint32_t tcount = 0;
        int prev_m0 = -1;
        for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x, ++tcount) {
            const int m0 = tile * kTileM;
            if (prefetch) {
                float4 v[8];
                load_tile_regs(p, m0, tid, cf, v);                         // overlaps the previous tile's GEMMs
                TRACE(tcount, 0);
                mbar_wait_backoff(smem_u32(&bar_actfree), (tcount & 1u) ^ 1u);   // previous tile's GEMM3 is complete
                TRACE(tcount, 1);
                store_tile_regs(act_buf, tid, v);
            } else {
                mbar_wait_backoff(smem_u32(&bar_actfree), (tcount & 1u) ^ 1u);
                for (int kc = 0; kc < D; kc += kKChunk)
                    produce_chunk<kProdThreads>(p, act_buf + (kc / kSub) * (kTileM * 128), nullptr, m0, 0, 0, kc, tid, cf);
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(smem_u32(&bar_a1));
            TRACE(tcount, 2);
            if (prev_m0 >= 0) {
                // the previous tile's output, while this tile is in GEMM1 / its first activation epilogue
                asm volatile("tcgen05.fence::after_thread_sync;");
                output_tile(P, ((tcount - 1u) & 1u) ? acc1 : acc0, warp, lane, prev_m0, scratch, sbias[2]);
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_outdone));
                TRACE(tcount - 1u, 14);
            }
            prev_m0 = m0;
        }
        if (prev_m0 >= 0) {
            mbar_wait_backoff(smem_u32(&bar_actfree), (tcount & 1u) ^ 1u);       // the last tile's GEMM3
            asm volatile("tcgen05.fence::after_thread_sync;");
            output_tile(P, ((tcount - 1u) & 1u) ? acc1 : acc0, warp, lane, prev_m0, scratch, sbias[2]);
            TRACE(tcount - 1u, 14);
        }
    } else if (warp == kLoaderWarp) {
        // ===== weight loader: one bulk copy per sub-chunk, completion counted in bytes on the stage's mbarrier =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x) {
                const uint8_t *src = P.packed;
                for (int l = 0; l < 3; ++l) {
                    const int Nl = l == 2 ? D : H;
                    const int nsub = l == 0 ? sub1 : sub2;
                    const int grp = P.stage_bytes / (Nl * 128);             // K blocks per stage (narrow layers: several)
                    for (int c = 0; c < nsub; c += grp, ++it) {
                        const uint32_t bytes = (uint32_t)((nsub - c < grp ? nsub - c : grp) * Nl) * 128u;
                        const uint32_t s = it % S, ph = (it / S) & 1u;
                        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
                        const uint32_t bar = smem_u32(&bar_full[s]);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                        // (request size is not the limiter: 1 x 32 KB, 2 x 16 KB and 8 x 4 KB per stage time the same,
                        // profiles/r01_mlp3_chained.md -- the SM's ingest from L2 is)
                        const uint32_t dst0 = smem_u32(ring + (size_t)s * P.stage_bytes);
                        for (uint32_t o = 0; o < bytes; o += kBulkPiece) {
                            const uint32_t nb = bytes - o < kBulkPiece ? bytes - o : kBulkPiece;
                            asm volatile(
                                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst0 + o),
                                "l"(src + o), "r"(nb), "r"(bar)
                                : "memory");
                        }
                        src += bytes;
                    }
                }
            }
        }
    } else if (warp == kProdWarps) {
        // ===== MMA issuer =====
        const uint32_t act_lo = desc_lo(smem_u32(act_buf)), ring_lo = desc_lo(smem_u32(ring));
        const uint32_t empty_s = smem_u32(&bar_empty[0]), hfree_s = smem_u32(&bar_hfree[0]);
        const uint32_t R = (uint32_t)P.act_slots;
        const bool ring_act = sub2 > P.act_slots;
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x, ++tcount) {
            const uint32_t tp = tcount & 1u;
            for (int l = 0; l < 3; ++l) {
                const int Nl = l == 2 ? D : H;
                const int nsub = l == 0 ? sub1 : sub2;
                const uint32_t idesc = make_idesc_tf32(Nl);
                // accumulator roles alternate per tile (GEMM1 and GEMM3 -> A, GEMM2 -> B, A/B swapped on odd tiles), so this
                // tile's GEMM1 can run while the previous tile's output is still being read out of ITS accumulator A.
                // A(t+1) = B(t) needs no barrier: E2(t) drained it before GEMM3(t) could finish (GEMM3 consumes what E2
                // writes).  B(t+1) = A(t) is released by the producer warps' output pass (bar_outdone) before GEMM2(t+1).
                const uint32_t tacc = ((l == 1) != (tp == 1u)) ? acc1 : acc0;
                if (l == 0) mbar_wait(smem_u32(&bar_a1), tp);              // input tile is in ACT
                if (l == 1 && tcount > 0) mbar_wait(smem_u32(&bar_outdone), tp ^ 1u);   // B(t) = A(t-1) has been read out
                TRACE(tcount, 3 + l * 2);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const int grp = P.stage_bytes / (Nl * 128);
                const uint32_t b_step = (uint32_t)(Nl * 128) >> 4;          // descriptor units per weight K block
                uint32_t slot = 0;                                          // ACT slot of K block c + kb (= (c + kb) % R)
                for (int c = 0; c < nsub; c += grp, ++it) {
                    const uint32_t s = it % S, ph = (it / S) & 1u;
                    const int nb = nsub - c < grp ? nsub - c : grp;
#ifndef B2_MLP3_NOWAIT_W
                    mbar_wait(smem_u32(&bar_full[s]), ph);
#endif
                    uint32_t b_lo = ring_lo + s * ((uint32_t)P.stage_bytes >> 4);
                    for (int kb = 0; kb < nb; ++kb) {
                        if (l == 1) mbar_wait(smem_u32(&bar_a2[c + kb]), tp);      // this K block of act(h1) is in ACT
                        if (l == 2) mbar_wait(smem_u32(&bar_a3[c + kb]), tp);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        mma_kblock_tf32<1>(tacc, act_lo + slot * (uint32_t)((kTileM * 128) >> 4), b_lo, idesc, (c + kb) > 0 ? 1u : 0u);
                        if (ring_act && l > 0) commit_elect(hfree_s + slot * 8u);   // the slot may be overwritten once these MMAs have read it
                        b_lo += b_step;
                        if (++slot == R) slot = 0;
                    }
                    commit_elect(empty_s + s * 8u);
                    // GEMM3 complete = ACT may be refilled AND the output accumulator is final
                    if (c + nb == nsub) commit_elect(smem_u32(l == 0 ? &bar_t1 : (l == 1 ? &bar_t2 : &bar_actfree)));
                }
                TRACE(tcount, 4 + l * 2);
            }
        }
    } else {
        // ===== epilogue warps =====
        const int q = warp & 3, half = (warp - (kProdWarps + 1)) >> 2;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x, ++tcount) {
            const uint32_t tp = tcount & 1u;
            // h1 -> ACT
            mbar_wait(smem_u32(&bar_t1), tp);
            if (q == 0 && half == 0) TRACE(tcount, 9);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t accA = tp ? acc1 : acc0, accB = tp ? acc0 : acc1;
            epilogue_to_act(smem_u32(act_buf), accA, q, half, lane, H, sbias[0], P.act, smem_u32(&bar_a2[0]), false, P.act_slots,
                            smem_u32(&bar_hfree[0]), 2u * tcount);
            if (q == 0 && half == 0) TRACE(tcount, 10);
            // h2 -> ACT
            mbar_wait(smem_u32(&bar_t2), tp);
            if (q == 0 && half == 0) TRACE(tcount, 11);
            asm volatile("tcgen05.fence::after_thread_sync;");
            epilogue_to_act(smem_u32(act_buf), accB, q, half, lane, H, sbias[1], P.act, smem_u32(&bar_a3[0]), false, P.act_slots,
                            smem_u32(&bar_hfree[0]), 2u * tcount + 1u);
            if (q == 0 && half == 0) TRACE(tcount, 12);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == kProdWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// ================================================================================================
// k_mlp3_tf32_pair: the same chain on a CTA PAIR (thread-block cluster of 2, tcgen05 cta_group::2).
// One 256-row tile per pair: each CTA owns 128 rows (its own ACT, its own epilogues, its own output) and streams only
// HALF of every weight K block (rows [rank * N/2, (rank+1) * N/2) of the packed image -- the 8-row swizzle atoms make
// each half a valid operand image on its own); the tensor cores of both SMs read both halves.  Per-SM weight ingest
// from L2, the single-CTA kernel's bound, halves.  CTA 0 issues every MMA; "operand ready" barriers (input tile,
// act(h) K blocks, output drained, the peer's weight stage) live in CTA 0 and are arrived on from both CTAs,
// "MMA done" barriers (stage free, accumulator ready, ACT free) exist in both CTAs and are arrived on by multicast commits.
// ================================================================================================
__device__ __forceinline__ uint32_t make_idesc_tf32_m256(int N) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= 2u << 7;
    d |= 2u << 10;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(256 >> 4) << 24;                       // M = 256 across the pair
    return d;
}

__device__ __forceinline__ void commit_pair(uint64_t *bar) {      // by one elected lane of the converged MMA warp
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
        "h"((unsigned short)3)
        : "memory");
}

__global__ void __launch_bounds__(kChainThreads, 1) k_mlp3_tf32_pair(const __grid_constant__ Mlp3Params P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *act_buf = smem;
    uint8_t *ring = smem + P.act_bytes;
    float *tiles = reinterpret_cast<float *>(ring + (size_t)P.stages * P.stage_bytes);
    __shared__ __align__(8) uint64_t bar_full[kMaxRing], bar_pfull[kMaxRing], bar_empty[kMaxRing];
    __shared__ __align__(8) uint64_t bar_a1, bar_actfree, bar_t1, bar_t2, bar_outdone;
    __shared__ __align__(8) uint64_t bar_a2[8], bar_a3[8];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[3][256];

    const DenseParams &p = P.in;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = P.D, H = P.H;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    for (int i = tid; i < 3 * 256; i += kChainThreads) {
        const int l = i >> 8, c = i & 255;
        const float *b = l == 0 ? P.b1 : (l == 1 ? P.b2 : P.b3);
        sbias[l][c] = (b && c < (l == 2 ? D : H)) ? b[c] : 0.f;
    }
    const int tiles_m = (P.M + kTileM - 1) / kTileM;
    const int n_pairs = (tiles_m + 1) / 2;                 // an odd tile count leaves the last pair's second CTA on rows >= M
    const int sub1 = mlp3_subs(D), sub2 = mlp3_subs(H);
    const uint32_t S = (uint32_t)P.stages;

    if (warp == kProdWarps) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    if (tid == 0) {
        auto init = [](uint64_t *b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); };
        for (int i = 0; i < kMaxRing; ++i) {
            init(&bar_full[i], 1u);              // own half of the stage: the loader's expect_tx arrive + the bytes
            init(&bar_pfull[i], 1u);             // (CTA 0) the peer's half has landed: relayed by the peer's MMA warp
            init(&bar_empty[i], 1u);             // multicast commit
        }
        init(&bar_a1, 2u * kProdWarps);          // (CTA 0) input tiles of both CTAs: one arrival per producer warp
        init(&bar_outdone, 2u * kProdWarps);     // (CTA 0) output tiles of both CTAs read out
        init(&bar_actfree, 1u);
        init(&bar_t1, 1u);
        init(&bar_t2, 1u);
        for (int i = 0; i < 8; ++i) {
            init(&bar_a2[i], 8u);                // (CTA 0) four epilogue warps per CTA
            init(&bar_a3[i], 8u);
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();                          // the peer's barriers exist before anything arrives on them
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_slot;
    const uint32_t acc0 = tmem_base, acc1 = tmem_base + 256u;

    if (warp < kProdWarps) {
        // ===== input producers (per CTA: its own 128 rows) =====
        float cf[kMaxNK];
#pragma unroll
        for (int j = 0; j < kMaxNK; ++j) cf[j] = 0.f;
        if (p.nk > 0) {
            const float dt = (float)p.st->dt;
#pragma unroll
            for (int j = 0; j < kMaxNK; ++j)
                if (j < p.nk) cf[j] = __fmul_rn(dt, (float)p.coef[j]);
        }
        const bool prefetch = D <= kKChunk && (D & 3) == 0;
        float *scratch = tiles + warp * (32 * 17);
        uint32_t tcount = 0;
        int prev_m0 = -1;
        for (int pair = cluster_id; pair < n_pairs; pair += n_clusters, ++tcount) {
            const int m0 = (2 * pair + (int)rank) * kTileM;
            if (prefetch) {
                float4 v[8];
                load_tile_regs(p, m0, tid, cf, v);
                TRACE(tcount, 0);
                mbar_wait_backoff(smem_u32(&bar_actfree), (tcount & 1u) ^ 1u);
                TRACE(tcount, 1);
                store_tile_regs(act_buf, tid, v);
            } else {
                mbar_wait_backoff(smem_u32(&bar_actfree), (tcount & 1u) ^ 1u);
                for (int kc = 0; kc < D; kc += kKChunk)
                    produce_chunk<kProdThreads>(p, act_buf + (kc / kSub) * (kTileM * 128), nullptr, m0, 0, 0, kc, tid, cf);
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cta0(smem_u32(&bar_a1));
            TRACE(tcount, 2);
            if (prev_m0 >= 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
                output_tile(P, ((tcount - 1u) & 1u) ? acc1 : acc0, warp, lane, prev_m0, scratch, sbias[2]);
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive_cta0(smem_u32(&bar_outdone));
                TRACE(tcount - 1u, 14);
            }
            prev_m0 = m0;
        }
        if (prev_m0 >= 0) {
            mbar_wait_backoff(smem_u32(&bar_actfree), (tcount & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;");
            output_tile(P, ((tcount - 1u) & 1u) ? acc1 : acc0, warp, lane, prev_m0, scratch, sbias[2]);
        }
    } else if (warp == kLoaderWarp) {
        // ===== weight loader: this CTA's half of every K block =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
                const uint8_t *src = P.packed;
                for (int l = 0; l < 3; ++l) {
                    const int Nl = l == 2 ? D : H;
                    const int nsub = l == 0 ? sub1 : sub2;
                    const uint32_t half = (uint32_t)(Nl / 2) * 128u;
                    const int grp = P.stage_bytes / (int)half;
                    for (int c = 0; c < nsub; c += grp, ++it) {
                        const int nb = nsub - c < grp ? nsub - c : grp;
                        const uint32_t s = it % S, ph = (it / S) & 1u;
                        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
                        const uint32_t bar = smem_u32(&bar_full[s]);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)nb * half) : "memory");
                        const uint32_t dst0 = smem_u32(ring + (size_t)s * P.stage_bytes);
                        for (int kb = 0; kb < nb; ++kb)
                            asm volatile(
                                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                    dst0 + (uint32_t)kb * half),
                                "l"(src + (size_t)(c + kb) * Nl * 128 + (size_t)rank * half), "r"(half), "r"(bar)
                                : "memory");
                    }
                    src += (size_t)nsub * Nl * 128;
                }
            }
        }
    } else if (warp == kProdWarps) {
        if (rank != 0) {
            // ===== peer CTA: tell CTA 0 when this CTA's half of a stage has landed =====
            uint32_t it = 0;
            for (int pair = cluster_id; pair < n_pairs; pair += n_clusters) {
                for (int l = 0; l < 3; ++l) {
                    const int Nl = l == 2 ? D : H;
                    const int nsub = l == 0 ? sub1 : sub2;
                    const int grp = P.stage_bytes / ((Nl / 2) * 128);
                    for (int c = 0; c < nsub; c += grp, ++it) {
                        const uint32_t s = it % S, ph = (it / S) & 1u;
                        mbar_wait(smem_u32(&bar_full[s]), ph);
                        if (lane == 0) mbar_arrive_cta0(smem_u32(&bar_pfull[s]));
                        __syncwarp();
                    }
                }
            }
        } else {
            // ===== MMA issuer for the pair =====
            const uint32_t act_lo = desc_lo(smem_u32(act_buf)), ring_lo = desc_lo(smem_u32(ring));
            uint32_t it = 0, tcount = 0;
            for (int pair = cluster_id; pair < n_pairs; pair += n_clusters, ++tcount) {
                const uint32_t tp = tcount & 1u;
                for (int l = 0; l < 3; ++l) {
                    const int Nl = l == 2 ? D : H;
                    const int nsub = l == 0 ? sub1 : sub2;
                    const uint32_t idesc = make_idesc_tf32_m256(Nl);
                    const uint32_t tacc = ((l == 1) != (tp == 1u)) ? acc1 : acc0;
                    if (l == 0) mbar_wait_cluster(smem_u32(&bar_a1), tp);
                    if (l == 1 && tcount > 0) mbar_wait_cluster(smem_u32(&bar_outdone), tp ^ 1u);
                    TRACE(tcount, 3 + l * 2);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t half = (uint32_t)(Nl / 2) * 128u;
                    const int grp = P.stage_bytes / (int)half;
                    for (int c = 0; c < nsub; c += grp, ++it) {
                        const uint32_t s = it % S, ph = (it / S) & 1u;
                        const int nb = nsub - c < grp ? nsub - c : grp;
                        mbar_wait(smem_u32(&bar_full[s]), ph);
                        mbar_wait_cluster(smem_u32(&bar_pfull[s]), ph);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        for (int kb = 0; kb < nb; ++kb) {
                            if (l == 1) mbar_wait_cluster(smem_u32(&bar_a2[c + kb]), tp);
                            if (l == 2) mbar_wait_cluster(smem_u32(&bar_a3[c + kb]), tp);
                            asm volatile("tcgen05.fence::after_thread_sync;");
                            mma_kblock_tf32<2>(tacc, act_lo + (uint32_t)(c + kb) * (uint32_t)((kTileM * 128) >> 4),
                                               ring_lo + s * ((uint32_t)P.stage_bytes >> 4) + (uint32_t)kb * (half >> 4), idesc,
                                               (c + kb) > 0 ? 1u : 0u);
                        }
                        commit_pair(&bar_empty[s]);
                        if (c + nb == nsub) commit_pair(l == 0 ? &bar_t1 : (l == 1 ? &bar_t2 : &bar_actfree));
                    }
                    TRACE(tcount, 4 + l * 2);
                }
            }
        }
    } else {
        // ===== epilogue warps (per CTA: its own 128 rows) =====
        const int q = warp & 3, half = (warp - (kProdWarps + 1)) >> 2;
        uint32_t tcount = 0;
        for (int pair = cluster_id; pair < n_pairs; pair += n_clusters, ++tcount) {
            const uint32_t tp = tcount & 1u;
            mbar_wait(smem_u32(&bar_t1), tp);
            if (q == 0 && half == 0) TRACE(tcount, 9);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t accA = tp ? acc1 : acc0, accB = tp ? acc0 : acc1;
            epilogue_to_act(smem_u32(act_buf), accA, q, half, lane, H, sbias[0], P.act, smem_u32(&bar_a2[0]), true, 8, 0u, 0u);
            if (q == 0 && half == 0) TRACE(tcount, 10);
            mbar_wait(smem_u32(&bar_t2), tp);
            if (q == 0 && half == 0) TRACE(tcount, 11);
            asm volatile("tcgen05.fence::after_thread_sync;");
            epilogue_to_act(smem_u32(act_buf), accB, q, half, lane, H, sbias[1], P.act, smem_u32(&bar_a3[0]), true, 8, 0u, 0u);
            if (q == 0 && half == 0) TRACE(tcount, 12);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();                          // neither CTA may free tensor memory or exit while the other still computes
    if (warp == kProdWarps) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// ================================================================================================
// host side
// ================================================================================================
// cudaFuncSetAttribute is per device and the persistent grids are sized from the SM count: keep both per device ordinal
// (a process may drive several GPUs)
constexpr int kMaxDev = 64;
struct MmaDevCfg {
    bool dense_simple, dense_ws, mlp3;
    int sms;
};
static MmaDevCfg g_mma_dev[kMaxDev];

static int mma_dev(MmaDevCfg **cfg) {
    int dev = 0;
    B2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) return b2_fail(B2ODE_EINVAL, "device ordinal %d out of range", dev);
    MmaDevCfg *c = &g_mma_dev[dev];
    if (c->sms == 0) B2_CUDA(cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, dev));
    *cfg = c;
    return 0;
}

static int dense_layer_impl(const void *x, const void *const *k, const double *coef, int nk, const void *state,
                            void *ystage, const void *W, const void *W_lo, const void *bias, void *out, int64_t M, int K, int N,
                            int act, void *cuda_stream) {
    if (!x || !W || !out || M < 1 || K < 1 || N < 1) return b2_fail(B2ODE_EINVAL, "bad dense-layer arguments");
    if (N % 16 != 0) return b2_fail(B2ODE_EINVAL, "dense layer: N must be a multiple of 16 (got %d)", N);
    if (nk < 0 || nk > kMaxNK || (nk > 0 && (!k || !coef || !state))) return b2_fail(B2ODE_EINVAL, "bad stage-combine arguments");
    if (act < 0 || act > 3) return b2_fail(B2ODE_EINVAL, "unknown activation %d", act);
    if (M > (int64_t)2147483647 - kTileM) return b2_fail(B2ODE_EINVAL, "M too large");
    DenseParams p;
    memset(&p, 0, sizeof(p));
    p.x = (const float *)x;
    for (int j = 0; j < nk; ++j) {
        if (!k[j]) return b2_fail(B2ODE_EINVAL, "k[%d] is null", j);
        p.k[j] = (const float *)k[j];
        p.coef[j] = coef[j];
    }
    p.nk = nk;
    p.st = (const b2ode_state *)state;
    p.ystage = (float *)ystage;
    p.W = (const float *)W;
    p.W_lo = (const float *)W_lo;
    p.bias = (const float *)bias;
    p.out = (float *)out;
    p.M = (int)M;
    p.K = K;
    p.N = N;
    p.act = act;
    MmaDevCfg *dc = nullptr;
    {
        const int rc = mma_dev(&dc);
        if (rc) return rc;
    }
    static int variant = -1;         // B2ODE_DENSE_SIMPLE=1 selects the non-pipelined kernel (kept as a cross-check)
    if (variant < 0) {
        const char *e = getenv("B2ODE_DENSE_SIMPLE");
        variant = (e && e[0] == '1') ? 1 : 0;
    }
    if (variant == 1) {
        const size_t smem = 2 * kTileM * 128 + 2 * 256 * 128 + 1024;      // A + B + alignment slack = 99 328 B
        if (!dc->dense_simple) {
            B2_CUDA(cudaFuncSetAttribute(k_dense_layer_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dc->dense_simple = true;
        }
        const int grid = (int)((M + kTileM - 1) / kTileM);
        k_dense_layer_tf32<<<grid, kMmaThreads, smem, (cudaStream_t)cuda_stream>>>(p);
    } else {
        const size_t smem = (size_t)kStages * kStageBytes + 4 * 32 * 33 * sizeof(float) + 1024;   // 2 x 96 KB ring + epilogue tiles + slack
        if (!dc->dense_ws) {
            B2_CUDA(cudaFuncSetAttribute(k_dense_layer_tf32_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dc->dense_ws = true;
        }
        const long long items = ((M + kTileM - 1) / kTileM) * (long long)((N + 255) / 256);
        const int grid = (int)(items < dc->sms ? items : dc->sms);         // persistent: one CTA per SM
        k_dense_layer_tf32_ws<<<grid, kWsThreads, smem, (cudaStream_t)cuda_stream>>>(p);
    }
    B2_CUDA(cudaGetLastError());
    b2_count_launch();
    return 0;
}

extern "C" int b2ode_dense_layer(const void *x, const void *const *k, const double *coef, int nk, const void *state,
                                 void *ystage, const void *W, const void *bias, void *out, int64_t M, int K, int N, int act,
                                 void *cuda_stream) {
    return dense_layer_impl(x, k, coef, nk, state, ystage, W, nullptr, bias, out, M, K, N, act, cuda_stream);
}

// 3xTF32: W_hi = tf32(W), W_lo = tf32(W - W_hi), both [N, K]; the kernel splits A the same way on the fly and accumulates
// A_lo.W_hi + A_hi.W_lo + A_hi.W_hi in fp32 -- products as accurate as fp32 FMAs at three times the tensor-core work.
extern "C" int b2ode_dense_layer_x3(const void *x, const void *const *k, const double *coef, int nk, const void *state,
                                    void *ystage, const void *W_hi, const void *W_lo, const void *bias, void *out, int64_t M, int K,
                                    int N, int act, void *cuda_stream) {
    if (!W_lo) return b2_fail(B2ODE_EINVAL, "W_lo is null");
    return dense_layer_impl(x, k, coef, nk, state, ystage, W_hi, W_lo, bias, out, M, K, N, act, cuda_stream);
}

// ---- fc1 -> act -> fc2 -> act -> fc3 in one launch (tfdiffeq/models/dense_odenet.py:85-92); see k_mlp3_tf32 ----
static bool mlp3_dims_ok(int D, int H) { return D >= 16 && H >= 16 && D <= 256 && H <= 256 && D % 16 == 0 && H % 16 == 0; }

extern "C" int64_t b2ode_mlp3_packed_bytes(int D, int H) {
    if (!mlp3_dims_ok(D, H)) return -1;
    return (int64_t)(mlp3_layer_bytes(H, D) + mlp3_layer_bytes(H, H) + mlp3_layer_bytes(D, H));
}

extern "C" int b2ode_mlp3_pack(const void *W1, const void *W2, const void *W3, int D, int H, void *packed, void *cuda_stream) {
    if (!W1 || !W2 || !W3 || !packed) return b2_fail(B2ODE_EINVAL, "null pointer");
    if (!mlp3_dims_ok(D, H)) return b2_fail(B2ODE_EINVAL, "mlp3: dim and hidden must be multiples of 16 in [16, 256] (got %d, %d)", D, H);
    if ((uintptr_t)packed & 15) return b2_fail(B2ODE_EINVAL, "packed image must be 16-byte aligned");
    const int64_t pieces = b2ode_mlp3_packed_bytes(D, H) / 16;
    const int grid = (int)((pieces + 255) / 256);
    k_mlp3_pack<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>((const float *)W1, (const float *)W2, (const float *)W3, D, H, (uint8_t *)packed);
    B2_CUDA(cudaGetLastError());
    b2_count_launch();
    return 0;
}

extern "C" int b2ode_mlp3(const void *x, const void *const *k, const double *coef, int nk, const void *state, void *ystage,
                          const void *packed, const void *b1, const void *b2, const void *b3, void *out, int64_t M, int D, int H,
                          int act, void *cuda_stream) {
    if (!x || !packed || !out || M < 1) return b2_fail(B2ODE_EINVAL, "bad mlp3 arguments");
    if (!mlp3_dims_ok(D, H)) return b2_fail(B2ODE_EINVAL, "mlp3: dim and hidden must be multiples of 16 in [16, 256] (got %d, %d)", D, H);
    if ((uintptr_t)packed & 15) return b2_fail(B2ODE_EINVAL, "packed image must be 16-byte aligned");
    if (nk < 0 || nk > kMaxNK || (nk > 0 && (!k || !coef || !state))) return b2_fail(B2ODE_EINVAL, "bad stage-combine arguments");
    if (act < 0 || act > 3) return b2_fail(B2ODE_EINVAL, "unknown activation %d", act);
    if (M > (int64_t)2147483647 - kTileM) return b2_fail(B2ODE_EINVAL, "M too large");
    Mlp3Params P;
    memset(&P, 0, sizeof(P));
    P.in.x = (const float *)x;
    for (int j = 0; j < nk; ++j) {
        if (!k[j]) return b2_fail(B2ODE_EINVAL, "k[%d] is null", j);
        P.in.k[j] = (const float *)k[j];
        P.in.coef[j] = coef[j];
    }
    P.in.nk = nk;
    P.in.st = (const b2ode_state *)state;
    P.in.ystage = (float *)ystage;
    P.in.M = (int)M;
    P.in.K = D;
    P.in.N = H;
    P.packed = (const uint8_t *)packed;
    P.b1 = (const float *)b1;
    P.b2 = (const float *)b2;
    P.b3 = (const float *)b3;
    P.out = (float *)out;
    P.M = (int)M;
    P.D = D;
    P.H = H;
    P.act = act;
    // CTA pairs (cta_group::2) when there are at least two tiles; B2ODE_MLP3_PAIR=0/1 overrides (A/B runs)
    const char *pair_env = getenv("B2ODE_MLP3_PAIR");
    const bool pair = pair_env ? (atoi(pair_env) != 0) : kMlp3PairDefault;
    // shared memory: ACT (input chunks are produced 64 columns = 2 blocks at a time), the weight ring, the output tiles
    const int blocks_in = 2 * ((D + kKChunk - 1) / kKChunk), blocks_h = mlp3_subs(H);
    int slots = blocks_in > blocks_h ? blocks_in : blocks_h;
    // the single-CTA kernel streams act(h) through a ring of R slots (GEMM2 / GEMM3 trail the epilogues by one K block)
    // so that the weight ring can be deeper: a stage refill takes ~1.3 us, the MMAs of a K block ~0.4 us
    const char *slots_env = getenv("B2ODE_MLP3_ACT_SLOTS");
    int want = slots_env ? atoi(slots_env) : kMlp3ActSlots;
    if (!pair && want >= 2 && want < slots && want >= blocks_in) slots = want;
    P.act_slots = pair ? 8 : slots;
    P.act_bytes = slots * (kTileM * 128);
    P.stage_bytes = (H > D ? H : D) * 128 / (pair ? 2 : 1);
    const int budget = 227 * 1024 - 4096 - 1024 - P.act_bytes - kTileScratchBytes;      // static (biases, barriers) + alignment slack
    int stages = budget / P.stage_bytes;
    if (stages > kMaxRing) stages = kMaxRing;
    if (stages < 2) return b2_fail(B2ODE_EINVAL, "mlp3: shared memory budget exhausted");
    P.stages = stages;
    const size_t smem = (size_t)P.act_bytes + (size_t)stages * P.stage_bytes + kTileScratchBytes + 1024;
    MmaDevCfg *dc = nullptr;
    {
        const int rc = mma_dev(&dc);
        if (rc) return rc;
    }
    if (!dc->mlp3) {
        B2_CUDA(cudaFuncSetAttribute(k_mlp3_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
        B2_CUDA(cudaFuncSetAttribute(k_mlp3_tf32_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
        dc->mlp3 = true;
    }
    const int sms = dc->sms;
    const long long tiles = (M + kTileM - 1) / kTileM;
    if (pair) {
        const long long pairs = (tiles + 1) / 2;
        const int clusters = (int)(pairs < sms / 2 ? pairs : sms / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters, 1, 1);
        cfg.blockDim = dim3(kChainThreads, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)cuda_stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        B2_CUDA(cudaLaunchKernelEx(&cfg, k_mlp3_tf32_pair, P));
    } else {
        const int grid = (int)(tiles < sms ? tiles : sms);
        k_mlp3_tf32<<<grid, kChainThreads, smem, (cudaStream_t)cuda_stream>>>(P);
    }
    B2_CUDA(cudaGetLastError());
    b2_count_launch();
    return 0;
}

#ifdef B2ODE_TRACE
extern "C" int b2ode_debug_mlp3_trace(unsigned long long *out) {
    B2_CUDA(cudaMemcpyFromSymbol(out, g_mlp3_trace, sizeof(unsigned long long) * 256));
    B2_CUDA(cudaMemcpyFromSymbol(out + 256, g_mlp3_clk, sizeof(unsigned long long) * 80));
    int zero = 0;
    B2_CUDA(cudaMemcpyToSymbol(g_mlp3_clk_n, &zero, sizeof(int)));
    return 0;
}
#endif
