// The 16-byte tagged partial of the persistent kernel's per-attempt exchange (b2ode_fused.cu), in a header of its own so that
// the host-side check (tests/host/pay16_check.cu, compiled by nvcc and run on the CPU) exercises the very functions the
// kernel is built from.
#pragma once
#include <cstring>

#define B2_HD __host__ __device__ __forceinline__

struct Pay {
    double a;                 // sum of squares (MODE 0: sum err^2; MODE 1: sum of the initial-step norms)
    unsigned long long b;     // bit pattern of a non-negative double maximum (compared as an integer)
    unsigned flag;            // non-finite somewhere
};

B2_HD unsigned long long pay_bits(double x) {
#ifdef __CUDA_ARCH__
    return (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long u;
    memcpy(&u, &x, 8);
    return u;
#endif
}
B2_HD double pay_double(unsigned long long u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}

// Message: {a | tag, b | tag}.  The 4-bit tag replaces the four lowest mantissa bits of both words (2^-48 relative: below the
// rounding noise of the sums it carries) so that ONE 16-byte load both fetches and validates a partial; the flag rides in the
// sign bit of `a` (a sum of squares; a NaN is made canonical first).
// tag(seq) = seq mod 15 + 1, never 0: zero-filled memory (the intra-GPU slot array is cleared before every launch) is never
// valid, nor is the poison pattern {tag 0, tag 1} of a mailbox slot that holds no partial.  A buffer is reused every second
// exchange and tag(seq - 2) != tag(seq), so a stale message never validates either.
B2_HD unsigned pay_tag(unsigned seq) { return seq % 15u + 1u; }

B2_HD void pay_pack16(const Pay &x, unsigned seq, unsigned long long &w0, unsigned long long &w1) {
    unsigned long long ab = pay_bits(x.a);
    if (x.a != x.a) ab = 0x7ff8000000000000ull;
    const unsigned long long tag = (unsigned long long)pay_tag(seq);
    w0 = (ab & 0x7ffffffffffffff0ull) | ((unsigned long long)(x.flag & 1u) << 63) | tag;
    w1 = (x.b & ~0xfull) | tag;
}
// 0 iff both words carry `tag` (the low words are enough: the polling loop's two-instruction test)
B2_HD unsigned pay_mismatch(unsigned long long w0, unsigned long long w1, unsigned tag) {
    return (((unsigned)w0 ^ tag) | ((unsigned)w1 ^ tag)) & 15u;
}
B2_HD bool pay_valid16(unsigned long long w0, unsigned long long w1, unsigned seq) { return pay_mismatch(w0, w1, pay_tag(seq)) == 0u; }
B2_HD Pay pay_unpack16(unsigned long long w0, unsigned long long w1) {
    Pay r;
    r.flag = (unsigned)(w0 >> 63);
    r.a = pay_double(w0 & 0x7ffffffffffffff0ull);
    r.b = w1 & ~0xfull;
    return r;
}

// what a mailbox slot without a partial holds (Mailbox::fused_part in b2ode_dev.cuh uses the same two words)
constexpr unsigned long long kPayPoisonW0 = 0ull, kPayPoisonW1 = 1ull;
