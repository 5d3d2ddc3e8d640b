// Host-side check of the tagged 16-byte partial (tfdiffeq_b200/csrc/b2ode_pay16.cuh): the same functions the persistent kernel
// is compiled from, run on the CPU.   nvcc -std=c++17 -I tfdiffeq_b200/csrc -o pay16_check tests/host/pay16_check.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>

#include "b2ode_pay16.cuh"

static int fails = 0;
#define CHECK(c)                                                      \
    do {                                                              \
        if (!(c)) {                                                   \
            ++fails;                                                  \
            if (fails < 20) printf("FAILED line %d: %s\n", __LINE__, #c); \
        }                                                             \
    } while (0)

int main() {
    unsigned long long lcg = 12345;
    auto rnd = [&]() {
        lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
        return lcg;
    };
    for (unsigned seq = 0; seq < 5000; ++seq) {
        const unsigned tag = pay_tag(seq);
        CHECK(tag >= 1 && tag <= 15);
        CHECK(pay_tag(seq + 2) != tag);                       // the buffer's previous occupant
        CHECK(pay_tag(seq + 1) != tag);
        Pay x;
        x.a = std::ldexp((double)(rnd() >> 11), (int)(rnd() % 200) - 150);      // a non-negative sum
        const double m = std::ldexp((double)(rnd() >> 11), (int)(rnd() % 200) - 150);
        x.b = pay_bits(m);
        x.flag = (unsigned)(rnd() & 1);
        unsigned long long w0, w1;
        pay_pack16(x, seq, w0, w1);
        CHECK(pay_valid16(w0, w1, seq));
        CHECK(pay_mismatch(w0, w1, tag) == 0u);
        for (unsigned d = 1; d < 15; ++d) CHECK(!pay_valid16(w0, w1, seq + d));
        CHECK(pay_valid16(w0, w1, seq + 15));                 // the tag comes round after 15 exchanges, never after 2
        const Pay y = pay_unpack16(w0, w1);
        CHECK(y.flag == x.flag);
        CHECK(y.a <= x.a && x.a - y.a <= x.a * std::ldexp(1.0, -48));            // 4 mantissa bits: 2^-48 relative, toward zero
        CHECK(pay_double(y.b) <= m && m - pay_double(y.b) <= m * std::ldexp(1.0, -48));
        // neither cleared memory nor the poison pattern ever validates
        CHECK(!pay_valid16(0ull, 0ull, seq));
        CHECK(!pay_valid16(kPayPoisonW0, kPayPoisonW1, seq));
        // a torn message (one fresh word, one stale word) does not validate
        unsigned long long s0, s1;
        pay_pack16(x, seq + 2, s0, s1);
        CHECK(!pay_valid16(w0, s1, seq) && !pay_valid16(s0, w1, seq));
    }
    Pay n;
    n.a = std::numeric_limits<double>::quiet_NaN();
    n.b = 0;
    n.flag = 1;
    unsigned long long w0, w1;
    pay_pack16(n, 7, w0, w1);
    CHECK(pay_valid16(w0, w1, 7));
    const Pay y = pay_unpack16(w0, w1);
    CHECK(y.a != y.a && y.flag == 1);                          // NaN survives, canonical, with the flag in the sign bit
    Pay z;
    z.a = 0.0;
    z.b = 0;
    z.flag = 0;
    pay_pack16(z, 14, w0, w1);                                 // tag(14) = 15; tag(15) = 1: an all-zero payload is still tagged
    CHECK(w0 != 0 && w1 != 0 && pay_valid16(w0, w1, 14));
    printf(fails ? "pay16: %d checks FAILED\n" : "pay16: ok\n", fails);
    return fails ? 1 : 0;
}
