// Host-side bit-comparison of audiomod_b200/csrc/pv_math.cuh's atan2f restatement with the
// libm atan2f of this image (glibc 2.39, the function the reference calls at
// src/common/dsp/FFT.cc:2629).  Build: g++ -O2 -ffp-contract=off -static (see tests/test_atan2f_host.py).
// usage: host_atan2f_check <n_random_pairs> [seed]   -> prints mismatches, exit 1 if any
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../audiomod_b200/csrc/pv_math.cuh"

static uint64_t s[2];
static inline uint64_t rng() { // xorshift128+
    uint64_t x = s[0], y = s[1];
    s[0] = y; x ^= x << 23; s[1] = x ^ y ^ (x >> 17) ^ (y >> 26);
    return s[1] + y;
}
static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float fromb(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static long bad = 0, total = 0;
static inline void check(float y, float x) {
    float a = pv_atan2f(y, x), b = atan2f(y, x), f = pv_atan2f_fast(y, x);
    if (bits(f) != bits(a) && !(f != f && a != a)) { if (bad < 20) printf("FAST MISMATCH y=%a x=%a fast=%a ref=%a\n", y, x, f, a); ++bad; }
    ++total;
    if (bits(a) != bits(b) && !(a != a && b != b)) {
        if (bad < 20) printf("MISMATCH y=%a x=%a mine=%a (%08x) libm=%a (%08x)\n", y, x, a, bits(a), b, bits(b));
        ++bad;
    }
}
int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 10000000;
    s[0] = 0x9E3779B97F4A7C15ull ^ (argc > 2 ? strtoull(argv[2], 0, 10) : 1); s[1] = 0xD1B54A32D192ED03ull;
    // special values, every pairing
    const uint32_t sp[] = {0x00000000, 0x80000000, 0x00000001, 0x80000001, 0x007fffff, 0x00800000, 0x3f800000, 0xbf800000,
                           0x3f000000, 0x3ee00000, 0x3edfffff, 0x3f300000, 0x3f2fffff, 0x3f980000, 0x3f97ffff, 0x401c0000,
                           0x401bffff, 0x4c800000, 0x4c7fffff, 0x31000000, 0x30ffffff, 0x7f7fffff, 0xff7fffff, 0x7f800000,
                           0xff800000, 0x7fc00000, 0x40490fdb, 0x5f800000, 0x1f800000, 0x42c80000, 0xc2c80000};
    const int ns = sizeof(sp) / sizeof(sp[0]);
    for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) check(fromb(sp[i]), fromb(sp[j]));
    // fully random bit patterns
    for (long i = 0; i < n / 4; ++i) { uint64_t r = rng(); check(fromb((uint32_t)r), fromb((uint32_t)(r >> 32))); }
    // audio-like: both in a moderate exponent range, random signs/mantissas (what FFT bins look like)
    for (long i = 0; i < n / 2; ++i) {
        uint64_t r = rng(), q = rng();
        uint32_t ey = 100 + (q % 40), ex = 100 + ((q >> 8) % 40);
        uint32_t by = ((uint32_t)r & 0x807fffff) | (ey << 23), bx = ((uint32_t)(r >> 32) & 0x807fffff) | (ex << 23);
        check(fromb(by), fromb(bx));
    }
    // ratios near the reduction boundaries
    const float edges[] = {0.4375f, 0.6875f, 1.1875f, 2.4375f, 1.0f, 0.5f, 1.5f};
    for (long i = 0; i < n / 4; ++i) {
        uint64_t r = rng();
        float x = fromb(((uint32_t)r & 0x807fffff) | ((100 + (r >> 40) % 50) << 23));
        float e = edges[(r >> 50) % 7];
        float y = x * e;
        int d = (int)((r >> 54) % 9) - 4;
        y = fromb(bits(y) + d);
        check(y, x);
        check(x, y);
    }
    printf("checked %ld pairs, %ld mismatches\n", total, bad);
    return bad ? 1 : 0;
}
