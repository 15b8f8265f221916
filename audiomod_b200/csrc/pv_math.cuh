// Exactly-rounded float helpers and a bit-exact atan2f for the analysis stage.
//
// Why: every discrete decision of the phase vocoder (strict-> peak picking, nearest-peak
// linking, princarg wraps, lrint source bins) is taken on the forward-FFT magnitude and
// on atan2f of the forward-FFT bins (reference: src/common/dsp/FFT.cc:2617-2631 feeding
// src/phasevocoder/phasevocoderprocess.cc:574-706).  The reference build has no FMA
// contraction and calls glibc 2.39's float atan2f, so the device must (a) never fuse
// a*b+c on that path and (b) reproduce that atan2f bit for bit.
//
// The atan2f below restates the classic fdlibm single-precision algorithm that glibc
// 2.39 ships (sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c -- third-party, not in
// /root/reference): argument reduction at 7/16, 11/16, 19/16, 39/16, an 11-term
// odd/even polynomial in x^2 and a hi/lo table correction.  It is compiled for the host
// as well (tests/host_atan2f_check.cc) and compared bit-for-bit with this image's libm
// over >1e9 inputs plus all special cases.
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define PV_HD __host__ __device__ __forceinline__
#define PV_MUL(a, b) __fmul_rn((a), (b))
#define PV_ADD(a, b) __fadd_rn((a), (b))
#define PV_SUB(a, b) __fsub_rn((a), (b))
#define PV_DIV(a, b) __fdiv_rn((a), (b))
#define PV_F2I(x) __float_as_int(x)
#define PV_I2F(x) __int_as_float(x)
#else
#if defined(__CUDACC__)
#define PV_HD __host__ __device__ inline
#else
#define PV_HD static inline
#endif
// host build: compile with -ffp-contract=off so these stay separate roundings
#define PV_MUL(a, b) ((float)((float)(a) * (float)(b)))
#define PV_ADD(a, b) ((float)((float)(a) + (float)(b)))
#define PV_SUB(a, b) ((float)((float)(a) - (float)(b)))
#define PV_DIV(a, b) ((float)((float)(a) / (float)(b)))
PV_HD int32_t pv_f2i_(float x) { union { float f; int32_t i; } u; u.f = x; return u.i; }
PV_HD float pv_i2f_(int32_t x) { union { float f; int32_t i; } u; u.i = x; return u.f; }
#define PV_F2I(x) pv_f2i_(x)
#define PV_I2F(x) pv_i2f_(x)
#endif

// atanf on the reduced argument; hx carries the sign of the original argument.
PV_HD float pv_atanf(float x) {
    const float hi0 = PV_I2F(0x3eed6338), hi1 = PV_I2F(0x3f490fda), hi2 = PV_I2F(0x3f7b985e), hi3 = PV_I2F(0x3fc90fda);
    const float lo0 = PV_I2F(0x31ac3769), lo1 = PV_I2F(0x33222168), lo2 = PV_I2F(0x33140fb4), lo3 = PV_I2F(0x33a22168);
    const float a0 = PV_I2F(0x3eaaaaab), a1 = PV_I2F(0xbe4ccccd), a2 = PV_I2F(0x3e124925), a3 = PV_I2F(0xbde38e38),
                a4 = PV_I2F(0x3dba2e6e), a5 = PV_I2F(0xbd9d8795), a6 = PV_I2F(0x3d886b35), a7 = PV_I2F(0xbd6ef16b),
                a8 = PV_I2F(0x3d4bda59), a9 = PV_I2F(0xbd15a221), a10 = PV_I2F(0x3c8569d7);
    const int32_t hx = PV_F2I(x);
    const int32_t ix = hx & 0x7fffffff;
    int id;
    float hi = 0.f, lo = 0.f;
    if (ix >= 0x4c000000) { // |x| >= 2^25
        if (ix > 0x7f800000) return PV_ADD(x, x); // NaN
        const float r = PV_ADD(hi3, lo3);
        return hx > 0 ? r : -r;
    }
    if (ix < 0x3ee00000) {      // |x| < 7/16
        if (ix < 0x31000000) return x; // |x| < 2^-29
        id = -1;
    } else {
        x = PV_I2F(ix); // fabsf
        if (ix < 0x3f980000) {  // |x| < 19/16
            if (ix < 0x3f300000) { // 7/16 <= |x| < 11/16
                id = 0; hi = hi0; lo = lo0;
                x = PV_DIV(PV_SUB(PV_MUL(2.0f, x), 1.0f), PV_ADD(2.0f, x));
            } else {               // 11/16 <= |x| < 19/16
                id = 1; hi = hi1; lo = lo1;
                x = PV_DIV(PV_SUB(x, 1.0f), PV_ADD(x, 1.0f));
            }
        } else {
            if (ix < 0x401c0000) { // |x| < 39/16
                id = 2; hi = hi2; lo = lo2;
                x = PV_DIV(PV_SUB(x, 1.5f), PV_ADD(1.0f, PV_MUL(1.5f, x)));
            } else {               // 39/16 <= |x| < 2^25
                id = 3; hi = hi3; lo = lo3;
                x = PV_DIV(-1.0f, x);
            }
        }
    }
    const float z = PV_MUL(x, x);
    const float w = PV_MUL(z, z);
    const float s1 = PV_MUL(z, PV_ADD(a0, PV_MUL(w, PV_ADD(a2, PV_MUL(w, PV_ADD(a4, PV_MUL(w, PV_ADD(a6, PV_MUL(w, PV_ADD(a8, PV_MUL(w, a10)))))))))));
    const float s2 = PV_MUL(w, PV_ADD(a1, PV_MUL(w, PV_ADD(a3, PV_MUL(w, PV_ADD(a5, PV_MUL(w, PV_ADD(a7, PV_MUL(w, a9)))))))));
    if (id < 0) return PV_SUB(x, PV_MUL(x, PV_ADD(s1, s2)));
    const float r = PV_SUB(hi, PV_SUB(PV_SUB(PV_MUL(x, PV_ADD(s1, s2)), lo), x));
    return hx < 0 ? -r : r;
}

#if defined(__CUDA_ARCH__)
__device__ __noinline__ float pv_atan2f_rare(float y, float x);
#endif
PV_HD float pv_atan2f(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi_o_4 = PV_I2F(0x3f490fdb), pi_o_2 = PV_I2F(0x3fc90fdb), pi = PV_I2F(0x40490fdb), pi_lo = PV_I2F(0xb3bbbd2e);
    const int32_t hx = PV_F2I(x), hy = PV_F2I(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return PV_ADD(x, y); // NaN
    if (hx == 0x3f800000) return pv_atanf(y);                    // x == 1
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);           // 2*sign(x) + sign(y)
    if (iy == 0) {                                               // y == 0
        switch (m) {
            case 0: case 1: return y;
            case 2: return PV_ADD(pi, tiny);
            default: return PV_SUB(-pi, tiny);
        }
    }
    if (ix == 0) return hy < 0 ? PV_SUB(-pi_o_2, tiny) : PV_ADD(pi_o_2, tiny);
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
                case 0: return PV_ADD(pi_o_4, tiny);
                case 1: return PV_SUB(-pi_o_4, tiny);
                case 2: return PV_ADD(PV_MUL(3.0f, pi_o_4), tiny);
                default: return PV_SUB(PV_MUL(-3.0f, pi_o_4), tiny);
            }
        } else {
            switch (m) {
                case 0: return 0.0f;
                case 1: return -0.0f;
                case 2: return PV_ADD(pi, tiny);
                default: return PV_SUB(-pi, tiny);
            }
        }
    }
    if (iy == 0x7f800000) return hy < 0 ? PV_SUB(-pi_o_2, tiny) : PV_ADD(pi_o_2, tiny);
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = PV_ADD(pi_o_2, PV_MUL(0.5f, pi_lo));  // |y/x| > 2^60
    else if (hx < 0 && k < -60) z = 0.0f;                 // |y|/x < -2^60
    else z = pv_atanf(PV_I2F(PV_F2I(PV_DIV(y, x)) & 0x7fffffff));
    switch (m) {
        case 0: return z;
        case 1: return PV_I2F(PV_F2I(z) ^ (int32_t)0x80000000);
        case 2: return PV_SUB(pi, PV_SUB(z, pi_lo));
        default: return PV_SUB(PV_SUB(z, pi_lo), pi);
    }
}

// Same function, arranged for SIMT: the common case (finite, non-zero arguments, x != 1, exponents within 2^60 of each
// other, |y/x| < 2^25) runs without data-dependent branches -- the range reduction is written as
// (a*z + b) / (c*z + d) with per-range constants, which performs exactly the reference's operations
// (2z-1)/(2+z), (z-1)/(z+1), (z-1.5)/(1+1.5z), -1/z because multiplying by 0, 1 or 2 and adding 0 are exact.
// Everything else falls through to pv_atan2f above.
PV_HD float pv_atan2f_fast(float y, float x) {
    const int32_t hx = PV_F2I(x), hy = PV_F2I(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    const int k = (iy - ix) >> 23;
    // iy-1 / ix-1 as unsigned: rejects zero and inf/NaN with one compare each
    bool common = ((uint32_t)(iy - 1) < 0x7f7fffffu) & ((uint32_t)(ix - 1) < 0x7f7fffffu) & (hx != 0x3f800000) & (k <= 60) & (k >= -60);
    const float q = PV_I2F(PV_F2I(PV_DIV(y, x)) & 0x7fffffff);  // fabsf(y/x); harmless when the arguments are not "common"
    const int32_t iq = PV_F2I(q);
    common &= iq < 0x4c000000;                                  // |y/x| >= 2^25 (or the quotient overflowed) goes the long way
    const bool r1 = iq >= 0x3f300000, r2 = iq >= 0x3f980000, r3 = iq >= 0x401c0000;   // 11/16, 19/16, 39/16
    const float a = r3 ? 0.0f : (r1 ? 1.0f : 2.0f);
    const float b = r2 && !r3 ? -1.5f : -1.0f;
    const float c = r2 && !r3 ? 1.5f : 1.0f;
    const float d = r3 ? 0.0f : (r1 ? 1.0f : 2.0f);
    const float hi = PV_I2F(r3 ? 0x3fc90fda : (r2 ? 0x3f7b985e : (r1 ? 0x3f490fda : 0x3eed6338)));
    const float lo = PV_I2F(r3 ? 0x33a22168 : (r2 ? 0x33140fb4 : (r1 ? 0x33222168 : 0x31ac3769)));
    const bool small = iq < 0x3ee00000;  // < 7/16: no reduction
    const float red = PV_DIV(PV_ADD(PV_MUL(a, q), b), PV_ADD(PV_MUL(c, q), d));
    const float t = small ? q : red;
    const float a0 = PV_I2F(0x3eaaaaab), a1 = PV_I2F(0xbe4ccccd), a2 = PV_I2F(0x3e124925), a3 = PV_I2F(0xbde38e38),
                a4 = PV_I2F(0x3dba2e6e), a5 = PV_I2F(0xbd9d8795), a6 = PV_I2F(0x3d886b35), a7 = PV_I2F(0xbd6ef16b),
                a8 = PV_I2F(0x3d4bda59), a9 = PV_I2F(0xbd15a221), a10 = PV_I2F(0x3c8569d7);
    const float z = PV_MUL(t, t);
    const float w = PV_MUL(z, z);
    const float s1 = PV_MUL(z, PV_ADD(a0, PV_MUL(w, PV_ADD(a2, PV_MUL(w, PV_ADD(a4, PV_MUL(w, PV_ADD(a6, PV_MUL(w, PV_ADD(a8, PV_MUL(w, a10)))))))))));
    const float s2 = PV_MUL(w, PV_ADD(a1, PV_MUL(w, PV_ADD(a3, PV_MUL(w, PV_ADD(a5, PV_MUL(w, PV_ADD(a7, PV_MUL(w, a9)))))))));
    const float ts = PV_MUL(t, PV_ADD(s1, s2));
    // |t| < 2^-29 returns t unchanged in the reference; t - t*(s1+s2) rounds to t there as well (t is non-zero here)
    const float zs = iq < 0x31000000 ? t : PV_SUB(t, ts);
    const float zb = PV_SUB(hi, PV_SUB(PV_SUB(ts, lo), t));
    const float zr = small ? zs : zb;
    const float pi = PV_I2F(0x40490fdb), pi_lo = PV_I2F(0xb3bbbd2e);
    // quadrant: x > 0: +-zr;  x < 0: +-(pi - (zr - pi_lo))   [(zr - pi_lo) - pi == -(pi - (zr - pi_lo)) exactly]
    const float v = hx >= 0 ? zr : PV_SUB(pi, PV_SUB(zr, pi_lo));
    const float res = PV_I2F(PV_F2I(v) ^ (hy & (int32_t)0x80000000));
#if defined(__CUDA_ARCH__)
    return common ? res : pv_atan2f_rare(y, x);
#else
    return common ? res : pv_atan2f(y, x);
#endif
}
