// Register-tiled complex FFT in KissFFT's exact butterfly order (src/common/kissfft/kiss_fft.c:36-104,250-286 of
// the reference), for the power-of-two sizes the phase vocoder uses (complex length NC = N/2 = 256..4096).
//
// KissFFT's recursion for NC = 4^a * 2^b (b in {0,1}) is: digit-reverse the input (kf_work, :250-286), then run the
// butterfly stages innermost first -- the radix-2 stage (if any) with span 1, then radix-4 stages with spans
// 2^b, 4*2^b, ...  Butterflies of one stage are independent, so only the arithmetic inside a butterfly is order
// sensitive; that arithmetic (bfly2 / bfly4 below) is the reference's, with __f*_rn intrinsics so nothing is fused.
//
// Mapping: T = NC/16 threads own one frame; every thread keeps 16 complex values in registers and runs TWO stages per
// pass (a 16-point or 8-point closed set), exchanging through one padded shared-memory buffer between passes:
//   NC = 256 : [4,4 | span 1] [4,4 | span 16]
//   NC = 512 : [2,4 | span 1] [4,4 | span 8]  [4 | span 128]
//   NC = 1024: [4,4 | span 1] [4,4 | span 16] [4 | span 256]
//   NC = 2048: [2,4 | span 1] [4,4 | span 8]  [4,4 | span 128]
//   NC = 4096: [4,4 | span 1] [4,4 | span 16] [4,4 | span 256]
// Buffer layout: element i lives at pad(i) = i + (i >> 4) + (i >> 8) (float2 units), which makes the 16-contiguous
// accesses of the first pass, the strided accesses of the later passes and the digit-reversed staging writes
// bank-conflict free (or 2-way at worst) for 64-bit shared-memory accesses.
#pragma once
#include <cuda_runtime.h>

namespace pvgpu {

__device__ __forceinline__ float2 cmul_rn(float2 a, float2 b) {  // C_MUL, _kiss_fft_guts.h:87-89
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cadd_rn(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 csub_rn(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }

template <bool kInverse>
__device__ __forceinline__ void bfly4(float2 &f0, float2 &f1, float2 &f2, float2 &f3, float2 t1, float2 t2, float2 t3) {  // kf_bfly4
    const float2 s0 = cmul_rn(f1, t1), s1 = cmul_rn(f2, t2), s2 = cmul_rn(f3, t3);
    const float2 s5 = csub_rn(f0, s1);
    f0 = cadd_rn(f0, s1);
    const float2 s3 = cadd_rn(s0, s2), s4 = csub_rn(s0, s2);
    f2 = csub_rn(f0, s3);
    f0 = cadd_rn(f0, s3);
    if (kInverse) {
        f1 = make_float2(__fsub_rn(s5.x, s4.y), __fadd_rn(s5.y, s4.x));
        f3 = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));
    } else {
        f1 = make_float2(__fadd_rn(s5.x, s4.y), __fsub_rn(s5.y, s4.x));
        f3 = make_float2(__fsub_rn(s5.x, s4.y), __fadd_rn(s5.y, s4.x));
    }
}

__device__ __forceinline__ void bfly2(float2 &f0, float2 &f1, float2 t) {  // kf_bfly2
    const float2 s = cmul_rn(f1, t);
    f1 = csub_rn(f0, s);
    f0 = cadd_rn(f0, s);
}

template <int NC> struct FftShape {
    static_assert(NC == 256 || NC == 512 || NC == 1024 || NC == 2048 || NC == 4096, "unsupported size");
    static constexpr bool kHasRadix2 = (NC == 512 || NC == 2048);
    static constexpr int kThreads = NC / 16;                       // threads per frame
    static constexpr int kMid = kHasRadix2 ? 8 : 16;               // span of the second pass
    static constexpr int kLast = kMid * 16;                        // span of the third pass (if any)
    static constexpr bool kThirdIsPair = (NC / kLast) == 16;       // third pass is [4,4]
    static constexpr bool kThirdIsSingle = (NC / kLast) == 4;      // third pass is [4]
    static constexpr int kPadded = NC + NC / 16 + NC / 256 + 2;    // float2 slots of the exchange buffer
};

__device__ __forceinline__ int fft_pad(int i) { return i + (i >> 4) + (i >> 8); }

// slot (position after KissFFT's input permutation) of complex input index c
template <int NC> __device__ __forceinline__ int fft_slot_of_input(int c) {
    constexpr bool r2 = FftShape<NC>::kHasRadix2;
    constexpr int bits4 = r2 ? (31 - __builtin_clz(NC)) - 1 : (31 - __builtin_clz(NC));  // bits covered by radix-4 digits
    unsigned lo = (unsigned)c & ((1u << bits4) - 1);
    unsigned r = __brev(lo) >> (32 - bits4);                           // bit reversal ...
    r = ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);           // ... with the bits of each base-4 digit swapped back
    return r2 ? (int)(r * 2 + ((unsigned)c >> bits4)) : (int)r;
}

// ---- passes: v[16] are this thread's values; buf is the frame's padded exchange buffer -------------------------------
// First pass, 16 contiguous slots [16*t, 16*t+16): stages (radix 4, span 1) + (radix 4, span 4), or for sizes with a
// radix-2 factor two sets of 8 slots: (radix 2, span 1) + (radix 4, span 2).
template <int NC, bool kInverse>
__device__ __forceinline__ void fft_first_pass(float2 (&v)[16], const float2 *__restrict__ tw) {
    if (!FftShape<NC>::kHasRadix2) {
        const float2 w0 = __ldg(&tw[0]);
#pragma unroll
        for (int g = 0; g < 4; ++g) bfly4<kInverse>(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3], w0, w0, w0);
        constexpr int F = NC / 16;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            bfly4<kInverse>(v[k], v[4 + k], v[8 + k], v[12 + k], __ldg(&tw[k * F]), __ldg(&tw[2 * k * F]), __ldg(&tw[3 * k * F]));
    } else {
        const float2 w0 = __ldg(&tw[0]);
#pragma unroll
        for (int g = 0; g < 8; ++g) bfly2(v[2 * g], v[2 * g + 1], w0);
        constexpr int F = NC / 8;
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int k = 0; k < 2; ++k)
                bfly4<kInverse>(v[8 * u + k], v[8 * u + 2 + k], v[8 * u + 4 + k], v[8 * u + 6 + k], __ldg(&tw[k * F]), __ldg(&tw[2 * k * F]),
                                __ldg(&tw[3 * k * F]));
    }
}

// Pair pass: stages (radix 4, span M) + (radix 4, span 4M) on the set {base + k + j*M, j < 16}; v[j] holds element j.
// The twiddles of the later passes are read from per-pass tables laid out [entry][k] (fft_pass_table on the host builds them
// from the KissFFT table, same values): the lanes of a warp have consecutive k, so every load is one contiguous line instead
// of a gather with a 128-byte stride through the original table (which made the L1 data pipe the limiter of both FFT kernels).
//   pair pass, span M:  entries 0..2 = tw[{1,2,3} * k * Fa], entries 3 + 3*j1 + {0,1,2} = tw[{1,2,3} * (k + j1*M) * Fb]
template <int NC, int M, bool kInverse>
__device__ __forceinline__ void fft_pair_pass(float2 (&v)[16], int k, const float2 *__restrict__ twp) {
    {
        const float2 t1 = __ldg(&twp[k]), t2 = __ldg(&twp[M + k]), t3 = __ldg(&twp[2 * M + k]);
#pragma unroll
        for (int g = 0; g < 4; ++g) bfly4<kInverse>(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3], t1, t2, t3);
    }
#pragma unroll
    for (int j1 = 0; j1 < 4; ++j1)
        bfly4<kInverse>(v[j1], v[4 + j1], v[8 + j1], v[12 + j1], __ldg(&twp[(3 + 3 * j1) * M + k]), __ldg(&twp[(4 + 3 * j1) * M + k]),
                        __ldg(&twp[(5 + 3 * j1) * M + k]));
}

// Single last stage (radix 4, span M = NC/4): four butterflies per thread, v[4q + j] = element (t + T*q) + j*M.
//   single pass: entries 3*q + {0,1,2} = tw[{1,2,3} * (t + T*q)], table laid out [entry][t]
template <int NC, bool kInverse>
__device__ __forceinline__ void fft_single_pass(float2 (&v)[16], int t, const float2 *__restrict__ twp) {
    constexpr int T = FftShape<NC>::kThreads;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        bfly4<kInverse>(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3], __ldg(&twp[(3 * q) * T + t]), __ldg(&twp[(3 * q + 1) * T + t]),
                        __ldg(&twp[(3 * q + 2) * T + t]));
}

// Barrier over the threads of one frame.  Frames owned by less than / exactly one warp use __syncwarp; larger groups a
// named barrier (id 1..15, all threads of the group must call it).
template <int T> __device__ __forceinline__ void frame_sync(int group) {
    if (T <= 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(T) : "memory");
    }
}

// fft_pad is additive over operands with disjoint bits ((a + b) >> s == (a >> s) + (b >> s) when a + b has no carries), and
// every access pattern below is "thread part + compile-time part" with disjoint bits, so each shared-memory access is one
// per-thread base plus an immediate offset.

// natural-order index of v[i] after fft_frame, split into its thread part and its compile-time part
template <int NC> __device__ __forceinline__ int fft_out_base(int t) {
    using S = FftShape<NC>;
    constexpr int M = (NC == 256) ? S::kMid : S::kLast;
    if (NC == 256 || S::kThirdIsPair) return (t / M) * (16 * M) + (t & (M - 1));
    return t;
}
template <int NC> __device__ __forceinline__ constexpr int fft_out_const(int i) {
    using S = FftShape<NC>;
    constexpr int M = (NC == 256) ? S::kMid : S::kLast;
    if (NC == 256 || S::kThirdIsPair) return i * M;
    return S::kThreads * (i >> 2) + (i & 3) * M;
}
template <int NC> __device__ __forceinline__ int fft_out_index(int t, int i) { return fft_out_base<NC>(t) + fft_out_const<NC>(i); }

// The whole NC-point FFT for one frame.  On entry buf holds the input in slot order (slot = fft_slot_of_input(c)) at
// padded positions and the frame's threads have synchronised; on exit v[] holds this thread's outputs and
// fft_out_index() tells which.  The caller decides whether the outputs go back to buf (analysis) or out to memory.
template <int NC, bool kInverse>
__device__ __forceinline__ void fft_frame(float2 (&v)[16], float2 *buf, int t, int group, const float2 *__restrict__ tw,
                                          const float2 *__restrict__ tw2 /*second pass*/, const float2 *__restrict__ tw3 /*third pass*/) {
    using S = FftShape<NC>;
    constexpr int T = S::kThreads;
    {
        float2 *b1 = buf + fft_pad(16 * t);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = b1[j];
        fft_first_pass<NC, kInverse>(v, tw);
#pragma unroll
        for (int j = 0; j < 16; ++j) b1[j] = v[j];
    }
    frame_sync<T>(group);
    {
        constexpr int M = S::kMid;
        const int k = t & (M - 1);
        float2 *b2 = buf + fft_pad((t / M) * (16 * M) + k);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = b2[fft_pad(j * M)];
        fft_pair_pass<NC, M, kInverse>(v, k, tw2);
        if (NC > 256) {
#pragma unroll
            for (int j = 0; j < 16; ++j) b2[fft_pad(j * M)] = v[j];
        }
    }
    if (NC > 256) {
        frame_sync<T>(group);
        constexpr int M = S::kLast;
        if (S::kThirdIsPair) {
            const int k = t & (M - 1);
            const float2 *b3 = buf + fft_pad((t / M) * (16 * M) + k);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = b3[fft_pad(j * M)];
            fft_pair_pass<NC, M, kInverse>(v, k, tw3);
        } else {
            const float2 *b3 = buf + fft_pad(t);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[4 * q + j] = b3[fft_pad(T * q + j * M)];
            fft_single_pass<NC, kInverse>(v, t, tw3);
        }
    }
}

}  // namespace pvgpu
