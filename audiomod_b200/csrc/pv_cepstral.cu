// k_cepstral: the cepstral spectral-envelope kernel (north_star "Formant path") -- an OPTIONAL pair of
// modes (PVGPU_GENDER_CEPSTRAL, PVGPU_FORMANT_CEPSTRAL).  In the reference this routine, formantShiftSlice
// (src/phasevocoder/phasevocoderprocess.cc:925-999) with FFT::inverseCepstral (src/common/dsp/FFT.cc:2723-2733), is dead code:
// every call site is commented out in favour of the nearest-bin warp freqCompSlice (:824-840), which is what the parity
// modes 1 and 2 implement.  The modes here do what the reference does when those three comments are swapped back
// (formantPreserveSlice -> formantShiftSlice(ch, 1), maleToFemale -> (ch, 0.85), femaleToMale -> (ch, 1.17)); the test
// infrastructure builds exactly that variant of the reference as the checker (tests/test_gpu_cepstral.py).
//
// Per frame, on the magnitudes the phase core left untouched:
//   c   = kiss_fftri(log(mag + 1e-6))                      the real cepstrum (unnormalised inverse real FFT)
//   c[0] /= 2, c[59] /= 2, c[60..] = 0, c[0..59] *= 1/N     the lifter: 60 quefrency bins
//   env = exp(Re kiss_fftr(c))                              the smooth spectral envelope
//   mag = mag / env * env[warp(bin)]                        whiten, then put a frequency-warped envelope back
// One frame per N/32-thread group, both FFTs through the register-tiled KissFFT-order transforms of pv_fft.cuh; spectra are
// scaled in place ((re, im) in the Cartesian pipelines, mag in the polar ones), between the phase core and the synthesis.
#include "pv_kernels.cuh"
#include "pv_fft.cuh"

namespace pvgpu {

constexpr int kCepCutoff = 60;   // phasevocoderprocess.cc:946

__device__ __forceinline__ float env_at_or_zero(const float *env, int src, int nc) { return src > nc ? 0.f : env[src]; }

template <int N>
__global__ void __launch_bounds__((N / 32 > 256 ? N / 32 : 256)) k_cepstral(const DevPlan p, const DevRows g, float env_comp, int nf, int total) {
    constexpr int NC = N / 2;
    using S = FftShape<NC>;
    constexpr int T = S::kThreads;
    constexpr int G = (T >= 256) ? 1 : 256 / T;
    constexpr int kEnv = (NC + 1 + 3) & ~3;           // floats of the envelope row
    extern __shared__ __align__(16) float2 sbuf[];
    const int group = threadIdx.x / T, t = threadIdx.x % T;
    float2 *buf = sbuf + group * S::kPadded;
    float *env = (float *)(sbuf + G * S::kPadded) + group * kEnv;
    const int fid = blockIdx.x * G + group;
    const bool active = fid < total;
    const int row = active ? fid / nf : 0, f = active ? fid % nf : 0;
    float *__restrict__ ga = g.mag + ((int64_t)row * g.F + f) * p.Hp;     // re or mag
    float *__restrict__ gb = g.phase + ((int64_t)row * g.F + f) * p.Hp;   // im or phase
    constexpr int Q = (NC / 2) / T;
    const int sa = fft_pad(fft_slot_of_input<NC>(t)), sb = fft_pad(fft_slot_of_input<NC>((T - t) & (T - 1)));
    auto mag_of = [&](int i) -> float {               // FFT.cc:2624 (Cartesian spectra: formed here with the same operations)
        const float a = ga[i];
        if (!g.spec) return a;
        const float b = gb[i];
        return __fsqrt_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
    };
    // ---- log-magnitude spectrum (imaginary parts 0) through the inverse real-FFT pre-pass (kiss_fftr.c:123-159) ----
    if (active) {
        const float2 *__restrict__ stw = p.stw_inv;
        constexpr int U = Q >= 4 ? 4 : Q;   // pairs per step: their (up to 16) loads are all in flight before the first logf
#pragma unroll 1
        for (int q0 = 0; q0 < Q; q0 += U) {
            float la[U], lb[U], ha[U], hb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int kk = t + T * (q0 + u);
                la[u] = ga[kk]; ha[u] = ga[NC - kk];
                lb[u] = g.spec ? gb[kk] : 0.f; hb[u] = g.spec ? gb[NC - kk] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = q0 + u, kk = t + T * q;
                const float ml = g.spec ? __fsqrt_rn(__fadd_rn(__fmul_rn(la[u], la[u]), __fmul_rn(lb[u], lb[u]))) : la[u];
                const float mh = g.spec ? __fsqrt_rn(__fadd_rn(__fmul_rn(ha[u], ha[u]), __fmul_rn(hb[u], hb[u]))) : ha[u];
                const float fk = logf(ml + 0.000001f), fq = logf(mh + 0.000001f);
                if (kk == 0) {
                    buf[fft_pad(fft_slot_of_input<NC>(0))] = make_float2(fk + fq, fk - fq);
                } else {
                    const float2 fek = make_float2(__fadd_rn(fk, fq), 0.f), d = make_float2(__fsub_rn(fk, fq), 0.f);
                    const float2 fok = cmul_rn(d, __ldg(&stw[kk]));
                    const float2 a = cadd_rn(fek, fok);
                    const float2 b = csub_rn(fek, fok);
                    buf[sa + fft_pad(fft_slot_of_input<NC>(T * q))] = a;
                    buf[sb + (t == 0 ? fft_pad(fft_slot_of_input<NC>((T * (16 - q)) & (NC - 1))) : fft_pad(fft_slot_of_input<NC>(T * (15 - q))))] =
                        make_float2(b.x, -b.y);
                }
            }
        }
        if (t == 0) {   // kk == NC/2 pairs with itself; the second write wins (kiss_fftr.c:150-155)
            const float fk = logf(mag_of(NC / 2) + 0.000001f);
            const float2 fek = make_float2(__fadd_rn(fk, fk), 0.f), d = make_float2(0.f, 0.f);
            const float2 fok = cmul_rn(d, __ldg(&stw[NC / 2]));
            const float2 b = csub_rn(fek, fok);
            buf[fft_pad(fft_slot_of_input<NC>(NC / 2))] = make_float2(b.x, -b.y);
        }
    }
    frame_sync<T>(group);
    float2 v[16];
    if (active) fft_frame<NC, true>(v, buf, t, group, p.tw_inv, p.tw2_inv, p.tw3_inv);
    else { if (NC > 256) { frame_sync<T>(group); } frame_sync<T>(group); }
    frame_sync<T>(group);   // every read of the inverse transform's last pass is done
    // ---- lifter: the first 60 cepstral coefficients (complex outputs 0..29 hold samples 0..59), the rest zero, as the
    // permuted input of the forward transform (complex c = samples 2c, 2c+1; no window, no shift) ----
    if (active) {
#pragma unroll
        for (int i = 0; i < 16; ++i) buf[fft_pad(fft_slot_of_input<NC>(t + T * i))] = make_float2(0.f, 0.f);
    }
    frame_sync<T>(group);
    if (active) {
        const int ob = fft_out_base<NC>(t);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int o = ob + fft_out_const<NC>(i);
            if (o < kCepCutoff / 2) {
                float c0 = v[i].x, c1 = v[i].y;
                if (o == 0) c0 = __fmul_rn(c0, 0.5f);                          // internalbuffer[0] /= 2
                if (o == kCepCutoff / 2 - 1) c1 = __fmul_rn(c1, 0.5f);         // internalbuffer[cutoff - 1] /= 2
                buf[fft_pad(fft_slot_of_input<NC>(o))] = make_float2(__fmul_rn(c0, p.inv_n), __fmul_rn(c1, p.inv_n));
            }
        }
    }
    frame_sync<T>(group);
    if (active) fft_frame<NC, false>(v, buf, t, group, p.tw_fwd, p.tw2_fwd, p.tw3_fwd);
    else { if (NC > 256) { frame_sync<T>(group); } frame_sync<T>(group); }
    frame_sync<T>(group);
    if (active) {
        float2 *bo = buf + fft_pad(fft_out_base<NC>(t));
#pragma unroll
        for (int i = 0; i < 16; ++i) bo[fft_pad(fft_out_const<NC>(i))] = v[i];
    }
    frame_sync<T>(group);
    // ---- real parts of the forward real FFT (kiss_fftr.c:67-121) -> envelope = exp(.) ----
    if (active) {
        const float2 *__restrict__ stw = p.stw_fwd;
        const int pa = fft_pad(t), pb = fft_pad((T - t) & (T - 1));
#pragma unroll 1
        for (int q = 0; q < Q; ++q) {
            const int kk = t + T * q;
            const float2 fpk = buf[pa + fft_pad(T * q)];
            const float2 fq = buf[pb + (t == 0 ? fft_pad((T * (16 - q)) & (NC - 1)) : fft_pad(T * (15 - q)))];
            float ar, br;
            if (kk == 0) {
                ar = __fadd_rn(fpk.x, fpk.y);
                br = __fsub_rn(fpk.x, fpk.y);
            } else {
                const float2 fpnk = make_float2(fq.x, -fq.y);
                const float2 f1k = cadd_rn(fpk, fpnk), f2k = csub_rn(fpk, fpnk);
                const float2 tw = cmul_rn(f2k, __ldg(&stw[kk]));
                ar = __fmul_rn(__fadd_rn(f1k.x, tw.x), 0.5f);
                br = __fmul_rn(__fsub_rn(f1k.x, tw.x), 0.5f);
            }
            env[kk] = expf(ar);
            env[NC - kk] = expf(br);
        }
        if (t == 0) {
            const float2 fpk = buf[fft_pad(NC / 2)];
            const float2 fpnk = make_float2(fpk.x, -fpk.y);
            const float2 f1k = cadd_rn(fpk, fpnk), f2k = csub_rn(fpk, fpnk);
            const float2 tw = cmul_rn(f2k, __ldg(&stw[NC / 2]));
            env[NC / 2] = expf(__fmul_rn(__fsub_rn(f1k.x, tw.x), 0.5f));
        }
    }
    frame_sync<T>(group);
    if (!active) return;
    // ---- whiten by the envelope, put the warped envelope back (:969-994), in place ----
    constexpr int kBins = NC / T;   // 16 bins per thread, plus bin NC for thread 0
#pragma unroll 1
    for (int j0 = 0; j0 < kBins; j0 += 4) {
        float a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = t + T * (j0 + u); a[u] = ga[i]; b[u] = g.spec ? gb[i] : 0.f; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = t + T * (j0 + u);
            const float e_new = env_comp > 1.0f ? env_at_or_zero(env, __float2int_rn(__fmul_rn((float)i, env_comp)), NC)
                                                 : env[__float2int_rn(__fmul_rn((float)i, env_comp))];
            const float e_old = env[i];
            ga[i] = __fmul_rn(__fdiv_rn(a[u], e_old), e_new);
            if (g.spec) gb[i] = __fmul_rn(__fdiv_rn(b[u], e_old), e_new);
        }
    }
    if (t == 0) {   // bin N/2: warped like the rest when compressing, left alone by the descending loop of the expanding direction
        const float e_old = env[NC];
        const float e_new = env_comp > 1.0f ? env_at_or_zero(env, __float2int_rn(__fmul_rn((float)NC, env_comp)), NC) : e_old;
        ga[NC] = __fmul_rn(__fdiv_rn(ga[NC], e_old), e_new);
        if (g.spec) gb[NC] = __fmul_rn(__fdiv_rn(gb[NC], e_old), e_new);
    }
}

template <int N> static void launch_cepstral_t(const DevPlan &p, const DevRows &g, float env_comp, int nframes, cudaStream_t st) {
    using S = FftShape<N / 2>;
    constexpr int T = S::kThreads, G = (T >= 256) ? 1 : 256 / T;
    const int total = nframes * g.rows;
    const int grid = (total + G - 1) / G, block = T >= 256 ? T : 256;
    const size_t sm = sizeof(float2) * G * S::kPadded + sizeof(float) * G * ((N / 2 + 1 + 3) & ~3);
    static bool configured[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (sm > 48 * 1024 && dev < 16 && !configured[dev]) {
        cudaFuncSetAttribute(k_cepstral<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
        configured[dev] = true;
    }
    k_cepstral<N><<<grid, block, sm, st>>>(p, g, env_comp, nframes, total);
}

bool launch_cepstral(const DevPlan &p, const DevRows &g, float env_comp, int nframes, cudaStream_t st) {
    switch (p.N) {
        case 512: launch_cepstral_t<512>(p, g, env_comp, nframes, st); return true;
        case 1024: launch_cepstral_t<1024>(p, g, env_comp, nframes, st); return true;
        case 2048: launch_cepstral_t<2048>(p, g, env_comp, nframes, st); return true;
        case 4096: launch_cepstral_t<4096>(p, g, env_comp, nframes, st); return true;
        case 8192: launch_cepstral_t<8192>(p, g, env_comp, nframes, st); return true;
        default: return false;   // the cepstral modes exist for the register-tiled FFT sizes only
    }
}

}  // namespace pvgpu
