// k_synth_ola_ws: the warp-specialised variant of the fused resynthesis kernel (see pv_fused.cu for the single-role kernel and
// the reference arithmetic).  Its own translation unit so that the two sets of instantiations compile in parallel.
#include "pv_kernels.cuh"
#include "pv_fft.cuh"
#include "pv_synth.cuh"
#include "pv_resample.cuh"
#include "pv_fused.cuh"

namespace pvgpu {

// ------------------------------------------------------------------------------------------------
// k_synth_ola_ws: the same work, warp-specialised.  The single-role kernel above alternates between an inverse-FFT phase
// (latency bound: dependent gathers, three shared-memory exchanges, frame barriers) and a resampler phase (FMA and LDS
// bound), three CTAs of 8 warps per SM, and issues on 53 % of the cycles.  Here a CTA has two roles that run CONCURRENTLY:
//   warps 0..7   producers: pre-pass, inverse FFT, ordered overlap-add, normalisation of the finished samples into one of two
//                resampler windows (with the history in front), ring clearing
//   warps 8..15  consumers: the Speex resampler over the window the producers filled one round earlier
// handing windows over through named barriers (full[2] / empty[2], bar.arrive on one side, bar.sync on the other).  Two
// CTAs of 16 warps per SM: the same 32 resident warps as the split kernels, but every scheduler always has both instruction
// mixes to pick from.  Only used when there is a resampler; results are bit-identical to k_synth_ola's.
// Barrier ids: frame groups 1..4 (T > 32 only), tokens kWsToken0.., producers 9, full 10-11, empty 12-13, consumers 14.
// ------------------------------------------------------------------------------------------------
constexpr int kWsProducerBar = 9, kWsFull0 = 10, kWsEmpty0 = 12, kWsConsumerBar = 14;

template <int N, int kPre, bool kOV8>
__global__ void __launch_bounds__(2 * FusedShape<N>::kThreads, N <= 2048 ? 2 : 1) k_synth_ola_ws(const DevPlan p, const DevRows g, const FusedArgs a) {
    using FS = FusedShape<N>;
    constexpr int NC = N / 2;
    using S = FftShape<NC>;
    constexpr int T = FS::T, G = FS::G, U = FS::U, UT = FS::UT, FU = FS::FU, kP = FS::kThreads;   // kP producer threads, kP consumer threads
    constexpr int kWsToken0 = T > 32 ? 5 : 1;
    extern __shared__ __align__(16) float4 smem4[];
    __shared__ ResampleRun s_hdr[2];
    __shared__ int64_t s_res_base[2];
    const int row = blockIdx.x;
    const bool producer = threadIdx.x < kP;
    const int tid = producer ? threadIdx.x : threadIdx.x - kP;   // index inside the role
    const bool quad = !p.rs_direct;
    const int L = (int)p.rs_filt_len, HL = a.hist_len;
    float4 *s_quad = smem4;
    float *s_in0 = (float *)(smem4 + (quad ? p.rs_table_len : 0));   // two windows of in_len floats
    float *s_acc = s_in0 + 2 * a.in_len;
    const int mask = a.acc_len - 1;
    const SliceRec *__restrict__ rr = a.recs - a.recs_base;
    const long k0 = a.k0;
    const int nf = a.nf, R = a.run;
    const int n_rounds = (nf + R - 1) / R;
    const int64_t row_out = (int64_t)row * g.out_stride - g.out_base;
    const int64_t row_limit = g.n_out[row];

    if (!producer) {
        // ------------------------------------------------ consumers ------------------------------------------------
        if (quad) {
            const float4 *__restrict__ tab4 = p.rs_quads;
            for (int e = tid; e < p.rs_table_len; e += kP) s_quad[e] = __ldg(&tab4[e]);
        }
        named_sync(kWsConsumerBar, kP);
        const int warp = tid >> 5, nwarp = kP >> 5;
        for (int r = 0; r < n_rounds; ++r) {
            const int b = r & 1;
            named_sync(kWsFull0 + b, 2 * kP);                       // producers filled window b (and s_hdr[b], s_res_base[b])
            const ResampleRun &hdr = s_hdr[b];
            const float *s_in = s_in0 + b * a.in_len;
            const int x_shift = (int)(hdr.u_lo - (s_res_base[b] - HL)) - kResPad;
            const int64_t orow = row_out + hdr.out_first, out_limit = row_limit - hdr.out_first;
            if (kOV8) resample_run<8>(p, g, hdr, s_quad, s_in, x_shift, orow, out_limit, a.rs_ent, a.rs_frac, a.rs_steps, L, warp, nwarp);
            else if (!quad) resample_run<0>(p, g, hdr, s_quad, s_in, x_shift, orow, out_limit, a.rs_ent, a.rs_frac, a.rs_steps, L, warp, nwarp);
            else if (p.rs_oversample == 4) resample_run<4>(p, g, hdr, s_quad, s_in, x_shift, orow, out_limit, a.rs_ent, a.rs_frac, a.rs_steps, L, warp, nwarp);
            else if (p.rs_oversample == 2) resample_run<2>(p, g, hdr, s_quad, s_in, x_shift, orow, out_limit, a.rs_ent, a.rs_frac, a.rs_steps, L, warp, nwarp);
            else if (p.rs_oversample == 8) resample_run<8>(p, g, hdr, s_quad, s_in, x_shift, orow, out_limit, a.rs_ent, a.rs_frac, a.rs_steps, L, warp, nwarp);
            else resample_run<1>(p, g, hdr, s_quad, s_in, x_shift, orow, out_limit, a.rs_ent, a.rs_frac, a.rs_steps, L, warp, nwarp);
            if (r + 2 < n_rounds) named_arrive(kWsEmpty0 + b, 2 * kP);   // window b may be refilled (nobody waits after the last two rounds)
        }
        return;
    }

    // ---------------------------------------------------- producers ----------------------------------------------------
    const int group = tid / T, t = tid % T, unit = tid / UT;
    float2 *buf = (float2 *)(s_acc + a.acc_len) + group * S::kPadded;
    const int64_t ola_base = rr[k0].ola_off;
    float *__restrict__ tail = g.ola_tail + (int64_t)row * N;
    float *__restrict__ hist = g.res_hist + (int64_t)row * HL;
    for (int i = tid; i < a.acc_len; i += kP) s_acc[i] = i < N ? tail[i] : 0.f;
    for (int i = tid; i < HL; i += kP) s_in0[i] = hist[i];
    named_sync(kWsProducerBar, kP);

    const float2 *__restrict__ w2 = (const float2 *)p.window;
    const int ob = fft_out_base<NC>(t);
    const long total_slots = ((long)(nf / R) * (R / G) + ((nf % R) + G - 1) / G) * U;
    long slot = unit;
    int used_prev = 0;   // normalised samples the previous round appended behind its history
    for (int r = 0; r < n_rounds; ++r) {
        const int r0 = r * R, b = r & 1;
        const int nfr = min(R, nf - r0);
        const long ka = k0 + r0, kb = ka + nfr;
        // ---- inverse FFTs of the run's frames, G at a time, added to the ring in frame order ----
        for (int j0 = 0; j0 < nfr; j0 += G, slot += U) {
            const bool active = j0 + group < nfr;
            const int f = r0 + j0 + group;
            const long k = k0 + f;
            if (active) {
                if (kPre == 0) synth_prepass_generic<N>(p, g, a.car_mag, a.car_phase, row, f, k, t, buf);
                else synth_prepass_lock<N, kPre == 2>(p, g, row, f, t, buf);
            }
            frame_sync<T>(group);
            float2 v[16];
            if (active) fft_frame<NC, true>(v, buf, t, group, p.tw_inv, p.tw2_inv, p.tw3_inv);
            else { if (NC > 256) { frame_sync<T>(group); } frame_sync<T>(group); }
            if (U > 1) named_sync(kWsToken0 + unit, slot != 0 ? 2 * UT : UT);
            else named_sync(kWsProducerBar, kP);
#pragma unroll
            for (int h = 0; h < FU; ++h) {
                if (active && (FU == 1 || (group & (FU - 1)) == h)) fused_add_frame<N>(v, s_acc, mask, (int)(rr[k].ola_off - ola_base), ob, w2);
                if (FU > 1) __syncwarp();
            }
            if (U > 1 && slot != total_slots - 1) named_arrive(kWsToken0 + (unit + 1) % U, 2 * UT);
        }
        named_sync(kWsProducerBar, kP);                            // every frame of the run is in the ring
        if (r >= 2) named_sync(kWsEmpty0 + b, 2 * kP);             // the consumers are done with window b (round r - 2)
        float *s_in = s_in0 + b * a.in_len;
        if (r > 0) {   // history: the last HL samples the previous round left in the other window
            const float *prev = s_in0 + (b ^ 1) * a.in_len + used_prev;
            for (int i = tid; i < HL; i += kP) s_in[i] = prev[i];
        }
        if (tid < (int)(sizeof(ResampleRun) / sizeof(int))) ((int *)&s_hdr[b])[tid] = ((const int *)&a.runs[(ka - a.run_origin) / R])[tid];
        // ---- the run's finished samples: normalise (:1152) into the window, clear the ring ----
        const int64_t res_base = rr[ka].res_off;
        if (tid == 0) s_res_base[b] = res_base;
        for (long k = ka; k < kb; ++k) {
            const SliceRec rc = rr[k];
            if (rc.flags & 1) continue;
            const int off = (int)(rc.ola_off - ola_base);
            const int rel = (int)(rc.res_off - res_base) + HL;
            const float *__restrict__ nrm = a.norm + (rc.ola_off - a.norm_base);
            for (int e = tid; e < rc.shift_inc; e += kP) {
                const int idx = (off + e) & mask;
                const float s = s_acc[idx];
                s_acc[idx] = 0.f;
                if (e < rc.consumed) s_in[rel + e] = s / nrm[e];
            }
        }
        {
            const SliceRec &last = rr[kb - 1];
            used_prev = (int)(last.res_off - res_base) + ((last.flags & 1) ? 0 : last.consumed);
        }
        named_arrive(kWsFull0 + b, 2 * kP);                        // window b is ready
        named_sync(kWsProducerBar, kP);                            // ring cleared before the next run's frames wrap onto it
    }
    // ---- state out (the final window is only read by the consumers from here on) ----
    {
        const SliceRec &last = rr[k0 + nf - 1];
        const int off_end = (int)(last.ola_off - ola_base) + ((last.flags & 1) ? 0 : last.shift_inc);
        for (int i = tid; i < N; i += kP) tail[i] = s_acc[(off_end + i) & mask];
        const float *fin = s_in0 + ((n_rounds - 1) & 1) * a.in_len + used_prev;
        for (int i = tid; i < HL; i += kP) hist[i] = fin[i];
    }
}

template <int N, int kPre, bool kOV8>
static cudaError_t launch_ws_one(const DevPlan &p, const DevRows &g, const FusedArgs &a, size_t sm, cudaStream_t st) {
    static bool configured[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_synth_ola_ws<N, kPre, kOV8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    k_synth_ola_ws<N, kPre, kOV8><<<g.rows, 2 * FusedShape<N>::kThreads, sm, st>>>(p, g, a);
    return cudaSuccess;
}

template <int N>
static cudaError_t launch_ws_n(const DevPlan &p, const DevRows &g, const FusedArgs &a, int pre, bool ov8, size_t sm, cudaStream_t st) {
    if (pre == 1) return ov8 ? launch_ws_one<N, 1, true>(p, g, a, sm, st) : launch_ws_one<N, 1, false>(p, g, a, sm, st);
    if (pre == 2) return ov8 ? launch_ws_one<N, 2, true>(p, g, a, sm, st) : launch_ws_one<N, 2, false>(p, g, a, sm, st);
    return ov8 ? launch_ws_one<N, 0, true>(p, g, a, sm, st) : launch_ws_one<N, 0, false>(p, g, a, sm, st);
}

cudaError_t launch_synth_ola_ws(const DevPlan &p, const DevRows &g, const FusedArgs &a, int pre, bool ov8, size_t sm, cudaStream_t st) {
    switch (p.N) {
        case 512: return launch_ws_n<512>(p, g, a, pre, ov8, sm, st);
        case 1024: return launch_ws_n<1024>(p, g, a, pre, ov8, sm, st);
        case 2048: return launch_ws_n<2048>(p, g, a, pre, ov8, sm, st);
        case 4096: return launch_ws_n<4096>(p, g, a, pre, ov8, sm, st);
        case 8192: return launch_ws_n<8192>(p, g, a, pre, ov8, sm, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace pvgpu
