// The phase-locked core (coremode 1, phasevocoderprocess.cc:574-706) on CARTESIAN spectra, split into its frame-parallel
// part and its frame-serial recursion.  Included by pv_kernels.cu inside namespace pvgpu.
//
// Locking a region to its peak adds one rotation `rot` to the phase of every bin of the region (:685-699), so the input of
// the synthesis, mag * (cos, sin)(princarg(phase + rot)), is the analysis bin (re, im) multiplied by (cos rot, sin rot): the
// analysis kernel stores (re, im) and does no sqrtf / atan2f, the synthesis kernel does no sincosf per bin.  The analysis
// phase itself is only needed where the recursion reads it -- at the peaks, phi[p2], and at the previous frame's linked
// bins, prev_phase[p1] / prev_outphase[p1] (:663-665) -- and the only quantity that really depends on the previous
// frame's *result* is prev_outphase[p1] = princarg(prev_phase[p1] + rot_prev[region of p1]).  Hence two kernels:
//
//   k_lock_peaks   parallel over (stream, run of frames): exact peak picking (:587-596) on squared magnitudes (with an
//                  exact-sqrt tie-break), region starts (:668-683), nearest-previous-peak linking (:641-652) through the
//                  previous iteration's bin->region map, and per peak everything that depends only on analysis data:
//                  phi[p2], prev_phase[p1] (restated atan2f), the peak's advance dphi * phaseInc / hop (:655-664) and the
//                  region of p1 in the channel's previous frame.  One 16-byte record per peak.
//   k_lock_chain   one CTA per stream, serial over frames, one thread per peak: prev_out -> target -> rot
//                  (three princargs, :663-667) and (cos rot, sin rot) for the synthesis kernel.
//
// Every decision (peaks, links, regions) and every peak phase / rotation is computed with the reference's operations in
// the reference's order, i.e. bit-identically; only the rotated output bins differ from the reference's
// atan2f -> princarg -> sincosf round trip, by float rounding (~1e-7 relative).
//
// State of a channel between frames, as the chain kernel sees it (prev_phase is always the analysis phase of the
// channel's previous frame, which k_lock_peaks recomputes from that frame's spectrum):
//   kind 0  nothing yet                     prev_phase = prev_outphase = 0   (channelinfo.cc:92-115)
//   kind 1  prev_outphase is an array       (after a frame of classic propagation, :617-636)
//   kind 2  prev_outphase[i] = princarg(prev_phase[i] + rot[region(i)])      (after a locked frame)
//   kind 3  prev_outphase = prev_phase      (after the pass-through first frame, :606-616)
// Frames are tagged by k_lock_peaks (lock_hdr[..].y): 0 pass-through first frame, 1 classic propagation (no peaks now or
// in the previous iteration), 2 locked.  Classic frames are rare (silence); their per-bin inputs go through the same record
// area and the chain kernel rewrites their spectrum in place as mag * (cos, sin)(outphase).
#pragma once

constexpr int kLockRun = 32;   // frames per CTA of k_lock_peaks (plus one warm-up frame that is only peak-picked)

// strict "mag[b] > every neighbour" (:589-592) decided on squared magnitudes q = fl(fl(re^2) + fl(im^2)): sqrtf is monotone,
// so q_b <= q_n means "not greater"; q_b > q_n (1 + 2^-21) means the rounded square roots differ; in between the exactly
// rounded square roots decide
__device__ __forceinline__ bool lock_is_peak(float qb, float qmax) {
    if (!(qb > qmax)) return false;
    if (qb > __fmul_rn(qmax, 1.00000048f) && qb > 1e-30f) return true;
    return __fsqrt_rn(qb) > __fsqrt_rn(qmax);
}

// shared-memory layout of k_lock_peaks, in floats: where the bin->region maps start (16-byte aligned) and the total
__host__ __device__ inline int lock_peaks_map_offset(int half, int C, int maxpk) { return (3 * half + 8 + 2 * C * half + 3 * maxpk + 1 + 32 + 3) & ~3; }
inline size_t lock_peaks_smem(int half, int C, int maxpk) {
    return sizeof(float) * (size_t)lock_peaks_map_offset(half, C, maxpk) + sizeof(unsigned short) * (size_t)C * half + 4 * sizeof(float) * (size_t)C * maxpk;
}
inline size_t lock_chain_smem(int half, int C, int maxpk) { return sizeof(float) * ((size_t)C * 2 * maxpk + (size_t)C * half) + sizeof(int) * 2 * (size_t)C; }

// launch shape of k_lock_peaks<N, kC>: E bins per thread, half / E threads, CTAs per SM the register budget allows
template <int N> struct LockShape {
    static constexpr int kHalf = N / 2;
    static constexpr int E = N == 8192 ? 8 : 4;
    static constexpr int kThreads = kHalf / E;
    static constexpr int kMinBlocks = 1024 / kThreads > 16 ? 16 : 1024 / kThreads;   // 64 registers per thread
    static constexpr int kMaxPk = kHalf / 3 + 2;   // == Pipeline::max_peaks(): peaks are at least 3 bins apart
};

template <int N, int kC>   // kC: channels per stream when 1 or 2, 0 = any (run time)
__global__ void __launch_bounds__(LockShape<N>::kThreads, LockShape<N>::kMinBlocks) k_lock_peaks(const DevPlan p, const DevRows g, const SliceRec *__restrict__ recs,
                                                                                                   long recs_base, long k0, int nframes) {
    constexpr int E = LockShape<N>::E;
    extern __shared__ float smem[];
    constexpr int half = LockShape<N>::kHalf, maxpk = LockShape<N>::kMaxPk, nthr = LockShape<N>::kThreads, nwarp = nthr / 32;
    const int C = kC ? kC : g.channels;
    float *s_q = smem;                                // half + 8   squared magnitudes, two guard bins each side
    float *s_cre = s_q + half + 8;                    // half       current frame
    float *s_cim = s_cre + half;                      // half
    float *s_A = s_cim + half;                        // C * half   the channel's previous frame (re)
    float *s_B = s_A + C * half;                      // C * half   (im)
    int *s_pk0 = (int *)(s_B + C * half);             // maxpk      peak list A (shared by the channels of the stream)
    int *s_pk1 = s_pk0 + maxpk;                       // maxpk      peak list B
    int *s_start = s_pk1 + maxpk;                     // maxpk + 1  region starts of the current frame
    int *s_wsum = s_start + maxpk + 1;                // 32
    unsigned short *s_map = (unsigned short *)(smem + lock_peaks_map_offset(half, C, maxpk));   // C * half   region of every bin in the channel's latest frame
    const int stream = blockIdx.y, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int b0 = tid * E;
    const int fa = blockIdx.x * kLockRun, fb = min(fa + kLockRun, nframes);
    const int f_first = (k0 + fa > 0) ? fa - 1 : fa;   // a frame before the run exists: warm up on it (peaks, regions, spectrum)
    float *s_cph = (float *)(s_map + C * half);       // C * 2 * maxpk  phi at the peaks of the channel's latest frame (ping-pong) ...
    int *s_cpk = (int *)(s_cph + C * 2 * maxpk);      // C * 2 * maxpk  ... and their bins (valid while bit c of cache_ok is set)
    unsigned cache_ok = 0, cache_side = 0;            // bit c of cache_side: which half holds the channel's latest frame

    // (frame slot, channel) of the next fetch; slot -1 is the last frame of the previous launch (lock_tail)
    float pre[E], pim[E];
    int ff = f_first, cf = 0;
    const int64_t row0 = (int64_t)stream * C;
    const int64_t ch_step = (int64_t)g.F * p.Hp;                 // next channel, same frame
    const int64_t fr_step = (int64_t)p.Hp - (C - 1) * ch_step;   // channel C-1 of frame f -> channel 0 of frame f+1
    const float *fre = g.mag + (row0 * g.F + max(f_first, 0)) * p.Hp + b0;   // (frame max(f_first, 0), channel 0, bin b0)
    const float *fim = g.phase + (row0 * g.F + max(f_first, 0)) * p.Hp + b0;
    auto fetch = [&]() {
        const float *sr, *si;
        if (ff < 0) {
            sr = g.lock_tail + (row0 + cf) * 2 * p.Hp + b0;
            si = sr + p.Hp;
        } else {
            sr = fre; si = fim;
            const int64_t step = (kC == 1 || cf == C - 1) ? fr_step : ch_step;
            fre += step; fim += step;
        }
        if (kC == 1 || ++cf == C) { cf = 0; ++ff; }
#pragma unroll
        for (int e = 0; e < E; e += 4) {
            const float4 r4 = *(const float4 *)(sr + e);
            const float4 i4 = *(const float4 *)(si + e);
            pre[e] = r4.x; pre[e + 1] = r4.y; pre[e + 2] = r4.z; pre[e + 3] = r4.w;
            pim[e] = i4.x; pim[e + 1] = i4.y; pim[e + 2] = i4.z; pim[e + 3] = i4.w;
        }
    };
    if (f_first < fb) fetch();
    if (tid < 4) { s_q[tid] = 0.f; s_q[half + 4 + tid] = 0.f; }   // guards (bins 0, 1, half-2, half-1 are never peaks anyway)
    int *s_prev = s_pk0, *s_cur = s_pk1;
    int nprev = 0;
    const float hopf = (float)p.hop;
    const SliceRec *__restrict__ rec_f = recs + (k0 - recs_base);

    for (int f = f_first; f < fb; ++f) {
        const bool warm = f < fa;
        const bool k_is_0 = k0 + f == 0;
        for (int c = 0; c < C; ++c) {
            float *A = s_A + c * half, *B = s_B + c * half;
            unsigned short *map_c = s_map + c * half;                              // this channel's previous frame
            const unsigned short *map_p = s_map + (c == 0 ? C - 1 : c - 1) * half;   // the previous iteration's frame
            float re[E], im[E], q[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                re[e] = pre[e]; im[e] = pim[e];
                q[e] = __fadd_rn(__fmul_rn(re[e], re[e]), __fmul_rn(im[e], im[e]));   // FFT.cc:2624 before the sqrtf
            }
#pragma unroll
            for (int e = 0; e < E; e += 4) {
                *(float4 *)(s_q + 4 + b0 + e) = make_float4(q[e], q[e + 1], q[e + 2], q[e + 3]);
                *(float4 *)(s_cre + b0 + e) = make_float4(re[e], re[e + 1], re[e + 2], re[e + 3]);
                *(float4 *)(s_cim + b0 + e) = make_float4(im[e], im[e + 1], im[e + 2], im[e + 3]);
            }
            __syncthreads();   // (A)
            float w[E + 4];
            w[0] = s_q[4 + b0 - 2]; w[1] = s_q[4 + b0 - 1];
#pragma unroll
            for (int e = 0; e < E; ++e) w[2 + e] = q[e];
            w[E + 2] = s_q[4 + b0 + E]; w[E + 3] = s_q[4 + b0 + E + 1];
            // branch-free first: candidates (q_b above all four neighbours) and sure peaks (above them by the margin that
            // separates the rounded square roots, lock_is_peak); the rare candidates in between take the exact-sqrt test
            unsigned flags = 0, unsure = 0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const float qmax = fmaxf(fmaxf(w[e], w[e + 1]), fmaxf(w[e + 3], w[e + 4]));
                const bool cand = w[e + 2] > qmax;
                const bool sure = w[e + 2] > fmaxf(__fmul_rn(qmax, 1.00000048f), 1e-30f);
                flags |= (unsigned)sure << e;
                unsure |= (unsigned)(cand && !sure) << e;
            }
            while (unsure) {
                const int e = __ffs(unsure) - 1;
                unsure &= unsure - 1;
                const float qmax = fmaxf(fmaxf(s_q[4 + b0 + e - 2], s_q[4 + b0 + e - 1]), fmaxf(s_q[4 + b0 + e + 1], s_q[4 + b0 + e + 2]));
                flags |= (unsigned)lock_is_peak(s_q[4 + b0 + e], qmax) << e;
            }
            // 2 <= b <= half - 3 (:587): only the first and the last thread own excluded bins
            if (tid == 0) flags &= ~3u;
            if (tid == nthr - 1) flags &= ~(3u << (E - 2));
            const int cnt = __popc(flags);
            // peaks in the lanes before this one: a thread has 0..3 peaks, so three ballots count them
            const unsigned lt = (1u << lane) - 1;
            const unsigned b1 = __ballot_sync(0xffffffffu, cnt >= 1), b2 = __ballot_sync(0xffffffffu, cnt >= 2);
            int excl = __popc(b1 & lt) + __popc(b2 & lt), wtot = __popc(b1) + __popc(b2);
            if (E == 8) { const unsigned b3 = __ballot_sync(0xffffffffu, cnt >= 3); excl += __popc(b3 & lt); wtot += __popc(b3); }
            const int incl = excl + cnt;
            if (lane == 31) s_wsum[warp] = wtot;
            if (ff < fb) fetch();   // prefetch the next (frame, channel)
            __syncthreads();   // (B)
            // peaks in the warps before this one, and in the whole frame: a second scan over the (at most 16) warp totals
            int ws = lane < nwarp ? s_wsum[lane] : 0;
#pragma unroll
            for (int d = 1; d < nwarp; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, ws, d);
                if (lane >= d) ws += v;
            }
            const int npk = __shfl_sync(0xffffffffu, ws, nwarp - 1);
            int base = __shfl_sync(0xffffffffu, ws, (warp + 31) & 31);   // inclusive total of warp - 1; warp 0 reads lane 31 (discarded)
            base = (warp == 0 ? 0 : base) + incl - cnt;   // peaks before this thread's first bin
            if (flags) {   // peaks are at least 3 bins apart: at most two in 4 bins, three in 8
                s_cur[base] = b0 + __ffs(flags) - 1;
                if (cnt > 1) s_cur[base + cnt - 1] = b0 + 31 - __clz(flags);
                if (E == 8 && cnt > 2) s_cur[base + 1] = b0 + __ffs(flags & (flags - 1)) - 1;
            }
            const int64_t slot = (row0 + c) * g.F + f;    // (row, frame slot); not used while warming up
            const bool first = k_is_0 && c == 0;          // first call of the process (:602-616)
            const int kind = first ? 0 : (npk == 0 || nprev == 0) ? 1 : 2;
            const bool lock_now = !warm && kind == 2;
            __syncthreads();   // (C) peak list complete
            for (int r = tid; r <= npk; r += nthr)   // region starts (:668-683): round((a + b) * 0.5), half away from zero
                s_start[r] = r == 0 ? 0 : r == npk ? half : (s_cur[r - 1] + s_cur[r] + 1) >> 1;
            if (lock_now) {
                const float phase_inc = (float)rec_f[f].phase_inc;
                float4 *__restrict__ out = g.lock_rec + slot * g.rec_stride;
                const bool cached = (cache_ok >> c) & 1;
                const int cs = (cache_side >> c) & 1;
                const float *cph = s_cph + (c * 2 + cs) * maxpk;
                const int *cpk = s_cpk + (c * 2 + cs) * maxpk;
                float *nph = s_cph + (c * 2 + (cs ^ 1)) * maxpk;   // this frame's peak phases: the next frame's cache
                int *npb = s_cpk + (c * 2 + (cs ^ 1)) * maxpk;
                for (int r = tid; r < npk; r += nthr) {
                    const int p2 = s_cur[r];
                    // nearest previous peak, ties keep the lower index (:641-652): the region of the previous iteration's
                    // frame that contains p2 belongs to the nearest peak; exactly half way the reference stays on the lower one
                    const int r1 = map_p[p2];
                    const int q1 = s_prev[r1];
                    const int j = (r1 > 0 && q1 - p2 == p2 - s_prev[r1 - 1]) ? r1 - 1 : r1;
                    const int p1 = s_prev[j];
                    const int ridx = map_c[p1];
                    const float php = pv_atan2f_fast(s_cim[p2], s_cre[p2]);
                    float a;   // prev_phase[p1]: phi of the channel's previous frame, cached when p1 was one of its peaks
                    if (k_is_0) a = 0.f;
                    else if (cached && cpk[ridx] == p1) a = cph[ridx];
                    else a = pv_atan2f_fast(B[p1], A[p1]);
                    const float avg_p = (float)((double)(p1 + p2) * 0.5);
                    // (2*pi*hop*(avg_p-1)) / N: N is a power of two, so the division is an exact scaling
                    const float pomega = (float)__dmul_rn(__dmul_rn(p.two_pi_hop, (double)__fsub_rn(avg_p, 1.0f)), (double)p.inv_n);
                    const float dphi = (float)__dadd_rn((double)pomega, princarg_fast((double)sub3_rn(php, a, pomega)));
                    const float adv = __fdiv_rn(__fmul_rn(dphi, phase_inc), hopf);
                    out[r] = make_float4(php, a, adv, __int_as_float(p1 | (ridx << 16)));
                    nph[r] = php; npb[r] = p2;
                }
            } else if (!warm && kind == 1) {
                // classic propagation (:617-636): per bin phi, prev_phase and the advance; prev_outphase is the chain's business
                const float phase_inc = (float)rec_f[f].phase_inc;
                float *__restrict__ adv_o = (float *)(g.lock_rec + slot * g.rec_stride);
                float *__restrict__ a_o = adv_o + half;
                unsigned short *__restrict__ ridx_o = g.lock_map + slot * half;
#pragma unroll 1
                for (int e = 0; e < E; ++e) {   // rolled: rare path, keep the kernel small
                    const int i = b0 + e;
                    const float omega = __ldg(&p.omega[i]);
                    const float phi = pv_atan2f_fast(s_cim[i], s_cre[i]);
                    const float a = k_is_0 ? 0.f : pv_atan2f_fast(B[i], A[i]);
                    const float dphi = (float)__dadd_rn((double)omega, princarg_fast((double)sub3_rn(phi, a, omega)));
                    adv_o[i] = __fdiv_rn(__fmul_rn(dphi, phase_inc), hopf);
                    a_o[i] = a;
                    ridx_o[i] = map_c[i];
                }
            }
            __syncthreads();   // (D) region starts complete; every read of the channel's previous frame is done
            if (npk > 0) {
                // region of bin i: that of the last peak at or before i (index lo; none: region 0), or the next one's when i
                // has passed their midpoint.  s_start[npk] = half keeps the last region open-ended.
                unsigned short rg[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int lo = base - 1 + __popc(flags & ((2u << e) - 1));
                    rg[e] = (unsigned short)(lo + (b0 + e >= s_start[lo + 1]));
                }
#pragma unroll
                for (int e = 0; e < E; e += 4) {
                    const uint2 v = make_uint2(rg[e] | ((unsigned)rg[e + 1] << 16), rg[e + 2] | ((unsigned)rg[e + 3] << 16));
                    *(uint2 *)(map_c + b0 + e) = v;
                    if (lock_now) *(uint2 *)(g.lock_map + slot * half + b0 + e) = v;
                }
            }
#pragma unroll
            for (int e = 0; e < E; e += 4) {
                *(float4 *)(A + b0 + e) = make_float4(re[e], re[e + 1], re[e + 2], re[e + 3]);
                *(float4 *)(B + b0 + e) = make_float4(im[e], im[e + 1], im[e + 2], im[e + 3]);
            }
            // the peak phases of a locked frame are the prev_phase of most of the next frame's links
            if (lock_now) { cache_ok |= 1u << c; cache_side ^= 1u << c; }
            else cache_ok &= ~(1u << c);
            if (!warm && tid == 0) g.lock_hdr[slot] = make_int2(npk, kind);
            { int *tsw = s_prev; s_prev = s_cur; s_cur = tsw; }
            nprev = npk;
            // (A)..(C) of the next iteration order these writes before their readers
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_lock_chain: the serial recursion.  One CTA per stream; a thread per peak.
// ------------------------------------------------------------------------------------------------
// classic propagation of one (frame, channel) in the chain kernel (:617-636); the spectrum is rewritten in place as
// mag * (cos, sin)(outphase).  Rare (silence), kept out of line.
__device__ __noinline__ void lock_chain_classic(int half, const float *__restrict__ adv_i, const unsigned short *__restrict__ ridx_i, float *__restrict__ gre,
                                                float *__restrict__ gim, int kp, const float *rot_p, float *out_c) {
    const float *__restrict__ a_i = adv_i + half;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        const float a = a_i[i];
        float po;
        if (kp == 2) po = (float)princarg_fast((double)__fadd_rn(a, rot_p[ridx_i[i]]));
        else if (kp == 1) po = out_c[i];
        else if (kp == 3) po = a;
        else po = 0.f;
        const float outp = (float)princarg_fast((double)__fadd_rn(po, adv_i[i]));
        out_c[i] = outp;
        const float re = gre[i], im = gim[i];
        const float m = __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
        float sn, cs;
        sincosf(outp, &sn, &cs);
        gre[i] = m * cs; gim[i] = m * sn;
    }
}

constexpr int kChainThreads = 128;
template <int kC>   // channels per stream when 1 or 2, 0 = any (run time)
__global__ void __launch_bounds__(kChainThreads, 10) k_lock_chain(const DevPlan p, const DevRows g, int nframes) {
    extern __shared__ float smem[];
    const int half = p.half, C = kC ? kC : g.channels, maxpk = g.maxpk;
    float *s_rot = smem;                          // C * 2 * maxpk   rotation of every region of the channel's previous frame (ping-pong)
    float *s_out = s_rot + C * 2 * maxpk;         // C * half        prev_outphase while the channel is in state kind 1
    int *s_kind = (int *)(s_out + C * half);      // C
    int *s_flip = s_kind + C;                     // C
    const int stream = blockIdx.x, tid = threadIdx.x;
    constexpr int nthr = kChainThreads;
    const int64_t row0 = (int64_t)stream * C;
    for (int c = 0; c < C; ++c) {
        const int64_t row = row0 + c;
        const int kd = g.lock_kind[row];
        if (kd == 2) for (int i = tid; i < maxpk; i += nthr) s_rot[c * 2 * maxpk + i] = g.lock_rot[row * maxpk + i];
        if (kd == 1) for (int i = tid; i < half; i += nthr) s_out[c * half + i] = g.prev_out[row * half + i];
        if (tid == 0) { s_kind[c] = kd; s_flip[c] = 0; }
    }
    __syncthreads();
    int kind_r[kC ? kC : 1], flip_r[kC ? kC : 1];
#pragma unroll
    for (int c = 0; c < kC; ++c) { kind_r[c] = s_kind[c]; flip_r[c] = 0; }
    // Running pointers to (frame, channel) of the next fetch and of the current iteration.  The header and this thread's
    // first two records of the next (frame, channel) are fetched one iteration ahead; the record area is padded by
    // 2 * kChainThreads entries, so the loads are unconditional (garbage beyond the frame's peaks, never used).
    const int rs = g.rec_stride;
    const int64_t ch_step = g.F, fr_step = 1 - (int64_t)(C - 1) * g.F;   // in (row, frame) slots
    const int2 *__restrict__ hdr_f = g.lock_hdr + row0 * g.F;
    const float4 *__restrict__ rec_f = g.lock_rec + row0 * g.F * rs + tid;
    const float4 *__restrict__ rec_c = rec_f;
    float2 *__restrict__ csn_c = g.lock_csn + row0 * g.F * maxpk;
    int2 hdr_n = make_int2(0, 0);
    float4 rec_n0 = make_float4(0.f, 0.f, 0.f, 0.f), rec_n1 = rec_n0;
    int cf = 0, left = nframes * C;   // iterations still to fetch
    auto fetch = [&]() {
        hdr_n = *hdr_f;
        rec_n0 = rec_f[0];
        rec_n1 = rec_f[nthr];
        const int64_t step = (kC == 1 || cf == C - 1) ? fr_step : ch_step;
        if (kC != 1) cf = cf == C - 1 ? 0 : cf + 1;
        hdr_f += step; rec_f += step * rs;
        --left;
    };
    if (left > 0) fetch();
    for (int f = 0; f < nframes; ++f) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int2 hdr = hdr_n;
            const float4 rec0 = rec_n0, rec1 = rec_n1;
            if (left > 0) fetch();
            // The channel's state (kind, ping-pong side) is replaced at the end of the iteration; with 1 or 2 channels every
            // thread keeps its own copy in registers, otherwise all threads must have read it before thread 0 replaces it.
            int kp, fl;
            if (kC) { kp = kind_r[c]; fl = flip_r[c]; }
            else { kp = s_kind[c]; fl = s_flip[c]; __syncthreads(); }
            const float *rot_p = s_rot + (c * 2 + fl) * maxpk;
            float *rot_n = s_rot + (c * 2 + (fl ^ 1)) * maxpk;
            float *out_c = s_out + c * half;
            int new_kind;
            if (hdr.y == 2) {
                const int npk = hdr.x;
                for (int r = tid; r < npk; r += nthr) {
                    const float4 rec = r == tid ? rec0 : r == tid + nthr ? rec1 : rec_c[r - tid];
                    const float php = rec.x, a = rec.y, adv = rec.z;
                    const int link = __float_as_int(rec.w);
                    float po;   // prev_outphase[p1]
                    if (kp == 2) po = (float)princarg_fast((double)__fadd_rn(a, rot_p[link >> 16]));   // locked_phase of that frame (:687-689)
                    else if (kp == 1) po = out_c[link & 0xffff];
                    else if (kp == 3) po = a;
                    else po = 0.f;
                    const float target = (float)princarg_fast((double)__fadd_rn(po, adv));   // :663-665
                    const float rot = (float)princarg_fast((double)__fsub_rn(target, php));   // :666-667
                    rot_n[r] = rot;
                    float sn, cs;
                    sincosf(rot, &sn, &cs);
                    csn_c[r] = make_float2(cs, sn);
                }
                new_kind = 2;
            } else if (hdr.y == 0) {
                new_kind = 3;   // pass-through (:606-616): prev_phase = prev_outphase = phase
            } else {
                const int64_t slot = (row0 + c) * g.F + f;
                lock_chain_classic(half, (const float *)(g.lock_rec + slot * rs), g.lock_map + slot * half, g.mag + slot * p.Hp, g.phase + slot * p.Hp, kp, rot_p, out_c);
                new_kind = 1;
            }
            if (kC) { kind_r[c] = new_kind; if (new_kind == 2) flip_r[c] = fl ^ 1; }
            else if (tid == 0) { s_kind[c] = new_kind; if (new_kind == 2) s_flip[c] = fl ^ 1; }
            const int64_t step = (kC == 1 || c == C - 1) ? fr_step : ch_step;
            rec_c += step * rs; csn_c += step * maxpk;
            __syncthreads();
        }
    }
    if (kC && tid == 0) {
#pragma unroll
        for (int c = 0; c < kC; ++c) { s_kind[c] = kind_r[c]; s_flip[c] = flip_r[c]; }
    }
    __syncthreads();
    for (int c = 0; c < C; ++c) {
        const int64_t row = row0 + c;
        const int kd = s_kind[c];
        if (kd == 2) for (int i = tid; i < maxpk; i += nthr) g.lock_rot[row * maxpk + i] = s_rot[(c * 2 + s_flip[c]) * maxpk + i];
        if (kd == 1) for (int i = tid; i < half; i += nthr) g.prev_out[row * half + i] = s_out[c * half + i];
        if (tid == 0) g.lock_kind[row] = kd;
    }
}
