// Speex quality-4 windowed-sinc resampler from a shared-memory input window (speex/resample.c:352-403 direct table,
// :462-560 cubic-interpolated table) -- the second half of k_ola_resample and of the fused k_synth_ola.
//
// The host has bucketed the outputs of a run of slices by their table phase (`offset` in resample.c:470), with buckets
// starting at multiples of kResBlock (ResampleRun, pv_kernels.cuh).  A warp step therefore works on one phase: the sinc
// "quad" (tab[e-2], tab[e-1], tab[e], tab[e+1]) of every tap is a single broadcast shared-memory access that feeds
// 4 x kResPerThread FMAs, and the host orders the entries so that the 32 input windows of a warp row start on different banks.
#pragma once
#include "pv_kernels.cuh"
#include "pv_synth.cuh"

namespace pvgpu {

// s_x[i + x_shift] is the normalised-stream sample an entry's packed tap-0 position i refers to; out positions are relative
// to orow (element index of the run's first output in g.out); outputs at or beyond out_limit are not stored.
template <int OV>   // sinc-table oversampling of the interpolated mode; 0 = direct table
__device__ __forceinline__ void resample_run(const DevPlan &p, const DevRows &g, const ResampleRun &hdr, const float4 *s_quad, const float *s_x, int x_shift,
                                             int64_t orow, int64_t out_limit, const unsigned *__restrict__ rs_ent, const float *__restrict__ rs_frac, int L) {
    constexpr int nb = OV > 0 ? OV : 1;
    const unsigned *__restrict__ ent_tab = rs_ent + hdr.ent_off;
    const float *__restrict__ frac_tab = rs_frac + hdr.ent_off;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    for (int blk = warp * kResBlock; blk < hdr.padded; blk += nwarp * kResBlock) {
        int bucket = 0;
#pragma unroll
        for (int q = 1; q < nb; ++q) bucket += hdr.start[q] <= blk;   // warp-uniform
        unsigned ent[kResPerThread];
        const float *xs[kResPerThread];
        bool live[kResPerThread];
#pragma unroll
        for (int u = 0; u < kResPerThread; ++u) {
            ent[u] = __ldg(&ent_tab[blk + lane + 32 * u]);
            live[u] = ent[u] != 0xffffffffu && (int64_t)(ent[u] & 0xffffu) < out_limit;
            // tap 0; dead entries read (and discard) from a safe place
            xs[u] = live[u] ? s_x + ((int)(ent[u] >> 16) + x_shift) : s_x + (kResPad + x_shift);
        }
        if (OV == 0) {
#pragma unroll
            for (int u = 0; u < kResPerThread; ++u) {
                if (!live[u]) continue;
                float sum = 0.f;
                const float *__restrict__ tt = p.rs_table + (size_t)__float_as_uint(__ldg(&frac_tab[blk + lane + 32 * u])) * L;
                for (int j = 0; j < L; ++j) sum += xs[u][j] * __ldg(&tt[j]);
                pcm_store(g.out, g.fmt, orow + (ent[u] & 0xffffu), sum);
            }
        } else {
            const int qoff = 4 + OV - bucket;
            auto quad_at = [&](int tap) -> float4 { return s_quad[qoff + tap * OV]; };   // warp-uniform: a broadcast load
            float acc[kResPerThread][4];
#pragma unroll
            for (int u = 0; u < kResPerThread; ++u) { acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f; }
            // two taps per stage, two stages in flight: the loads of the next stage are issued before the 32 FMAs of the
            // current one (filt_len is a multiple of 4, resample.c:712)
            float4 tqa[2], tqb[2];
            float xa[2][kResPerThread], xb[2][kResPerThread];
            auto load2 = [&](int j, float4 (&tq)[2], float (&x)[2][kResPerThread]) {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    tq[jj] = quad_at(j + jj);
#pragma unroll
                    for (int u = 0; u < kResPerThread; ++u) x[jj][u] = xs[u][j + jj];
                }
            };
            auto fma2 = [&](const float4 (&tq)[2], const float (&x)[2][kResPerThread]) {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int u = 0; u < kResPerThread; ++u) {
                        acc[u][0] += x[jj][u] * tq[jj].x;
                        acc[u][1] += x[jj][u] * tq[jj].y;
                        acc[u][2] += x[jj][u] * tq[jj].z;
                        acc[u][3] += x[jj][u] * tq[jj].w;
                    }
            };
            load2(0, tqa, xa);
#pragma unroll 1
            for (int j = 0; j < L; j += 4) {
                load2(j + 2, tqb, xb);
                fma2(tqa, xa);
                if (j + 4 < L) load2(j + 4, tqa, xa);
                fma2(tqb, xb);
            }
#pragma unroll
            for (int u = 0; u < kResPerThread; ++u) {
                if (!live[u]) continue;
                const float frac = __ldg(&frac_tab[blk + lane + 32 * u]);
                // cubic_coef (resample.c:339-351)
                const float i0 = -0.16667f * frac + 0.16667f * frac * frac * frac;
                const float i1 = frac + 0.5f * frac * frac - 0.5f * frac * frac * frac;
                const float i3 = -0.33333f * frac + 0.5f * frac * frac - 0.16667f * frac * frac * frac;
                const float i2 = (float)(1. - i0 - i1 - i3);
                pcm_store(g.out, g.fmt, orow + (ent[u] & 0xffffu), (i0 * acc[u][0]) + (i1 * acc[u][1]) + (i2 * acc[u][2]) + (i3 * acc[u][3]));
            }
        }
    }
}

}  // namespace pvgpu
