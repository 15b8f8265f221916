// k_ola_resample_ws: the overlap-add + normalisation + Speex resampler stage (k_ola_resample, pv_kernels.cu) as a persistent,
// warp-specialised kernel.
//
// Why: k_ola_resample runs its two halves one after the other in every CTA -- a memory-latency-bound gather (the frames covering
// the run, ~190 KB from the frame ring) and then an fp32-pipe-bound filter (96 taps x 4 accumulators per output).  ncu's
// per-instruction samples put 38 % of the kernel's time in the gather + normalisation (10 % of its instructions), 14 % in the
// per-CTA set-up and 40 % in the filter loops: with four CTAs per SM in arbitrary phases the SM is too often left with nothing
// but waiting warps.  Here every CTA is both at once: 4 producer warps build the normalised stream of work item i+1 into one
// of two shared-memory input windows while 4 consumer warps filter item i from the other, and the CTA walks a contiguous list
// of (row, run) items, so the sinc quads are staged once per CTA instead of once per run.
//
//   producers                                   consumers
//   wait empty[b]   (items >= 2)                wait full[b]
//   tables of the item, OLA gather, / norm      resample_run from s_in[b]
//   arrive full[b]                              arrive empty[b]   (if an item i+2 follows)
//
// Arithmetic and its order are k_ola_resample's (same gather loop, same resample_run): outputs are bit-identical.
#include <cstdlib>

#include "pv_fused.cuh"
#include "pv_kernels.cuh"
#include "pv_resample.cuh"

namespace pvgpu {

namespace {
constexpr int kWsSlices = 48;   // = kOlaMaxSlices / kOlaMaxFrames of pv_kernels.cu (ola_max_table_slices / _frames report them)
constexpr int kWsFrames = 96;
constexpr int kProdWarps = 4, kConsWarps = 4;
constexpr int kProd = 32 * kProdWarps, kCons = 32 * kConsWarps;
enum { kBarFull0 = 1, kBarEmpty0 = 3, kBarProd = 5 };

struct WsTables {
    int res_rel[kWsSlices + 1];
    int ola_rel[kWsSlices];
    int j0[kWsSlices];
    int fr_off[kWsFrames];
    int fr_pos[kWsFrames];
};
}  // namespace

template <int OV>
__global__ void __launch_bounds__(256, 4) k_ola_resample_ws(const DevPlan p, const DevRows g, const SliceRec *__restrict__ recs, const float *__restrict__ norm,
                                                         int64_t norm_base, long recs_base, long k0, int nf, int run, int max_in, const ResampleRun *__restrict__ runs,
                                                         const unsigned *__restrict__ rs_ent, const float *__restrict__ rs_frac, const unsigned *__restrict__ rs_steps,
                                                         long run_origin, int runs_per_row, int items_per_cta, int total_items, int dbg) {
    extern __shared__ float4 smem4[];
    __shared__ WsTables T;
    const int L = (int)p.rs_filt_len;
    const bool quad = !p.rs_direct;
    float4 *s_quad = smem4;
    const int win = L + max_in;   // floats of one input window: L zeros (history before the stream) + the run's span
    float *s_win = (float *)(smem4 + (quad ? p.rs_table_len : 0));
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * win; i += blockDim.x) s_win[i] = 0.f;
    if (quad) {
        const float4 *__restrict__ tab4 = p.rs_quads;
        for (int e = tid; e < p.rs_table_len; e += blockDim.x) s_quad[e] = __ldg(&tab4[e]);
    }
    __syncthreads();

    const int item0 = blockIdx.x * items_per_cta;
    const int n_items = min(items_per_cta, total_items - item0);
    if (n_items <= 0) return;
    const int N = p.N;
    const SliceRec *__restrict__ rr = recs - recs_base;

    if (tid < kProd) {
        // ------------------------------------------------ producers ------------------------------------------------
        const int lane = tid & 31, warp = tid >> 5;
        for (int it = 0; it < n_items; ++it) {
            const int item = item0 + it, b = it & 1;
            const int row = item / runs_per_row;
            const long ka = k0 + (long)(item - row * runs_per_row) * run;
            const long kb = min(ka + run, k0 + (long)nf);
            float *s_in = s_win + b * win + L;
            // the slices whose normalised samples the run needs (its own and the resampler history before it)
            const int64_t u_lo_raw = rr[ka].res_off + rr[ka].rs_last - L + 1;
            long kmin = ka;
            while (kmin > recs_base && kmin > 0 && rr[kmin].res_off > u_lo_raw && ka - kmin < kWsSlices - run - 1) --kmin;
            const long jmin = rr[kmin].jlo;
            const int nsl = (int)(kb - kmin);
            const int nfr = min((int)(kb - jmin), kWsFrames);
            const int64_t ola_base = rr[jmin].ola_off;
            const int64_t u_lo = u_lo_raw < 0 ? 0 : u_lo_raw;
            const int64_t u_hi = rr[kb - 1].res_off + ((rr[kb - 1].flags & 1) ? 0 : rr[kb - 1].consumed);
            const int span = (int)min(u_hi - u_lo, (int64_t)max_in);
            named_sync(kBarProd, kProd);   // every producer is done with the previous item's tables
            for (int n = tid; n < nsl; n += kProd) {
                const SliceRec &r = rr[kmin + n];
                T.res_rel[n] = (int)(r.res_off - u_lo);
                T.ola_rel[n] = (int)(r.ola_off - ola_base);
                T.j0[n] = (int)(r.jlo - jmin);
            }
            if (tid == 0) T.res_rel[nsl] = span;
            for (int i = tid; i < nfr; i += kProd) {
                const long j = jmin + i;
                T.fr_off[i] = (int)(rr[j].ola_off - ola_base);
                T.fr_pos[i] = (int)(j % g.Fr) * N;
            }
            if (it >= 2) named_sync(kBarEmpty0 + b, kProd + kCons);   // the consumers have finished with this window
            named_sync(kBarProd, kProd);

            const float *__restrict__ fr = g.frames + (int64_t)row * g.Fr * N;
            const float *__restrict__ nrm = norm + (ola_base - norm_base);
            const int nchunk = (max_in / run + 127) >> 7;
            const int nitem = nsl * nchunk;
            for (int w = warp; w < ((dbg & 1) ? 0 : nitem); w += kProdWarps) {
                const int sl = w / nchunk, eb0 = (w - sl * nchunk) << 7;
                const int r0 = T.res_rel[sl];
                const int e_lo = max(r0, 0) + eb0, e_hi = min(T.res_rel[sl + 1], span);
                if (e_lo >= e_hi) continue;
                const int rel0 = T.ola_rel[sl] - r0;
                const int jb = (int)(kmin - jmin) + sl;
                const int e0 = e_lo + lane;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                for (int j = T.j0[sl]; j <= jb; j += 4) {   // frames in slice order = the reference's accumulator sequence
                    float v[4][4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int jc = min(j + jj, jb);
                        const int o = rel0 + e0 - T.fr_off[jc];
                        const float *__restrict__ src = fr + (T.fr_pos[jc] + o);
                        const int lim = j + jj <= jb ? N : 0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[jj][u] = o + 32 * u < lim ? src[32 * u] : 0.f;
                    }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                        for (int u = 0; u < 4; ++u) acc[u] += v[jj][u];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + 32 * u;
                    if (e >= e_hi) break;
                    s_in[e] = acc[u] / nrm[rel0 + e];
                }
            }
            named_arrive(kBarFull0 + b, kProd + kCons);
        }
    } else {
        // ------------------------------------------------ consumers ------------------------------------------------
        const int ctid = tid - kProd;
        for (int it = 0; it < n_items; ++it) {
            const int item = item0 + it, b = it & 1;
            const int row = item / runs_per_row;
            const long ka = k0 + (long)(item - row * runs_per_row) * run;
            const ResampleRun *__restrict__ hp = &runs[(ka - run_origin) / run];
            ResampleRun hdr;
            hdr.ent_off = __ldg(&hp->ent_off);
            hdr.step_off = __ldg(&hp->step_off);
            hdr.n_steps = __ldg(&hp->n_steps);
            const int64_t out_first = __ldg((const long long *)&hp->out_first);
            const int64_t row_out = (int64_t)row * g.out_stride - g.out_base;
            const int64_t orow = row_out + out_first;
            const int64_t out_limit = g.n_out[row] - out_first;
            const float *s_in = s_win + b * win + L;
            named_sync(kBarFull0 + b, kProd + kCons);
            if (!(dbg & 2)) resample_run<OV>(p, g, hdr, s_quad, s_in, -kResPad, orow, out_limit, rs_ent, rs_frac, rs_steps, L, ctid >> 5, kConsWarps);
            if (it + 2 < n_items) named_arrive(kBarEmpty0 + b, kProd + kCons);
        }
    }
}

template <int OV>
static cudaError_t launch_t(const DevPlan &p, const DevRows &g, const SliceRec *recs, const float *norm, int64_t norm_base, long recs_base, long k0, int nframes,
                            int run, int max_in, size_t sm, const ResampleRun *runs, const unsigned *rs_ent, const float *rs_frac, const unsigned *rs_steps,
                            long run_origin, cudaStream_t st) {
    static bool configured[16] = {};
    static int sms[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_ola_resample_ws<OV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
        if (e != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int nsm = dev < 16 && sms[dev] > 0 ? sms[dev] : 148;
    const int runs_per_row = (nframes + run - 1) / run;
    const long total = (long)runs_per_row * g.rows;
    if (total <= 0) return cudaSuccess;
    // one wave of resident CTAs (4 per SM), every CTA a contiguous stretch of the item list
    static const int dbg_env = []() { const char *v = getenv("PVGPU_OLA_WS_DBG"); return v ? atoi(v) : 0; }();   // timing experiments: 1 no gather, 2 no filter
    const int slots = nsm * 4;
    const int per = (int)((total + slots - 1) / slots);
    const int grid = (int)((total + per - 1) / per);
    k_ola_resample_ws<OV><<<grid, 256, sm, st>>>(p, g, recs, norm, norm_base, recs_base, k0, nframes, run, max_in, runs, rs_ent, rs_frac, rs_steps, run_origin,
                                               runs_per_row, per, (int)total, dbg_env);
    return cudaGetLastError();
}

// false: this plan / run shape has no warp-specialised kernel (the caller launches k_ola_resample)
bool launch_ola_resample_ws(const DevPlan &p, const DevRows &g, const SliceRec *recs, const float *norm, int64_t norm_base, long recs_base, long k0, int nframes,
                            int run, int max_consumed, const ResampleRun *runs, const unsigned *rs_ent, const float *rs_frac, const unsigned *rs_steps,
                            long run_origin, cudaStream_t st, cudaError_t *err) {
    *err = cudaSuccess;
    if (!p.rs_active || run > kWsSlices - 16) return false;
    const int L = (int)p.rs_filt_len;
    const bool quad = !p.rs_direct;
    const int max_in = ((run * max_consumed + L + 8) + 3) & ~3;
    const size_t sm = (quad ? sizeof(float4) * (size_t)p.rs_table_len : 0) + sizeof(float) * 2 * (size_t)(max_in + L);
    if (sm > (size_t)54 * 1024) return false;   // four CTAs per SM (with their static tables and the per-CTA reserve) or not at all
#define PV_WS(OVV) *err = launch_t<OVV>(p, g, recs, norm, norm_base, recs_base, k0, nframes, run, max_in, sm, runs, rs_ent, rs_frac, rs_steps, run_origin, st)
    if (!quad) PV_WS(0);
    else if (p.rs_oversample == 8) PV_WS(8);
    else if (p.rs_oversample == 4) PV_WS(4);
    else if (p.rs_oversample == 2) PV_WS(2);
    else PV_WS(1);
#undef PV_WS
    return true;
}

}  // namespace pvgpu
