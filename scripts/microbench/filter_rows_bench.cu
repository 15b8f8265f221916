// Micro-benchmark for a warp-specialised resampler: how fast can 16 consumer warps per SM (4 CTAs x 4 warps, the other half of
// each CTA's warps being gather producers) run the filter loop, as a function of the rows per warp step (the warp-uniform quad
// is loaded once per tap per step: more rows = fewer shared-memory wavefronts per output) and the software-pipeline depth?
// Compile once per register budget:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -maxrregcount=R -DREGS=R ...
// Reported: SM clocks per tap per 128 outputs (the unit of quad_source_bench: 6.7 with 32 warps per SM and 4 rows; fp32 floor 4.0).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kTaps = 96, kOV = 8, kTab = kTaps * kOV + 8, kIn = 4096 + 512;
#ifndef REGS
#define REGS 64
#endif

__device__ __forceinline__ void ffma2_bcast(float2 &acc, float x, float tx, float ty) {
    unsigned long long a, b, c;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(x));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(tx), "f"(ty));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(c));
}

// ROWS rows per step; TS taps per pipeline stage; two stages in flight (as pv_resample.cuh)
template <int ROWS, int TS, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 4) k(const float4 *__restrict__ gtab, const unsigned *__restrict__ steps, float *out, int nsteps) {
    extern __shared__ float4 smem4[];
    float4 *s_quad = smem4;
    float *s_x = (float *)(smem4 + kTab);
    for (int i = threadIdx.x; i < kTab; i += blockDim.x) s_quad[i] = gtab[i];
    for (int i = threadIdx.x; i < kIn; i += blockDim.x) s_x[i] = (float)(i & 255) * 1e-3f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float total = 0.f;
    for (int s = warp; s < nsteps; s += WARPS) {
        const unsigned desc = __ldg(&steps[s]);
        const int bucket = (int)(desc >> 24), start = (int)(desc & 0xfff);
        const int qoff = 4 + kOV - bucket;
        const float *xs[ROWS];
#pragma unroll
        for (int u = 0; u < ROWS; ++u) xs[u] = s_x + start + lane + 33 * 8 * u;
        float2 acc[ROWS][2];
#pragma unroll
        for (int u = 0; u < ROWS; ++u) { acc[u][0] = make_float2(0.f, 0.f); acc[u][1] = make_float2(0.f, 0.f); }
        float4 tqa[TS], tqb[TS];
        float xa[TS][ROWS], xb[TS][ROWS];
        auto load = [&](int j, float4 (&tq)[TS], float (&x)[TS][ROWS]) {
#pragma unroll
            for (int jj = 0; jj < TS; ++jj) {
                tq[jj] = s_quad[qoff + (j + jj) * kOV];
#pragma unroll
                for (int u = 0; u < ROWS; ++u) x[jj][u] = xs[u][j + jj];
            }
        };
        auto fma = [&](const float4 (&tq)[TS], const float (&x)[TS][ROWS]) {
#pragma unroll
            for (int jj = 0; jj < TS; ++jj)
#pragma unroll
                for (int u = 0; u < ROWS; ++u) {
                    ffma2_bcast(acc[u][0], x[jj][u], tq[jj].x, tq[jj].y);
                    ffma2_bcast(acc[u][1], x[jj][u], tq[jj].z, tq[jj].w);
                }
        };
        load(0, tqa, xa);
#pragma unroll 1
        for (int j = 0; j < kTaps; j += 2 * TS) {
            load(j + TS, tqb, xb);
            fma(tqa, xa);
            if (j + 2 * TS < kTaps) load(j + 2 * TS, tqa, xa);
            fma(tqb, xb);
        }
#pragma unroll
        for (int u = 0; u < ROWS; ++u) total += acc[u][0].x + acc[u][0].y + acc[u][1].x + acc[u][1].y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
}

template <int ROWS, int TS, int WARPS> static void run(const float4 *gtab, const unsigned *steps, float *out, int sms, int khz) {
    const int nsteps = 2048 * 4 / ROWS, grid = sms * 4;   // the same number of outputs for every ROWS
    const size_t sm = sizeof(float4) * kTab + sizeof(float) * kIn;
    cudaFuncSetAttribute(k<ROWS, TS, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<ROWS, TS, WARPS>);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ROWS, TS, WARPS><<<grid, 32 * WARPS, sm>>>(gtab, steps, out, 64);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<ROWS, TS, WARPS><<<grid, 32 * WARPS, sm>>>(gtab, steps, out, nsteps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double units = 4.0 * nsteps * kTaps * ROWS / 4.0;   // taps x 128 outputs per SM
    const double clocks = ms * 1e-3 * khz * 1e3;
    printf("{\"maxrregcount\": %d, \"warps_per_sm\": %d, \"rows_per_step\": %d, \"taps_per_stage\": %d, \"regs\": %d, \"spill_bytes\": %zu, \"clocks_per_tap_per_128_outputs\": %.2f, \"err\": \"%s\"}\n",
           REGS, 4 * WARPS, ROWS, TS, fa.numRegs, fa.localSizeBytes, clocks / units, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    static float4 tab[1024];
    for (int i = 0; i < 1024; ++i) tab[i] = make_float4(1e-3f * i, 2e-3f * i, -1e-3f * i, 5e-4f * i);
    float4 *gtab; cudaMalloc(&gtab, sizeof(tab)); cudaMemcpy(gtab, tab, sizeof(tab), cudaMemcpyHostToDevice);
    unsigned hs[4096];
    for (int i = 0; i < 4096; ++i) hs[i] = ((unsigned)(i * 5 % 8) << 24) | (unsigned)((i * 37) % 2400);
    unsigned *steps; cudaMalloc(&steps, sizeof(hs)); cudaMemcpy(steps, hs, sizeof(hs), cudaMemcpyHostToDevice);
    float *out; cudaMalloc(&out, sizeof(float) * pr.multiProcessorCount * 4 * 256);
    const int sms = pr.multiProcessorCount;
    run<4, 2, 8>(gtab, steps, out, sms, khz);
    run<4, 2, 4>(gtab, steps, out, sms, khz);
    run<4, 4, 4>(gtab, steps, out, sms, khz);
    run<6, 2, 4>(gtab, steps, out, sms, khz);
    run<8, 1, 4>(gtab, steps, out, sms, khz);
    run<8, 2, 4>(gtab, steps, out, sms, khz);
    run<8, 2, 8>(gtab, steps, out, sms, khz);
    return 0;
}
