// Micro-benchmark: where should the resampler's warp-uniform sinc "quads" come from?  The filter loop of k_ola_resample
// (pv_resample.cuh) per tap and warp step: one warp-uniform float4 (the quad), ROWS = 4 conflict-free 32-bit shared loads
// (the rows' input samples) and 8 FFMA2.  ncu shows the L1 data pipe saturated in that loop (a uniform LDS.128 costs two
// wavefronts, every row load one).  Variants of the quad source:
//   smem    LDS.128 broadcast (what the kernel does)
//   param   a 16 KB table passed by value as a __grid_constant__ kernel parameter, indexed per warp: LDC from constant bank 0
//   ldg     __ldg of the global table (L1-cached; same data pipe as shared memory -- control)
// Reported: ns per tap step per SM and tap steps per clock per SM; 4 CTAs of 256 threads per SM, 96 taps, 8 phases.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o quad_source_bench quad_source_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kTaps = 96, kOV = 8, kTab = kTaps * kOV + 8, kIn = 4096 + 128;
struct QuadTable { float4 q[1024]; };

__device__ __forceinline__ void ffma2_bcast(float2 &acc, float x, float tx, float ty) {
    unsigned long long a, b, c;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(x));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(tx), "f"(ty));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(c));
}

template <int MODE>
__global__ void __launch_bounds__(256, 4) k(const __grid_constant__ QuadTable qt, const float4 *__restrict__ gtab, const unsigned *__restrict__ steps, float *out, int nsteps) {
    extern __shared__ float4 smem4[];
    float4 *s_quad = smem4;
    float *s_x = (float *)(smem4 + kTab);
    for (int i = threadIdx.x; i < kTab; i += blockDim.x) s_quad[i] = gtab[i];
    for (int i = threadIdx.x; i < kIn; i += blockDim.x) s_x[i] = (float)(i & 255) * 1e-3f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float total = 0.f;
    for (int s = warp; s < nsteps; s += 8) {
        const unsigned desc = __ldg(&steps[s]);
        const int bucket = (int)(desc >> 24), start = (int)(desc & 0xfff);
        const int qoff = 4 + kOV - bucket;
        const float *xs[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xs[u] = s_x + start + lane + 33 * 32 * u / 4;   // lanes on distinct banks
        float2 acc[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc[u][0] = make_float2(0.f, 0.f); acc[u][1] = make_float2(0.f, 0.f); }
        auto quad_at = [&](int tap) -> float4 {
            if (MODE == 0) return s_quad[qoff + tap * kOV];
            if (MODE == 1) return qt.q[qoff + tap * kOV];
            return __ldg(&gtab[qoff + tap * kOV]);
        };
#pragma unroll 1
        for (int j = 0; j < kTaps; j += 4) {
            float4 tq[4];
            float x[4][4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                tq[jj] = quad_at(j + jj);
#pragma unroll
                for (int u = 0; u < 4; ++u) x[jj][u] = xs[u][j + jj];
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ffma2_bcast(acc[u][0], x[jj][u], tq[jj].x, tq[jj].y);
                    ffma2_bcast(acc[u][1], x[jj][u], tq[jj].z, tq[jj].w);
                }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) total += acc[u][0].x + acc[u][0].y + acc[u][1].x + acc[u][1].y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
}

template <int MODE> static void run(const char *name, const QuadTable &qt, const float4 *gtab, const unsigned *steps, float *out, int sms, int khz) {
    const int nsteps = 4096, grid = sms * 4;
    const size_t sm = sizeof(float4) * kTab + sizeof(float) * kIn;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256, sm>>>(qt, gtab, steps, out, 64);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, 256, sm>>>(qt, gtab, steps, out, nsteps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double tap_steps_per_sm = 4.0 * nsteps * kTaps;   // 4 CTAs per SM, each nsteps warp steps of kTaps taps
    const double clocks = ms * 1e-3 * khz * 1e3;
    printf("{\"quads_from\": \"%s\", \"ms\": %.3f, \"clocks_per_tap_step_per_sm\": %.2f, \"fp32_pipe_floor_clocks\": 4.0, \"err\": \"%s\"}\n", name, ms, clocks / tap_steps_per_sm,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    static QuadTable qt;
    for (int i = 0; i < 1024; ++i) qt.q[i] = make_float4(1e-3f * i, 2e-3f * i, -1e-3f * i, 5e-4f * i);
    float4 *gtab; cudaMalloc(&gtab, sizeof(qt)); cudaMemcpy(gtab, &qt, sizeof(qt), cudaMemcpyHostToDevice);
    unsigned hs[4096];
    for (int i = 0; i < 4096; ++i) hs[i] = ((unsigned)(i * 5 % 8) << 24) | (unsigned)((i * 37) % 3000);
    unsigned *steps; cudaMalloc(&steps, sizeof(hs)); cudaMemcpy(steps, hs, sizeof(hs), cudaMemcpyHostToDevice);
    float *out; cudaMalloc(&out, sizeof(float) * pr.multiProcessorCount * 4 * 256);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_nominal\": %d}\n", pr.name, pr.multiProcessorCount, khz);
    for (int r = 0; r < 2; ++r) {
        run<0>("shared memory (LDS.128 broadcast)", qt, gtab, steps, out, pr.multiProcessorCount, khz);
        run<1>("kernel parameter (LDC, constant bank 0)", qt, gtab, steps, out, pr.multiProcessorCount, khz);
        run<2>("global table (__ldg)", qt, gtab, steps, out, pr.multiProcessorCount, khz);
    }
    return 0;
}
