// Micro-benchmark behind DESIGN.md's "warp shuffles for the intra-warp butterflies" decision (north_star FFT item; VERDICT r01
// "measure, don't argue"): the exchange between the first and the second pass of the register-tiled 1024-point complex FFT
// (16 complex values per thread, a 16 x 16 transpose among 16 consecutive lanes) done three ways, 64-thread frame groups,
// 4 CTAs of 256 threads per SM:
//   smem    what pv_fft.cuh does: 16 STS.64 into the padded buffer, a 64-thread named barrier, 16 LDS.64 (stride 16)
//   shfl    the same transpose with __shfl_sync: the value a lane must SEND in round r depends on its lane index, so the
//           registers are rotated by the lane index first (4 stages of conditional moves), 16 rounds x 2 shuffles move the
//           data, and a second rotation puts it in order -- no shared memory, no barrier
//   shfl_raw  32 shuffles per exchange and nothing else (a lower bound: what the exchange would cost if the rotations were free)
// Prints ns per exchange per SM-resident thread group and, under ncu, the data-pipe wavefronts of each variant show whether
// SHFL shares the LSU data pipe with LDS/STS on sm_100.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o exchange_bench exchange_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int fft_pad(int i) { return i + (i >> 4) + (i >> 8); }
constexpr int kPadded = 1024 + 64 + 4 + 2;

template <int MODE>
__global__ void __launch_bounds__(256, 4) k_exchange(float2 *out, int iters) {
    extern __shared__ float2 sbuf[];
    const int group = threadIdx.x / 64, t = threadIdx.x % 64, lane = threadIdx.x & 31;
    float2 *buf = sbuf + group * kPadded;
    float2 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = make_float2((float)(threadIdx.x * 16 + j), (float)blockIdx.x);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            float2 *b1 = buf + fft_pad(16 * t);
#pragma unroll
            for (int j = 0; j < 16; ++j) b1[j] = v[j];
            asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(64) : "memory");
            const int k = t & 15;
            const float2 *b2 = buf + fft_pad((t / 16) * 256 + k);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = b2[fft_pad(j * 16)];
            asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(64) : "memory");
        } else if (MODE == 1) {
            const int k = lane & 15;
            // rotate so that register r holds the element for lane (k + r) & 15: v'[r] = v[(k + r) & 15]
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const bool on = (k >> s) & 1;
                float2 w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = on ? v[(j + (1 << s)) & 15] : v[j];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = w[j];
            }
            // round r: lane k sends v'[r] (= its element for lane (k + r) & 15) and receives from lane (k - r) & 15 that lane's
            // element for k, i.e. source element index (k - r) & 15 of the transposed row
            float2 u[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int src = (lane & 16) | ((k - r) & 15);
                u[r].x = __shfl_sync(0xffffffffu, v[r].x, src);
                u[r].y = __shfl_sync(0xffffffffu, v[r].y, src);
            }
            // u[r] came from lane (k - r) & 15: rotate back so that v[j] is the element from lane j: v[j] = u[(k - j) & 15]
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = u[(16 - j) & 15];   // v[j] = u[-j]; then rotate by k
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const bool on = (k >> s) & 1;
                float2 w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = on ? v[(j - (1 << s)) & 15] : v[j];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = w[j];
            }
        } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int src = (lane & 16) | ((lane - r) & 15);
                v[r].x = __shfl_sync(0xffffffffu, v[r].x, src);
                v[r].y = __shfl_sync(0xffffffffu, v[r].y, src);
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j].x += 1.0f;   // keep a dependence between iterations
    }
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 16; ++j) { acc.x += v[j].x; acc.y += v[j].y; }
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int MODE> static float run(float2 *d_out, int grid, int iters) {
    const size_t sm = sizeof(float2) * 4 * kPadded;
    cudaFuncSetAttribute(k_exchange<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_exchange<MODE><<<grid, 256, sm>>>(d_out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k_exchange<MODE><<<grid, 256, sm>>>(d_out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 4, iters = 4000;
    float2 *d_out;
    cudaMalloc(&d_out, sizeof(float2) * grid * 256);
    const float a = run<0>(d_out, grid, iters), b = run<1>(d_out, grid, iters), c = run<2>(d_out, grid, iters);
    // one exchange moves 16 complex values per thread; a frame (64 threads) does one exchange per pass boundary
    const double frames = (double)grid * 4 * iters;
    printf("{\"sms\": %d, \"exchanges\": %.0f, \"smem_ms\": %.3f, \"shfl_transpose_ms\": %.3f, \"shfl_raw_ms\": %.3f, "
           "\"ns_per_frame_exchange\": {\"smem\": %.2f, \"shfl_transpose\": %.2f, \"shfl_raw\": %.2f}}\n",
           sms, frames, a, b, c, 1e6 * a / frames * sms, 1e6 * b / frames * sms, 1e6 * c / frames * sms);
    if (cudaGetLastError() != cudaSuccess) { printf("cuda error\n"); return 1; }
    return 0;
}
