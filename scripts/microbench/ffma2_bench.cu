// Micro-benchmark: does the packed fp32 instruction of sm_100 (FFMA2 / FADD2 / FMUL2, PTX fma/add/mul.rn.f32x2) raise the fp32
// rate per issue slot?  Every thread keeps 16 independent accumulator pairs and runs a long unrolled chain; the grid fills the
// GPU (4 CTAs of 256 threads per SM, as the product kernels).  Reported: FMAs (or adds / muls) per clock per SM and warp
// instructions per clock per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

// MODE 0: 32 scalar FFMA   1: 16 FFMA2 (all operands pairs)   2: 16 FFMA2 with a scalar-broadcast multiplicand
// MODE 3: 32 scalar FADD   4: 16 FADD2   5: 32 scalar FMUL+FADD pairs (complex-style, unfused) 6: FMUL2 + FADD2
template <int MODE>
__global__ void __launch_bounds__(256, 4) k(float *out, int iters, float s0, float s1) {
    float a[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = (float)(threadIdx.x + i);
    float x = s0, y = s1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) a[i] = __fmaf_rn(a[i], x, y);
            } else if (MODE == 1) {
                const u64 xx = pack(x, y), yy = pack(y, x);
#pragma unroll
                for (int i = 0; i < 16; ++i) { u64 v = pack(a[2 * i], a[2 * i + 1]); asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(xx), "l"(yy)); unpack(v, a[2 * i], a[2 * i + 1]); }
            } else if (MODE == 2) {
                const u64 xx = pack(x, x), yy = pack(y, x);
#pragma unroll
                for (int i = 0; i < 16; ++i) { u64 v = pack(a[2 * i], a[2 * i + 1]); asm("fma.rn.f32x2 %0, %1, %0, %2;" : "+l"(v) : "l"(xx), "l"(yy)); unpack(v, a[2 * i], a[2 * i + 1]); }
            } else if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 32; ++i) a[i] = __fadd_rn(a[i], x);
            } else if (MODE == 4) {
                const u64 xx = pack(x, y);
#pragma unroll
                for (int i = 0; i < 16; ++i) { u64 v = pack(a[2 * i], a[2 * i + 1]); asm("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(xx)); unpack(v, a[2 * i], a[2 * i + 1]); }
            } else if (MODE == 5) {
#pragma unroll
                for (int i = 0; i < 32; ++i) a[i] = __fadd_rn(__fmul_rn(a[i], x), y);
            } else {
                const u64 xx = pack(x, y), yy = pack(y, x);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    u64 v = pack(a[2 * i], a[2 * i + 1]);
                    asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(xx));
                    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(yy));
                    unpack(v, a[2 * i], a[2 * i + 1]);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> static void run(const char *name, int flop_instr_per_rep, int warp_instr_per_rep, float *out, int sms, int khz) {
    const int iters = 2000, grid = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 10, 0.999f, 1e-3f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double clocks = ms * 1e-3 * khz * 1e3;
    const double warps = 4.0 * 8;                      // warps per SM
    const double ops = warps * 32 * (double)iters * 4 * flop_instr_per_rep;   // scalar fp ops per SM
    const double wi = warps * (double)iters * 4 * warp_instr_per_rep;
    printf("{\"variant\": \"%s\", \"ms\": %.3f, \"fp_ops_per_clk_per_sm\": %.1f, \"warp_instr_per_clk_per_sm\": %.2f, \"err\": \"%s\"}\n", name, ms, ops / clocks, wi / clocks,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, sizeof(float) * pr.multiProcessorCount * 4 * 256);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_nominal\": %d}\n", pr.name, pr.multiProcessorCount, khz);
    for (int r = 0; r < 2; ++r) {
        run<0>("ffma scalar", 32, 32, out, pr.multiProcessorCount, khz);
        run<1>("ffma2 pairs", 32, 16, out, pr.multiProcessorCount, khz);
        run<2>("ffma2 scalar-broadcast operand", 32, 16, out, pr.multiProcessorCount, khz);
        run<3>("fadd scalar", 32, 32, out, pr.multiProcessorCount, khz);
        run<4>("fadd2 pairs", 32, 16, out, pr.multiProcessorCount, khz);
        run<5>("fmul+fadd scalar", 64, 64, out, pr.multiProcessorCount, khz);
        run<6>("fmul2+fadd2 pairs", 64, 32, out, pr.multiProcessorCount, khz);
    }
    return 0;
}
