// Micro-benchmark: per-SM issue throughput of FFMA, FFMA2 (fma.rn.f32x2), MUFU.EX2, MUFU.RCP on sm_100a.
// Decides how the GELU epilogue and the attention softmax are written.  nvcc -arch=sm_100a -O3 alu.cu -o alu && ./alu
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
#define ITERS 4096
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<u64&>(a)), "l"(reinterpret_cast<u64&>(b)), "l"(reinterpret_cast<u64&>(c)));
    return d;
}
template <int MODE>
__global__ void k(float* out, float seed) {
    float a[8];
    float2 p[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; p[i] = make_float2(a[i], a[i] + 1.f); }
    const float2 b2 = make_float2(1.0001f, 0.9999f), c2 = make_float2(0.001f, -0.001f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.001f));
            if (MODE == 1) p[i] = ffma2(p[i], b2, c2);
            if (MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(p[i].x) : "f"(1.0001f), "f"(0.001f));
                             asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(p[i].y) : "f"(1.0001f), "f"(0.001f)); }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, double ops_per_iter_thread) {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 1024>>>(out, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148, 1024>>>(out, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int mhz; cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    double clk = ms * 1e-3 * mhz * 1e3;
    double per_sm_per_clk = 1024.0 * ITERS * ops_per_iter_thread / clk;
    printf("%-28s %.3f ms  %.1f thread-instr/clk/SM (at %d MHz nominal)\n", name, ms, per_sm_per_clk, mhz / 1000);
    cudaFree(out);
}
int main() {
    run<0>("FFMA", 8);
    run<1>("FFMA2 (2 flop-pairs)", 8);
    run<2>("MUFU.EX2", 8);
    run<3>("MUFU.RCP", 8);
    run<4>("EX2 + 2 FFMA interleaved", 24);
    return 0;
}
