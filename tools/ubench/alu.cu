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
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<u64&>(a)), "l"(reinterpret_cast<u64&>(b)));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<u64&>(a)), "l"(reinterpret_cast<u64&>(b)));
    return d;
}
__device__ __forceinline__ float2 fma2n(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<u64&>(a)), "l"(reinterpret_cast<u64&>(b)), "l"(reinterpret_cast<u64&>(c)));
    return d;
}
// the GEMM epilogue's GELU value + derivative for two values (polus_b200/csrc/common.cuh gelu_fwd_grad2)
__device__ __forceinline__ void gelu_pair(float2 x, float2& y, float2& d) {
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 den = fma2n(ax, make_float2(0.3275911f * 0.70710678118654752f, 0.3275911f * 0.70710678118654752f), make_float2(1.0f, 1.0f));
    const float2 arg = fma2n(mul2(x, x), make_float2(-0.7213475108146667f, -0.7213475108146667f), make_float2(-1.325748085975647f, -1.325748085975647f));
    float2 t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(arg.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(arg.y));
    float2 np = fma2n(t, make_float2(-1.3302744626998901f, -1.3302744626998901f), make_float2(1.8212559223175049f, 1.8212559223175049f));
    np = fma2n(t, np, make_float2(-1.781477928161621f, -1.781477928161621f));
    np = fma2n(t, np, make_float2(0.3565637767314911f, 0.3565637767314911f));
    np = fma2n(t, np, make_float2(-0.3193815350532532f, -0.3193815350532532f));
    const float2 r = fma2n(mul2(np, t), e, make_float2(0.5f, 0.5f));
    const float2 cdf = add2(make_float2(0.5f, 0.5f), make_float2(copysignf(r.x, x.x), copysignf(r.y, x.y)));
    y = mul2(x, cdf);
    d = fma2n(x, e, cdf);
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
            if (MODE == 5) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(p[i].x)); a[i] = __uint_as_float(r | 0x3f800000u); }
            if (MODE == 6) { float2 y, d; gelu_pair(p[i], y, d); p[i] = make_float2(y.x + d.x, y.y + d.y); }
            if (MODE == 7) { float2 y, d; gelu_pair(p[i], y, d); unsigned r0, r1;
                             asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r0) : "f"(y.y), "f"(y.x));
                             asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r1) : "f"(d.y), "f"(d.x));
                             p[i] = make_float2(__uint_as_float(r0), __uint_as_float(r1)); }
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
    run<5>("F2FP (cvt.rn.bf16x2.f32)", 8);
    run<6>("GELU value+derivative (elements)", 16);
    run<7>("GELU v+d + 2 bf16x2 packs (elements)", 16);
    return 0;
}
