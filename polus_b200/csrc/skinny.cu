// "Skinny" dense layer for very narrow outputs (N <= 32): the K-tag projection of the NER head
// (polus/ner/models.py:37: Dense(hidden -> output_classes), N = 4) and the 10-way head of
// tutorials/classifier_example.py:47.  A tcgen05 tile cannot take N = 4; the generic CUDA-core GEMM spent
// 1.2 ms per step on these three products (profiles/r01_launch_summary_v1.txt).  HBM-bound: x is read once.
//   fwd : y[M,N]  = act(x[M,K] . W[K,N] + b)           (one warp per row, W in shared memory)
//   bwd : dx[M,K] = dz[M,N] . W^T ; dW[K,N] += x^T . dz ; db[N] += colsum(dz)   (one pass over x and dz)
#include "common.cuh"
#include <atomic>
extern std::atomic<long long> g_launch_count;

namespace {

__device__ __forceinline__ float ldx(const void* p, long long i, int is_bf16) {
    return is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}

template <int NT>
__global__ void __launch_bounds__(256)
skinny_fwd_kernel(const void* __restrict__ x, int x_bf16, const float* __restrict__ W, const float* __restrict__ b,
                  int M, int K, int N, int act, float* __restrict__ y, float* __restrict__ z) {
    extern __shared__ float sW[];  // [K][NT]
    for (int i = threadIdx.x; i < K * NT; i += blockDim.x) {
        const int k = i / NT, n = i % NT;
        sW[i] = n < N ? W[(long long)k * N + n] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
        float acc[NT];
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[n] = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float xv = ldx(x, (long long)row * K + k, x_bf16);
#pragma unroll
            for (int n = 0; n < NT; ++n) acc[n] = fmaf(xv, sW[k * NT + n], acc[n]);
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[n] = warp_sum(acc[n]);
        if (lane < N) {
            float v = 0.f;
#pragma unroll
            for (int n = 0; n < NT; ++n) v = (lane == n) ? acc[n] : v;
            v += b ? b[lane] : 0.f;
            if (z) z[(long long)row * N + lane] = v;
            y[(long long)row * N + lane] = act_fwd(act, v);
        }
    }
}

// KPL = ceil(K/32) columns of x per lane
template <int NT, int KPL>
__global__ void __launch_bounds__(256)
skinny_bwd_kernel(const void* __restrict__ x, int x_bf16, const float* __restrict__ W, const float* __restrict__ dz,
                  int M, int K, int N, void* __restrict__ dx, int dx_bf16, float* __restrict__ gW, float* __restrict__ gb) {
    extern __shared__ float sm[];
    float* sW = sm;            // [K][NT]
    float* sG = sm + K * NT;   // [K][NT] block accumulator
    float* sB = sG + K * NT;   // [NT]
    for (int i = threadIdx.x; i < K * NT; i += blockDim.x) {
        const int k = i / NT, n = i % NT;
        sW[i] = n < N ? W[(long long)k * N + n] : 0.f;
        sG[i] = 0.f;
    }
    if (threadIdx.x < NT) sB[threadIdx.x] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float gw[KPL][NT];
    float gbl[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        gbl[n] = 0.f;
#pragma unroll
        for (int i = 0; i < KPL; ++i) gw[i][n] = 0.f;
    }
    for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
        float d[NT];
#pragma unroll
        for (int n = 0; n < NT; ++n) d[n] = n < N ? __ldg(dz + (long long)row * N + n) : 0.f;  // warp broadcast
#pragma unroll
        for (int n = 0; n < NT; ++n) gbl[n] += d[n];
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int k = lane + 32 * i;
            if (k < K) {
                const float xv = ldx(x, (long long)row * K + k, x_bf16);
                float g = 0.f;
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    gw[i][n] = fmaf(xv, d[n], gw[i][n]);
                    g = fmaf(d[n], sW[k * NT + n], g);
                }
                if (dx != nullptr) {
                    if (dx_bf16) reinterpret_cast<bf16*>(dx)[(long long)row * K + k] = __float2bfloat16(g);
                    else reinterpret_cast<float*>(dx)[(long long)row * K + k] = g;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        const int k = lane + 32 * i;
        if (k < K) {
#pragma unroll
            for (int n = 0; n < NT; ++n) atomicAdd(&sG[k * NT + n], gw[i][n]);
        }
    }
    if (lane == 0) {  // every lane holds the same column sums of its warp's rows
#pragma unroll
        for (int n = 0; n < NT; ++n) atomicAdd(&sB[n], gbl[n]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * NT; i += blockDim.x) {
        const int k = i / NT, n = i % NT;
        if (n < N) atomicAdd(gW + (long long)k * N + n, sG[i]);
    }
    if (gb != nullptr && threadIdx.x < N) atomicAdd(gb + threadIdx.x, sB[threadIdx.x]);
}

int pick_nt(int N) { return N <= 4 ? 4 : (N <= 8 ? 8 : (N <= 16 ? 16 : 32)); }

}  // namespace

extern "C" int polus_skinny_supported(int K, int N) {
    if (N < 1 || N > 32 || K < 1) return 0;
    const int nt = pick_nt(N);
    const int kpl = (K + 31) / 32;
    if (kpl * nt > 128) return 0;                       // per-lane dW accumulators must stay in registers
    if ((size_t)(2 * K * nt + nt) * 4 > 160 * 1024) return 0;
    return 1;
}

extern "C" int polus_skinny_fwd(const void* x, int x_dtype, const float* W, const float* b, int M, int K, int N, int act,
                                float* y, float* z, void* stream) {
    POLUS_REQUIRE(polus_skinny_supported(K, N), "polus_skinny_fwd: unsupported K=%d N=%d", K, N);
    POLUS_REQUIRE(x_dtype == POLUS_F32 || x_dtype == POLUS_BF16, "polus_skinny_fwd: x must be f32 or bf16");
    if (M == 0) return 0;
    const int nt = pick_nt(N);
    const size_t smem = (size_t)K * nt * 4;
    int grid = cdiv(M, 8);
    const int cap = polus_num_sms() * 8;
    if (grid > cap) grid = cap;
    cudaStream_t st = (cudaStream_t)stream;
#define SK_FWD(NT_)                                                                                                   \
    {                                                                                                                 \
        if (smem > 48 * 1024) POLUS_CHECK_CUDA(cudaFuncSetAttribute(skinny_fwd_kernel<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        skinny_fwd_kernel<NT_><<<grid, 256, smem, st>>>(x, x_dtype == POLUS_BF16, W, b, M, K, N, act, y, z);            \
    }
    if (nt == 4) SK_FWD(4) else if (nt == 8) SK_FWD(8) else if (nt == 16) SK_FWD(16) else SK_FWD(32)
#undef SK_FWD
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_skinny_bwd(const void* x, int x_dtype, const float* W, const float* dz, int M, int K, int N, void* dx,
                                int dx_dtype, float* gW, float* gb, void* stream) {
    POLUS_REQUIRE(polus_skinny_supported(K, N), "polus_skinny_bwd: unsupported K=%d N=%d", K, N);
    POLUS_REQUIRE(gW != nullptr, "polus_skinny_bwd: gW required");
    if (M == 0) return 0;
    const int nt = pick_nt(N);
    const int kpl = (K + 31) / 32;
    const size_t smem = (size_t)(2 * K * nt + nt) * 4;
    int grid = cdiv(M, 64);  // >= 8 rows per warp so the block-level atomics amortise
    const int cap = polus_num_sms() * 2;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    cudaStream_t st = (cudaStream_t)stream;
    const int xb = x_dtype == POLUS_BF16, dxb = dx_dtype == POLUS_BF16;
#define SK_BWD(NT_, KPL_)                                                                                             \
    {                                                                                                                 \
        if (smem > 48 * 1024) POLUS_CHECK_CUDA(cudaFuncSetAttribute(skinny_bwd_kernel<NT_, KPL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        skinny_bwd_kernel<NT_, KPL_><<<grid, 256, smem, st>>>(x, xb, W, dz, M, K, N, dx, dxb, gW, gb);                  \
    }
#define SK_BWD_K(NT_)                                                                                                 \
    {                                                                                                                 \
        if (kpl <= 1) SK_BWD(NT_, 1) else if (kpl <= 2) SK_BWD(NT_, 2) else if (kpl <= 4) SK_BWD(NT_, 4)               \
        else if (kpl * NT_ <= 32) SK_BWD(NT_, 32 / NT_) else if (kpl * NT_ <= 64) SK_BWD(NT_, 64 / NT_) else SK_BWD(NT_, 128 / NT_) \
    }
    if (nt == 4) SK_BWD_K(4) else if (nt == 8) SK_BWD_K(8) else if (nt == 16) SK_BWD_K(16) else SK_BWD_K(32)
#undef SK_BWD_K
#undef SK_BWD
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}
