// CUDA-core GEMM with the same contract as polus_gemm_tc, for shapes a tcgen05 tile cannot take
// (the K=4 tag projection of polus/ner/models.py:37, the 10-way head of tutorials/classifier_example.py:47)
// and as the on-device checker the GPU tests compare polus_gemm_tc against.
#include "common.cuh"
#include <atomic>
extern std::atomic<long long> g_launch_count;

namespace {

struct SmallParams {
    int M, N, K, batch0;
    const void* A;
    const void* B;
    long long lda, abs0, abs1, ldb, bbs0, bbs1;
    int a_mn, b_mn, a_bf16, b_bf16;
    void* C;
    void* C2;
    long long ldc, cbs0, cbs1;
    int c_f32;
    const float* bias;
    float alpha;
    int act;
    int accumulate;
    int c2_kind;
    const bf16* emul;
    float* colsum;
};

__device__ __forceinline__ float load_elem(const void* p, long long idx, int is_bf16) {
    return is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p)[idx])
                   : reinterpret_cast<const float*>(p)[idx];
}

constexpr int TS = 32;

__global__ void __launch_bounds__(256) gemm_small_kernel(const SmallParams p) {
    __shared__ float sA[TS][TS + 1];  // [m][k]
    __shared__ float sB[TS][TS + 1];  // [n][k]
    const int b = blockIdx.z;
    const int b0 = b % p.batch0, b1 = b / p.batch0;
    const long long a_off = (long long)b0 * p.abs0 + (long long)b1 * p.abs1;
    const long long b_off = (long long)b0 * p.bbs0 + (long long)b1 * p.bbs1;
    const long long c_off = (long long)b0 * p.cbs0 + (long long)b1 * p.cbs1;
    const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty in [0,8)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < p.K; k0 += TS) {
        // load tiles; pick the thread mapping that keeps the contiguous dim on tx
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty + 8 * i;
            {
                int m, k;
                if (p.a_mn) { m = m0 + tx; k = k0 + r; } else { m = m0 + r; k = k0 + tx; }
                float v = 0.f;
                if (m < p.M && k < p.K)
                    v = load_elem(p.A, a_off + (p.a_mn ? (long long)k * p.lda + m : (long long)m * p.lda + k), p.a_bf16);
                if (p.a_mn) sA[tx][r] = v; else sA[r][tx] = v;
            }
            {
                int n, k;
                if (p.b_mn) { n = n0 + tx; k = k0 + r; } else { n = n0 + r; k = k0 + tx; }
                float v = 0.f;
                if (n < p.N && k < p.K)
                    v = load_elem(p.B, b_off + (p.b_mn ? (long long)k * p.ldb + n : (long long)n * p.ldb + k), p.b_bf16);
                if (p.b_mn) sB[tx][r] = v; else sB[r][tx] = v;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < TS; ++k) {
            const float bv = sB[tx][k];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(sA[ty + 8 * i][k], bv, acc[i]);
        }
        __syncthreads();
    }
    const int n = n0 + tx;
    if (n >= p.N) return;
    const float bias = p.bias ? p.bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty + 8 * i;
        if (m >= p.M) continue;
        float z = acc[i] * p.alpha + bias;
        const long long idx = c_off + (long long)m * p.ldc + n;
        if (p.emul) z *= __bfloat162float(p.emul[idx]);
        if (p.C2) {
            const float z2 = p.c2_kind == 1 ? act_grad(p.act, z) : z;
            if (p.c_f32) reinterpret_cast<float*>(p.C2)[idx] = z2;
            else reinterpret_cast<bf16*>(p.C2)[idx] = __float2bfloat16(z2);
        }
        float y = act_fwd(p.act, z);
        if (p.colsum) atomicAdd(p.colsum + n, p.c_f32 ? y : __bfloat162float(__float2bfloat16(y)));
        if (p.c_f32) {
            float* c = reinterpret_cast<float*>(p.C) + idx;
            *c = p.accumulate ? *c + y : y;
        } else {
            reinterpret_cast<bf16*>(p.C)[idx] = __float2bfloat16(y);
        }
    }
}

}  // namespace

extern "C" int polus_gemm_small(const polus_gemm_t* g, void* stream) {
    POLUS_REQUIRE(g->M >= 1 && g->N >= 1 && g->K >= 1, "polus_gemm_small: empty problem");
    POLUS_REQUIRE((g->A.dtype == POLUS_F32 || g->A.dtype == POLUS_BF16) &&
                      (g->B.dtype == POLUS_F32 || g->B.dtype == POLUS_BF16),
                  "polus_gemm_small: operands must be f32 or bf16");
    POLUS_REQUIRE(g->c_dtype == POLUS_F32 || g->c_dtype == POLUS_BF16, "polus_gemm_small: C dtype");
    POLUS_REQUIRE(!(g->accumulate && g->c_dtype != POLUS_F32), "polus_gemm_small: accumulate needs fp32 C");
    SmallParams p;
    p.M = g->M; p.N = g->N; p.K = g->K;
    p.batch0 = g->batch0 < 1 ? 1 : g->batch0;
    const int batch1 = g->batch1 < 1 ? 1 : g->batch1;
    p.A = g->A.ptr; p.B = g->B.ptr;
    p.lda = g->A.ld; p.abs0 = g->A.bs0; p.abs1 = g->A.bs1;
    p.ldb = g->B.ld; p.bbs0 = g->B.bs0; p.bbs1 = g->B.bs1;
    p.a_mn = g->A.mn_major; p.b_mn = g->B.mn_major;
    p.a_bf16 = g->A.dtype == POLUS_BF16; p.b_bf16 = g->B.dtype == POLUS_BF16;
    p.C = g->C; p.C2 = g->C2; p.ldc = g->ldc; p.cbs0 = g->cbs0; p.cbs1 = g->cbs1;
    p.c_f32 = g->c_dtype == POLUS_F32;
    p.bias = g->bias; p.alpha = g->alpha; p.act = g->act; p.accumulate = g->accumulate;
    p.c2_kind = g->c2_kind; p.emul = (const bf16*)g->Emul; p.colsum = g->colsum;
    dim3 grid(cdiv(g->N, TS), cdiv(g->M, TS), p.batch0 * batch1);
    POLUS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "polus_gemm_small: problem too large (M=%d batch=%d)", g->M, (int)grid.z);
    gemm_small_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}
