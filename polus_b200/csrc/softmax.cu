// Masked softmax + dropout over attention scores, forward and backward (HBM-bound, one warp per row).
//
// Replaces `tf.nn.softmax(scores/sqrt(dh) + (1-mask)*-10000)` + `tf.nn.dropout` of HF
// TFBertSelfAttention, with the additive mask polus builds in polus/models.py:175-195.
#include "common.cuh"
#include <atomic>
extern std::atomic<long long> g_launch_count;

namespace {
constexpr int kWarps = 8;

template <int MAXC>
__global__ void __launch_bounds__(kWarps * 32)
softmax_fwd_kernel(const bf16* __restrict__ scores, const int32_t* __restrict__ mask, long long rows, int nh, int Sq,
                   int Sk, float scale, unsigned long long seed, uint32_t site, uint32_t thresh16, float inv_keep,
                   const uint32_t* __restrict__ d_step, bf16* __restrict__ P, bf16* __restrict__ Pd) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunks = Sk >> 3;
    const uint32_t step = thresh16 ? *d_step : 0u;
    for (long long row = (long long)blockIdx.x * kWarps + warp; row < rows; row += (long long)gridDim.x * kWarps) {
        const int b = (int)(row / ((long long)nh * Sq));
        const long long base = row * Sk;
        float v[MAXC][8];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < chunks) {
                unpack8(ld_stream8(scores + base + c * 8), v[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float x = v[i][j] * scale;
                    if (mask != nullptr) x += (1.0f - (float)__ldg(mask + (long long)b * Sk + c * 8 + j)) * -10000.0f;
                    v[i][j] = x;
                    mx = fmaxf(mx, x);
                }
            }
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i)
            if (lane + 32 * i < chunks) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[i][j] = __expf(v[i][j] - mx);
                    sum += v[i][j];
                }
            }
        const float inv = 1.0f / warp_sum(sum);
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < chunks) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[i][j] *= inv;
                *reinterpret_cast<bf16x8*>(P + base + c * 8) = pack8(v[i]);
                if (thresh16) {
                    const uint32_t keep = dropout_keep8(seed, site, step, (unsigned long long)row * chunks + c, thresh16);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[i][j] = ((keep >> j) & 1u) ? v[i][j] * inv_keep : 0.f;
                    *reinterpret_cast<bf16x8*>(Pd + base + c * 8) = pack8(v[i]);
                }
            }
        }
    }
}

template <int MAXC>
__global__ void __launch_bounds__(kWarps * 32)
softmax_bwd_kernel(bf16* __restrict__ P, bf16* __restrict__ dPd, long long rows, int Sk, float scale,
                   unsigned long long seed, uint32_t site, uint32_t thresh16, float inv_keep,
                   const uint32_t* __restrict__ d_step) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunks = Sk >> 3;
    const uint32_t step = thresh16 ? *d_step : 0u;
    for (long long row = (long long)blockIdx.x * kWarps + warp; row < rows; row += (long long)gridDim.x * kWarps) {
        const long long base = row * Sk;
        float p[MAXC][8], dp[MAXC][8];
        uint32_t keepm[MAXC];
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = lane + 32 * i;
            keepm[i] = 0xFFu;
            if (c < chunks) {
                unpack8(*reinterpret_cast<const bf16x8*>(P + base + c * 8), p[i]);
                unpack8(*reinterpret_cast<const bf16x8*>(dPd + base + c * 8), dp[i]);
                if (thresh16) {
                    keepm[i] = dropout_keep8(seed, site, step, (unsigned long long)row * chunks + c, thresh16);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dp[i][j] = ((keepm[i] >> j) & 1u) ? dp[i][j] * inv_keep : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) dot += p[i][j] * dp[i][j];
            }
        }
        dot = warp_sum(dot);
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < chunks) {
                float ds[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) ds[j] = scale * p[i][j] * (dp[i][j] - dot);
                *reinterpret_cast<bf16x8*>(dPd + base + c * 8) = pack8(ds);
                if (thresh16) {
                    float pd[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) pd[j] = ((keepm[i] >> j) & 1u) ? p[i][j] * inv_keep : 0.f;
                    *reinterpret_cast<bf16x8*>(P + base + c * 8) = pack8(pd);
                }
            }
        }
    }
}

}  // namespace

extern "C" int polus_softmax_fwd(const polus_bf16_t* scores, const int32_t* mask, int B, int nh, int Sq, int Sk,
                                 float scale, float p_drop, uint64_t seed, uint32_t site, const uint32_t* d_step,
                                 polus_bf16_t* P, polus_bf16_t* Pd, void* stream) {
    POLUS_REQUIRE(Sk > 0 && Sk % 8 == 0 && Sk <= 4096, "polus_softmax_fwd: Sk must be a multiple of 8 and <= 4096 (got %d)", Sk);
    POLUS_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "polus_softmax_fwd: bad dropout %f", p_drop);
    POLUS_REQUIRE(p_drop == 0.f || (d_step != nullptr && Pd != nullptr && Pd != P), "polus_softmax_fwd: dropout needs d_step and a separate Pd");
    const long long rows = (long long)B * nh * Sq;
    if (rows == 0) return 0;
    const uint32_t th = p_drop > 0.f ? (uint32_t)lrintf(p_drop * 65536.0f) : 0u;
    const float ik = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    long long want = (rows + kWarps - 1) / kWarps;
    const int grid = (int)(want < (long long)polus_num_sms() * 16 ? want : (long long)polus_num_sms() * 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (Sk <= 256)
        softmax_fwd_kernel<1><<<grid, kWarps * 32, 0, st>>>((const bf16*)scores, mask, rows, nh, Sq, Sk, scale, seed, site, th, ik, d_step, (bf16*)P, (bf16*)Pd);
    else if (Sk <= 512)
        softmax_fwd_kernel<2><<<grid, kWarps * 32, 0, st>>>((const bf16*)scores, mask, rows, nh, Sq, Sk, scale, seed, site, th, ik, d_step, (bf16*)P, (bf16*)Pd);
    else if (Sk <= 1024)
        softmax_fwd_kernel<4><<<grid, kWarps * 32, 0, st>>>((const bf16*)scores, mask, rows, nh, Sq, Sk, scale, seed, site, th, ik, d_step, (bf16*)P, (bf16*)Pd);
    else
        softmax_fwd_kernel<16><<<grid, kWarps * 32, 0, st>>>((const bf16*)scores, mask, rows, nh, Sq, Sk, scale, seed, site, th, ik, d_step, (bf16*)P, (bf16*)Pd);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_softmax_bwd(polus_bf16_t* P, polus_bf16_t* dPd, int B, int nh, int Sq, int Sk, float scale,
                                 float p_drop, uint64_t seed, uint32_t site, const uint32_t* d_step, void* stream) {
    POLUS_REQUIRE(Sk > 0 && Sk % 8 == 0 && Sk <= 4096, "polus_softmax_bwd: Sk must be a multiple of 8 and <= 4096 (got %d)", Sk);
    POLUS_REQUIRE(p_drop == 0.f || d_step != nullptr, "polus_softmax_bwd: dropout needs d_step");
    const long long rows = (long long)B * nh * Sq;
    if (rows == 0) return 0;
    const uint32_t th = p_drop > 0.f ? (uint32_t)lrintf(p_drop * 65536.0f) : 0u;
    const float ik = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
    long long want = (rows + kWarps - 1) / kWarps;
    const int grid = (int)(want < (long long)polus_num_sms() * 16 ? want : (long long)polus_num_sms() * 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (Sk <= 256)
        softmax_bwd_kernel<1><<<grid, kWarps * 32, 0, st>>>((bf16*)P, (bf16*)dPd, rows, Sk, scale, seed, site, th, ik, d_step);
    else if (Sk <= 512)
        softmax_bwd_kernel<2><<<grid, kWarps * 32, 0, st>>>((bf16*)P, (bf16*)dPd, rows, Sk, scale, seed, site, th, ik, d_step);
    else if (Sk <= 1024)
        softmax_bwd_kernel<4><<<grid, kWarps * 32, 0, st>>>((bf16*)P, (bf16*)dPd, rows, Sk, scale, seed, site, th, ik, d_step);
    else
        softmax_bwd_kernel<16><<<grid, kWarps * 32, 0, st>>>((bf16*)P, (bf16*)dPd, rows, Sk, scale, seed, site, th, ik, d_step);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}
