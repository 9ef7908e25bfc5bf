// Linear-chain CRF: negative log-likelihood with gradients (forward-backward in one launch) and Viterbi
// decode.  One warp per sequence, lane j owns tag j (K <= 32), transitions in shared memory, the
// T-step recurrence kept in registers + warp shuffles.
//
// Replaces tensorflow_addons.text.crf_log_likelihood / crf_decode, which the reference calls from
// polus/layers.py:78-80 (decode, also run in training), :91-96 and :109-114 (loss); TF executes them
// as a tf.scan / RNN loop of ~6 tiny kernels per time step.
//
// tfa semantics restated here (and in oracle/crf.py):
//   score(b)    = sum_{t<len} x[t,y_t] + sum_{t<len-1} A[y_t,y_{t+1}]
//   alpha_0     = x_0 ; alpha_t[j] = x_t[j] + logsumexp_i(alpha_{t-1}[i] + A[i,j]) ; logZ = logsumexp_j alpha_{len-1}[j]
//   decode      : delta_t[j] = x_t[j] + max_i(delta_{t-1}[i] + A[i,j]), argmax ties -> lowest index,
//                 positions >= len decode to tag 0.
#include "common.cuh"
#include <atomic>
extern std::atomic<long long> g_launch_count;

namespace {

__device__ __forceinline__ float lse2(float a, float b) {
    const float m = fmaxf(a, b);
    if (m == -INFINITY) return -INFINITY;
    return m + logf(expf(a - m) + expf(b - m));
}

// gemis doubles as scratch for the alphas: forward pass writes alpha_t there, backward pass replaces
// each entry by the gradient.
__global__ void crf_nll_kernel(const float* __restrict__ emis, const int32_t* __restrict__ tags,
                               const int32_t* __restrict__ lens, const float* __restrict__ trans,
                               const float* __restrict__ weights, int B, int T, int K, float* __restrict__ nll_out,
                               float* __restrict__ loss, float* __restrict__ gemis, float* __restrict__ gtrans) {
    extern __shared__ float sm[];
    float* sA = sm;           // [K][K]
    float* sG = sm + K * K;   // [K][K] gradient accumulator for this block's sequence
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    for (int i = lane; i < K * K; i += 32) {
        sA[i] = trans[i];
        sG[i] = 0.f;
    }
    __syncwarp();
    int len = lens ? lens[b] : T;
    len = len < 0 ? 0 : (len > T ? T : len);
    const float w = (weights ? weights[b] : 1.0f);
    const float gscale = w / (float)B;
    const float* x = emis + (long long)b * T * K;
    const int32_t* y = tags + (long long)b * T;
    float* g = gemis + (long long)b * T * K;
    const bool act = lane < K;

    if (len == 0) {
        for (int i = lane; i < T * K; i += 32) g[i] = 0.f;
        if (lane == 0 && nll_out) nll_out[b] = 0.f;
        return;
    }
    // ---- forward: alphas
    float alpha = act ? x[lane] : -INFINITY;
    if (act) g[lane] = alpha;
    for (int t = 1; t < len; ++t) {
        float m = -INFINITY;
        for (int i = 0; i < K; ++i) {
            const float ai = __shfl_sync(0xffffffffu, alpha, i);
            if (act) m = fmaxf(m, ai + sA[i * K + lane]);
        }
        float s = 0.f;
        for (int i = 0; i < K; ++i) {
            const float ai = __shfl_sync(0xffffffffu, alpha, i);
            if (act) s += expf(ai + sA[i * K + lane] - m);
        }
        alpha = act ? x[t * K + lane] + m + logf(s) : -INFINITY;
        if (act) g[t * K + lane] = alpha;
    }
    float mz = warp_max(alpha);
    float logZ = mz + logf(warp_sum(act ? expf(alpha - mz) : 0.f));
    // ---- gold path score
    float sc = 0.f;
    for (int t = lane; t < len; t += 32) {
        const int yt = min(max(y[t], 0), K - 1);
        sc += x[t * K + yt];
        if (t + 1 < len) sc += sA[yt * K + min(max(y[t + 1], 0), K - 1)];
    }
    sc = warp_sum(sc);
    const float nll = logZ - sc;
    if (lane == 0) {
        if (nll_out) nll_out[b] = nll;
        if (loss) atomicAdd(loss, nll * gscale);
    }
    // ---- backward: betas, marginals, gradients
    __syncwarp();
    float beta = act ? 0.f : -INFINITY;  // beta_{len-1}
    for (int t = len - 1; t >= 0; --t) {
        const float a_t = act ? g[t * K + lane] : -INFINITY;
        const int yt = min(max(y[t], 0), K - 1);
        if (act) {
            const float marg = expf(a_t + beta - logZ);
            g[t * K + lane] = gscale * (marg - (lane == yt ? 1.0f : 0.f));
        }
        if (t > 0) {
            // pairwise marginals for (t-1 -> t): exp(alpha_{t-1}[i] + A[i,j] + x_t[j] + beta_t[j] - logZ)
            const float xb = act ? x[t * K + lane] + beta : -INFINITY;  // indexed by j = lane
            float newbeta = -INFINITY;                                 // beta_{t-1}[i = lane]
            for (int j = 0; j < K; ++j) {
                const float xbj = __shfl_sync(0xffffffffu, xb, j);
                if (act) newbeta = lse2(newbeta, sA[lane * K + j] + xbj);
            }
            const float a_prev = act ? g[(t - 1) * K + lane] : -INFINITY;  // alpha_{t-1}[i = lane]
            for (int j = 0; j < K; ++j) {
                const float xbj = __shfl_sync(0xffffffffu, xb, j);
                if (act) sG[lane * K + j] += expf(a_prev + sA[lane * K + j] + xbj - logZ);
            }
            __syncwarp();
            if (lane == 0) {
                const int yp = min(max(y[t - 1], 0), K - 1);
                sG[yp * K + yt] -= 1.0f;
            }
            __syncwarp();
            beta = newbeta;
        }
    }
    for (int i = len * K + lane; i < T * K; i += 32) g[i] = 0.f;
    __syncwarp();
    if (gtrans != nullptr)
        for (int i = lane; i < K * K; i += 32) atomicAdd(gtrans + i, gscale * sG[i]);
}

__global__ void crf_decode_kernel(const float* __restrict__ emis, const int32_t* __restrict__ lens,
                                  const float* __restrict__ trans, int T, int K, int32_t* __restrict__ tags_out,
                                  float* __restrict__ score_out) {
    extern __shared__ float sm[];
    float* sA = sm;                                              // [K][K]
    uint8_t* bp = reinterpret_cast<uint8_t*>(sm + K * K);        // [T][K] back-pointers
    const int b = blockIdx.x, lane = threadIdx.x;
    for (int i = lane; i < K * K; i += 32) sA[i] = trans[i];
    __syncwarp();
    int len = lens ? lens[b] : T;
    len = len < 0 ? 0 : (len > T ? T : len);
    const float* x = emis + (long long)b * T * K;
    int32_t* out = tags_out + (long long)b * T;
    const bool act = lane < K;
    for (int t = len + lane; t < T; t += 32) out[t] = 0;
    if (len == 0) {
        if (lane == 0 && score_out) score_out[b] = 0.f;
        return;
    }
    float delta = act ? x[lane] : -INFINITY;
    for (int t = 1; t < len; ++t) {
        float best = -INFINITY;
        int arg = 0;
        for (int i = 0; i < K; ++i) {
            const float di = __shfl_sync(0xffffffffu, delta, i);
            if (act) {
                const float c = di + sA[i * K + lane];
                if (c > best) {  // strict: ties keep the lowest i (tf.argmax)
                    best = c;
                    arg = i;
                }
            }
        }
        if (act) {
            bp[t * K + lane] = (uint8_t)arg;
            delta = x[t * K + lane] + best;
        }
    }
    // final argmax over j, ties -> lowest j
    float bv = delta;
    int bi = act ? lane : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
        }
    }
    __syncwarp();
    if (lane == 0) {
        if (score_out) score_out[b] = bv;
        int cur = bi;
        out[len - 1] = cur;
        for (int t = len - 1; t >= 1; --t) {
            cur = bp[t * K + cur];
            out[t - 1] = cur;
        }
    }
}

__global__ void crf_mask_kernel(const float* __restrict__ trans, const float* __restrict__ mask, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = trans[i] * mask[i] + (float)((int)(1.0f - mask[i]) * -10000);
}

// CRF.loss_sample_weights (polus/layers.py:116-121): w_b = any(y*m == 1) + all(y*m == 0) * negative_weight
__global__ void crf_sample_weights_kernel(const float* __restrict__ y, const float* __restrict__ maskpos, float negw,
                                          int T, int K, float* __restrict__ out) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const float* yb = y + (long long)b * T * K;
    int any_one = 0, all_zero = 1;
    for (int i = lane; i < T * K; i += 32) {
        const float v = yb[i] * maskpos[i % K];
        any_one |= (v == 1.0f);
        all_zero &= (v == 0.0f);
    }
    any_one = __any_sync(0xffffffffu, any_one);
    all_zero = __all_sync(0xffffffffu, all_zero);
    if (lane == 0) out[b] = (any_one ? 1.0f : 0.0f) + (all_zero ? negw : 0.0f);
}

}  // namespace

extern "C" int polus_crf_sample_weights(const float* y_true, const float* mask_positive, float negative_weight, int B,
                                        int T, int K, float* out, void* stream) {
    if (B == 0) return 0;
    crf_sample_weights_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(y_true, mask_positive, negative_weight, T, K, out);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_crf_nll(const float* emis, const int32_t* tags, const int32_t* lens, const float* trans,
                             const float* weights, int B, int T, int K, float* nll, float* loss, float* gemis,
                             float* gtrans, void* stream) {
    POLUS_REQUIRE(K >= 1 && K <= 32, "polus_crf_nll: K must be in [1,32] (got %d)", K);
    POLUS_REQUIRE(gemis != nullptr, "polus_crf_nll: gemis is required (it doubles as the alpha scratch)");
    POLUS_REQUIRE(T >= 1, "polus_crf_nll: T must be >= 1");
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (loss) POLUS_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    crf_nll_kernel<<<B, 32, 2 * K * K * sizeof(float), st>>>(emis, tags, lens, trans, weights, B, T, K, nll, loss, gemis, gtrans);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_crf_decode(const float* emis, const int32_t* lens, const float* trans, int B, int T, int K,
                                int32_t* tags, float* score, void* stream) {
    POLUS_REQUIRE(K >= 1 && K <= 32, "polus_crf_decode: K must be in [1,32] (got %d)", K);
    POLUS_REQUIRE(T >= 1, "polus_crf_decode: T must be >= 1");
    const size_t smem = (size_t)K * K * sizeof(float) + (size_t)T * K;
    POLUS_REQUIRE(smem <= 200 * 1024, "polus_crf_decode: T*K=%d too large for the shared-memory back-pointer table", T * K);
    if (B == 0) return 0;
    if (smem > 48 * 1024) POLUS_CHECK_CUDA(cudaFuncSetAttribute(crf_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    crf_decode_kernel<<<B, 32, smem, (cudaStream_t)stream>>>(emis, lens, trans, T, K, tags, score);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_crf_mask_transitions(const float* trans, const float* mask, int K, float* out, void* stream) {
    crf_mask_kernel<<<cdiv(K * K, 128), 128, 0, (cudaStream_t)stream>>>(trans, mask, K * K, out);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}
