// Fused multi-tensor Adam over the flat parameter arena (HBM-bound: 4 g + 8 m + 8 v + 8 p + 2 shadow +
// 4 zeroed g = 34 B/param; SURVEY section 8d counts 30 B without the gradient zeroing).
//
// Replaces the ~200 per-variable ResourceApplyAdam launches behind `optimizer.apply_gradients`
// (polus/training.py:191; Keras Adam in tutorials/classifier_example.py:54, HF AdamWeightDecay via
// polus/schedulers.py:2).  Keras semantics: t = iterations+1, lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
// p -= lr_t*m/(sqrt(v)+eps) -- epsilon OUTSIDE the bias-corrected sqrt.  The learning-rate schedule
// of polus/schedulers.py:5-23 (HF WarmUp over PolynomialDecay(power=1)) is evaluated on the device
// from the step counter, so a replayed CUDA graph needs no host value per step.
#include "common.cuh"
#include <atomic>
extern std::atomic<long long> g_launch_count;

namespace {

__device__ __forceinline__ float schedule_lr(const polus_adam_cfg_t& c, uint32_t it) {
    if (c.schedule == 0) return c.lr;
    const float step = (float)it;
    if ((int)it < c.warmup_steps) return c.lr * (step / (float)c.warmup_steps);
    const float d = fminf(step - (float)c.warmup_steps, (float)c.decay_steps);
    return (c.lr - c.end_lr) * (1.0f - d / (float)c.decay_steps) + c.end_lr;
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            bf16* __restrict__ pb, const uint8_t* __restrict__ decay_mask, long long n, polus_adam_cfg_t c,
            const float* __restrict__ d_hyper, const uint32_t* __restrict__ d_step) {
    // {lr, grad_scale, weight_decay, end_lr} live in device memory when the caller passes d_hyper: a captured graph
    // then follows optimizer.learning_rate.assign() (polus/training.py:90-94) without being re-captured
    if (d_hyper != nullptr) { c.lr = d_hyper[0]; c.grad_scale = d_hyper[1]; c.weight_decay = d_hyper[2]; c.end_lr = d_hyper[3]; }
    const uint32_t it = *d_step;
    const float t = (float)(it + 1);
    const float lr = schedule_lr(c, it);
    const float lr_t = lr * sqrtf(1.0f - powf(c.beta2, t)) / (1.0f - powf(c.beta1, t));
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 gv = reinterpret_cast<float4*>(g)[i];
        float4 mv = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float4 pv = reinterpret_cast<float4*>(p)[i];
        float ga[4] = {gv.x, gv.y, gv.z, gv.w}, ma[4] = {mv.x, mv.y, mv.z, mv.w};
        float va[4] = {vv.x, vv.y, vv.z, vv.w}, pa[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gj = ga[j] * c.grad_scale;
            if (c.weight_decay > 0.f && (decay_mask == nullptr || decay_mask[i * 4 + j])) pa[j] -= lr * c.weight_decay * pa[j];
            ma[j] = c.beta1 * ma[j] + (1.0f - c.beta1) * gj;
            va[j] = c.beta2 * va[j] + (1.0f - c.beta2) * gj * gj;
            pa[j] -= lr_t * ma[j] / (sqrtf(va[j]) + c.eps);
        }
        reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
        reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
        reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pb != nullptr) {
            reinterpret_cast<__nv_bfloat162*>(pb)[2 * i] = __floats2bfloat162_rn(pa[0], pa[1]);
            reinterpret_cast<__nv_bfloat162*>(pb)[2 * i + 1] = __floats2bfloat162_rn(pa[2], pa[3]);
        }
    }
    // tail (n % 4)
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        const float gj = g[i] * c.grad_scale;
        float pj = p[i];
        if (c.weight_decay > 0.f && (decay_mask == nullptr || decay_mask[i])) pj -= lr * c.weight_decay * pj;
        const float mj = c.beta1 * m[i] + (1.0f - c.beta1) * gj;
        const float vj = c.beta2 * v[i] + (1.0f - c.beta2) * gj * gj;
        pj -= lr_t * mj / (sqrtf(vj) + c.eps);
        m[i] = mj; v[i] = vj; p[i] = pj; g[i] = 0.f;
        if (pb != nullptr) pb[i] = __float2bfloat16(pj);
    }
}

__global__ void step_inc_kernel(uint32_t* d_step) { *d_step += 1; }

}  // namespace

extern "C" int polus_adam(float* p, float* g, float* m, float* v, polus_bf16_t* pb, const uint8_t* decay_mask,
                          int64_t n, const polus_adam_cfg_t* cfg, const float* d_hyper, uint32_t* d_step, int increment_step,
                          void* stream) {
    POLUS_REQUIRE(cfg != nullptr && d_step != nullptr, "polus_adam: cfg and d_step required");
    POLUS_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0, "polus_adam: arenas must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (n > 0) {
        long long blocks = ((n >> 2) + 255) / 256;
        long long cap = (long long)polus_num_sms() * 8;
        adam_kernel<<<(int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap)), 256, 0, st>>>(p, g, m, v, (bf16*)pb, decay_mask, n, *cfg, d_hyper, d_step);
        g_launch_count++;
        POLUS_LAUNCH_CHECK();
    }
    if (increment_step) {
        step_inc_kernel<<<1, 1, 0, st>>>(d_step);
        g_launch_count++;
        POLUS_LAUNCH_CHECK();
    }
    return 0;
}
