// Cross-entropy family with fused gradient (HBM-bound, tiny): one warp per row.
//   kind 0  tf.keras.losses.SparseCategoricalCrossentropy(from_logits=True)  (tutorials/classifier_example.py:55)
//   kind 1  polus.losses.weighted_softmax_cross_entropy_from_logits          (polus/losses.py:5-18)
//   kind 2  polus.losses.weighted_sigmoid_cross_entropy_from_logits          (polus/losses.py:21-42)
// loss = mean over rows (tf.reduce_mean); glogits = d loss / d logits.
#include "common.cuh"
#include <atomic>
extern std::atomic<long long> g_launch_count;

namespace {

__global__ void xent_kernel(int kind, const float* __restrict__ logits, const void* __restrict__ labels,
                            const float* __restrict__ cw, float negw, int rows, int C, float* __restrict__ loss,
                            float* __restrict__ glogits) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* x = logits + (long long)row * C;
    float* g = glogits ? glogits + (long long)row * C : nullptr;
    const float inv_rows = 1.0f / (float)rows;
    float row_loss;
    if (kind == 2) {
        // sum_c [max(x,0) - x*z + log1p(exp(-|x|))] * (sum_c w_c z_c + negw*[all z == 0])
        const float* z = reinterpret_cast<const float*>(labels) + (long long)row * C;
        float l = 0.f, wsum = 0.f, any = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float xv = x[c], zv = z[c];
            l += fmaxf(xv, 0.f) - xv * zv + log1pf(expf(-fabsf(xv)));
            wsum += (cw ? cw[c] : 1.0f) * zv;
            any += (zv != 0.f) ? 1.f : 0.f;
        }
        l = warp_sum(l);
        wsum = warp_sum(wsum);
        any = warp_sum(any);
        const float w = wsum + (any == 0.f ? negw : 0.f);
        row_loss = l * w;
        if (g)
            for (int c = lane; c < C; c += 32) g[c] = (1.0f / (1.0f + expf(-x[c])) - z[c]) * w * inv_rows;
    } else {
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += expf(x[c] - mx);
        const float lse = mx + logf(warp_sum(se));
        if (kind == 0) {
            int y = reinterpret_cast<const int32_t*>(labels)[row];
            y = y < 0 ? 0 : (y >= C ? C - 1 : y);
            row_loss = lse - x[y];
            if (g)
                for (int c = lane; c < C; c += 32) g[c] = (expf(x[c] - lse) - (c == y ? 1.0f : 0.f)) * inv_rows;
        } else {
            // CE(z, x) = sum_c z_c * (lse - x_c); weight = sum_c w_c z_c
            const float* z = reinterpret_cast<const float*>(labels) + (long long)row * C;
            float ce = 0.f, wsum = 0.f, zsum = 0.f;
            for (int c = lane; c < C; c += 32) {
                ce += z[c] * (lse - x[c]);
                wsum += (cw ? cw[c] : 1.0f) * z[c];
                zsum += z[c];
            }
            ce = warp_sum(ce);
            wsum = warp_sum(wsum);
            zsum = warp_sum(zsum);
            row_loss = ce * wsum;
            if (g)
                for (int c = lane; c < C; c += 32) g[c] = (expf(x[c] - lse) * zsum - z[c]) * wsum * inv_rows;
        }
    }
    if (lane == 0 && loss) atomicAdd(loss, row_loss * inv_rows);
}

}  // namespace

extern "C" int polus_xent(int kind, const float* logits, const void* labels, const float* cw, float negw, int rows,
                          int C, float* loss, float* glogits, void* stream) {
    POLUS_REQUIRE(kind >= 0 && kind <= 2, "polus_xent: kind must be 0, 1 or 2");
    POLUS_REQUIRE(C >= 1, "polus_xent: C must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    if (loss) POLUS_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    if (rows == 0) return 0;
    xent_kernel<<<cdiv(rows, 8), 256, 0, st>>>(kind, logits, labels, cw, negw, rows, C, loss, glogits);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}
