// Small HBM-bound utilities: activation backward + bias-gradient column sums, dropout, casts, fp32
// element-wise / reduction helpers used by the heads, losses and metrics
// (polus/ner/models.py:26-67, polus/losses.py, polus/metrics.py:55-63, polus/models.py:148-150).
#include "common.cuh"
#include <atomic>
extern std::atomic<long long> g_launch_count;

namespace {

constexpr int kColsumSlabs = 128;

// grid (ceil(N/256), slabs); CTA = 8 warps; lane owns 8 consecutive columns, warps stride over rows.
__global__ void __launch_bounds__(256)
act_bwd_colsum_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ z, int M, int N, int act,
                      bf16* __restrict__ dz, float* __restrict__ partial, int rows_per_slab) {
    __shared__ float red[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 256 + lane * 8;
    const int r0 = blockIdx.y * rows_per_slab;
    const int r1 = min(M, r0 + rows_per_slab);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    pdl_trigger();
    pdl_wait();
    if (col < N) {
        // 4 rows per iteration: all loads are issued before the (MUFU-heavy) activation-gradient math
        constexpr int U = 4;
        for (int r = r0 + warp; r < r1; r += 8 * U) {
            bf16x8 gq[U], zq[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int rr = r + 8 * u;
                if (rr < r1) {
                    const long long off = (long long)rr * N + col;
                    gq[u] = ld_stream8(dy + off);
                    if (act == POLUS_ACT_DERIV_U8) {   // 8 one-byte derivatives (polus_gemm_t.c2_kind = 2): the first two words
                        const uint2 w = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(z) + off));
                        uint32_t* zu = reinterpret_cast<uint32_t*>(&zq[u]);
                        zu[0] = w.x;
                        zu[1] = w.y;
                    } else if (act != POLUS_ACT_NONE) zq[u] = ld_stream8(z + off);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int rr = r + 8 * u;
                if (rr < r1) {
                    float g[8];
                    unpack8(gq[u], g);
                    if (act == POLUS_ACT_DERIV_U8) {
                        float zv[8];
                        const uint32_t* zu = reinterpret_cast<const uint32_t*>(&zq[u]);
                        d8_unpack4(zu[0], zv);
                        d8_unpack4(zu[1], zv + 4);
#pragma unroll
                        for (int j = 0; j < 8; ++j) g[j] *= zv[j];
                    } else if (act != POLUS_ACT_NONE) {
                        float zv[8];
                        unpack8(zq[u], zv);
#pragma unroll
                        for (int j = 0; j < 8; ++j) g[j] *= (act == POLUS_ACT_DERIV ? zv[j] : act_grad(act, zv[j]));
                    }
                    if (dz != nullptr) *reinterpret_cast<bf16x8*>(dz + (long long)rr * N + col) = pack8(g);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += g[j];
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int c = threadIdx.x;
    if (partial != nullptr && blockIdx.x * 256 + c < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][c];
        // `partial` is the bias-gradient vector itself: one fp32 reduction per column per slab (<= 128 per address);
        // replaces a workspace + second kernel (74 extra launches per step, profiles/r01_launch_summary_v3_fused_attn.txt)
        atomicAdd(partial + blockIdx.x * 256 + c, s);
    }
}

// out[c] += sum_r partial[r][c]; columns < split go to out0, the rest to out1 (either may be NULL).
// Block = 32 columns x 8 row groups: coalesced 128-byte row reads, shared-memory tree for the 8 groups.
__global__ void __launch_bounds__(256)
colsum_reduce_kernel(const float* __restrict__ partial, int rows, int cols, float* __restrict__ out0,
                     float* __restrict__ out1, int split) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (c < cols)
        for (int r = ty; r < rows; r += 8) s += partial[(long long)r * cols + c];
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += red[g][tx];
        if (c < split) {
            if (out0) out0[c] += t;
        } else if (out1) {
            out1[c - split] += t;
        }
    }
}

__global__ void dropout_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n8,
                               unsigned long long seed, uint32_t site, uint32_t thresh16, float inv_keep,
                               const uint32_t* __restrict__ d_step) {
    const uint32_t step = *d_step;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        unpack8(*reinterpret_cast<const bf16x8*>(x + i * 8), v);
        const uint32_t keep = dropout_keep8(seed, site, step, (unsigned long long)i, thresh16);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * inv_keep : 0.f;
        *reinterpret_cast<bf16x8*>(y + i * 8) = pack8(v);
    }
}

template <typename S, typename D>
__device__ __forceinline__ D conv(S v);
template <> __device__ __forceinline__ float conv<float, float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 conv<float, bf16>(float v) { return __float2bfloat16(v); }
template <> __device__ __forceinline__ float conv<bf16, float>(bf16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float conv<int32_t, float>(int32_t v) { return (float)v; }
template <> __device__ __forceinline__ int32_t conv<float, int32_t>(float v) { return (int32_t)v; }
template <> __device__ __forceinline__ float conv<uint8_t, float>(uint8_t v) { return (float)v; }
template <> __device__ __forceinline__ bf16 conv<uint8_t, bf16>(uint8_t v) { return __float2bfloat16((float)v); }
template <> __device__ __forceinline__ bf16 conv<int32_t, bf16>(int32_t v) { return __float2bfloat16((float)v); }

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        d[i] = conv<S, D>(s[i]);
}

__global__ void fill_kernel(float* d, float v, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) d[i] = v;
}

__global__ void binary_kernel(int op, const float* __restrict__ a, const float* __restrict__ b, long long n, long long bn,
                              float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = a[i];
        const float y = b[bn == n ? i : (bn == 1 ? 0 : i % bn)];
        float r;
        switch (op) {
            case 0: r = x + y; break;
            case 1: r = x - y; break;
            case 2: r = x * y; break;
            default: r = x / y; break;
        }
        out[i] = r;
    }
}

__device__ __forceinline__ float unary_fwd(int op, float x, float alpha) {
    switch (op) {
        case 16: return expf(x);
        case 17: return logf(x);
        case 18: return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
        case 19: return 1.0f / (1.0f + expf(-x));
        case 20: return -x;
        case 21: return x * x;
        case 22: return x * alpha;
        default: return act_fwd(op, x);
    }
}
__device__ __forceinline__ float unary_grad(int op, float x, float alpha) {
    switch (op) {
        case 16: return expf(x);
        case 17: return 1.0f / x;
        case 18: return 1.0f / (1.0f + expf(-x));
        case 19: { float s = 1.0f / (1.0f + expf(-x)); return s * (1.0f - s); }
        case 20: return -1.0f;
        case 21: return 2.0f * x;
        case 22: return alpha;
        default: return act_grad(op, x);
    }
}
__global__ void unary_kernel(int op, const float* __restrict__ x, const float* __restrict__ dy, int grad, long long n,
                             float* __restrict__ out, float alpha) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = grad ? dy[i] * unary_grad(op, x[i], alpha) : unary_fwd(op, x[i], alpha);
}

// axis=1: one warp per row
__global__ void reduce_rows_kernel(const float* __restrict__ x, int rows, int cols, float scale, float* __restrict__ out, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += x[(long long)row * cols + c];
    s = warp_sum(s) * scale;
    if (lane == 0) out[row] = accumulate ? out[row] + s : s;
}
// axis=0: one thread per column (coalesced across threads), deterministic
__global__ void reduce_cols_kernel(const float* __restrict__ x, int rows, int cols, float scale, float* __restrict__ out, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += x[(long long)r * cols + c];
    s *= scale;
    out[c] = accumulate ? out[c] + s : s;
}

__global__ void argmax_kernel(const float* __restrict__ x, int rows, int cols, int32_t* __restrict__ out) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const float* p = x + (long long)row * cols;
    float best = p[0];
    int bi = 0;
    for (int c = 1; c < cols; ++c)
        if (p[c] > best) { best = p[c]; bi = c; }  // ties -> lowest index (tf.argmax)
    out[row] = bi;
}

__global__ void one_hot_kernel(const int32_t* __restrict__ idx, int rows, int cols, float* __restrict__ out) {
    const long long n = (long long)rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (idx[i / cols] == (int)(i % cols)) ? 1.0f : 0.0f;
}

__global__ void confusion_kernel(const int32_t* __restrict__ yt, const int32_t* __restrict__ yp, long long n, int K,
                                 int32_t* __restrict__ cm) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = yt[i], q = yp[i];
        if (t >= 0 && t < K && q >= 0 && q < K) atomicAdd(cm + t * K + q, 1);
    }
}

__global__ void gather_rows_kernel(const uint8_t* __restrict__ x, long long row_stride, long long row_bytes,
                                   long long first, long long step, long long n_rows, uint8_t* __restrict__ out) {
    const long long vec = row_bytes / 16;
    const long long total = n_rows * vec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / vec, v = i % vec;
        reinterpret_cast<uint4*>(out + r * row_bytes)[v] =
            reinterpret_cast<const uint4*>(x + (first + r * step) * row_stride)[v];
    }
}

__global__ void scatter_rows_add_kernel(const bf16* __restrict__ g, long long cols, long long first, long long step,
                                        long long n_rows, bf16* __restrict__ out) {
    const long long total = n_rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols, c = i % cols;
        bf16* o = out + (first + r * step) * cols + c;
        *o = __float2bfloat16(__bfloat162float(*o) + __bfloat162float(g[i]));
    }
}

__global__ void add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float x[8], y[8];
        unpack8(*reinterpret_cast<const bf16x8*>(a + i * 8), x);
        unpack8(*reinterpret_cast<const bf16x8*>(b + i * 8), y);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += y[j];
        *reinterpret_cast<bf16x8*>(out + i * 8) = pack8(x);
    }
}

__global__ void sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += x[i] * x[i];
    s = warp_sum(s);
    __shared__ float red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(out, t);
    }
}
__global__ void clip_scale_kernel(float* __restrict__ x, long long n, const float* __restrict__ sumsq, float max_norm) {
    const float norm = sqrtf(*sumsq);
    const float f = norm > max_norm ? max_norm / norm : 1.0f;  // tf.clip_by_global_norm
    if (f == 1.0f) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= f;
}

inline int ew_grid(long long n, int threads = 256) {
    long long g = (n + threads - 1) / threads;
    long long cap = (long long)polus_num_sms() * 16;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

}  // namespace

int polus_launch_colsum_reduce(const float* partial, int rows, int cols, float* out0, float* out1, int split, cudaStream_t st) {
    colsum_reduce_kernel<<<cdiv(cols, 32), 256, 0, st>>>(partial, rows, cols, out0, out1, split);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t polus_colsum_ws_floats(int N) { return (size_t)kColsumSlabs * (size_t)N; }

extern "C" int polus_act_bwd_colsum(const polus_bf16_t* dy, const polus_bf16_t* z, int M, int N, int act,
                                    polus_bf16_t* dz, float* gbias, float* ws, void* stream) {
    POLUS_REQUIRE(N > 0 && N % 8 == 0, "polus_act_bwd_colsum: N must be a multiple of 8 (got %d)", N);
    POLUS_REQUIRE(act == POLUS_ACT_NONE || z != nullptr, "polus_act_bwd_colsum: activation backward needs z");
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int slabs = cdiv(M, 64);
    if (slabs > kColsumSlabs) slabs = kColsumSlabs;
    const int rows_per_slab = cdiv(M, slabs);
    slabs = cdiv(M, rows_per_slab);
    dim3 grid(cdiv(N, 256), slabs);
    (void)ws;
    POLUS_CHECK_CUDA(polus_launch_pdl(act_bwd_colsum_kernel, grid, dim3(256), 0, st, (const bf16*)dy, (const bf16*)z, M, N, act, (bf16*)dz, gbias, rows_per_slab));
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_dropout(const polus_bf16_t* x, polus_bf16_t* y, int64_t n, float p_drop, uint64_t seed,
                             uint32_t site, const uint32_t* d_step, void* stream) {
    POLUS_REQUIRE(n % 8 == 0, "polus_dropout: n must be a multiple of 8");
    POLUS_REQUIRE(p_drop > 0.f && p_drop < 1.f && d_step != nullptr, "polus_dropout: need 0 < p < 1 and d_step");
    if (n == 0) return 0;
    dropout_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, n / 8, seed, site,
                                                                      (uint32_t)lrintf(p_drop * 65536.0f), 1.0f / (1.0f - p_drop), d_step);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_cast(const void* s, int sd, void* d, int dd, int64_t n, void* stream) {
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ew_grid(n);
#define CASTCASE(SD, DD, ST, DT) \
    if (sd == SD && dd == DD) { cast_kernel<ST, DT><<<grid, 256, 0, st>>>((const ST*)s, (DT*)d, n); g_launch_count++; POLUS_LAUNCH_CHECK(); return 0; }
    CASTCASE(POLUS_F32, POLUS_BF16, float, bf16)
    CASTCASE(POLUS_BF16, POLUS_F32, bf16, float)
    CASTCASE(POLUS_I32, POLUS_F32, int32_t, float)
    CASTCASE(POLUS_F32, POLUS_I32, float, int32_t)
    CASTCASE(POLUS_U8, POLUS_F32, uint8_t, float)
    CASTCASE(POLUS_U8, POLUS_BF16, uint8_t, bf16)
    CASTCASE(POLUS_I32, POLUS_BF16, int32_t, bf16)
    CASTCASE(POLUS_F32, POLUS_F32, float, float)
#undef CASTCASE
    polus_set_error("polus_cast: unsupported conversion %d -> %d", sd, dd);
    return POLUS_ERR_INVALID;
}

extern "C" int polus_fill_f32(float* d, float v, int64_t n, void* stream) {
    if (n == 0) return 0;
    fill_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d, v, n);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_binary_f32(int op, const float* a, const float* b, int64_t n, int64_t bn, float* out, void* stream) {
    POLUS_REQUIRE(op >= 0 && op <= 3, "polus_binary_f32: bad op %d", op);
    POLUS_REQUIRE(bn >= 1 && n % bn == 0, "polus_binary_f32: broadcast length %lld does not divide %lld", (long long)bn, (long long)n);
    if (n == 0) return 0;
    binary_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(op, a, b, n, bn, out);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_unary_f32(int op, const float* x, const float* dy, int grad, int64_t n, float* out, float alpha, void* stream) {
    POLUS_REQUIRE((op >= 0 && op <= 5) || (op >= 16 && op <= 22), "polus_unary_f32: bad op %d", op);
    POLUS_REQUIRE(!grad || dy != nullptr, "polus_unary_f32: grad needs dy");
    if (n == 0) return 0;
    unary_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(op, x, dy, grad, n, out, alpha);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_reduce_sum_f32(const float* x, int rows, int cols, int axis, float scale, float* out, int accumulate, void* stream) {
    POLUS_REQUIRE(axis == 0 || axis == 1, "polus_reduce_sum_f32: axis must be 0 or 1");
    if (rows == 0 || cols == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (axis == 1) reduce_rows_kernel<<<cdiv(rows, 8), 256, 0, st>>>(x, rows, cols, scale, out, accumulate);
    else reduce_cols_kernel<<<cdiv(cols, 128), 128, 0, st>>>(x, rows, cols, scale, out, accumulate);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_argmax_f32(const float* x, int rows, int cols, int32_t* out, void* stream) {
    POLUS_REQUIRE(cols >= 1, "polus_argmax_f32: cols must be >= 1");
    if (rows == 0) return 0;
    argmax_kernel<<<cdiv(rows, 128), 128, 0, (cudaStream_t)stream>>>(x, rows, cols, out);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_one_hot_f32(const int32_t* idx, int rows, int cols, float* out, void* stream) {
    if (rows == 0 || cols == 0) return 0;
    one_hot_kernel<<<ew_grid((long long)rows * cols), 256, 0, (cudaStream_t)stream>>>(idx, rows, cols, out);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_confusion_matrix(const int32_t* yt, const int32_t* yp, int64_t n, int K, int32_t* cm, void* stream) {
    if (n == 0) return 0;
    confusion_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(yt, yp, n, K, cm);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_gather_rows(const void* x, int64_t row_stride_bytes, int64_t row_bytes, int64_t first, int64_t step,
                                 int64_t n_rows, void* out, void* stream) {
    POLUS_REQUIRE(row_bytes % 16 == 0 && row_stride_bytes % 16 == 0, "polus_gather_rows: rows must be 16-byte multiples");
    if (n_rows == 0) return 0;
    gather_rows_kernel<<<ew_grid(n_rows * (row_bytes / 16)), 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)x, row_stride_bytes, row_bytes, first, step, n_rows, (uint8_t*)out);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_scatter_rows_add_bf16(const polus_bf16_t* g, int64_t cols, int64_t first, int64_t step,
                                           int64_t n_rows, polus_bf16_t* out, void* stream) {
    if (n_rows == 0) return 0;
    scatter_rows_add_kernel<<<ew_grid(n_rows * cols), 256, 0, (cudaStream_t)stream>>>((const bf16*)g, cols, first, step, n_rows, (bf16*)out);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_add_bf16(const polus_bf16_t* a, const polus_bf16_t* b, polus_bf16_t* out, int64_t n, void* stream) {
    POLUS_REQUIRE(n % 8 == 0, "polus_add_bf16: n must be a multiple of 8");
    if (n == 0) return 0;
    add_bf16_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n / 8);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_sumsq_f32(const float* x, int64_t n, float* out, void* stream) {
    if (n == 0) return 0;
    sumsq_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, n, out);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

extern "C" int polus_scale_by_clip(float* x, int64_t n, const float* sumsq, float max_norm, void* stream) {
    if (n == 0) return 0;
    clip_scale_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, n, sumsq, max_norm);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}
