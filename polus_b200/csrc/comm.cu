// Data-parallel collectives over NCCL / NVLink 5.  One communicator per process (one process per GPU).
//
// Replaces horovod.tensorflow as polus uses it -- the six functions of polus/mock/horovod.py:5-24:
//   hvd.init/size/local_rank            -> polus_comm_init / polus_comm_size / polus_comm_rank
//   hvd.DistributedGradientTape (Average allreduce of every gradient, polus/training.py:182-185)
//                                       -> polus_comm_allreduce_f32 on buckets of the flat gradient arena
//                                          (sum; the 1/N is folded into polus_adam's grad_scale)
//   hvd.broadcast_variables (training.py:208-211) -> polus_comm_broadcast over the flat arenas
//   hvd.allgather_object (callbacks.py:249)        -> polus_comm_allgather (fixed-size payloads)
#include "common.cuh"
#include <nccl.h>
#include <atomic>

static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_size = 1;

#define POLUS_CHECK_NCCL(expr)                                                                          \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != ncclSuccess) {                                                                        \
            polus_set_error("%s:%d NCCL error %d (%s) in `%s`", __FILE__, __LINE__, (int)_r,            \
                            ncclGetErrorString(_r), #expr);                                             \
            return POLUS_ERR_NCCL;                                                                      \
        }                                                                                               \
    } while (0)

namespace {
// fp32 <-> bf16 wire format of the gradient exchange, 8 elements per thread per trip (32 B in, 16 B out)
__global__ void __launch_bounds__(256) pack_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
    const long long n8 = n >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
        __nv_bfloat162 o[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w),
                               __floats2bfloat162_rn(b.x, b.y), __floats2bfloat162_rn(b.z, b.w)};
        reinterpret_cast<uint4*>(dst)[i] = *reinterpret_cast<uint4*>(o);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) dst[(n8 << 3) + threadIdx.x] = __float2bfloat16(src[(n8 << 3) + threadIdx.x]);
}
__global__ void __launch_bounds__(256) unpack_bf16_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
    const long long n8 = n >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        uint4 raw = reinterpret_cast<const uint4*>(src)[i];
        const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&raw);
        const float2 f0 = __bfloat1622float2(v[0]), f1 = __bfloat1622float2(v[1]), f2 = __bfloat1622float2(v[2]), f3 = __bfloat1622float2(v[3]);
        reinterpret_cast<float4*>(dst)[2 * i] = make_float4(f0.x, f0.y, f1.x, f1.y);
        reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(f2.x, f2.y, f3.x, f3.y);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) dst[(n8 << 3) + threadIdx.x] = __bfloat162float(src[(n8 << 3) + threadIdx.x]);
}
inline int wire_grid(long long n) {
    long long g = ((n >> 3) + 255) / 256;
    const long long cap = (long long)polus_num_sms() * 4;   // small grid: these run next to tensor-bound GEMMs
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}
}  // namespace

extern std::atomic<long long> g_launch_count;

extern "C" {

int polus_comm_unique_id(void* h_id128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    POLUS_CHECK_NCCL(ncclGetUniqueId(&id));
    memcpy(h_id128, &id, sizeof(id));
    return 0;
}

int polus_comm_init_cfg(int rank, int size, const void* h_id128, int max_ctas) {
    POLUS_REQUIRE(size >= 1 && rank >= 0 && rank < size, "polus_comm_init: bad rank %d / size %d", rank, size);
    POLUS_REQUIRE(g_comm == nullptr, "polus_comm_init: communicator already initialised");
    ncclUniqueId id;
    memcpy(&id, h_id128, sizeof(id));
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    // The collectives run UNDER backward: every SM an NCCL CTA holds is one a persistent 148-CTA GEMM has to wait
    // for.  max_ctas > 0 caps the communicator's CTAs (NVLink 5 needs few to saturate); 0 leaves NCCL's default.
    if (max_ctas > 0) { cfg.maxCTAs = max_ctas; cfg.minCTAs = 1; }
    POLUS_CHECK_NCCL(ncclCommInitRankConfig(&g_comm, size, id, rank, &cfg));
    g_rank = rank;
    g_size = size;
    return 0;
}

int polus_comm_init(int rank, int size, const void* h_id128) { return polus_comm_init_cfg(rank, size, h_id128, 0); }

int polus_comm_size(void) { return g_size; }
int polus_comm_rank(void) { return g_rank; }

int polus_comm_allreduce_f32(float* d_buf, int64_t n, void* stream) {
    if (g_size == 1 || n == 0) return 0;
    POLUS_REQUIRE(g_comm != nullptr, "polus_comm_allreduce_f32: communicator not initialised");
    POLUS_CHECK_NCCL(ncclAllReduce(d_buf, d_buf, (size_t)n, ncclFloat, ncclSum, g_comm, (cudaStream_t)stream));
    return 0;
}

// Gradient exchange with a bf16 wire format: pack the fp32 bucket to bf16 (one pass, 6 B/element), sum in bf16 over
// NVLink (half the bytes of the fp32 exchange), unpack into the fp32 arena the optimizer reads.  Every rank ends with
// the same values, so replicas stay bit-identical; the sum carries one bf16 rounding per rank pair (relative 2^-9).
int polus_comm_allreduce_bf16(float* d_buf, polus_bf16_t* d_scratch, int64_t n, void* stream) {
    if (g_size == 1 || n == 0) return 0;
    POLUS_REQUIRE(g_comm != nullptr, "polus_comm_allreduce_bf16: communicator not initialised");
    POLUS_REQUIRE(d_scratch != nullptr, "polus_comm_allreduce_bf16: scratch required");
    POLUS_REQUIRE(((uintptr_t)d_buf | (uintptr_t)d_scratch) % 16 == 0, "polus_comm_allreduce_bf16: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    pack_bf16_kernel<<<wire_grid(n), 256, 0, st>>>(d_buf, (bf16*)d_scratch, n);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    POLUS_CHECK_NCCL(ncclAllReduce(d_scratch, d_scratch, (size_t)n, ncclBfloat16, ncclSum, g_comm, st));
    unpack_bf16_kernel<<<wire_grid(n), 256, 0, st>>>((const bf16*)d_scratch, d_buf, n);
    g_launch_count++;
    POLUS_LAUNCH_CHECK();
    return 0;
}

int polus_comm_broadcast(void* d_buf, size_t bytes, int root, void* stream) {
    if (g_size == 1 || bytes == 0) return 0;
    POLUS_REQUIRE(g_comm != nullptr, "polus_comm_broadcast: communicator not initialised");
    POLUS_CHECK_NCCL(ncclBroadcast(d_buf, d_buf, bytes, ncclChar, root, g_comm, (cudaStream_t)stream));
    return 0;
}

int polus_comm_allgather(const void* d_send, void* d_recv, size_t bytes_per_rank, void* stream) {
    if (bytes_per_rank == 0) return 0;
    if (g_size == 1) {
        if (d_send != d_recv) POLUS_CHECK_CUDA(cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return 0;
    }
    POLUS_REQUIRE(g_comm != nullptr, "polus_comm_allgather: communicator not initialised");
    POLUS_CHECK_NCCL(ncclAllGather(d_send, d_recv, bytes_per_rank, ncclChar, g_comm, (cudaStream_t)stream));
    return 0;
}

int polus_comm_destroy(void) {
    if (g_comm != nullptr) {
        // ncclCommAbort, not ncclCommDestroy: Destroy waits for every reference to the communicator, and a captured
        // CUDA graph that holds NCCL kernels keeps one for as long as the graph exists (measured: a 2-rank run that
        // called Destroy after training through captured steps blocked until the watchdog fired).  All device work
        // has been synchronised by the caller's last read; Abort releases the resources without that wait.
        cudaDeviceSynchronize();
        ncclCommAbort(g_comm);
        g_comm = nullptr;
    }
    g_rank = 0;
    g_size = 1;
    return 0;
}

}  // extern "C"
