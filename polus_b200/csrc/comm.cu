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

extern "C" {

int polus_comm_unique_id(void* h_id128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    POLUS_CHECK_NCCL(ncclGetUniqueId(&id));
    memcpy(h_id128, &id, sizeof(id));
    return 0;
}

int polus_comm_init(int rank, int size, const void* h_id128) {
    POLUS_REQUIRE(size >= 1 && rank >= 0 && rank < size, "polus_comm_init: bad rank %d / size %d", rank, size);
    POLUS_REQUIRE(g_comm == nullptr, "polus_comm_init: communicator already initialised");
    ncclUniqueId id;
    memcpy(&id, h_id128, sizeof(id));
    POLUS_CHECK_NCCL(ncclCommInitRank(&g_comm, size, id, rank));
    g_rank = rank;
    g_size = size;
    return 0;
}

int polus_comm_size(void) { return g_size; }
int polus_comm_rank(void) { return g_rank; }

int polus_comm_allreduce_f32(float* d_buf, int64_t n, void* stream) {
    if (g_size == 1 || n == 0) return 0;
    POLUS_REQUIRE(g_comm != nullptr, "polus_comm_allreduce_f32: communicator not initialised");
    POLUS_CHECK_NCCL(ncclAllReduce(d_buf, d_buf, (size_t)n, ncclFloat, ncclSum, g_comm, (cudaStream_t)stream));
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
        ncclCommDestroy(g_comm);
        g_comm = nullptr;
    }
    g_rank = 0;
    g_size = 1;
    return 0;
}

}  // extern "C"
