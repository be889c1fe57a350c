// comm_nccl.cpp - an ncclComm_t wrapped into the scb_comm the sharded flush asks for (include/scalce_b200.h): the two
// collectives of scb_shard_flush over NCCL / NVLink, no Python anywhere. Built into libscalce_b200_nccl.so (links libnccl);
// libscalce_b200.so itself has no NCCL dependency - MPI or any other transport can fill the same struct.
//   scb_nccl_unique_id   rank 0 makes the id and hands it to the other ranks (file, pipe, MPI_Bcast, torch.distributed ...)
//   scb_nccl_comm_create one communicator per process / GPU (ncclCommInitRank), returned as an scb_comm
//   scb_nccl_comm_destroy
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/scalce_b200.h"

namespace {
struct NcclCtx {
    ncclComm_t comm = nullptr;
    cudaStream_t st = nullptr;      // for the host-buffer collectives and the barrier
    int n = 0, device = 0;
    void *stage = nullptr; size_t stage_cap = 0;   // device staging of host all-gathers
    scb_comm iface;
};

int ensure_stage(NcclCtx *c, size_t bytes) {
    if (c->stage_cap >= bytes) return 0;
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr; c->stage_cap = 0;
    if (cudaMalloc(&c->stage, bytes) != cudaSuccess) return 1;
    c->stage_cap = bytes;
    return 0;
}

int nccl_allgather(void *ctx, const void *send, void *recv, int64_t bytes, int32_t device, void *stream) {
    NcclCtx *c = (NcclCtx *)ctx;
    if (bytes <= 0) return 0;
    if (cudaSetDevice(c->device) != cudaSuccess) return 1;
    if (device) return ncclAllGather(send, recv, (size_t)bytes, ncclUint8, c->comm, (cudaStream_t)stream) == ncclSuccess ? 0 : 1;
    // host buffers: stage through device memory
    const size_t total = (size_t)bytes * (size_t)(c->n + 1);
    if (ensure_stage(c, total)) return 1;
    char *d_send = (char *)c->stage, *d_recv = d_send + bytes;
    if (cudaMemcpyAsync(d_send, send, (size_t)bytes, cudaMemcpyHostToDevice, c->st) != cudaSuccess) return 1;
    if (ncclAllGather(d_send, d_recv, (size_t)bytes, ncclUint8, c->comm, c->st) != ncclSuccess) return 1;
    if (cudaMemcpyAsync(recv, d_recv, (size_t)bytes * (size_t)c->n, cudaMemcpyDeviceToHost, c->st) != cudaSuccess) return 1;
    return cudaStreamSynchronize(c->st) == cudaSuccess ? 0 : 1;
}

int nccl_barrier(void *ctx) {
    NcclCtx *c = (NcclCtx *)ctx;
    if (cudaSetDevice(c->device) != cudaSuccess) return 1;
    if (ensure_stage(c, 256)) return 1;
    if (ncclAllReduce(c->stage, c->stage, 1, ncclInt32, ncclSum, c->comm, c->st) != ncclSuccess) return 1;
    return cudaStreamSynchronize(c->st) == cudaSuccess ? 0 : 1;
}
}  // namespace

extern "C" {

int scb_nccl_unique_id(uint8_t *id128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return SCB_ECUDA;
    memcpy(id128, &id, 128);
    return SCB_OK;
}

int scb_nccl_comm_create(const uint8_t *id128, int32_t rank, int32_t n_ranks, int32_t device, scb_comm **out) {
    if (!id128 || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return SCB_EINVAL;
    NcclCtx *c = new (std::nothrow) NcclCtx();
    if (!c) return SCB_ENOMEM;
    c->n = n_ranks; c->device = device;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess ||
        ncclCommInitRank(&c->comm, n_ranks, id, rank) != ncclSuccess) { delete c; return SCB_ECUDA; }
    memset(&c->iface, 0, sizeof c->iface);
    c->iface.rank = rank; c->iface.n_ranks = n_ranks; c->iface.same_process = 0; c->iface.ctx = c;
    c->iface.allgather = nccl_allgather; c->iface.barrier = nccl_barrier;
    *out = &c->iface;
    return SCB_OK;
}

void scb_nccl_comm_destroy(scb_comm *cm) {
    if (!cm) return;
    NcclCtx *c = (NcclCtx *)cm->ctx;
    cudaSetDevice(c->device);
    if (c->comm) ncclCommDestroy(c->comm);
    if (c->stage) cudaFree(c->stage);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

}  // extern "C"
