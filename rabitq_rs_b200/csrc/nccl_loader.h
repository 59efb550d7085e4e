// nccl_loader.h -- the few NCCL entry points the one-call sharded search needs, resolved at run time (dlopen libnccl.so.2),
// so that librbq.so has no link-time dependency on NCCL and single-GPU users never load it.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace rbq {
struct NcclApi {
    int (*get_unique_id)(void* id128) = nullptr;                                                    // ncclGetUniqueId
    int (*comm_destroy)(void* comm) = nullptr;                                                      // ncclCommDestroy
    int (*all_gather)(const void* send, void* recv, size_t sendcount, int dtype, void* comm, cudaStream_t st) = nullptr;
    int (*all_reduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st) = nullptr;
    int (*send)(const void* buf, size_t count, int dtype, int peer, void* comm, cudaStream_t st) = nullptr;
    int (*recv)(void* buf, size_t count, int dtype, int peer, void* comm, cudaStream_t st) = nullptr;
    int (*group_start)() = nullptr;
    int (*group_end)() = nullptr;
    const char* (*get_error_string)(int) = nullptr;
    int comm_init_rank(void** comm, int nranks, const uint8_t* id128, int rank) const;
};
int nccl_load();             // RBQ_OK or an error (library missing / symbol missing)
const NcclApi& nccl_api();   // valid after a successful nccl_load()
}  // namespace rbq
