// nccl_loader.cc -- run-time binding of NCCL (see nccl_loader.h).  Types and enum values follow <nccl.h> (2.x ABI):
// ncclUniqueId = 128 opaque bytes passed BY VALUE to ncclCommInitRank, ncclDataType_t ncclChar = 0 / ncclFloat32 = 7,
// ncclRedOp_t ncclMin = 3, ncclResult_t 0 = success.
#include "nccl_loader.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>

#include "rbq_internal.h"

namespace rbq {
namespace {
struct UniqueId {
    char internal[128];
};
typedef int (*InitRankFn)(void** comm, int nranks, UniqueId id, int rank);
NcclApi g_api;
InitRankFn g_init_rank = nullptr;
int g_state = 0;  // 0 not tried, 1 loaded, -1 failed
std::string g_why;
std::mutex g_mu;
}  // namespace

int NcclApi::comm_init_rank(void** comm, int nranks, const uint8_t* id128, int rank) const {
    UniqueId id;
    std::memcpy(id.internal, id128, 128);
    return g_init_rank(comm, nranks, id, rank);
}

int nccl_load() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_state == 1) return RBQ_OK;
    if (g_state == -1) return fail(RBQ_CUDA_ERROR, g_why);
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);  // the copy already loaded by the process (e.g. torch's) wins
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        g_state = -1;
        g_why = std::string("NCCL is not available: ") + dlerror();
        return fail(RBQ_CUDA_ERROR, g_why);
    }
    auto sym = [&](const char* name) -> void* {
        void* p = dlsym(lib, name);
        if (!p && g_why.empty()) g_why = std::string("NCCL symbol missing: ") + name;
        return p;
    };
    g_api.get_unique_id = reinterpret_cast<int (*)(void*)>(sym("ncclGetUniqueId"));
    g_init_rank = reinterpret_cast<InitRankFn>(sym("ncclCommInitRank"));
    g_api.comm_destroy = reinterpret_cast<int (*)(void*)>(sym("ncclCommDestroy"));
    g_api.all_gather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(sym("ncclAllGather"));
    g_api.all_reduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(sym("ncclAllReduce"));
    g_api.send = reinterpret_cast<int (*)(const void*, size_t, int, int, void*, cudaStream_t)>(sym("ncclSend"));
    g_api.recv = reinterpret_cast<int (*)(void*, size_t, int, int, void*, cudaStream_t)>(sym("ncclRecv"));
    g_api.group_start = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
    g_api.group_end = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
    g_api.get_error_string = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
    if (!g_why.empty()) {
        g_state = -1;
        return fail(RBQ_CUDA_ERROR, g_why);
    }
    g_state = 1;
    return RBQ_OK;
}

const NcclApi& nccl_api() { return g_api; }
}  // namespace rbq
