// NCCL is resolved at run time, and only when a multi-GPU entry point is used: linking libnccl at build
// time would load the system libnccl.so.2 into any process that loads this library, which then shadows
// the (newer) NCCL bundled with PyTorch for the rest of the process.  Here the already-loaded NCCL is
// preferred (RTLD_NOLOAD: the one torch.distributed initialised), falling back to the default search path.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace acvd {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    std::string error;

    bool load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) if (!handle) handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        for (const char* n : names) if (!handle) handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { error = "libnccl.so.2 not found"; return false; }
        auto sym = [&](const char* s) { void* p = dlsym(handle, s); if (!p) error = std::string("missing NCCL symbol ") + s; return p; };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        Broadcast = (decltype(Broadcast))sym("ncclBroadcast");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        return error.empty();
    }
};

inline NcclApi& nccl() {
    static NcclApi api;
    return api;
}

}  // namespace acvd
