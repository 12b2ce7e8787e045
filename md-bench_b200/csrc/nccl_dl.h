// nccl_dl.h -- NCCL bound at run time (dlopen), so libmdb200.so has no link-time dependency on it:
// single-GPU use never touches NCCL, and inside a process that already carries an NCCL (e.g. the one
// bundled with PyTorch, loaded by bench.py for torch.distributed) dlopen("libnccl.so.2") resolves to
// that same copy.  Only the stable subset of the API is declared (types as in nccl.h, NCCL >= 2.10).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include "mdb_util.cuh"

namespace mdb {

struct NcclApi {
    typedef void* comm_t;
    struct unique_id {
        char internal[128];
    };
    enum { Int8 = 0, Int32 = 2, Float64 = 8 }; // ncclDataType_t
    enum { Sum = 0 };                           // ncclRedOp_t

    void* handle = nullptr;
    int (*GetUniqueId)(unique_id*)                                                      = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int)                                   = nullptr;
    int (*CommDestroy)(comm_t)                                                          = nullptr;
    const char* (*GetErrorString)(int)                                                  = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t)        = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t)             = nullptr;
    int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t)                    = nullptr;
    int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t)                          = nullptr;
    int (*GroupStart)()                                                                 = nullptr;
    int (*GroupEnd)()                                                                   = nullptr;

    void load()
    {
        if (handle) return;
        const char* names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) throw Error(fmt("spatial decomposition over several processes needs NCCL: %s", dlerror()));
        auto sym = [&](const char* n) {
            void* p = dlsym(handle, n);
            if (!p) throw Error(fmt("NCCL symbol %s not found", n));
            return p;
        };
        GetUniqueId    = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank   = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommDestroy    = (decltype(CommDestroy))sym("ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        AllReduce      = (decltype(AllReduce))sym("ncclAllReduce");
        AllGather      = (decltype(AllGather))sym("ncclAllGather");
        Send           = (decltype(Send))sym("ncclSend");
        Recv           = (decltype(Recv))sym("ncclRecv");
        GroupStart     = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd       = (decltype(GroupEnd))sym("ncclGroupEnd");
    }
    void check(int r, const char* what) const
    {
        if (r != 0) throw Error(fmt("[NCCL Error]: %s: %s", what, GetErrorString ? GetErrorString(r) : "?"));
    }
};

inline NcclApi& nccl_api()
{
    static NcclApi api;
    return api;
}

#define MDB_NCCL(expr) ::mdb::nccl_api().check((expr), #expr)

} // namespace mdb
