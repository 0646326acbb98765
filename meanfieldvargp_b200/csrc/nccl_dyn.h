// NCCL bound at run time (dlopen): the library stays loadable on a box with no
// NCCL, and inside a PyTorch process it reuses the libnccl.so.2 that torch has
// already loaded instead of pulling in a second copy.  Only the handful of
// entry points the multi-GPU path needs are resolved.
#pragma once

#include <cstddef>
#include <cuda_runtime.h>
#include <dlfcn.h>

namespace mfvgp {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0, kNcclFloat64 = 8 };   // ncclResult_t / ncclDataType_t values

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;

    bool load() {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
        if (!lib) return false;
#define MFVGP_NCCL_SYM(field, name)                                     \
        field = reinterpret_cast<decltype(field)>(dlsym(lib, name));    \
        if (!field) return false;
        MFVGP_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        MFVGP_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        MFVGP_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        MFVGP_NCCL_SYM(AllGather, "ncclAllGather")
        MFVGP_NCCL_SYM(Send, "ncclSend")
        MFVGP_NCCL_SYM(Recv, "ncclRecv")
        MFVGP_NCCL_SYM(GroupStart, "ncclGroupStart")
        MFVGP_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        MFVGP_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef MFVGP_NCCL_SYM
        return true;
    }
};

inline NcclApi &nccl() {
    static NcclApi api;
    return api;
}

}  // namespace mfvgp
