// comm.cuh -- NCCL inside libtoygpu.so: the all-gather of the moved slices and the
// reductions of the error statistics run on the library's own stream, so a C caller -- the
// reference's single-process driver through gpu_shim.c, or one process per GPU -- gets the
// multi-GPU path without any Python.
//
// NCCL is bound at run time (dlopen): the library has no link-time dependency on it, a
// one-GPU run never touches it, and inside a process that already carries an NCCL (PyTorch
// bundles one) that same copy is used instead of loading a second one.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    bool ok = false;
    const char *why = "";
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already in the process?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.why = "libnccl.so.2 not found"; return &api; }
#define TG_NCCL_SYM(field, name)                                                    \
    *(void **)(&api.field) = dlsym(h, name);                                        \
    if (!api.field) { api.why = "symbol " name " missing in libnccl"; return &api; }
    TG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    TG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    TG_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    TG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    TG_NCCL_SYM(AllGather, "ncclAllGather")
    TG_NCCL_SYM(AllReduce, "ncclAllReduce")
    TG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef TG_NCCL_SYM
    api.ok = true;
    return &api;
}
