// nccl_dyn.h -- NCCL, loaded at run time.  The only exchange step of the path is the global dB range of
// update_spec_greys (lib.rs:194-209): a handful of floats, max-reduced across the GPUs that hold the tracks.
// libsgx.so carries no link-time dependency on NCCL: the entry points are resolved from libnccl.so.2 with dlopen
// the first time a communicator is asked for (inside a process that already loaded NCCL -- PyTorch does -- that is
// the copy already in memory).  Types are restated from nccl.h 2.x (stable ABI: 128-byte unique id, opaque comm).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace sgx {

struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
enum { kNcclFloat32 = 7, kNcclMax = 2 }; // ncclDataType_t / ncclRedOp_t values of nccl.h

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId *);
    int (*CommInitRank)(NcclComm *, int nranks, NcclUniqueId id, int rank);
    int (*CommInitAll)(NcclComm *, int ndev, const int *devlist);
    int (*CommDestroy)(NcclComm);
    int (*AllReduce)(const void *send, void *recv, size_t count, int dtype, int op, NcclComm, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char *(*GetErrorString)(int);
    int (*GetVersion)(int *);
};

// Throws Error(SGX_ERR_NCCL) when libnccl.so.2 cannot be loaded or lacks a symbol.
const NcclApi &nccl();
void nccl_check(int result, const char *what);

} // namespace sgx
