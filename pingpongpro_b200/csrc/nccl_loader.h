// nccl_loader.h -- run-time binding of the four NCCL entry points the built-in collective needs (ppp_nccl_attach).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "ppp_b200.h"

namespace pppk {

struct NcclId { char internal[PPP_NCCL_UNIQUE_ID_BYTES]; };
struct NcclApi {
    int (*getUniqueId)(NcclId*) = nullptr;
    int (*commInitRank)(void** comm, int nranks, NcclId id, int rank) = nullptr;
    int (*allReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t stream) = nullptr;
    int (*commDestroy)(void* comm) = nullptr;
    const char* (*getErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};
// Loads libnccl.so.2 on first use (PPP_NCCL_LIB overrides the path; a copy already in the process is preferred).
NcclApi& nccl();

}  // namespace pppk
