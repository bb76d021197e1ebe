// nccl_loader.cpp -- see nccl_loader.h.  Only four entry points (and ncclGetErrorString) are needed; their signatures
// and the enum values used in api.cu have been stable since NCCL 2.0.  libnccl is not a link-time dependency: a
// single-GPU run never touches it.
#include "nccl_loader.h"

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>

namespace pppk {

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        const char* env = getenv("PPP_NCCL_LIB");
        if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already uses (e.g. torch's)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { api.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        api.getUniqueId = reinterpret_cast<decltype(api.getUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        api.commInitRank = reinterpret_cast<decltype(api.commInitRank)>(dlsym(h, "ncclCommInitRank"));
        api.allReduce = reinterpret_cast<decltype(api.allReduce)>(dlsym(h, "ncclAllReduce"));
        api.commDestroy = reinterpret_cast<decltype(api.commDestroy)>(dlsym(h, "ncclCommDestroy"));
        api.getErrorString = reinterpret_cast<decltype(api.getErrorString)>(dlsym(h, "ncclGetErrorString"));
        api.ok = api.getUniqueId && api.commInitRank && api.allReduce && api.commDestroy && api.getErrorString;
        if (!api.ok) api.why = "libnccl.so.2 lacks an expected entry point";
    });
    return api;
}

}  // namespace pppk
