// nccl_dyn.cpp -- see nccl_dyn.h.
#include "nccl_dyn.h"

#include <dlfcn.h>

#include <mutex>
#include <string>

#include "engine.h"

namespace sgx {

namespace {
NcclApi g_api{};
bool g_loaded = false;
std::string g_load_error;
std::once_flag g_once;

void load()
{
    void *h = nullptr;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { g_load_error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
    bool ok = true;
    auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) { ok = false; g_load_error = std::string("libnccl lacks ") + n; } return p; };
    g_api.GetUniqueId = reinterpret_cast<decltype(g_api.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_api.CommInitRank = reinterpret_cast<decltype(g_api.CommInitRank)>(sym("ncclCommInitRank"));
    g_api.CommInitAll = reinterpret_cast<decltype(g_api.CommInitAll)>(sym("ncclCommInitAll"));
    g_api.CommDestroy = reinterpret_cast<decltype(g_api.CommDestroy)>(sym("ncclCommDestroy"));
    g_api.AllReduce = reinterpret_cast<decltype(g_api.AllReduce)>(sym("ncclAllReduce"));
    g_api.GroupStart = reinterpret_cast<decltype(g_api.GroupStart)>(sym("ncclGroupStart"));
    g_api.GroupEnd = reinterpret_cast<decltype(g_api.GroupEnd)>(sym("ncclGroupEnd"));
    g_api.GetErrorString = reinterpret_cast<decltype(g_api.GetErrorString)>(sym("ncclGetErrorString"));
    g_api.GetVersion = reinterpret_cast<decltype(g_api.GetVersion)>(sym("ncclGetVersion"));
    g_loaded = ok;
}
} // namespace

const NcclApi &nccl()
{
    std::call_once(g_once, load);
    if (!g_loaded) throw Error(SGX_ERR_NCCL, g_load_error);
    return g_api;
}

void nccl_check(int result, const char *what)
{
    if (result == 0) return;
    const char *msg = g_loaded && g_api.GetErrorString ? g_api.GetErrorString(result) : "?";
    throw Error(SGX_ERR_NCCL, std::string("NCCL: ") + msg + " in " + what);
}

} // namespace sgx
