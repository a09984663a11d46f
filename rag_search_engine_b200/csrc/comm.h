// comm.h — the handle's NCCL communicator for the row-sharded path (SURVEY §8b "the handle owns ... the NCCL
// comm", §8e).  NCCL is bound at RUN time (dlopen of libnccl.so.2, the soname both the system NCCL and the one
// torch bundles carry — whichever the process already loaded is reused), so librse.so has no link-time
// dependency on it and single-GPU users never touch it.  Only the handful of entry points the path needs.
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types only (ncclComm_t, ncclUniqueId, ncclDataType_t); no symbol is linked

#include <mutex>
#include <string>

namespace rse {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

// Process-wide, loaded on first use.  `path` (optional, or $RSE_NCCL_LIB) overrides the soname lookup.
inline NcclApi* nccl_api(const char* path = nullptr) {
  static NcclApi api;
  static std::mutex mu;                       // two handles may ask for the communicator API at the same time
  std::lock_guard<std::mutex> lock(mu);
  if (api.lib) return &api;
  const char* env = std::getenv("RSE_NCCL_LIB");
  const char* names[3] = {path, env, "libnccl.so.2"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
    const char* why = dlerror();               // (dlerror() clears the message: a second call returns NULL)
    api.error = why ? why : "dlopen failed";
  }
  if (!api.lib) {
    if (api.error.empty()) api.error = "libnccl.so.2 not found";
    return nullptr;
  }
  bool ok = true;
  auto sym = [&](const char* name) -> void* {
    void* p = dlsym(api.lib, name);
    if (!p) { ok = false; api.error = std::string("missing NCCL symbol ") + name; }
    return p;
  };
  api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  if (!ok) {
    dlclose(api.lib);
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}

// rank r owns queries [query_slice(nq, n, r), query_slice(nq, n, r + 1)) — the same split as sharded.query_slices
inline int query_slice(int nq, int n_ranks, int r) {
  const int per = nq / n_ranks, rem = nq % n_ranks;
  return r * per + (r < rem ? r : rem);
}

}  // namespace rse
