// NCCL for the devices of ONE process (bixcor_init(n), the mode an R session reaches): libnccl is opened at run time,
// so libbixcor.so has no link-time dependency on it and a single-GPU user never loads it.  Only the long-stable C API
// is used (ncclCommInitAll / AllReduce / Broadcast / Group*), declared here against <nccl.h> of the image for the types.
//
// The collectives replace what round 1 did through the host or with peer copies:
//   TOM         connectivity vector k  -> ncclAllReduce(sum);  affinity slabs -> grouped ncclBroadcast (one fused launch)
//   sparse Gram partial X'X + column sums of the cell shards -> ncclAllReduce(sum)
#pragma once

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>

namespace bixcor {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  int version = 0;
  char error[256] = "";

  bool load() {
    if (handle) return true;
    const char* names[] = {getenv("BIXCOR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (handle) break;
    }
    if (!handle) {
      snprintf(error, sizeof error, "libnccl.so.2 not found (%s)", dlerror());
      return false;
    }
    bool ok = true;
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(handle, name);
      if (!p) {
        snprintf(error, sizeof error, "symbol %s missing in libnccl", name);
        ok = false;
      }
      return p;
    };
    GetVersion = (decltype(GetVersion))sym("ncclGetVersion");
    CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
    Broadcast = (decltype(Broadcast))sym("ncclBroadcast");
    AllGather = (decltype(AllGather))sym("ncclAllGather");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!ok) {
      dlclose(handle);
      handle = nullptr;
      return false;
    }
    GetVersion(&version);
    return true;
  }
};

}  // namespace bixcor
