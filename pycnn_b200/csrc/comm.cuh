// comm.cuh -- NCCL, bound at run time with dlopen so that (a) the library loads on machines without
// NCCL/GPU (the CPU test tier checks the exported symbols) and (b) inside a torchrun process we share
// the libnccl.so.2 that torch already mapped instead of pulling a second copy.
// One process per GPU; the 128-byte unique id is created on rank 0 and broadcast by the host layer
// (pycnn_b200/parallel.py) over torch.distributed.
#pragma once
#include <dlfcn.h>

#include "common.cuh"

namespace pb {

struct NcclUniqueId {
  char internal[128];
};
typedef void* NcclComm;
enum { kNcclInt64 = 4, kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*CommGetAsyncError)(NcclComm, int*) = nullptr;
  int (*CommAbort)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

static inline NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already mapped by torch?
    if (h) break;
  }
  if (!h) {
    const char* env = getenv("PYCNN_B200_NCCL_LIB");
    if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  }
  for (const char* n : names) {
    if (h) break;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!h) return nullptr;
  api.GetUniqueId = reinterpret_cast<int (*)(NcclUniqueId*)>(dlsym(h, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclUniqueId, int)>(dlsym(h, "ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t)>(
      dlsym(h, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(dlsym(h, "ncclCommDestroy"));
  api.CommAbort = reinterpret_cast<int (*)(NcclComm)>(dlsym(h, "ncclCommAbort"));
  api.CommGetAsyncError = reinterpret_cast<int (*)(NcclComm, int*)>(dlsym(h, "ncclCommGetAsyncError"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
  api.GetVersion = reinterpret_cast<int (*)(int*)>(dlsym(h, "ncclGetVersion"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) return nullptr;
  api.handle = h;
  return &api;
}

#define PB_NCCL(api, expr)                                                                         \
  do {                                                                                             \
    int _r = (expr);                                                                               \
    if (_r != 0)                                                                                   \
      return pb::fail("%s failed: %s", #expr, (api)->GetErrorString ? (api)->GetErrorString(_r) : "?"); \
  } while (0)

}  // namespace pb
