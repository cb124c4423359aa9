// Minimal run-time binding to NCCL (dlopen).  The library is resolved by soname so that, inside a process that
// already imported torch, the torch-bundled libnccl.so.2 is the one used; a Julia host gets the system one.
// Only the handful of entry points the row-sharded path needs: unique id, communicator, fp64 sum all-reduce.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace scs {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int /*dtype*/, int /*op*/, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  static constexpr int kFloat64 = 8;  // ncclFloat64
  static constexpr int kSum = 0;      // ncclSum

  bool load(const char** why) {
    if (handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) {
      *why = "libnccl.so.2 not found (dlopen)";
      return false;
    }
    GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
    GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString) {
      *why = "libnccl is missing a required symbol";
      return false;
    }
    return true;
  }
};

}  // namespace scs
