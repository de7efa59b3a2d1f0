// nccl_dyn.h -- NCCL bound at run time with dlopen, so that libax1gpu.so links against nothing but
// the CUDA runtime and uses the very libnccl.so.2 that torch.distributed has already loaded.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstddef>

extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclInt = 2, ncclUint32 = 3, ncclInt64 = 4,
               ncclUint64 = 5, ncclFloat16 = 6, ncclHalf = 6, ncclFloat32 = 7, ncclFloat = 7, ncclFloat64 = 8,
               ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
}

namespace ax1 {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;

  bool load(const char* path) {
    if (lib) return true;
    lib = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
#define AX1_SYM(name) *(void**)(&name) = dlsym(lib, "nccl" #name); if (!name) return false;
    AX1_SYM(GetUniqueId) AX1_SYM(CommInitRank) AX1_SYM(CommDestroy) AX1_SYM(AllReduce) AX1_SYM(AllGather) AX1_SYM(Send) AX1_SYM(Recv)
    AX1_SYM(GroupStart) AX1_SYM(GroupEnd) AX1_SYM(GetErrorString)
#undef AX1_SYM
    return true;
  }
};

}  // namespace ax1
