// NCCL through the instance that is already loaded in the process (torch's), looked up with dlopen(RTLD_NOLOAD):
// the library is not linked, the communicator belongs to the caller (torch.distributed's ProcessGroupNCCL hands out
// its ncclComm_t).  Shared by the partial-Gram exchange (comm.cu) and the distributed eigensolver rounds (eig.cu).
#pragma once
#include <dlfcn.h>

#include "common.cuh"

namespace vvt {

typedef int (*NcclAllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*NcclAllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*NcclSendFn)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*NcclRecvFn)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*NcclPreMulFn)(int*, void*, int, int, void*);
typedef int (*NcclOpDestroyFn)(int, void*);
typedef int (*NcclGroupFn)(void);
typedef int (*NcclCommIntFn)(void*, int*);
typedef const char* (*NcclErrStrFn)(int);

struct Nccl {
  NcclAllReduceFn all_reduce = nullptr;
  NcclAllGatherFn all_gather = nullptr;
  NcclSendFn send = nullptr;
  NcclRecvFn recv = nullptr;
  NcclPreMulFn premul = nullptr;
  NcclOpDestroyFn op_destroy = nullptr;
  NcclGroupFn group_start = nullptr, group_end = nullptr;
  NcclCommIntFn comm_count = nullptr, comm_rank = nullptr;
  NcclErrStrFn err = nullptr;
  bool ok = false;
};

inline const Nccl& nccl() {
  static Nccl n = [] {
    Nccl r;
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_LAZY | RTLD_NOLOAD);
      if (h) break;
    }
    if (!h) return r;
    r.all_reduce = reinterpret_cast<NcclAllReduceFn>(dlsym(h, "ncclAllReduce"));
    r.all_gather = reinterpret_cast<NcclAllGatherFn>(dlsym(h, "ncclAllGather"));
    r.send = reinterpret_cast<NcclSendFn>(dlsym(h, "ncclSend"));
    r.recv = reinterpret_cast<NcclRecvFn>(dlsym(h, "ncclRecv"));
    r.premul = reinterpret_cast<NcclPreMulFn>(dlsym(h, "ncclRedOpCreatePreMulSum"));
    r.op_destroy = reinterpret_cast<NcclOpDestroyFn>(dlsym(h, "ncclRedOpDestroy"));
    r.group_start = reinterpret_cast<NcclGroupFn>(dlsym(h, "ncclGroupStart"));
    r.group_end = reinterpret_cast<NcclGroupFn>(dlsym(h, "ncclGroupEnd"));
    r.comm_count = reinterpret_cast<NcclCommIntFn>(dlsym(h, "ncclCommCount"));
    r.comm_rank = reinterpret_cast<NcclCommIntFn>(dlsym(h, "ncclCommUserRank"));
    r.err = reinterpret_cast<NcclErrStrFn>(dlsym(h, "ncclGetErrorString"));
    r.ok = r.all_reduce && r.all_gather && r.send && r.recv && r.premul && r.op_destroy && r.group_start && r.group_end && r.comm_count &&
           r.comm_rank;
    return r;
  }();
  return n;
}

constexpr int kNcclSum = 0, kNcclUint64 = 5, kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclScalarHostImmediate = 1;

inline int check_nccl(int status, const char* where) {
  if (status == 0) return VVT_OK;
  return fail(VVT_ERR_CUDA, "%s: NCCL: %s", where, nccl().err ? nccl().err(status) : "error");
}

}  // namespace vvt
