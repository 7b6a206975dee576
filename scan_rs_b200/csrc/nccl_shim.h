// nccl_shim.h -- NCCL is resolved with dlopen at the first communicator call instead of being a
// link-time dependency: a single-GPU process never loads it, and a process that already holds a
// libnccl.so.2 (e.g. the one bundled with PyTorch, which the bench uses for rendezvous plumbing)
// shares that copy instead of clashing with the system one.
#pragma once
#include <nccl.h>

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
};

// returns nullptr (and sets the error string) if libnccl.so.2 cannot be loaded
const NcclApi *sb_nccl();

#define ncclGetUniqueId sb_nccl()->GetUniqueId
#define ncclCommInitRank sb_nccl()->CommInitRank
#define ncclCommDestroy sb_nccl()->CommDestroy
#define ncclAllReduce sb_nccl()->AllReduce
#define ncclAllGather sb_nccl()->AllGather
#define ncclBroadcast sb_nccl()->Broadcast
#define ncclGroupStart sb_nccl()->GroupStart
#define ncclGroupEnd sb_nccl()->GroupEnd
#define ncclGetErrorString sb_nccl()->GetErrorString
