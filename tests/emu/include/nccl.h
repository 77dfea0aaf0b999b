// nccl.h -- TEST INFRASTRUCTURE (tests/emu): the few NCCL types hb_api.cu names.  The product resolves NCCL with dlopen at
// hb_comm_init; the CPU model only ever runs single-rank communicators, which never touch it.
#pragma once
#include <stddef.h>
#include "cuda_runtime.h"
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclUint64 = 5 } ncclDataType_t;
