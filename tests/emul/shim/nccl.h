// TEST INFRASTRUCTURE — types of nccl.h that libcpppd names; the emulated build never calls NCCL (world 1).
#pragma once
#include <cstddef>
typedef struct { char internal[128]; } ncclUniqueId;
typedef struct ncclComm *ncclComm_t;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2 } ncclRedOp_t;
