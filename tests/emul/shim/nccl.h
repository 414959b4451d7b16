// TEST INFRASTRUCTURE — the part of nccl.h libcpppd names.  The emulated build resolves these entry points
// with dlopen/dlsym like the product does; tests point CPPPD_NCCL_LIB at libcpppd_emul.so itself, whose
// tests/emul/emul_nccl.cpp implements them for ranks that are threads of one process.
#pragma once
#include <cstddef>
typedef struct { char internal[128]; } ncclUniqueId;
typedef struct ncclComm *ncclComm_t;
typedef enum { ncclSuccess = 0, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint64 = 5, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2 } ncclRedOp_t;
