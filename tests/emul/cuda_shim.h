// TEST INFRASTRUCTURE — just enough of the CUDA device language to compile the hot kernels of
// pysparselp_b200/csrc (cpppd_device_types.cuh + cpppd_hot_kernels.cuh) with g++ and run them on the
// CPU, one CUDA thread after the other.  Nothing here is part of the product.
//
// What is modelled: threadIdx / blockIdx / blockDim / gridDim, __shared__ (one block runs at a time, so
// a function-local static is the block's shared memory), the load intrinsics (plain loads), the
// round-to-nearest arithmetic intrinsics (plain IEEE operations; the emulator is compiled with
// -ffp-contract=off, matching nvcc -fmad=false), and __syncthreads() for kernels whose only work
// before the barrier is an idempotent shared-memory fill: the block is run twice, the first time every
// thread stops at the barrier (exception), the second time the barrier is a no-op.
#pragma once
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>

struct EmulDim3 { unsigned x = 0, y = 0, z = 0; };
static EmulDim3 threadIdx, blockIdx, blockDim, gridDim;
static bool g_emul_stop_at_barrier = false;
struct EmulBarrier {};

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

inline void __syncthreads() { if (g_emul_stop_at_barrier) throw EmulBarrier(); }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcs(const T *p) { return *p; }
template <typename T> inline T __ldca(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *p; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline void __nanosleep(unsigned) {}
inline void __syncwarp() {}
inline void __threadfence_system() {}
inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p += v; return o; }

// device-wide nanosecond clock (%globaltimer on the GPU)
inline unsigned long long global_timer_ns() {
  return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}
