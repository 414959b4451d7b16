// Device-side types shared by every kernel: constants, the SELL view, scalar-foldable vector operands.
// Part of libcpppd (single translation unit, included by cpppd.cu).  Deliberately free of any CUDA
// runtime / NCCL / CUB dependency: tests/emul/ compiles this header and cpppd_hot_kernels.cuh for the
// CPU (with a small shim for the CUDA keywords) to run the real kernel code without a GPU.
#pragma once

#include <climits>
#include <cstdint>
#include <type_traits>

namespace {

constexpr int kSlice = 32;       // SELL slice height C (= warp size)
constexpr int kBlock = 256;      // threads per CTA for the streaming kernels (8 slices)
constexpr int kGraphChunk = 50;  // iterations captured per CUDA graph
constexpr int kColQ = 5;         // column-pass partials per CTA: 4 sums + max bound violation
constexpr int kRowQ = 9;         // row-pass partials per CTA: 4 sums + 5 maxima (the last two with the row offsets)
constexpr int kGtQ = 2;          // ground-truth pass: sum |gt - x|, sum |gt - round(x)|
constexpr int kStatQ = kColQ + kRowQ + kGtQ;
// entries of the per-rank stats vector that are folded with a (NaN-propagating) max instead of a sum
__host__ __device__ constexpr bool stat_is_max(int q) { return q == 4 || (q >= kColQ + 4 && q < kColQ + kRowQ); }
constexpr int32_t kEqBit = 0x40000000;   // A^T entries: set when the source row is an equality
constexpr int32_t kIdxMask = 0x3fffffff;
// padding entry of a slice: negative, and its masked index is 0 so that a gather the compiler
// hoists above the `idx >= 0` test still reads a valid address
constexpr int32_t kPad = INT32_MIN;
constexpr int kMaxWorld = 64;

struct SellView {
  const int64_t *__restrict__ slice_ptr;
  const int32_t *__restrict__ idx;
  const double *__restrict__ val;
  int64_t nrows, nslices;
  int64_t uniform_width;  // -1: read slice_ptr
  const double *__restrict__ dict;
  int32_t idx_mask;       // low bits of an entry word that hold the gather index
  int32_t idx_bits, code_mask, ndict;
};

// a vector operand that may have been folded into a scalar (CPPPD_FLAG_CONST_VECTORS)
struct Vec {
  const double *p;
  double c;
  __device__ __forceinline__ double at(int64_t i) const { return p ? __ldcs(p + i) : c; }
};

// first / one-past-last element offset of slice s
__device__ __forceinline__ void slice_range(const SellView &S, int64_t s, int64_t &p0, int64_t &p1) {
  if (S.uniform_width >= 0) {
    p0 = s * S.uniform_width * 32;
    p1 = p0 + S.uniform_width * 32;
  } else {
    p0 = __ldg(S.slice_ptr + s);
    p1 = __ldg(S.slice_ptr + s + 1);
  }
}

// value of the entry stored at position p whose index word is w (non-hot kernels)
__device__ __forceinline__ double entry_value(const SellView &S, int64_t p, int32_t w) {
  return S.dict ? S.dict[(w >> S.idx_bits) & S.code_mask] : S.val[p];
}

// bookkeeping of the peer-memory halo exchange (device resident)
struct SyncState {
  unsigned long long push_stamp[2];  // halos pushed so far        ([0] xbar, [1] y)
  unsigned long long wait_stamp[2];  // halos consumed so far
  unsigned int ticket[2];            // CTA arrival counter of k_push
  unsigned int timed_out;            // a wait gave up (see wait_for_stamp)
  unsigned int pad;
  unsigned long long timeout_ns;     // 0: wait for ever
};

// nanoseconds of a clock that is common to all threads of the device
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif
}  // namespace
