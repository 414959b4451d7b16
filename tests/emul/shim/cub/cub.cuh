// TEST INFRASTRUCTURE — CPU stand-ins for the CUB device-wide primitives libcpppd uses at setup.
// Same calling convention: first call with d_temp == nullptr returns the workspace size.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "../cuda_runtime.h"

namespace cub {

template <typename T>
struct DoubleBuffer {
  T *d_buffers[2];
  int selector = 0;
  DoubleBuffer(T *a, T *b) { d_buffers[0] = a; d_buffers[1] = b; }
  T *Current() const { return d_buffers[selector]; }
  T *Alternate() const { return d_buffers[selector ^ 1]; }
};

template <typename K>
inline std::vector<int64_t> stable_order(const K *keys, int64_t n, int begin_bit, int end_bit) {
  const int bits = end_bit - begin_bit;
  const unsigned long long mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    return (((unsigned long long)keys[a] >> begin_bit) & mask) < (((unsigned long long)keys[b] >> begin_bit) & mask);
  });
  return order;
}

struct DeviceRadixSort {
  template <typename K, typename V, typename N>
  static cudaError_t SortPairs(void *tmp, size_t &bytes, DoubleBuffer<K> &keys, DoubleBuffer<V> &vals, N n,
                               int begin_bit, int end_bit, cudaStream_t) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    auto order = stable_order(keys.Current(), (int64_t)n, begin_bit, end_bit);
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      keys.Alternate()[i] = keys.Current()[order[i]];
      vals.Alternate()[i] = vals.Current()[order[i]];
    }
    keys.selector ^= 1;
    vals.selector ^= 1;
    return cudaSuccess;
  }
  template <typename K, typename N>
  static cudaError_t SortKeys(void *tmp, size_t &bytes, const K *in, K *out, N n, int begin_bit, int end_bit,
                              cudaStream_t) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    auto order = stable_order(in, (int64_t)n, begin_bit, end_bit);
    for (int64_t i = 0; i < (int64_t)n; ++i) out[i] = in[order[i]];
    return cudaSuccess;
  }
};

struct DeviceScan {
  template <typename T, typename N>
  static cudaError_t ExclusiveSum(void *tmp, size_t &bytes, const T *in, T *out, N n, cudaStream_t) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    T acc = 0;
    for (int64_t i = 0; i < (int64_t)n; ++i) { T v = in[i]; out[i] = acc; acc += v; }
    return cudaSuccess;
  }
};

struct DeviceReduce {
  template <typename T, typename N>
  static cudaError_t Min(void *tmp, size_t &bytes, const T *in, T *out, N n, cudaStream_t) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    *out = *std::min_element(in, in + n);
    return cudaSuccess;
  }
  template <typename T, typename N>
  static cudaError_t Max(void *tmp, size_t &bytes, const T *in, T *out, N n, cudaStream_t) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    *out = *std::max_element(in, in + n);
    return cudaSuccess;
  }
};

struct DeviceSelect {
  template <typename T, typename N>
  static cudaError_t Unique(void *tmp, size_t &bytes, const T *in, T *out, int *num_out, N n, cudaStream_t) {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    int k = 0;
    for (int64_t i = 0; i < (int64_t)n; ++i)
      if (i == 0 || in[i] != in[i - 1]) out[k++] = in[i];
    *num_out = k;
    return cudaSuccess;
  }
};

}  // namespace cub
