// Device helpers: NaN-aware max, |a|^p, warp / block reductions.
// Part of libcpppd (single translation unit, included by cpppd.cu).
#pragma once

namespace {

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double nan_max(double a, double b) {
  // numpy.max semantics: NaN wins
  if (a != a) return a;
  if (b != b) return b;
  return a > b ? a : b;
}

__device__ __forceinline__ double abs_pow(double a, double p) {
  // numpy: np.abs(data) ** p.  numpy special-cases the scalar exponents 1 and 2 (exact), so do we.
  double v = fabs(a);
  if (p == 1.0) return v;
  if (p == 2.0) return __dmul_rn(v, v);
  if (p == 0.0) return 1.0;
  return pow(v, p);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_nanmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = nan_max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Reduce Q per-thread values over the CTA; thread 0 writes them to out[0..Q).
// is_max bit q set -> NaN-propagating max, else sum.
template <int Q>
__device__ __forceinline__ void block_reduce_write(double (&v)[Q], unsigned is_max, double *out) {
  __shared__ double sh[Q][kBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    double r = (is_max >> q) & 1u ? warp_nanmax(v[q]) : warp_sum(v[q]);
    if (lane == 0) sh[q][warp] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      double r = sh[q][0];
      for (int w = 1; w < kBlock / 32; ++w)
        r = (is_max >> q) & 1u ? nan_max(r, sh[q][w]) : __dadd_rn(r, sh[q][w]);
      out[q] = r;
    }
  }
}

// one warp per slice: width = longest row of the slice; out[s] = 32 * width
__global__ void k_slice_extent(const int64_t *__restrict__ rowptr, int64_t nrows, int64_t nslices,
                               int64_t *__restrict__ extent) {
  int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= nslices) return;
  int64_t r = s * kSlice + lane;
  int64_t len = r < nrows ? rowptr[r + 1] - rowptr[r] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if (lane == 0) extent[s] = len * kSlice;
}

// position of `v` (compared by bit pattern) in the sorted dictionary, or -1
__device__ __forceinline__ int dict_find(const unsigned long long *__restrict__ dict, int ndict, double v) {
  const unsigned long long key = (unsigned long long)__double_as_longlong(v);
  int lo = 0, hi = ndict - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const unsigned long long d = dict[mid];
    if (d == key) return mid;
    if (d < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// one warp per slice: copy CSR entries into the column-major slice, pad with idx = kPad.
// With a dictionary the value is folded into the index word as a code and `val` is not written.
__global__ void k_fill_sell(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices,
                            const double *__restrict__ values, int64_t nrows, int64_t nslices,
                            const int64_t *__restrict__ slice_ptr, int32_t *__restrict__ idx,
                            double *__restrict__ val, const unsigned long long *__restrict__ dict, int ndict,
                            int idx_bits) {
  int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= nslices) return;
  int64_t r = s * kSlice + lane;
  int64_t p0 = slice_ptr[s], p1 = slice_ptr[s + 1];
  int64_t e0 = 0, len = 0;
  if (r < nrows) {
    e0 = rowptr[r];
    len = rowptr[r + 1] - e0;
  }
  int64_t width = (p1 - p0) / kSlice;
  for (int64_t k = 0; k < width; ++k) {
    int64_t p = p0 + k * kSlice + lane;
    if (k < len) {
      int32_t w = indices[e0 + k];
      if (dict) {
        const int code = dict_find(dict, ndict, values[e0 + k]);
        w = (w & kEqBit) | (w & ((1 << idx_bits) - 1)) | (code << idx_bits);
      } else {
        val[p] = values[e0 + k];
      }
      idx[p] = w;
    } else {
      idx[p] = kPad;
      if (!dict) val[p] = 0.0;
    }
  }
}

// flag[0] = 1 when some value is not in the dictionary
__global__ void k_dict_check(const double *__restrict__ values, int64_t nnz, const unsigned long long *__restrict__ dict,
                             int ndict, int *__restrict__ flag) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
    if (dict_find(dict, ndict, values[e]) < 0) *flag = 1;
}

// flag[0] = 1 when some element differs (bitwise) from the first one
__global__ void k_not_constant(const double *__restrict__ v, int64_t count, int *__restrict__ flag) {
  const long long first = __double_as_longlong(v[0]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    if (__double_as_longlong(v[i]) != first) *flag = 1;
}

}  // namespace
