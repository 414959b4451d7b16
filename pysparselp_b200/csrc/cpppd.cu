// libcpppd — B200 (sm_100a) core of the Chambolle-Pock PPD LP solver.  C ABI in include/cpppd.h.
//
// What runs here is the reference's pysparselp/ChambollePockPPD.py:122-343, restated
// for the GPU (file:line citations refer to that file unless noted):
//   * operator storage : A = [A_eq; A_ineq] and A^T, each in SELL-32 (sliced ELLPACK, slice
//                        height = warp size).  One thread owns one row; lane l of a warp
//                        reads element k of its row at  base + 32*k + l, so every warp-level
//                        load of values / indices is one contiguous 256 B / 128 B segment.
//   * k_primal         : d = c + A^T y (:198-217) fused with x - T d, the box clip and the
//                        theta extrapolation (:220-228).  d never reaches memory except on
//                        stats iterations.
//   * k_dual           : r = A xbar - b (:231-240) fused with y + Sigma r and the projection
//                        of y_ineq on >= 0 (:333-341).  r never reaches memory.
//   * k_precond_*      : column / row abs-power sums -> diag_t, diag_sigma (:122-179).
//   * k_stats_*        : the stats block (:248-291) as warp-shuffle + block reductions with a
//                        deterministic two-level tree, the best-integer bookkeeping on device.
//
// Floating point: IEEE fp64, compiled with -fmad=false and written with explicit
// __dmul_rn/__dadd_rn so products and sums round exactly like the numpy/scipy code of the
// reference.  A row (column) sum is accumulated sequentially in the stored entry order from
// 0.0 — the same order as scipy's csr_matvec (csc_matvec) — so x, xbar, y, T and Sigma are
// bit-identical to the reference on a single GPU.  Only the scalar dot products of the
// stats block use a different (tree) order than numpy.dot.
#include "../../include/cpppd.h"

#include <cuda_runtime.h>

#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

constexpr int kSlice = 32;       // SELL slice height C (= warp size)
constexpr int kBlock = 256;      // threads per CTA for the streaming kernels (8 slices)
constexpr int kGraphChunk = 50;  // iterations captured per CUDA graph
constexpr int kColQ = 4;         // column-pass partial sums per CTA
constexpr int kRowQ = 7;         // row-pass partial sums per CTA

thread_local std::string g_create_error;

struct Sell {
  int64_t nrows = 0, nslices = 0, padded = 0;
  int64_t uniform_width = -1;    // >= 0 when every slice has this width (slice_ptr is then implicit)
  int64_t *slice_ptr = nullptr;  // nslices+1 element offsets
  int32_t *idx = nullptr;        // padded entries, -1 = padding
  double *val = nullptr;
};

struct SellView {
  const int64_t *__restrict__ slice_ptr;
  const int32_t *__restrict__ idx;
  const double *__restrict__ val;
  int64_t nrows, nslices;
  int64_t uniform_width;  // -1: read slice_ptr
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

struct StatsDev {  // device-resident, copied verbatim into cpppd_stats
  cpppd_stats s;
};

}  // namespace

struct cpppd_solver {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t n = 0, m_eq = 0, m_ineq = 0, m = 0, nnz = 0;
  double alpha = 1, theta = 1, one_plus_theta = 2;
  uint32_t flags = 0;
  cpppd_alloc_fn alloc = nullptr;
  cpppd_free_fn free_fn = nullptr;
  void *alloc_user = nullptr;
  std::vector<void *> owned;
  int64_t device_bytes = 0;
  Sell A, AT;
  double *c = nullptr, *T = nullptr, *lb = nullptr, *ub = nullptr, *x = nullptr, *xbar = nullptr;
  double *b = nullptr, *sigma = nullptr, *y = nullptr, *dbuf = nullptr, *best = nullptr;
  double *colpart = nullptr, *rowpart = nullptr, *xr_scratch = nullptr;
  int stat_blocks_c = 0, stat_blocks_r = 0;
  StatsDev *stats_dev = nullptr;
  cpppd_stats *stats_host = nullptr;
  int64_t niter = 0;
  bool mid_iteration = false;  // primal step issued, dual step pending
  bool stats_pending = false;
  bool have_d = false;
  int sm_count = 148;
  std::map<int64_t, cudaGraphExec_t> graphs;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int sticky = 0;
};

namespace {

int fail(cpppd_solver *h, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) {
    h->err = buf;
    if (code == CPPPD_ERR_CUDA) h->sticky = code;
  }
  g_create_error = buf;
  return code;
}

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(h, CPPPD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                        \
  } while (0)

#define CHECK_HANDLE(h)                                  \
  do {                                                   \
    if (!(h)) return CPPPD_ERR_INVALID;                  \
    if ((h)->sticky) return (h)->sticky;                 \
    cudaSetDevice((h)->device);                          \
  } while (0)

void *dev_alloc(cpppd_solver *h, size_t bytes, bool persistent) {
  if (bytes == 0) bytes = 256;
  void *p = nullptr;
  if (h->alloc) {
    p = h->alloc(bytes, h->alloc_user);
  } else if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    p = nullptr;
  }
  if (p && persistent) {
    h->owned.push_back(p);
    h->device_bytes += (int64_t)bytes;
  }
  return p;
}

void dev_free(cpppd_solver *h, void *p) {
  if (!p) return;
  if (h->alloc) {
    if (h->free_fn) h->free_fn(p, h->alloc_user);
  } else {
    cudaFree(p);
  }
}

template <typename T>
int alloc_array(cpppd_solver *h, T **out, int64_t count, bool persistent = true) {
  *out = static_cast<T *>(dev_alloc(h, sizeof(T) * (size_t)std::max<int64_t>(count, 1), persistent));
  if (!*out) return fail(h, CPPPD_ERR_NOMEM, "device allocation of %lld bytes failed", (long long)(sizeof(T) * count));
  return 0;
}

inline int grid_for(int64_t items, int block = kBlock) { return (int)((items + block - 1) / block); }

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

// ------------------------------------------------------------------------------------------
// setup kernels: CSR -> SELL-32, transpose, preconditioners
// ------------------------------------------------------------------------------------------
__global__ void k_widen_indptr(const int32_t *__restrict__ in, int64_t *__restrict__ out, int64_t count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = in[i];
}

// flags[0] |= 1 when a row has negative length, |= 2 when a column index is out of range
__global__ void k_validate(const int64_t *__restrict__ rowptr, int64_t m, const int32_t *__restrict__ indices,
                           int64_t nnz, int64_t n, int *flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0;
  if (i < m && rowptr[i + 1] < rowptr[i]) bad |= 1;
  for (int64_t e = i; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
    int32_t j = indices[e];
    if (j < 0 || j >= n) bad |= 2;
  }
  if (bad) atomicOr(flags, bad);
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

// one warp per slice: copy CSR entries into the column-major slice, pad with idx = -1
__global__ void k_fill_sell(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices,
                            const double *__restrict__ values, int64_t nrows, int64_t nslices,
                            const int64_t *__restrict__ slice_ptr, int32_t *__restrict__ idx,
                            double *__restrict__ val) {
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
      idx[p] = indices[e0 + k];
      val[p] = values[e0 + k];
    } else {
      idx[p] = -1;
      val[p] = 0.0;
    }
  }
}

__global__ void k_row_of_entry(const int64_t *__restrict__ rowptr, int64_t m, int64_t nnz,
                               uint32_t *__restrict__ row_of, uint32_t *__restrict__ entry_id) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int64_t lo = 0, hi = m;  // last row with rowptr[row] <= e
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (rowptr[mid] <= e) lo = mid; else hi = mid;
  }
  row_of[e] = (uint32_t)lo;
  entry_id[e] = (uint32_t)e;
}

__global__ void k_colptr_from_sorted(const uint32_t *__restrict__ sorted_cols, int64_t nnz, int64_t n,
                                     int64_t *__restrict__ colptr) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n) return;
  int64_t lo = 0, hi = nnz;  // first position with sorted_cols[pos] >= j
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)sorted_cols[mid] < j) lo = mid + 1; else hi = mid;
  }
  colptr[j] = lo;
}

__global__ void k_gather_transposed(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ row_of,
                                    const double *__restrict__ values, int64_t nnz, int32_t *__restrict__ t_idx,
                                    double *__restrict__ t_val) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  uint32_t e = perm[p];
  t_idx[p] = (int32_t)row_of[e];
  t_val[p] = values[e];
}

// diag_t (:122-153): thread per column of A, sequential over the column in row order,
// equality and inequality parts accumulated separately then  (0 + s_eq) + s_ineq.
__global__ void k_precond_cols(SellView AT, int64_t n, int64_t m_eq, int has_eq, int has_ineq, double power,
                               double *__restrict__ T) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t s = j >> 5;
  if (s >= AT.nslices) return;
  int lane = threadIdx.x & 31;
  int64_t p0, p1;
  slice_range(AT, s, p0, p1);
  double s_eq = 0.0, s_in = 0.0;
  for (int64_t p = p0 + lane; p < p1; p += kSlice) {
    int32_t r = AT.idx[p];
    if (r >= 0) {
      double t = __dmul_rn(abs_pow(AT.val[p], power), 1.0);
      if (r < m_eq) s_eq = __dadd_rn(s_eq, t); else s_in = __dadd_rn(s_in, t);
    }
  }
  if (j < n) {
    double tmp = 0.0;
    if (has_eq) tmp = __dadd_rn(tmp, s_eq);
    if (has_ineq) tmp = __dadd_rn(tmp, s_in);
    if (tmp == 0.0) tmp = 1.0;
    T[j] = __ddiv_rn(1.0, tmp);
  }
}

// diag_sigma (:158-179): thread per row, sequential in stored order.
__global__ void k_precond_rows(SellView A, int64_t m, double power, double *__restrict__ sigma) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t s = i >> 5;
  if (s >= A.nslices) return;
  int lane = threadIdx.x & 31;
  int64_t p0, p1;
  slice_range(A, s, p0, p1);
  double acc = 0.0;
  for (int64_t p = p0 + lane; p < p1; p += kSlice) {
    if (A.idx[p] >= 0) acc = __dadd_rn(acc, __dmul_rn(abs_pow(A.val[p], power), 1.0));
  }
  if (i < m) {
    if (acc == 0.0) acc = 1.0;
    sigma[i] = __ddiv_rn(1.0, acc);
  }
}

// ------------------------------------------------------------------------------------------
// the two hot kernels
// ------------------------------------------------------------------------------------------
// Primal half-iteration (:198-228).  Thread j owns column j of A (row j of A^T).
// Loads that do not depend on the matrix (c, T, x, lb, ub) are issued first so that they are
// in flight together with the slice entries; matrix entries are read once (ld.global.cs).
template <bool kWriteD>
__global__ void __launch_bounds__(kBlock, 8)
k_primal(SellView AT, const double *__restrict__ y, const double *__restrict__ c, const double *__restrict__ T,
         const double *__restrict__ lb, const double *__restrict__ ub, double *__restrict__ x,
         double *__restrict__ xbar, double *__restrict__ d_out, int64_t n, int64_t m_eq, int has_eq, int has_ineq,
         double theta, double one_plus_theta) {
  const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t s = j >> 5;
  if (s >= AT.nslices) return;
  const int lane = threadIdx.x & 31;
  int64_t p0, p1;
  slice_range(AT, s, p0, p1);
  const bool live = j < n;
  double cj = 0.0, tj = 0.0, xo = 0.0;
  if (live) {
    cj = __ldcs(c + j);
    tj = __ldcs(T + j);
    xo = __ldcs(x + j);
  }
  double s_eq = 0.0, s_in = 0.0;
  {
    const int32_t *ip = AT.idx + p0 + lane;
    const double *vp = AT.val + p0 + lane;
    const int width = (int)((p1 - p0) >> 5);
    const int32_t meq = (int32_t)m_eq;
#pragma unroll 4
    for (int k = 0; k < width; ++k) {
      const int32_t r = __ldcs(ip + k * kSlice);
      const double a = __ldcs(vp + k * kSlice);
      if (r >= 0) {
        const double t = __dmul_rn(a, __ldg(y + r));
        if (r < meq) s_eq = __dadd_rn(s_eq, t); else s_in = __dadd_rn(s_in, t);
      }
    }
  }
  if (!live) return;
  double d = cj;
  if (has_eq) d = __dadd_rn(d, s_eq);
  if (has_ineq) d = __dadd_rn(d, s_in);
  const double l = __ldcs(lb + j), u = __ldcs(ub + j);
  double x2 = __dsub_rn(xo, __dmul_rn(tj, d));
  x2 = (l > x2) ? l : x2;  // np.maximum(x2, lb)  (NaN in x2 propagates)
  x2 = (u < x2) ? u : x2;  // np.minimum(x2, ub)
  xbar[j] = __dsub_rn(__dmul_rn(one_plus_theta, x2), __dmul_rn(theta, xo));
  x[j] = x2;
  if (kWriteD) d_out[j] = d;
}

// Dual half-iteration (:231-240, :333-341).  Thread i owns row i of A.
__global__ void __launch_bounds__(kBlock, 8)
k_dual(SellView A, const double *__restrict__ xbar, const double *__restrict__ b,
       const double *__restrict__ sigma, double *__restrict__ y, int64_t m, int64_t m_eq) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t s = i >> 5;
  if (s >= A.nslices) return;
  const int lane = threadIdx.x & 31;
  int64_t p0, p1;
  slice_range(A, s, p0, p1);
  const bool live = i < m;
  double bi = 0.0, si = 0.0, yi = 0.0;
  if (live) {
    bi = __ldcs(b + i);
    si = __ldcs(sigma + i);
    yi = __ldcs(y + i);
  }
  double acc = 0.0;
  {
    const int32_t *ip = A.idx + p0 + lane;
    const double *vp = A.val + p0 + lane;
    const int width = (int)((p1 - p0) >> 5);
#pragma unroll 4
    for (int k = 0; k < width; ++k) {
      const int32_t jc = __ldcs(ip + k * kSlice);
      const double a = __ldcs(vp + k * kSlice);
      if (jc >= 0) acc = __dadd_rn(acc, __dmul_rn(a, __ldg(xbar + jc)));
    }
  }
  if (!live) return;
  const double r = __dsub_rn(acc, bi);
  double yn = __dadd_rn(yi, __dmul_rn(si, r));
  if (i >= m_eq) yn = (yn < 0.0) ? 0.0 : yn;  // np.maximum(y_ineq, 0): NaN stays NaN, -0.0 stays
  y[i] = yn;
}

// ------------------------------------------------------------------------------------------
// stats block (:248-291)
// ------------------------------------------------------------------------------------------
// Column pass: c.x, c.x4, c.xr, #(xbar == 0); turns the d buffer into x4 in place and
// (force_integer) stores xr into xr_out.
__global__ void __launch_bounds__(kBlock)
k_stats_cols(const double *__restrict__ c, const double *__restrict__ x, const double *__restrict__ xbar,
             const double *__restrict__ lb, const double *__restrict__ ub, double *__restrict__ d_x4,
             double *__restrict__ xr_out, int64_t n, int force_integer, double *__restrict__ part) {
  double v[kColQ] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x; j < n; j += (int64_t)gridDim.x * kBlock) {
    const double cj = c[j], xj = x[j];
    const double x4 = d_x4[j] < 0.0 ? ub[j] : lb[j];  // x4 = lb; x4[d < 0] = ub[d < 0]  (:260-261)
    d_x4[j] = x4;
    double xr = xj;
    if (force_integer) {
      xr = rint(xj);  // np.round: half to even
      xr_out[j] = xr;
    }
    v[0] = __dadd_rn(v[0], __dmul_rn(cj, xj));
    v[1] = __dadd_rn(v[1], __dmul_rn(cj, x4));
    v[2] = __dadd_rn(v[2], __dmul_rn(cj, xr));
    v[3] = __dadd_rn(v[3], xbar[j] == 0.0 ? 1.0 : 0.0);
  }
  block_reduce_write<kColQ>(v, 0u, part + (int64_t)blockIdx.x * kColQ);
}

// Row pass: A x, A x4, A xbar, A xr per row -> energy terms and violation maxima.
__global__ void __launch_bounds__(kBlock)
k_stats_rows(SellView A, const double *__restrict__ x, const double *__restrict__ x4,
             const double *__restrict__ xbar, const double *__restrict__ xr, const double *__restrict__ b,
             const double *__restrict__ y, int64_t m, int64_t m_eq, int force_integer,
             double *__restrict__ part) {
  const double ninf = -INFINITY;
  double v[kRowQ] = {0.0, 0.0, 0.0, 0.0, ninf, ninf, ninf};
  const int lane = threadIdx.x & 31;
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; (i >> 5) < A.nslices;
       i += (int64_t)gridDim.x * kBlock) {
    const int64_t s = i >> 5;
    int64_t p0, p1;
    slice_range(A, s, p0, p1);
    double ax = 0.0, ax4 = 0.0, axb = 0.0, axr = 0.0;
    for (int64_t p = p0 + lane; p < p1; p += kSlice) {
      const int32_t jc = A.idx[p];
      if (jc >= 0) {
        const double a = A.val[p];
        ax = __dadd_rn(ax, __dmul_rn(a, x[jc]));
        ax4 = __dadd_rn(ax4, __dmul_rn(a, x4[jc]));
        if (i < m_eq) axb = __dadd_rn(axb, __dmul_rn(a, xbar[jc]));
        if (force_integer) axr = __dadd_rn(axr, __dmul_rn(a, xr[jc]));
      }
    }
    if (i < m) {
      if (!force_integer) axr = ax;
      const double bi = b[i], yi = y[i];
      const double t1 = __dmul_rn(yi, __dsub_rn(ax, bi));
      const double t2 = __dmul_rn(yi, __dsub_rn(ax4, bi));
      if (i < m_eq) {
        v[0] = __dadd_rn(v[0], t1);
        v[2] = __dadd_rn(v[2], t2);
        v[4] = nan_max(v[4], fabs(__dsub_rn(axb, bi)));
        v[5] = nan_max(v[5], fabs(__dsub_rn(axr, bi)));
      } else {
        v[1] = __dadd_rn(v[1], t1);
        v[3] = __dadd_rn(v[3], t2);
        v[6] = nan_max(v[6], __dsub_rn(axr, bi));
      }
    }
  }
  block_reduce_write<kRowQ>(v, 0x70u, part + (int64_t)blockIdx.x * kRowQ);
}

// One CTA: fold the per-CTA partials in a fixed order, apply :248-291's scalar logic.
__global__ void __launch_bounds__(kBlock)
k_stats_final(const double *__restrict__ colpart, int nbc, const double *__restrict__ rowpart, int nbr,
              int64_t n, int has_eq, int has_ineq, int64_t niter, StatsDev *out) {
  double cv[kColQ] = {0.0, 0.0, 0.0, 0.0};
  const double ninf = -INFINITY;
  double rv[kRowQ] = {0.0, 0.0, 0.0, 0.0, ninf, ninf, ninf};
  for (int bi = threadIdx.x; bi < nbc; bi += kBlock)
#pragma unroll
    for (int q = 0; q < kColQ; ++q) cv[q] = __dadd_rn(cv[q], colpart[(int64_t)bi * kColQ + q]);
  for (int bi = threadIdx.x; bi < nbr; bi += kBlock) {
#pragma unroll
    for (int q = 0; q < 4; ++q) rv[q] = __dadd_rn(rv[q], rowpart[(int64_t)bi * kRowQ + q]);
#pragma unroll
    for (int q = 4; q < kRowQ; ++q) rv[q] = nan_max(rv[q], rowpart[(int64_t)bi * kRowQ + q]);
  }
  __shared__ double fin[kColQ + kRowQ];
  block_reduce_write<kColQ>(cv, 0u, fin);
  __syncthreads();
  block_reduce_write<kRowQ>(rv, 0x70u, fin + kColQ);
  __syncthreads();
  if (threadIdx.x != 0) return;
  cpppd_stats &s = out->s;
  double e1 = fin[0], e2 = fin[1];
  if (has_eq) {
    e1 = __dadd_rn(e1, fin[kColQ + 0]);
    e2 = __dadd_rn(e2, fin[kColQ + 2]);
  }
  if (has_ineq) {
    e1 = __dadd_rn(e1, fin[kColQ + 1]);
    e2 = __dadd_rn(e2, fin[kColQ + 3]);
  }
  s.niter = niter;
  s.energy1 = e1;
  s.energy2 = e2;
  s.max_violated_equality = has_eq ? fin[kColQ + 4] : 0.0;
  s.max_violated_equality_rounded = has_eq ? fin[kColQ + 5] : 0.0;
  s.max_violated_inequality = fin[kColQ + 6];  // -inf when there is no inequality row
  s.energy_rounded = fin[2];
  s.frac_zero_xbar = n > 0 ? fin[3] / (double)n : 0.0;
  const int feasible = (s.max_violated_equality_rounded == 0.0) && (s.max_violated_inequality <= 0.0);
  s.feasible = feasible;
  s.improved = 0;
  if (feasible && s.energy_rounded < s.best_integer_energy) {  // :284-291
    s.best_integer_energy = s.energy_rounded;
    s.improved = 1;
    s.have_best_integer = 1;
  }
}

__global__ void k_snapshot_best(const StatsDev *st, const double *__restrict__ src, double *__restrict__ best,
                                int64_t n) {
  if (!st->s.improved) return;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    best[j] = src[j];
}

__global__ void k_init_stats(StatsDev *st) {
  memset(&st->s, 0, sizeof(cpppd_stats));
  st->s.best_integer_energy = INFINITY;  // :192
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
SellView view(const Sell &s) { return SellView{s.slice_ptr, s.idx, s.val, s.nrows, s.nslices, s.uniform_width}; }

// CSR (device, int64 rowptr) -> SELL-32 (device)
int build_sell(cpppd_solver *h, const int64_t *rowptr, const int32_t *indices, const double *values, int64_t nrows,
               Sell *out) {
  out->nrows = nrows;
  out->nslices = (nrows + kSlice - 1) / kSlice;
  int64_t ns = out->nslices;
  int64_t *extent = nullptr;
  if (int rc = alloc_array(h, &extent, ns + 1, false)) return rc;
  if (int rc = alloc_array(h, &out->slice_ptr, ns + 1)) return rc;
  CK(cudaMemsetAsync(extent, 0, sizeof(int64_t) * (ns + 1), h->stream));
  if (ns) k_slice_extent<<<grid_for(ns * 32), kBlock, 0, h->stream>>>(rowptr, nrows, ns, extent);
  size_t tmp_bytes = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, extent, out->slice_ptr, ns + 1, h->stream));
  void *tmp = dev_alloc(h, tmp_bytes, false);
  if (!tmp) return fail(h, CPPPD_ERR_NOMEM, "scan workspace allocation failed");
  CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, extent, out->slice_ptr, ns + 1, h->stream));
  CK(cudaMemcpyAsync(&out->padded, out->slice_ptr + ns, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  {  // uniform slice width <=> min extent == max extent
    int64_t *mm = nullptr;
    if (int rc = alloc_array(h, &mm, 2, false)) return rc;
    size_t b1 = 0, b2 = 0;
    CK(cub::DeviceReduce::Min(nullptr, b1, extent, mm, ns, h->stream));
    CK(cub::DeviceReduce::Max(nullptr, b2, extent, mm + 1, ns, h->stream));
    void *t2 = dev_alloc(h, std::max(b1, b2), false);
    if (!t2) return fail(h, CPPPD_ERR_NOMEM, "reduce workspace allocation failed");
    int64_t host_mm[2] = {0, 1};
    if (ns) {
      CK(cub::DeviceReduce::Min(t2, b1, extent, mm, ns, h->stream));
      CK(cub::DeviceReduce::Max(t2, b2, extent, mm + 1, ns, h->stream));
      CK(cudaMemcpyAsync(host_mm, mm, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    out->uniform_width = (ns && host_mm[0] == host_mm[1]) ? host_mm[0] / kSlice : -1;
    dev_free(h, t2);
    dev_free(h, mm);
  }
  CK(cudaStreamSynchronize(h->stream));
  dev_free(h, tmp);
  dev_free(h, extent);
  if (int rc = alloc_array(h, &out->idx, out->padded)) return rc;
  if (int rc = alloc_array(h, &out->val, out->padded)) return rc;
  if (ns) k_fill_sell<<<grid_for(ns * 32), kBlock, 0, h->stream>>>(rowptr, indices, values, nrows, ns, out->slice_ptr,
                                                                  out->idx, out->val);
  CK(cudaGetLastError());
  return 0;
}

int upload(cpppd_solver *h, double *dst, const double *src, int64_t count) {
  if (count == 0) return 0;
  CK(cudaMemcpyAsync(dst, src, sizeof(double) * count, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

int setup(cpppd_solver *h, const cpppd_problem *P) {
  const int64_t n = h->n, m = h->m, nnz = h->nnz;
  cudaStream_t st = h->stream;
  // ---- vectors
  for (double **v : {&h->c, &h->T, &h->lb, &h->ub, &h->x, &h->xbar, &h->dbuf, &h->best})
    if (int rc = alloc_array(h, v, n)) return rc;
  for (double **v : {&h->b, &h->sigma, &h->y})
    if (int rc = alloc_array(h, v, m)) return rc;
  if (int rc = upload(h, h->c, P->c, n)) return rc;
  if (int rc = upload(h, h->lb, P->lb, n)) return rc;
  if (int rc = upload(h, h->ub, P->ub, n)) return rc;
  if (int rc = upload(h, h->b, P->b, m)) return rc;
  if (P->x0) {
    if (int rc = upload(h, h->x, P->x0, n)) return rc;
  } else {
    CK(cudaMemsetAsync(h->x, 0, sizeof(double) * std::max<int64_t>(n, 1), st));
  }
  CK(cudaMemcpyAsync(h->xbar, h->x, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));  // x3 = x (:190)
  CK(cudaMemsetAsync(h->y, 0, sizeof(double) * std::max<int64_t>(m, 1), st));            // :166,:177
  // ---- CSR of A on the device (temporary)
  int64_t *rowptr = nullptr;
  int32_t *indices = nullptr;
  double *values = nullptr;
  if (int rc = alloc_array(h, &rowptr, m + 1, false)) return rc;
  if (int rc = alloc_array(h, &indices, nnz, false)) return rc;
  if (int rc = alloc_array(h, &values, nnz, false)) return rc;
  if (P->indptr_bits == 64) {
    CK(cudaMemcpyAsync(rowptr, P->indptr, sizeof(int64_t) * (m + 1), cudaMemcpyHostToDevice, st));
  } else {
    int32_t *tmp32 = nullptr;
    if (int rc = alloc_array(h, &tmp32, m + 1, false)) return rc;
    CK(cudaMemcpyAsync(tmp32, P->indptr, sizeof(int32_t) * (m + 1), cudaMemcpyHostToDevice, st));
    k_widen_indptr<<<grid_for(m + 1), kBlock, 0, st>>>(tmp32, rowptr, m + 1);
    CK(cudaStreamSynchronize(st));
    dev_free(h, tmp32);
  }
  if (nnz) {
    CK(cudaMemcpyAsync(indices, P->indices, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(values, P->values, sizeof(double) * nnz, cudaMemcpyHostToDevice, st));
  }
  {  // validation on the device: monotone row pointers, column indices in range
    int *flag = nullptr;
    if (int rc = alloc_array(h, &flag, 1, false)) return rc;
    CK(cudaMemsetAsync(flag, 0, sizeof(int), st));
    int64_t items = std::max<int64_t>(m, std::min<int64_t>(nnz, (int64_t)h->sm_count * 64 * kBlock));
    if (items) k_validate<<<grid_for(items), kBlock, 0, st>>>(rowptr, m, indices, nnz, n, flag);
    int host_flag = 0;
    CK(cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    dev_free(h, flag);
    if (host_flag & 1) return fail(h, CPPPD_ERR_INVALID, "indptr is not non-decreasing");
    if (host_flag & 2) return fail(h, CPPPD_ERR_INVALID, "column index outside [0, n)");
  }
  // ---- A in SELL-32
  if (int rc = build_sell(h, rowptr, indices, values, m, &h->A)) return rc;
  // ---- transpose: stable radix sort of the entries by column keeps, inside each column, the
  //      row order of the CSR — exactly the accumulation order of scipy's csc_matvec.
  {
    uint32_t *row_of = nullptr, *keys_a = nullptr, *keys_b = nullptr, *ids_a = nullptr, *ids_b = nullptr;
    int64_t *colptr = nullptr;
    int32_t *t_idx = nullptr;
    double *t_val = nullptr;
    if (int rc = alloc_array(h, &row_of, nnz, false)) return rc;
    if (int rc = alloc_array(h, &keys_a, nnz, false)) return rc;
    if (int rc = alloc_array(h, &keys_b, nnz, false)) return rc;
    if (int rc = alloc_array(h, &ids_a, nnz, false)) return rc;
    if (int rc = alloc_array(h, &ids_b, nnz, false)) return rc;
    if (int rc = alloc_array(h, &colptr, n + 1, false)) return rc;
    if (nnz) {
      k_row_of_entry<<<grid_for(nnz), kBlock, 0, st>>>(rowptr, m, nnz, row_of, ids_a);
      CK(cudaMemcpyAsync(keys_a, indices, sizeof(uint32_t) * nnz, cudaMemcpyDeviceToDevice, st));
    }
    int end_bit = 1;
    while (end_bit < 32 && ((int64_t)1 << end_bit) < n) ++end_bit;
    cub::DoubleBuffer<uint32_t> keys(keys_a, keys_b), ids(ids_a, ids_b);
    if (nnz) {
      size_t tmp_bytes = 0;
      CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, ids, nnz, 0, end_bit, st));
      void *tmp = dev_alloc(h, tmp_bytes, false);
      if (!tmp) return fail(h, CPPPD_ERR_NOMEM, "sort workspace allocation failed");
      CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, ids, nnz, 0, end_bit, st));
      CK(cudaStreamSynchronize(st));
      dev_free(h, tmp);
    }
    k_colptr_from_sorted<<<grid_for(n + 1), kBlock, 0, st>>>(keys.Current(), nnz, n, colptr);
    // reuse the now dead CSR index/key buffers?  keep it simple: dedicated outputs
    if (int rc = alloc_array(h, &t_idx, nnz, false)) return rc;
    if (int rc = alloc_array(h, &t_val, nnz, false)) return rc;
    if (nnz) k_gather_transposed<<<grid_for(nnz), kBlock, 0, st>>>(ids.Current(), row_of, values, nnz, t_idx, t_val);
    CK(cudaStreamSynchronize(st));
    dev_free(h, row_of);
    dev_free(h, keys_a);
    dev_free(h, keys_b);
    dev_free(h, ids_a);
    dev_free(h, ids_b);
    dev_free(h, indices);
    dev_free(h, values);
    dev_free(h, rowptr);
    if (int rc = build_sell(h, colptr, t_idx, t_val, n, &h->AT)) return rc;
    CK(cudaStreamSynchronize(st));
    dev_free(h, colptr);
    dev_free(h, t_idx);
    dev_free(h, t_val);
  }
  // ---- preconditioners (:122-179)
  const int has_eq = h->m_eq > 0, has_ineq = h->m_ineq > 0;
  if (h->AT.nslices)
    k_precond_cols<<<grid_for(h->AT.nslices * 32), kBlock, 0, st>>>(view(h->AT), n, h->m_eq, has_eq, has_ineq,
                                                                     2.0 - h->alpha, h->T);
  if (h->A.nslices)
    k_precond_rows<<<grid_for(h->A.nslices * 32), kBlock, 0, st>>>(view(h->A), m, h->alpha, h->sigma);
  // ---- stats plumbing
  h->stat_blocks_c = (int)std::max<int64_t>(1, std::min<int64_t>(grid_for(n), (int64_t)h->sm_count * 8));
  h->stat_blocks_r = (int)std::max<int64_t>(1, std::min<int64_t>(grid_for(h->A.nslices * 32), (int64_t)h->sm_count * 8));
  if (int rc = alloc_array(h, &h->colpart, (int64_t)h->stat_blocks_c * kColQ)) return rc;
  if (int rc = alloc_array(h, &h->rowpart, (int64_t)h->stat_blocks_r * kRowQ)) return rc;
  if (int rc = alloc_array(h, &h->stats_dev, 1)) return rc;
  k_init_stats<<<1, 1, 0, st>>>(h->stats_dev);
  CK(cudaMallocHost(&h->stats_host, sizeof(cpppd_stats)));
  memset(h->stats_host, 0, sizeof(cpppd_stats));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  return 0;
}

int launch_primal(cpppd_solver *h, bool write_d) {
  if (h->AT.nslices == 0) return 0;
  const int grid = grid_for(h->AT.nslices * 32);
  const int has_eq = h->m_eq > 0, has_ineq = h->m_ineq > 0;
  if (write_d)
    k_primal<true><<<grid, kBlock, 0, h->stream>>>(view(h->AT), h->y, h->c, h->T, h->lb, h->ub, h->x, h->xbar, h->dbuf,
                                                   h->n, h->m_eq, has_eq, has_ineq, h->theta, h->one_plus_theta);
  else
    k_primal<false><<<grid, kBlock, 0, h->stream>>>(view(h->AT), h->y, h->c, h->T, h->lb, h->ub, h->x, h->xbar, h->dbuf,
                                                    h->n, h->m_eq, has_eq, has_ineq, h->theta, h->one_plus_theta);
  return 0;
}

int launch_dual(cpppd_solver *h) {
  if (h->A.nslices == 0) return 0;
  k_dual<<<grid_for(h->A.nslices * 32), kBlock, 0, h->stream>>>(view(h->A), h->xbar, h->b, h->sigma, h->y, h->m, h->m_eq);
  return 0;
}

int get_graph(cpppd_solver *h, int64_t k, cudaGraphExec_t *out) {
  auto it = h->graphs.find(k);
  if (it != h->graphs.end()) {
    *out = it->second;
    return 0;
  }
  cudaGraph_t g = nullptr;
  CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  for (int64_t i = 0; i < k; ++i) {
    launch_primal(h, false);
    launch_dual(h);
  }
  CK(cudaStreamEndCapture(h->stream, &g));
  cudaGraphExec_t ge = nullptr;
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaGraphDestroy(g);
  h->graphs[k] = ge;
  *out = ge;
  return 0;
}

int run_iterations(cpppd_solver *h, int64_t k) {
  const bool use_graph = !(h->flags & CPPPD_FLAG_NO_GRAPH);
  while (k > 0) {
    int64_t step = std::min<int64_t>(k, kGraphChunk);
    if (use_graph && step >= 2) {
      cudaGraphExec_t ge = nullptr;
      if (int rc = get_graph(h, step, &ge)) return rc;
      CK(cudaGraphLaunch(ge, h->stream));
    } else {
      for (int64_t i = 0; i < step; ++i) {
        launch_primal(h, false);
        launch_dual(h);
      }
      CK(cudaGetLastError());
    }
    k -= step;
    h->niter += step;
  }
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int cpppd_abi_version(void) { return CPPPD_ABI_VERSION; }

const char *cpppd_last_error(cpppd_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int cpppd_create(const cpppd_problem *P, cpppd_handle *out) {
  cpppd_solver *h = nullptr;
  if (!P || !out) return fail(h, CPPPD_ERR_INVALID, "null argument");
  *out = nullptr;
  if (P->abi_version != CPPPD_ABI_VERSION)
    return fail(h, CPPPD_ERR_INVALID, "ABI version mismatch: caller %d, library %d", P->abi_version, CPPPD_ABI_VERSION);
  if (P->n < 0 || P->m_eq < 0 || P->m_ineq < 0 || P->nnz < 0) return fail(h, CPPPD_ERR_INVALID, "negative size");
  if (P->n >= (int64_t)1 << 31 || P->m_eq + P->m_ineq >= (int64_t)1 << 31)
    return fail(h, CPPPD_ERR_INVALID, "n and m must be below 2^31 (32-bit indices)");
  if (P->nnz >= (int64_t)1 << 32) return fail(h, CPPPD_ERR_INVALID, "nnz must be below 2^32");
  if (P->index_bits != 32) return fail(h, CPPPD_ERR_INVALID, "column indices must be int32 (narrow them on the host)");
  if (P->indptr_bits != 32 && P->indptr_bits != 64) return fail(h, CPPPD_ERR_INVALID, "indptr_bits must be 32 or 64");
  const int64_t m = P->m_eq + P->m_ineq;
  if (!P->indptr || (P->nnz && (!P->indices || !P->values)) || (P->n && (!P->c || !P->lb || !P->ub)) || (m && !P->b))
    return fail(h, CPPPD_ERR_INVALID, "null array pointer");
  {
    int64_t first = P->indptr_bits == 64 ? ((const int64_t *)P->indptr)[0] : ((const int32_t *)P->indptr)[0];
    int64_t last = P->indptr_bits == 64 ? ((const int64_t *)P->indptr)[m] : ((const int32_t *)P->indptr)[m];
    if (first != 0 || last != P->nnz) return fail(h, CPPPD_ERR_INVALID, "indptr[0] must be 0 and indptr[m] == nnz");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(h, CPPPD_ERR_NODEVICE, "no CUDA device available: this solver has no CPU path");
  }
  if (P->device < 0 || P->device >= ndev) return fail(h, CPPPD_ERR_INVALID, "device %d out of range (%d devices)", P->device, ndev);
  h = new cpppd_solver();
  h->device = P->device;
  h->n = P->n;
  h->m_eq = P->m_eq;
  h->m_ineq = P->m_ineq;
  h->m = m;
  h->nnz = P->nnz;
  h->alpha = P->alpha;
  h->theta = P->theta;
  h->one_plus_theta = P->one_plus_theta;
  h->flags = P->flags;
  h->alloc = P->alloc;
  h->free_fn = P->free;
  h->alloc_user = P->alloc_user;
  int rc = 0;
  do {
    if (cudaSetDevice(h->device) != cudaSuccess) {
      rc = fail(h, CPPPD_ERR_CUDA, "cudaSetDevice(%d) failed", h->device);
      break;
    }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    if (P->stream) {
      h->stream = (cudaStream_t)P->stream;
    } else {
      if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        rc = fail(h, CPPPD_ERR_CUDA, "cudaStreamCreate failed");
        break;
      }
      h->own_stream = true;
    }
    rc = setup(h, P);
  } while (0);
  if (rc) {
    g_create_error = h->err;
    cpppd_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

int cpppd_destroy(cpppd_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second);
  for (void *p : h->owned) dev_free(h, p);
  if (h->stats_host) cudaFreeHost(h->stats_host);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
  return 0;
}

int cpppd_iterate(cpppd_handle h, int64_t k) {
  CHECK_HANDLE(h);
  if (k < 0) return fail(h, CPPPD_ERR_INVALID, "negative iteration count");
  if (h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "cpppd_dual_step must close the open iteration first");
  return run_iterations(h, k);
}

int cpppd_primal_step(cpppd_handle h, int32_t keep_d) {
  CHECK_HANDLE(h);
  if (h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "primal step already issued for this iteration");
  launch_primal(h, keep_d != 0);
  h->mid_iteration = true;
  h->have_d = keep_d != 0;
  CK(cudaGetLastError());
  return 0;
}

int cpppd_stats_step(cpppd_handle h, int32_t force_integer) {
  CHECK_HANDLE(h);
  if (!h->mid_iteration || !h->have_d)
    return fail(h, CPPPD_ERR_STATE, "cpppd_stats_step needs a preceding cpppd_primal_step(keep_d=1)");
  const int has_eq = h->m_eq > 0, has_ineq = h->m_ineq > 0;
  // column pass turns dbuf into x4 in place; with force_integer it also materialises xr = rint(x)
  double *xr = h->x;
  if (force_integer) {
    if (!h->xr_scratch)
      if (int rc = alloc_array(h, &h->xr_scratch, h->n)) return rc;
    xr = h->xr_scratch;
  }
  k_stats_cols<<<h->stat_blocks_c, kBlock, 0, h->stream>>>(h->c, h->x, h->xbar, h->lb, h->ub, h->dbuf, xr, h->n,
                                                          force_integer, h->colpart);
  k_stats_rows<<<h->stat_blocks_r, kBlock, 0, h->stream>>>(view(h->A), h->x, h->dbuf, h->xbar, xr, h->b, h->y, h->m,
                                                          h->m_eq, force_integer, h->rowpart);
  k_stats_final<<<1, kBlock, 0, h->stream>>>(h->colpart, h->stat_blocks_c, h->rowpart, h->stat_blocks_r, h->n, has_eq,
                                             has_ineq, h->niter, h->stats_dev);
  k_snapshot_best<<<std::max(1, std::min(grid_for(h->n), h->sm_count * 8)), kBlock, 0, h->stream>>>(h->stats_dev, xr,
                                                                                                  h->best, h->n);
  CK(cudaMemcpyAsync(h->stats_host, h->stats_dev, sizeof(cpppd_stats), cudaMemcpyDeviceToHost, h->stream));
  h->stats_pending = true;
  h->have_d = false;  // dbuf now holds x4
  CK(cudaGetLastError());
  return 0;
}

int cpppd_dual_step(cpppd_handle h) {
  CHECK_HANDLE(h);
  if (!h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "no primal step open");
  launch_dual(h);
  h->mid_iteration = false;
  h->niter += 1;
  CK(cudaGetLastError());
  return 0;
}

int cpppd_sync(cpppd_handle h) {
  CHECK_HANDLE(h);
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int cpppd_read_stats(cpppd_handle h, cpppd_stats *out) {
  CHECK_HANDLE(h);
  if (!out) return fail(h, CPPPD_ERR_INVALID, "null output");
  if (!h->stats_pending) return fail(h, CPPPD_ERR_STATE, "no stats step has been issued");
  CK(cudaStreamSynchronize(h->stream));
  *out = *h->stats_host;
  return 0;
}

int cpppd_time_iterations(cpppd_handle h, int64_t k, float *elapsed_ms) {
  CHECK_HANDLE(h);
  if (!elapsed_ms || k < 0) return fail(h, CPPPD_ERR_INVALID, "bad argument");
  if (h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "cpppd_dual_step must close the open iteration first");
  CK(cudaEventRecord(h->ev0, h->stream));
  if (int rc = run_iterations(h, k)) return rc;
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
  return 0;
}

int cpppd_time_kernels(cpppd_handle h, int64_t k, float *primal_ms, float *dual_ms) {
  CHECK_HANDLE(h);
  if (!primal_ms || !dual_ms || k < 0 || k > 64) return fail(h, CPPPD_ERR_INVALID, "bad argument (k must be in [0, 64])");
  if (h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "cpppd_dual_step must close the open iteration first");
  std::vector<cudaEvent_t> ev(2 * k + 1);
  for (auto &e : ev) CK(cudaEventCreate(&e));
  CK(cudaEventRecord(ev[0], h->stream));
  for (int64_t i = 0; i < k; ++i) {
    launch_primal(h, false);
    CK(cudaEventRecord(ev[2 * i + 1], h->stream));
    launch_dual(h);
    CK(cudaEventRecord(ev[2 * i + 2], h->stream));
  }
  h->niter += k;
  CK(cudaEventSynchronize(ev[2 * k]));
  *primal_ms = *dual_ms = 0.f;
  for (int64_t i = 0; i < k; ++i) {
    float a = 0.f, b = 0.f;
    CK(cudaEventElapsedTime(&a, ev[2 * i], ev[2 * i + 1]));
    CK(cudaEventElapsedTime(&b, ev[2 * i + 1], ev[2 * i + 2]));
    *primal_ms += a;
    *dual_ms += b;
  }
  for (auto &e : ev) cudaEventDestroy(e);
  return 0;
}

static int vector_ptr(cpppd_solver *h, int32_t which, double **p, int64_t *count) {
  switch (which) {
    case CPPPD_VEC_X: *p = h->x; *count = h->n; return 0;
    case CPPPD_VEC_XBAR: *p = h->xbar; *count = h->n; return 0;
    case CPPPD_VEC_Y: *p = h->y; *count = h->m; return 0;
    case CPPPD_VEC_T: *p = h->T; *count = h->n; return 0;
    case CPPPD_VEC_SIGMA: *p = h->sigma; *count = h->m; return 0;
    case CPPPD_VEC_BEST_INTEGER: *p = h->best; *count = h->n; return 0;
    case CPPPD_VEC_D: *p = h->dbuf; *count = h->n; return 0;
    default: return fail(h, CPPPD_ERR_INVALID, "unknown vector id %d", which);
  }
}

int cpppd_get_vector(cpppd_handle h, int32_t which, double *host_dst) {
  CHECK_HANDLE(h);
  double *p = nullptr;
  int64_t count = 0;
  if (int rc = vector_ptr(h, which, &p, &count)) return rc;
  if (!host_dst) return fail(h, CPPPD_ERR_INVALID, "null destination");
  if (count) CK(cudaMemcpyAsync(host_dst, p, sizeof(double) * count, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int cpppd_set_vector(cpppd_handle h, int32_t which, const double *host_src) {
  CHECK_HANDLE(h);
  if (which != CPPPD_VEC_X && which != CPPPD_VEC_XBAR && which != CPPPD_VEC_Y)
    return fail(h, CPPPD_ERR_INVALID, "only x, xbar and y can be set");
  double *p = nullptr;
  int64_t count = 0;
  if (int rc = vector_ptr(h, which, &p, &count)) return rc;
  if (!host_src) return fail(h, CPPPD_ERR_INVALID, "null source");
  if (count) CK(cudaMemcpyAsync(p, host_src, sizeof(double) * count, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int cpppd_get_info(cpppd_handle h, cpppd_info *out) {
  CHECK_HANDLE(h);
  if (!out) return fail(h, CPPPD_ERR_INVALID, "null output");
  memset(out, 0, sizeof *out);
  out->n = h->n;
  out->m_eq = h->m_eq;
  out->m_ineq = h->m_ineq;
  out->nnz = h->nnz;
  out->a_padded_entries = h->A.padded;
  out->at_padded_entries = h->AT.padded;
  out->device_bytes = h->device_bytes;
  const int64_t P = h->nnz >= ((int64_t)1 << 31) ? 8 : 4;
  out->bytes_per_iteration_algorithmic = 2 * h->nnz * 12 + P * (h->m + 1) + P * (h->n + 1) + 8 * (8 * h->n + 5 * h->m);
  out->bytes_per_iteration_actual = (h->A.padded + h->AT.padded) * 12 + 8 * (h->A.nslices + h->AT.nslices + 2) +
                                    8 * (8 * h->n + 5 * h->m);
  out->value_bytes = 8;
  out->const_vector_mask = 0;
  out->sm_count = h->sm_count;
  out->world_size = 1;
  out->row_begin = 0;
  out->row_end = h->m;
  return 0;
}

int64_t cpppd_iteration_count(cpppd_handle h) { return h ? h->niter : -1; }

}  // extern "C"
