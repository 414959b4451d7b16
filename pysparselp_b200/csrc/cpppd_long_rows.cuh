// Long rows: rows of A (or columns of A, i.e. rows of A^T) with more entries than one thread should walk.
// Part of libcpppd (single translation unit, included by cpppd.cu).
//
// The hot kernels give one thread one row.  That is the right shape for the LPs of the reference's tests
// and benchmarks (Potts: 3 entries per row, 2-8 per column) but not for a skewed pattern: every weight
// column of the L1-SVM LP (reference pysparselp/examples/example_l1_svm.py:36-66) has ~4N/3 entries for N
// samples, and a single "budget" row can touch every variable.  One thread would walk such a row alone
// while its slice pads 31 other rows to the same length.
//
// A row longer than the threshold is therefore cut out of the SELL operand at setup.  It keeps `virt`
// VIRTUAL entries with value 1.0 whose gather index points behind the ghosts of the gathered vector (the
// "tail"): before the hot kernel runs, k_long_partial / k_long_finish compute the row's sum(s) with one CTA
// per 16384-entry segment and store them in the tail, and the unchanged hot kernel picks them up as
// 0 + 1.0 * sum.  A^T keeps two virtual entries per long column (equality part with kEqBit, inequality
// part), A keeps one.
//
// Accuracy: a long row is summed as a fixed tree (strided per thread, warp shuffles, warps in order,
// segments in order) — deterministic, independent of the GPU count, but not the sequential order of
// scipy's csr_matvec / csc_matvec.  Iterates of LPs WITH long rows agree with the reference to rounding
// (tests: 1e-9 relative after 100 iterations, BASELINE.json's bound) instead of bit for bit; LPs without
// long rows are untouched.  The "replaced by 1" masks of the preconditioners stay exact (a sum of
// non-negative terms is zero iff every term is).
#pragma once

namespace {

constexpr int kLongSeg = 16384;              // default entries per segment (one CTA: 192 KB of entries — with 4096 the three
                                             // dependent loads that locate a segment took longer than streaming it, ncu: 39 %
                                             // of HBM); CPPPD_LONG_SEG overrides (the segment length is part of the fixed tree)
constexpr int64_t kLongDefault = 2048;       // default threshold: rows with more entries are long

struct LongRows {
  int64_t count = 0, nseg = 0, nnz = 0;
  int virt = 1;                  // virtual entries per long row left in the SELL operand
  int64_t tail_base = 0;         // index of the first tail slot in the gathered vector
  int32_t *row = nullptr;        // [count]   local row id
  int64_t *ptr = nullptr;        // [count+1] entry offsets
  int64_t *seg_ptr = nullptr;    // [count+1] first segment of each long row
  int32_t *seg_row = nullptr;    // [nseg]    long row of a segment
  int32_t *seg_order = nullptr;  // [nseg]    launch order of the segments: by position inside their row, then by row
  int seg_len = kLongSeg;        // entries per segment
  int shape = 0;                 // compiled shape of k_long_partial (CPPPD_LONG_SHAPE)
  int32_t *idx = nullptr;        // entries in stored order (A^T: with kEqBit), plain values
  double *val = nullptr;
  double *partial = nullptr;     // [2 * nseg]
};

// flag[r] = 1 when row r has more than `threshold` entries (flag[nrows] = 0)
__global__ void k_long_flag(const int64_t *__restrict__ rowptr, int64_t nrows, int64_t threshold,
                            int32_t *__restrict__ flag) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nrows) return;
  flag[r] = (r < nrows && rowptr[r + 1] - rowptr[r] > threshold) ? 1 : 0;
}

// compact the long rows: id and length, in row order
__global__ void k_long_list(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ flag,
                            const int32_t *__restrict__ slot, int64_t nrows, int32_t *__restrict__ row_out,
                            int64_t *__restrict__ len_out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows || !flag[r]) return;
  row_out[slot[r]] = (int32_t)r;
  len_out[slot[r]] = rowptr[r + 1] - rowptr[r];
}

// length of every row once the long ones are reduced to their virtual entries (newlen[nrows] = 0)
__global__ void k_long_newlen(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ flag, int64_t nrows,
                              int virt, int64_t *__restrict__ newlen) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nrows) return;
  newlen[r] = r == nrows ? 0 : (flag[r] ? virt : rowptr[r + 1] - rowptr[r]);
}

// the CSR without the long rows' entries: short rows are copied, long rows get their virtual entries
__global__ void k_long_rewrite(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices,
                               const double *__restrict__ values, const int32_t *__restrict__ flag,
                               const int32_t *__restrict__ slot, const int64_t *__restrict__ new_rowptr, int64_t nrows,
                               int virt, int64_t tail_base, int32_t *__restrict__ out_idx,
                               double *__restrict__ out_val) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int64_t dst = new_rowptr[r];
  if (flag[r]) {
    for (int v = 0; v < virt; ++v) {
      // A^T (virt == 2): first the equality part, tagged like a real equality entry
      out_idx[dst + v] = (int32_t)(tail_base + (int64_t)virt * slot[r] + v) | ((virt == 2 && v == 0) ? kEqBit : 0);
      out_val[dst + v] = 1.0;
    }
    return;
  }
  const int64_t src = rowptr[r], len = rowptr[r + 1] - src;
  for (int64_t k = 0; k < len; ++k) {
    out_idx[dst + k] = indices[src + k];
    out_val[dst + k] = values[src + k];
  }
}

// one CTA per long row: copy its entries (stored order kept)
__global__ void __launch_bounds__(kBlock)
k_long_copy(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices, const double *__restrict__ values,
            const int32_t *__restrict__ row, const int64_t *__restrict__ ptr, int32_t *__restrict__ out_idx,
            double *__restrict__ out_val) {
  const int64_t src = rowptr[row[blockIdx.x]], dst = ptr[blockIdx.x], len = ptr[blockIdx.x + 1] - dst;
  for (int64_t k = threadIdx.x; k < len; k += kBlock) {
    out_idx[dst + k] = indices[src + k];
    out_val[dst + k] = values[src + k];
  }
}

// One CTA per segment: partial[2 s] / partial[2 s + 1] = equality / other part of sum_k a_k * vec[idx_k] over the
// segment (kGather false: sum_k |a_k|^power, the preconditioner sums).  Thread t takes entries t, t + 256, ...
// CTA b takes segment seg_order[b] (row-major by default; see split_long_rows).
template <int kLongUnroll, int kMinCtas, bool kGather>
__global__ void __launch_bounds__(kBlock, kMinCtas)
k_long_partial(const int64_t *__restrict__ ptr, const int64_t *__restrict__ seg_ptr, const int32_t *__restrict__ seg_row,
               const int32_t *__restrict__ seg_order, int seg_len, const int32_t *__restrict__ idx,
               const double *__restrict__ val, const double *__restrict__ vec, double power,
               double *__restrict__ partial) {
  const int64_t s = seg_order[blockIdx.x];
  const int32_t r = seg_row[s];
  const int64_t e0 = ptr[r] + (s - seg_ptr[r]) * seg_len;
  const int64_t e1 = min(e0 + (int64_t)seg_len, ptr[r + 1]);
  double v[2] = {0.0, 0.0};
  // kLongUnroll entries of the thread per trip, software-pipelined: the index / value loads of trip t + 1 are issued
  // before the gathers of trip t are consumed, so the entry stream never drains while a trip waits for its gathers.
  // The additions stay in the thread's entry order (the same fixed tree as a one-at-a-time loop).
  int32_t w[kLongUnroll], wn[kLongUnroll];
  double a[kLongUnroll], an[kLongUnroll], g[kLongUnroll];
  constexpr int64_t kTrip = (int64_t)kBlock * kLongUnroll;
  int64_t base = e0 + threadIdx.x;
#pragma unroll
  for (int u = 0; u < kLongUnroll; ++u) {
    const int64_t e = base + (int64_t)u * kBlock;
    wn[u] = e < e1 ? __ldcs(idx + e) : 0;
    an[u] = e < e1 ? __ldcs(val + e) : 0.0;
  }
#pragma unroll 1
  for (; base < e1; base += kTrip) {
#pragma unroll
    for (int u = 0; u < kLongUnroll; ++u) {
      w[u] = wn[u];
      a[u] = an[u];
    }
#pragma unroll
    for (int u = 0; u < kLongUnroll; ++u) g[u] = (kGather && base + (int64_t)u * kBlock < e1) ? __ldg(vec + (w[u] & kIdxMask)) : 0.0;
#pragma unroll
    for (int u = 0; u < kLongUnroll; ++u) {
      const int64_t e = base + kTrip + (int64_t)u * kBlock;
      wn[u] = e < e1 ? __ldcs(idx + e) : 0;
      an[u] = e < e1 ? __ldcs(val + e) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kLongUnroll; ++u) {
      if (base + (int64_t)u * kBlock >= e1) break;
      const double t = kGather ? __dmul_rn(a[u], g[u]) : abs_pow(a[u], power);
      if (w[u] & kEqBit) v[0] = __dadd_rn(v[0], t); else v[1] = __dadd_rn(v[1], t);
    }
  }
  block_reduce_write<2>(v, 0u, partial + 2 * s);
}

// compiled shapes of k_long_partial (same bits): entries in flight per thread x resident CTAs per SM
constexpr int kLongShapes = 3;
inline const char *long_shape_name(int shape) {
  return shape == 1 ? "pipe4/5cta" : shape == 2 ? "pipe8/3cta" : "pipe6/4cta";
}

enum { kLongSumsAT = 0, kLongSumsA = 1, kLongPrecondT = 2, kLongPrecondSigma = 3 };

// One warp per long row: fold its segments (lane-strided, then the shuffle tree) and store the result.
//   kLongSumsAT       out[2 r] = equality part, out[2 r + 1] = inequality part   (tail of y)
//   kLongSumsA        out[r] = sum                                               (tail of xbar / x / x4 / xr)
//   kLongPrecondT     out[row[r]] = 1 / ((0 + s_eq) + s_ineq), 0 -> 1            (:145-153)
//   kLongPrecondSigma out[row[r]] = 1 / sum, 0 -> 1                              (:162-164, :173-175)
__global__ void __launch_bounds__(kBlock)
k_long_finish(const int64_t *__restrict__ seg_ptr, const int32_t *__restrict__ row, int64_t count,
              const double *__restrict__ partial, int mode, int has_eq, int has_ineq, double *__restrict__ out) {
  const int64_t r = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= count) return;
  double s_eq = 0.0, s_in = 0.0;
  for (int64_t s = seg_ptr[r] + lane; s < seg_ptr[r + 1]; s += 32) {
    s_eq = __dadd_rn(s_eq, partial[2 * s]);
    s_in = __dadd_rn(s_in, partial[2 * s + 1]);
  }
  s_eq = warp_sum(s_eq);
  s_in = warp_sum(s_in);
  if (lane != 0) return;
  if (mode == kLongSumsAT) {
    out[2 * r] = s_eq;
    out[2 * r + 1] = s_in;
  } else if (mode == kLongSumsA) {
    out[r] = s_in;
  } else {
    double tmp = 0.0;
    if (mode == kLongPrecondSigma) {
      tmp = s_in;
    } else {
      if (has_eq) tmp = __dadd_rn(tmp, s_eq);
      if (has_ineq) tmp = __dadd_rn(tmp, s_in);
    }
    if (tmp == 0.0) tmp = 1.0;
    out[row[r]] = __ddiv_rn(1.0, tmp);
  }
}

}  // namespace
