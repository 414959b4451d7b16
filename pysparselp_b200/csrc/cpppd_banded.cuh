// Banded operands: the hot path for patterns WITHOUT locality (randomLP.py:14-75 — BASELINE configs[3]).
// Part of libcpppd (single translation unit, included by cpppd.cu).
//
// Why.  k_dual / k_primal gather xbar / y with one 8-byte load per entry.  When the gathered vector is far larger
// than the 126 MB L2 and the column pattern is random (20 M x 40 M random LP: xbar 160 MB, y 320 MB), nearly every
// gather misses L2 and pulls 64-96 bytes from DRAM for 8 useful ones: ncu measured 32.4 GB (k_primal) and 18.2 GB
// (k_dual) of DRAM reads per launch against 5.0 / 5.3 GB algorithmic — the kernels sit AT the DRAM roofline moving
// 3.5-6.5x the bytes (profiles/r02_random_lp.md).  Shared-memory staging cannot help (a slice's column window is the
// whole vector); what helps is making the gathers hit L2: tools/probe/gather_probe.cu measured 4.7 ns per 1000 random
// gathers while the gathered window is <= 48-64 MB, 2.2x / 3.7x / 4.2x that at 128 / 320 / 512 MB.
//
// What.  The gathered vector is cut into windows of <= 48 MB; the entries of the operand are stored window-major
// (all entries that gather from window 0, then window 1, ...), row-minor inside a window, in caller order inside
// a row, with one CSR pointer array per window.  A half-iteration is one launch PER WINDOW: thread i continues
// the sum of row i where the previous window left it (an fp64 carry in HBM, 16 bytes per row and window), so the
// working set of the gathers of one launch is one window, resident in L2, while entries, pointers and carries
// stream past it with evict-first loads.  The last window runs the fused epilogue of k_dual / k_primal.
//
// Bit-exactness.  A row sum must be accumulated sequentially in the caller's entry order (scipy csr_matvec /
// csc_matvec).  Splitting a row by window keeps that order iff the window index is non-decreasing
// along the row — true for rows with ascending column indices (every CSR that scipy canonicalised, the random
// LP generator, all of A^T, whose columns are sorted by source row by construction).  build_band() verifies it on the
// device and the operand stays in the SELL kernels when it does not hold.  The carry is the exact fp64 partial
// sum, so x, xbar, y are the same bits as with every other kernel variant.
#pragma once

namespace {

// window of element r of the gathered vector.  A^T gathers y = [y_eq; y_ineq]: equality and inequality rows are
// summed apart (:206, :216), so no window straddles m_eq (`split`); A gathers xbar: split = 0, eq_windows = 0.
struct BandGeometry {
  int64_t split, eq_elems, in_elems;
  int eq_windows, windows;
  __host__ __device__ int window_of(int64_t r) const {
    return r < split ? (int)(r / eq_elems) : eq_windows + (int)((r - split) / in_elems);
  }
};

// thread per row: entries per (window, row) into cnt[w * (nrows + 1) + row]; flag[0] |= 1 when some row visits its
// windows out of order (the operand then cannot be banded without changing the summation order)
__global__ void k_band_count(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices, int64_t nrows,
                             BandGeometry geo, uint32_t *__restrict__ cnt, int *__restrict__ flag) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const int64_t e0 = rowptr[row], e1 = rowptr[row + 1];
  int cur = -1;
  uint32_t run = 0;
  for (int64_t e = e0; e < e1; ++e) {
    const int w = geo.window_of(indices[e] & kIdxMask);
    if (w != cur) {
      if (cur >= 0) cnt[(int64_t)cur * (nrows + 1) + row] = run;
      if (w < cur) *flag = 1;
      cur = w;
      run = 0;
    }
    ++run;
  }
  if (cur >= 0) cnt[(int64_t)cur * (nrows + 1) + row] = run;
}

// thread per row: copy the entries to their window-major places (ptr = exclusive scan of cnt)
__global__ void k_band_fill(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices,
                            const double *__restrict__ values, int64_t nrows, BandGeometry geo,
                            const uint32_t *__restrict__ ptr, int32_t *__restrict__ idx, double *__restrict__ val) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const int64_t e0 = rowptr[row], e1 = rowptr[row + 1];
  int cur = -1;
  uint32_t dst = 0;
  for (int64_t e = e0; e < e1; ++e) {
    const int32_t r = indices[e] & kIdxMask;
    const int w = geo.window_of(r);
    if (w != cur) {
      cur = w;
      dst = ptr[(int64_t)w * (nrows + 1) + row];
    }
    idx[dst] = r;
    val[dst] = values[e];
    ++dst;
  }
}

// Locality of the gathers of a thread-per-row kernel: for sampled warps (32 consecutive rows) and every entry
// position k, the number of distinct 32-byte sectors the 32 lanes touch.  out[0] += sectors, out[1] += entries.
// Potts (lanes walk consecutive pixels): about 0.27 sectors per entry; random pattern: 1.0.
__global__ void k_band_locality(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices, int64_t nrows,
                                int64_t warp_stride, unsigned long long *__restrict__ out) {
  __shared__ int32_t sect[kBlock];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * (kBlock / 32) + wib) * warp_stride;
  const int64_t row = warp * 32 + lane;
  int64_t e0 = 0, len = 0;
  if (row < nrows) {
    e0 = rowptr[row];
    len = rowptr[row + 1] - e0;
  }
  int64_t widest = len;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) widest = max(widest, __shfl_xor_sync(0xffffffffu, widest, o));
  if (widest > 64) widest = 64;
  unsigned long long sectors = 0, entries = 0;
  for (int64_t k = 0; k < widest; ++k) {
    const int32_t s = k < len ? ((indices[e0 + k] & kIdxMask) >> 2) : -1 - lane;
    sect[threadIdx.x] = s;
    __syncwarp();
    if (k < len) {
      bool first = true;
      for (int l = 0; l < lane; ++l)
        if (sect[wib * 32 + l] == s) first = false;
      sectors += first ? 1 : 0;
      entries += 1;
    }
    __syncwarp();
  }
  if (entries) {
    atomicAdd(out, sectors);
    atomicAdd(out + 1, entries);
  }
}

// acc + sum of the entries [p0, p1) against vec, sequentially in stored order, kC entries in flight
template <int kC>
__device__ __forceinline__ double band_accumulate(const int32_t *__restrict__ idx, const double *__restrict__ val,
                                                  const double *__restrict__ vec, uint32_t p0, uint32_t p1, double acc) {
#pragma unroll 1
  for (uint32_t p = p0; p < p1; p += kC) {
    int32_t j[kC];
    double a[kC], g[kC];
#pragma unroll
    for (int u = 0; u < kC; ++u) {
      const bool ok = p + u < p1;
      j[u] = ok ? __ldcs(idx + p + u) : 0;
      a[u] = ok ? __ldcs(val + p + u) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kC; ++u) g[u] = p + u < p1 ? __ldg(vec + j[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < kC; ++u)
      if (p + u < p1) acc = __dadd_rn(acc, __dmul_rn(a[u], g[u]));
  }
  return acc;
}

constexpr int kBandChunk = 4;

// One window of the dual half-iteration (:231-240, :333-341).  Thread i owns row i of A.
// kFirst: the sum starts from 0.0 (csr_matvec), otherwise from the carry of the previous window.
// kLast : fused dual step + projection, otherwise the partial sum goes to the carry.
template <bool kFirst, bool kLast>
__global__ void __launch_bounds__(kBlock, 8)
k_dual_band(const uint32_t *__restrict__ ptr, const int32_t *__restrict__ idx, const double *__restrict__ val,
            const double *__restrict__ xbar, double *__restrict__ carry, Vec b, Vec sigma, double *__restrict__ y,
            int64_t m, int64_t m_eq) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= m) return;
  const uint32_t p0 = __ldcs(ptr + i), p1 = __ldcs(ptr + i + 1);
  double acc = 0.0, bi = 0.0, si = 0.0, yi = 0.0;
  if (!kFirst) acc = __ldcs(carry + i);
  if (kLast) {
    bi = b.at(i);
    si = sigma.at(i);
    yi = __ldcs(y + i);
  }
  acc = band_accumulate<kBandChunk>(idx, val, xbar, p0, p1, acc);
  if (!kLast) {
    __stcs(carry + i, acc);
    return;
  }
  const double r = __dsub_rn(acc, bi);
  double yn = __dadd_rn(yi, __dmul_rn(si, r));
  if (i >= m_eq) yn = (yn < 0.0) ? 0.0 : yn;
  y[i] = yn;
}

// One window of the primal half-iteration (:198-228).  Thread j owns column j of A.
// mode bit 0: the sum of this window's kind (equality / inequality rows) starts from 0.0
//      bit 1: last window — fused primal step, clip, extrapolation
//      bit 2: the window gathers equality duals (its sum is s_eq, kept apart from s_ineq as in :206, :216)
constexpr int kBandStart = 1, kBandLast = 2, kBandEq = 4;
template <bool kWriteD>
__global__ void __launch_bounds__(kBlock, 6)
k_primal_band(const uint32_t *__restrict__ ptr, const int32_t *__restrict__ idx, const double *__restrict__ val,
              const double *__restrict__ y, double *__restrict__ carry_eq, double *__restrict__ carry_in, int mode,
              Vec c, Vec T, Vec lb, Vec ub, double *__restrict__ x, double *__restrict__ xbar,
              double *__restrict__ d_out, int64_t n, int has_eq, int has_ineq, double theta, double one_plus_theta) {
  const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const uint32_t p0 = __ldcs(ptr + j), p1 = __ldcs(ptr + j + 1);
  double *carry = (mode & kBandEq) ? carry_eq : carry_in;
  double acc = 0.0, cj = 0.0, tj = 0.0, xo = 0.0, l = 0.0, u = 0.0, other = 0.0;
  if (!(mode & kBandStart)) acc = __ldcs(carry + j);
  if (mode & kBandLast) {
    cj = c.at(j);
    tj = T.at(j);
    xo = __ldcs(x + j);
    if (!(mode & kBandEq) && has_eq) other = __ldcs(carry_eq + j);
  }
  acc = band_accumulate<kBandChunk>(idx, val, y, p0, p1, acc);
  if (!(mode & kBandLast)) {
    __stcs(carry + j, acc);
    return;
  }
  const double s_eq = (mode & kBandEq) ? acc : other, s_in = (mode & kBandEq) ? 0.0 : acc;
  double d = cj;
  if (has_eq) d = __dadd_rn(d, s_eq);
  if (has_ineq) d = __dadd_rn(d, s_in);
  l = lb.at(j);
  u = ub.at(j);
  double x2 = __dsub_rn(xo, __dmul_rn(tj, d));
  x2 = (l > x2) ? l : x2;
  x2 = (u < x2) ? u : x2;
  xbar[j] = __dsub_rn(__dmul_rn(one_plus_theta, x2), __dmul_rn(theta, xo));
  x[j] = x2;
  if (kWriteD) d_out[j] = d;
}

}  // namespace
