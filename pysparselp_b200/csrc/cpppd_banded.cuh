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
// (all entries that gather from window 0, then window 1, ...); inside a window the 32 rows of a warp keep their entries
// together, packed column-major in caller order (see k_band_fill), with one count byte per row and window.  A half-iteration is one launch PER WINDOW: thread i continues
// the sum of row i where the previous window left it (an fp64 carry in HBM, 16 bytes per row and window), so the
// working set of the gathers of one launch is one window, resident in L2, while entries, counts and carries
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

// Storage of one window: the 32 rows of a warp ("tile") keep their entries of this window together, entry position
// k of every row that has one before position k + 1 of any row, rows in lane order ("packed" column-major: no padding,
// and the lanes that are active at position k read consecutive addresses).  Per row and window one byte holds the
// number of entries (an operand with more than 255 entries of one row in one window is not banded), per tile and
// window one 32-bit offset.

// thread per row: entries per (window, row) into cnt[w * rows_pad + row]; flag[0] |= 1 when some row visits its
// windows out of order (the operand then cannot be banded without changing the summation order), |= 2 when a
// count does not fit a byte
__global__ void k_band_count(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices, int64_t nrows,
                             int64_t rows_pad, BandGeometry geo, unsigned char *__restrict__ cnt, int *__restrict__ flag) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const int64_t e0 = rowptr[row], e1 = rowptr[row + 1];
  int cur = -1, bad = 0;
  uint32_t run = 0;
  for (int64_t e = e0; e < e1; ++e) {
    const int w = geo.window_of(indices[e] & kIdxMask);
    if (w != cur) {
      if (cur >= 0) cnt[(int64_t)cur * rows_pad + row] = (unsigned char)run;
      if (w < cur) bad |= 1;
      cur = w;
      run = 0;
    }
    if (++run > 255u) bad |= 2;
  }
  if (cur >= 0) cnt[(int64_t)cur * rows_pad + row] = (unsigned char)run;
  if (bad) atomicOr(flag, bad);
}

// thread per (window, tile): entries of the tile in the window (cnt is laid out window-major, so cell c covers
// the 32 bytes cnt[32 c .. 32 c + 31])
__global__ void k_band_tile_totals(const unsigned char *__restrict__ cnt, int64_t cells, uint32_t *__restrict__ total) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  const uint32_t *w = reinterpret_cast<const uint32_t *>(cnt + 32 * c);
  uint32_t t = 0;
  for (int q = 0; q < 8; ++q) {
    const uint32_t v = w[q];
    t += (v & 0xffu) + ((v >> 8) & 0xffu) + ((v >> 16) & 0xffu) + (v >> 24);
  }
  total[c] = t;
}

__device__ __forceinline__ int warp_max_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// warp per tile: copy the entries to their packed places (tile_base = exclusive scan of the tile totals)
__global__ void k_band_fill(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices,
                            const double *__restrict__ values, int64_t nrows, int64_t rows_pad, int windows,
                            const unsigned char *__restrict__ cnt, const uint32_t *__restrict__ tile_base,
                            int32_t *__restrict__ idx, double *__restrict__ val) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows_pad) return;  // (whole warps: rows_pad is a multiple of 32)
  const int lane = threadIdx.x & 31;
  const int64_t tile = row >> 5, ntiles = rows_pad >> 5;
  const unsigned lt = (1u << lane) - 1u;
  int64_t e = row < nrows ? rowptr[row] : 0;
  for (int w = 0; w < windows; ++w) {
    const int c = cnt[(int64_t)w * rows_pad + row];
    uint32_t off = tile_base[(int64_t)w * ntiles + tile];
    const int widest = warp_max_i32(c);
    for (int k = 0; k < widest; ++k) {
      const unsigned mask = __ballot_sync(0xffffffffu, k < c);
      if (k < c) {
        const uint32_t pos = off + (uint32_t)__popc(mask & lt);
        idx[pos] = indices[e + k] & kIdxMask;
        val[pos] = values[e + k];
      }
      off += (uint32_t)__popc(mask);
    }
    e += c;
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

// acc + sum of this lane's `c` entries of the tile against vec, sequentially in stored order.  The warp walks the
// entry positions kC at a time: the ballots give every active lane its packed address, then all index / value loads
// of the chunk are issued, then the gathers, then the sequential additions.
template <int kC>
__device__ __forceinline__ double band_accumulate(const int32_t *__restrict__ idx, const double *__restrict__ val,
                                                  const double *__restrict__ vec, uint32_t off, int c, int lane, double acc) {
  const unsigned lt = (1u << lane) - 1u;
  const int widest = warp_max_i32(c);
#pragma unroll 1
  for (int k = 0; k < widest; k += kC) {
    int32_t j[kC];
    double a[kC], g[kC];
#pragma unroll
    for (int u = 0; u < kC; ++u) {
      const bool ok = k + u < c;
      const unsigned mask = __ballot_sync(0xffffffffu, ok);
      const uint32_t pos = off + (uint32_t)__popc(mask & lt);
      off += (uint32_t)__popc(mask);
      j[u] = ok ? __ldcs(idx + pos) : 0;
      a[u] = ok ? __ldcs(val + pos) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kC; ++u) g[u] = k + u < c ? __ldg(vec + j[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < kC; ++u)
      if (k + u < c) acc = __dadd_rn(acc, __dmul_rn(a[u], g[u]));
  }
  return acc;
}

constexpr int kBandChunk = 4;

// One window of the dual half-iteration (:231-240, :333-341).  Thread i owns row i of A; the warp owns a tile.
// kFirst: the sum starts from 0.0 (csr_matvec), otherwise from the carry of the previous window.
// kLast : fused dual step + projection, otherwise the partial sum goes to the carry.
template <bool kFirst, bool kLast>
__global__ void __launch_bounds__(kBlock, 8)
k_dual_band(const unsigned char *__restrict__ cnt, const uint32_t *__restrict__ tile_base, const int32_t *__restrict__ idx,
            const double *__restrict__ val, const double *__restrict__ xbar, double *__restrict__ carry, Vec b, Vec sigma,
            double *__restrict__ y, int64_t m, int64_t rows_pad, int64_t m_eq) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= rows_pad) return;  // (whole warps)
  const int lane = threadIdx.x & 31;
  const bool live = i < m;
  const int c = __ldcs(cnt + i);  // (zero for the padding rows of the last tile)
  const uint32_t off = __ldg(tile_base + (i >> 5));
  double acc = 0.0, bi = 0.0, si = 0.0, yi = 0.0;
  if (!kFirst && live) acc = __ldcs(carry + i);
  if (kLast && live) {
    bi = b.at(i);
    si = sigma.at(i);
    yi = __ldcs(y + i);
  }
  acc = band_accumulate<kBandChunk>(idx, val, xbar, off, c, lane, acc);
  if (!live) return;
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
k_primal_band(const unsigned char *__restrict__ cnt, const uint32_t *__restrict__ tile_base, const int32_t *__restrict__ idx,
              const double *__restrict__ val, const double *__restrict__ y, double *__restrict__ carry_eq,
              double *__restrict__ carry_in, int mode, Vec c, Vec T, Vec lb, Vec ub, double *__restrict__ x,
              double *__restrict__ xbar, double *__restrict__ d_out, int64_t n, int64_t rows_pad, int has_eq, int has_ineq,
              double theta, double one_plus_theta) {
  const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (j >= rows_pad) return;  // (whole warps)
  const int lane = threadIdx.x & 31;
  const bool live = j < n;
  const int cn = __ldcs(cnt + j);
  const uint32_t off = __ldg(tile_base + (j >> 5));
  double *carry = (mode & kBandEq) ? carry_eq : carry_in;
  double acc = 0.0, cj = 0.0, tj = 0.0, xo = 0.0, other = 0.0;
  if (!(mode & kBandStart) && live) acc = __ldcs(carry + j);
  if ((mode & kBandLast) && live) {
    cj = c.at(j);
    tj = T.at(j);
    xo = __ldcs(x + j);
    if (!(mode & kBandEq) && has_eq) other = __ldcs(carry_eq + j);
  }
  acc = band_accumulate<kBandChunk>(idx, val, y, off, cn, lane, acc);
  if (!live) return;
  if (!(mode & kBandLast)) {
    __stcs(carry + j, acc);
    return;
  }
  const double s_eq = (mode & kBandEq) ? acc : other, s_in = (mode & kBandEq) ? 0.0 : acc;
  double d = cj;
  if (has_eq) d = __dadd_rn(d, s_eq);
  if (has_ineq) d = __dadd_rn(d, s_in);
  const double l = lb.at(j), u = ub.at(j);
  double x2 = __dsub_rn(xo, __dmul_rn(tj, d));
  x2 = (l > x2) ? l : x2;
  x2 = (u < x2) ? u : x2;
  xbar[j] = __dsub_rn(__dmul_rn(one_plus_theta, x2), __dmul_rn(theta, xo));
  x[j] = x2;
  if (kWriteD) d_out[j] = d;
}

}  // namespace
