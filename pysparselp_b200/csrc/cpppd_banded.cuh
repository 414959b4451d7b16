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
// What.  The gathered vector is cut into windows of <= 56 MB; the entries of the operand are stored window-major
// (all entries that gather from window 0, then window 1, ...), in CSR order inside a window, with one count byte per
// row and window and one offset per tile of 128 rows and window.  A half-iteration is one launch PER WINDOW: thread i continues
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

// Storage of one window: plain CSR order (row-major, caller order inside a row), described by one count byte per
// row (an operand with more than 255 entries of one row in one window is not banded) and one 32-bit offset per
// "tile" of kBandTile = 128 consecutive rows — the rows one warp handles, four consecutive rows per lane.
constexpr int kBandTile = 128;
constexpr int kBandRowsPerLane = kBandTile / 32;

// thread per row: entries per (window, row) into cnt[w * rows_pad + row]; flag[0] |= 1 when some row visits its
// windows out of order (the operand then cannot be banded without changing the summation order), |= 2 when a
// count does not fit a byte
// `orig` (may be nullptr): original id of a local element of the gathered vector — with more than one GPU the windows
// are ranges of ORIGINAL ids (the order the caller's rows are sorted in), which the local layout keeps in a few
// contiguous pieces (own range, then the ghosts of every peer in original order; see setup()).
__global__ void k_band_count(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices, int64_t nrows,
                             int64_t rows_pad, BandGeometry geo, const int32_t *__restrict__ orig,
                             unsigned char *__restrict__ cnt, int *__restrict__ flag) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const int64_t e0 = rowptr[row], e1 = rowptr[row + 1];
  int cur = -1, bad = 0;
  uint32_t run = 0;
  for (int64_t e = e0; e < e1; ++e) {
    const int32_t loc = indices[e] & kIdxMask;
    const int w = geo.window_of(orig ? orig[loc] : loc);
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

__device__ __forceinline__ uint32_t sum_bytes(uint32_t v) {
  return (v & 0xffu) + ((v >> 8) & 0xffu) + ((v >> 16) & 0xffu) + (v >> 24);
}

// thread per (window, tile): entries of the tile in the window (cnt is laid out window-major and rows_pad is a
// multiple of kBandTile, so cell c covers the bytes cnt[128 c .. 128 c + 127])
__global__ void k_band_tile_totals(const unsigned char *__restrict__ cnt, int64_t cells, uint32_t *__restrict__ total) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  const uint32_t *w = reinterpret_cast<const uint32_t *>(cnt + (int64_t)kBandTile * c);
  uint32_t t = 0;
  for (int q = 0; q < kBandTile / 4; ++q) t += sum_bytes(w[q]);
  total[c] = t;
}

// thread per row: copy the entries of the row to their places (tile_base = exclusive scan of the tile totals; the
// rows of a tile before this one are summed from their count bytes)
__global__ void k_band_fill(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices,
                            const double *__restrict__ values, int64_t nrows, int64_t rows_pad, int windows,
                            const unsigned char *__restrict__ cnt, const uint32_t *__restrict__ tile_base,
                            int32_t *__restrict__ idx, double *__restrict__ val) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const int64_t tile = row / kBandTile, ntiles = rows_pad / kBandTile, first = tile * kBandTile;
  int64_t e = rowptr[row];
  for (int w = 0; w < windows; ++w) {
    const unsigned char *cw = cnt + (int64_t)w * rows_pad;
    const int c = cw[row];
    if (c == 0) continue;
    uint32_t dst = tile_base[(int64_t)w * ntiles + tile];
    for (int64_t r = first; r < row; ++r) dst += cw[r];
    for (int k = 0; k < c; ++k) {
      idx[dst + k] = indices[e + k] & kIdxMask;
      val[dst + k] = values[e + k];
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

// The window a launch gathers from starts cold: until its sectors have been touched once, the gathers miss L2 and
// run at DRAM random-access speed.  The first CTAs of every launch therefore ask the L2 for the whole window up front
// (one bulk prefetch of 32 KB per CTA: a 56 MB window is requested by the first 1792 CTAs, within the first waves).
struct BandPrefetch {
  const char *base;  // first byte of the window in the gathered vector (nullptr: no prefetch)
  int64_t bytes;
};
constexpr int64_t kBandPrefetchSlice = 32768;
__device__ __forceinline__ void band_prefetch(const BandPrefetch &pf) {
#ifdef __CUDACC__
  const int64_t off = (int64_t)blockIdx.x * kBandPrefetchSlice;
  if (pf.base && threadIdx.x == 0 && off < pf.bytes) {
    const int64_t left = pf.bytes - off;
    const uint32_t size = (uint32_t)((left < kBandPrefetchSlice ? left : kBandPrefetchSlice) & ~(int64_t)15);
    if (size) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf.base + off), "r"(size) : "memory");
  }
#endif
}

// ---- the hot kernels ----------------------------------------------------------------------------------------
// A warp owns a tile of 128 rows, lane l the rows 4 l .. 4 l + 3.  The entries of the tile in this window are one
// contiguous run: the warp reads them FLAT, 128 at a time (lane l takes entries q + l, q + 32 + l, ...: every load
// instruction is one contiguous 128 B / 256 B segment and every lane has four independent index loads, then four
// independent gathers in flight, whatever the row lengths), multiplies value and gathered element, and parks the
// products in shared memory.  Then every lane adds the products of its own rows to their sums — sequentially, in
// stored order, exactly the additions of csr_matvec / csc_matvec.  (CSR-stream, without giving up the order.)
// kChunk = flat entries per lane and trip (32 kChunk per warp); the kernels are compiled in a few (kChunk, CTAs per SM)
// shapes — more entries in flight per warp against more warps per SM — and cpppd_create() times them (kBandShapes).
constexpr int kBandWarps = kBlock / 32;

// Per-row vectors: a lane owns four consecutive elements (32-byte aligned: its first row / column id is a multiple of
// 4), moved as two 16-byte accesses; elements behind `limit` (last tile of an operand only) are skipped.
__device__ __forceinline__ double2 band_load2(const double *__restrict__ p, int64_t first, int64_t limit) {
  if (first + 2 <= limit) return __ldcs(reinterpret_cast<const double2 *>(p + first));
  return make_double2(first < limit ? __ldcs(p + first) : 0.0, 0.0);
}
__device__ __forceinline__ double2 band_load2(const Vec &vec, int64_t first, int64_t limit) {
  return vec.p ? band_load2(vec.p, first, limit) : make_double2(vec.c, vec.c);
}
template <bool kStream>
__device__ __forceinline__ void band_store2(double *__restrict__ p, int64_t first, int64_t limit, double2 v) {
  if (first + 2 <= limit) {
    if (kStream) __stcs(reinterpret_cast<double2 *>(p + first), v); else *reinterpret_cast<double2 *>(p + first) = v;
  } else if (first < limit) {
    p[first] = v.x;
  }
}
__device__ __forceinline__ void band_load4(const double *__restrict__ p, int64_t first, int64_t limit, double (&v)[4]) {
  const double2 lo = band_load2(p, first, limit), hi = band_load2(p, first + 2, limit);
  v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
}
__device__ __forceinline__ void band_store4(double *__restrict__ p, int64_t first, int64_t limit, const double (&v)[4]) {
  band_store2<true>(p, first, limit, make_double2(v[0], v[1]));
  band_store2<true>(p, first + 2, limit, make_double2(v[2], v[3]));
}

struct BandRows {          // what a lane knows about its four rows in this window
  uint32_t counts;         // one byte per row
  uint32_t begin;          // tile-relative offset of the first entry of the first row
  uint32_t base, total;    // first entry of the tile (into idx / val), entries of the tile
};

__device__ __forceinline__ BandRows band_rows(const unsigned char *__restrict__ cnt, const uint32_t *__restrict__ tile_base,
                                              int64_t tile, int lane) {
  BandRows R;
  R.counts = __ldcs(reinterpret_cast<const uint32_t *>(cnt + tile * kBandTile) + lane);
  R.base = __ldg(tile_base + tile);
  R.total = __ldg(tile_base + tile + 1) - R.base;  // (tile_base holds one more element than there are tiles x windows)
  const uint32_t mine = sum_bytes(R.counts);
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  R.begin = incl - mine;
  return R;
}

// acc[r] += products of row r, for the four rows of this lane
template <int kChunk>
__device__ __forceinline__ void band_accumulate(const BandRows &R, const int32_t *__restrict__ idx,
                                                const double *__restrict__ val, const double *__restrict__ vec,
                                                double *__restrict__ prod, int lane, double (&acc)[kBandRowsPerLane]) {
  constexpr int kBandSpan = 32 * kChunk;
  const int32_t *ip = idx + R.base;
  const double *vp = val + R.base;
#pragma unroll 1
  for (uint32_t q = 0; q < R.total; q += kBandSpan) {
    int32_t j[kChunk];
    double a[kChunk], g[kChunk];
#pragma unroll
    for (int u = 0; u < kChunk; ++u) {
      const uint32_t e = q + u * 32 + lane;
      const bool ok = e < R.total;
      j[u] = ok ? __ldcs(ip + e) : 0;
      a[u] = ok ? __ldcs(vp + e) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kChunk; ++u) g[u] = q + u * 32 + lane < R.total ? __ldg(vec + j[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < kChunk; ++u) prod[u * 32 + lane] = __dmul_rn(a[u], g[u]);
    __syncwarp();
    uint32_t first = R.begin;
#pragma unroll
    for (int r = 0; r < kBandRowsPerLane; ++r) {
      const uint32_t last = first + ((R.counts >> (8 * r)) & 0xffu);
      const uint32_t lo = first > q ? first : q, hi = last < q + kBandSpan ? last : q + kBandSpan;
      for (uint32_t e = lo; e < hi; ++e) acc[r] = __dadd_rn(acc[r], prod[e - q]);
      first = last;
    }
    __syncwarp();
  }
}

// One window of the dual half-iteration (:231-240, :333-341).
// kFirst: the sums start from 0.0 (csr_matvec), otherwise from the carries of the previous window.
// kLast : fused dual step + projection, otherwise the partial sums go to the carry.
template <bool kFirst, bool kLast, int kChunk, int kMinB>
__global__ void __launch_bounds__(kBlock, kMinB)
k_dual_band(const unsigned char *__restrict__ cnt, const uint32_t *__restrict__ tile_base, const int32_t *__restrict__ idx,
            const double *__restrict__ val, const double *__restrict__ xbar, double *__restrict__ carry, Vec b, Vec sigma,
            double *__restrict__ y, int64_t m, int64_t ntiles, int64_t m_eq, BandPrefetch pf) {
  band_prefetch(pf);
  __shared__ double prod_all[kBandWarps][32 * kChunk];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t tile = (int64_t)blockIdx.x * kBandWarps + wib;
  if (tile >= ntiles) return;  // (whole warps)
  const int64_t i0 = tile * kBandTile + lane * kBandRowsPerLane;
  const BandRows R = band_rows(cnt, tile_base, tile, lane);
  double acc[kBandRowsPerLane] = {0.0, 0.0, 0.0, 0.0};
  if (!kFirst) band_load4(carry, i0, m, acc);
  band_accumulate<kChunk>(R, idx, val, xbar, prod_all[wib], lane, acc);
  if (!kLast) {
    band_store4(carry, i0, m, acc);
    return;
  }
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {  // two rows at a time
    const int64_t i = i0 + 2 * h2;
    const double2 bi = band_load2(b, i, m), si = band_load2(sigma, i, m), yi = band_load2(y, i, m);
    double2 yn;
    yn.x = __dadd_rn(yi.x, __dmul_rn(si.x, __dsub_rn(acc[2 * h2], bi.x)));
    yn.y = __dadd_rn(yi.y, __dmul_rn(si.y, __dsub_rn(acc[2 * h2 + 1], bi.y)));
    if (i >= m_eq) yn.x = (yn.x < 0.0) ? 0.0 : yn.x;
    if (i + 1 >= m_eq) yn.y = (yn.y < 0.0) ? 0.0 : yn.y;
    band_store2<false>(y, i, m, yn);
  }
}

// One window of the primal half-iteration (:198-228).
// mode bit 0: the sums of this window's kind (equality / inequality rows) start from 0.0
//      bit 2: the window gathers equality duals (its sums are s_eq, kept apart from s_ineq as in :206, :216)
// kLast     : last window — fused primal step, clip, extrapolation (kWriteD: d is kept for the stats block)
constexpr int kBandStart = 1, kBandEq = 4;
template <bool kLast, bool kWriteD, int kChunk, int kMinB>
__global__ void __launch_bounds__(kBlock, kMinB)
k_primal_band(const unsigned char *__restrict__ cnt, const uint32_t *__restrict__ tile_base, const int32_t *__restrict__ idx,
              const double *__restrict__ val, const double *__restrict__ y, double *__restrict__ carry_eq,
              double *__restrict__ carry_in, int mode, Vec c, Vec T, Vec lb, Vec ub, double *__restrict__ x,
              double *__restrict__ xbar, double *__restrict__ d_out, int64_t n, int64_t ntiles, int has_eq, int has_ineq,
              double theta, double one_plus_theta, BandPrefetch pf) {
  band_prefetch(pf);
  __shared__ double prod_all[kBandWarps][32 * kChunk];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t tile = (int64_t)blockIdx.x * kBandWarps + wib;
  if (tile >= ntiles) return;  // (whole warps)
  const int64_t j0 = tile * kBandTile + lane * kBandRowsPerLane;
  const BandRows R = band_rows(cnt, tile_base, tile, lane);
  double *carry = (mode & kBandEq) ? carry_eq : carry_in;
  double acc[kBandRowsPerLane] = {0.0, 0.0, 0.0, 0.0};
  if (!(mode & kBandStart)) band_load4(carry, j0, n, acc);
  band_accumulate<kChunk>(R, idx, val, y, prod_all[wib], lane, acc);
  if (!kLast) {
    band_store4(carry, j0, n, acc);
    return;
  }
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {  // two columns at a time
    const int64_t j = j0 + 2 * h2;
    double2 other = make_double2(0.0, 0.0);
    if (!(mode & kBandEq) && has_eq) other = band_load2(carry_eq, j, n);
    const double2 xo = band_load2(x, j, n), cj = band_load2(c, j, n), tj = band_load2(T, j, n);
    const double2 lo = band_load2(lb, j, n), up = band_load2(ub, j, n);
    double2 x2, xb, dd;
    {
      const double s = acc[2 * h2];
      double d = cj.x;
      if (has_eq) d = __dadd_rn(d, (mode & kBandEq) ? s : other.x);
      if (has_ineq) d = __dadd_rn(d, (mode & kBandEq) ? 0.0 : s);
      double v = __dsub_rn(xo.x, __dmul_rn(tj.x, d));
      v = (lo.x > v) ? lo.x : v;
      v = (up.x < v) ? up.x : v;
      x2.x = v;
      xb.x = __dsub_rn(__dmul_rn(one_plus_theta, v), __dmul_rn(theta, xo.x));
      dd.x = d;
    }
    {
      const double s = acc[2 * h2 + 1];
      double d = cj.y;
      if (has_eq) d = __dadd_rn(d, (mode & kBandEq) ? s : other.y);
      if (has_ineq) d = __dadd_rn(d, (mode & kBandEq) ? 0.0 : s);
      double v = __dsub_rn(xo.y, __dmul_rn(tj.y, d));
      v = (lo.y > v) ? lo.y : v;
      v = (up.y < v) ? up.y : v;
      x2.y = v;
      xb.y = __dsub_rn(__dmul_rn(one_plus_theta, v), __dmul_rn(theta, xo.y));
      dd.y = d;
    }
    band_store2<false>(xbar, j, n, xb);
    band_store2<false>(x, j, n, x2);
    if (kWriteD) band_store2<false>(d_out, j, n, dd);
  }
}

// ---- bulk-async staged variant (TMA 1-D bulk copy + mbarrier) ------------------------------------------------------
// The register-path kernels above pay two dependent memory round trips per 64-128 entries of a tile (index / value
// loads, then the gathers) and walk a tile's ~300 entries in three or four such trips: ncu shows them bound by
// latency, not by any unit (profiles/r02_random_lp.md).  Here one elected lane asks the copy engine for the tile's
// whole run of indices and values (two cp.async.bulk of up to kBandPiece entries into shared memory, completion on
// an mbarrier); no registers hold entries in flight, so the lanes can then keep kG gathers each (up to 384 per warp)
// in flight at once, multiply in place in shared memory and finish with the same sequential per-row sums.
constexpr int kBandPiece = 384;  // entries staged per bulk copy pair (a tile holds ~256-340 with the default windows)

struct alignas(16) BandStage {
  double val[kBandPiece + 2];    // (+ alignment slack: bulk copies move whole 16-byte units)
  int32_t idx[kBandPiece + 4];
  unsigned long long bar;        // mbarrier
  unsigned long long pad;
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t band_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void band_bar_init(unsigned long long *bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(band_smem_addr(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// orders the lanes' generic-proxy accesses to the stage before the copy engine's next writes to it
__device__ __forceinline__ void band_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void band_bar_expect(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(band_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void band_bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(band_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(band_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void band_bar_wait(unsigned long long *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(band_smem_addr(bar)),
      "r"(parity)
      : "memory");
}
#else  // CPU emulation (tests/emul): the copy is synchronous, the wait is the warp rendezvous that makes it visible
inline void band_bar_init(unsigned long long *) {}
inline void band_fence_async() {}
inline void band_bar_expect(unsigned long long *, uint32_t) {}
inline void band_bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *) { memcpy(dst, src, bytes); }
inline void band_bar_wait(unsigned long long *, uint32_t) { __syncwarp(); }
#endif

// acc[r] += products of row r for the four rows of this lane, entries staged through `st` by bulk copies
template <int kG>
__device__ __forceinline__ void band_accumulate_staged(const BandRows &R, const int32_t *__restrict__ idx,
                                                       const double *__restrict__ val, const double *__restrict__ vec,
                                                       BandStage *st, int lane, double (&acc)[kBandRowsPerLane]) {
  uint32_t parity = 0;
#pragma unroll 1
  for (uint32_t p0 = 0; p0 < R.total; p0 += kBandPiece) {
    const uint32_t count = R.total - p0 < (uint32_t)kBandPiece ? R.total - p0 : (uint32_t)kBandPiece;
    const uint32_t first = R.base + p0;
    // whole 16-byte units around [first, first + count): 4 indices / 2 values per unit
    const uint32_t i_lo = first & ~3u, i_hi = (first + count + 3u) & ~3u, v_lo = first & ~1u, v_hi = (first + count + 1u) & ~1u;
    const uint32_t i_off = first - i_lo, v_off = first - v_lo;
    if (lane == 0) {
      band_fence_async();
      band_bar_expect(&st->bar, (i_hi - i_lo) * 4u + (v_hi - v_lo) * 8u);
      band_bulk_load(st->idx, idx + i_lo, (i_hi - i_lo) * 4u, &st->bar);
      band_bulk_load(st->val, val + v_lo, (v_hi - v_lo) * 8u, &st->bar);
    }
    band_bar_wait(&st->bar, parity);
    parity ^= 1u;
#pragma unroll 1
    for (uint32_t q = 0; q < count; q += 32 * kG) {
      double g[kG];
#pragma unroll
      for (int u = 0; u < kG; ++u) {
        const uint32_t e = q + u * 32 + lane;
        g[u] = e < count ? __ldg(vec + st->idx[i_off + e]) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kG; ++u) {
        const uint32_t e = q + u * 32 + lane;
        if (e < count) st->val[v_off + e] = __dmul_rn(st->val[v_off + e], g[u]);
      }
    }
    __syncwarp();
    uint32_t begin = R.begin;
#pragma unroll
    for (int r = 0; r < kBandRowsPerLane; ++r) {
      const uint32_t end = begin + ((R.counts >> (8 * r)) & 0xffu);
      const uint32_t lo = begin > p0 ? begin : p0, hi = end < p0 + count ? end : p0 + count;
      for (uint32_t e = lo; e < hi; ++e) acc[r] = __dadd_rn(acc[r], st->val[v_off + e - p0]);
      begin = end;
    }
    __syncwarp();
  }
}

template <bool kFirst, int kG, int kMinB>
__global__ void __launch_bounds__(kBlock, kMinB)
k_dual_band_staged(const unsigned char *__restrict__ cnt, const uint32_t *__restrict__ tile_base,
                   const int32_t *__restrict__ idx, const double *__restrict__ val, const double *__restrict__ xbar,
                   double *__restrict__ carry, int64_t m, int64_t ntiles, BandPrefetch pf) {
  band_prefetch(pf);
  __shared__ BandStage stage[kBandWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t tile = (int64_t)blockIdx.x * kBandWarps + wib;
  if (tile >= ntiles) return;  // (whole warps)
  if (lane == 0) band_bar_init(&stage[wib].bar);
  __syncwarp();
  const int64_t i0 = tile * kBandTile + lane * kBandRowsPerLane;
  const BandRows R = band_rows(cnt, tile_base, tile, lane);
  double acc[kBandRowsPerLane] = {0.0, 0.0, 0.0, 0.0};
  if (!kFirst) band_load4(carry, i0, m, acc);
  band_accumulate_staged<kG>(R, idx, val, xbar, &stage[wib], lane, acc);
  band_store4(carry, i0, m, acc);
}

template <int kG, int kMinB>
__global__ void __launch_bounds__(kBlock, kMinB)
k_primal_band_staged(const unsigned char *__restrict__ cnt, const uint32_t *__restrict__ tile_base,
                     const int32_t *__restrict__ idx, const double *__restrict__ val, const double *__restrict__ y,
                     double *__restrict__ carry, int start, int64_t n, int64_t ntiles, BandPrefetch pf) {
  band_prefetch(pf);
  __shared__ BandStage stage[kBandWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t tile = (int64_t)blockIdx.x * kBandWarps + wib;
  if (tile >= ntiles) return;  // (whole warps)
  if (lane == 0) band_bar_init(&stage[wib].bar);
  __syncwarp();
  const int64_t j0 = tile * kBandTile + lane * kBandRowsPerLane;
  const BandRows R = band_rows(cnt, tile_base, tile, lane);
  double acc[kBandRowsPerLane] = {0.0, 0.0, 0.0, 0.0};
  if (!start) band_load4(carry, j0, n, acc);
  band_accumulate_staged<kG>(R, idx, val, y, &stage[wib], lane, acc);
  band_store4(carry, j0, n, acc);
}

// ---- compiled shapes of the window kernels ----------------------------------------------------------------------
// The windows before the last one only stream entries, counts and carries: lean kernels, many warps per SM.  The last
// window also runs the fused epilogue (up to eight more vectors): it keeps the 64-register shape.
struct BandShape {
  int chunk, min_blocks, staged;
  const char *name;
};
constexpr int kNumBandShapes = 8;
constexpr BandShape kBandShapes[kNumBandShapes] = {
    {4, 4, 0, "flat4/4cta"},    {3, 6, 0, "flat3/6cta"},     {2, 6, 0, "flat2/6cta"},     {2, 8, 0, "flat2/8cta"},
    {8, 5, 1, "bulk-g8/5cta"}, {10, 5, 1, "bulk-g10/5cta"}, {12, 4, 1, "bulk-g12/4cta"}, {6, 5, 1, "bulk-g6/5cta"}};

using DualBandFn = void (*)(const unsigned char *, const uint32_t *, const int32_t *, const double *, const double *, double *,
                            Vec, Vec, double *, int64_t, int64_t, int64_t, BandPrefetch);
using PrimalBandFn = void (*)(const unsigned char *, const uint32_t *, const int32_t *, const double *, const double *,
                              double *, double *, int, Vec, Vec, Vec, Vec, double *, double *, double *, int64_t, int64_t, int,
                              int, double, double, BandPrefetch);
using DualStagedFn = void (*)(const unsigned char *, const uint32_t *, const int32_t *, const double *, const double *, double *,
                              int64_t, int64_t, BandPrefetch);
using PrimalStagedFn = void (*)(const unsigned char *, const uint32_t *, const int32_t *, const double *, const double *,
                                double *, int, int64_t, int64_t, BandPrefetch);

template <bool kFirst, bool kLast>
DualBandFn dual_band_shape(int shape) {
  switch (shape) {
    case 1: return k_dual_band<kFirst, kLast, 3, 6>;
    case 2: return k_dual_band<kFirst, kLast, 2, 6>;
    case 3: return k_dual_band<kFirst, kLast, 2, 8>;
    default: return k_dual_band<kFirst, kLast, 4, 4>;
  }
}
// kernels of the windows before the last one; the last window always runs the register path with the fused epilogue
inline DualBandFn dual_band_kernel(bool first, bool last, int shape) {
  if (last) return first ? k_dual_band<true, true, 3, 5> : k_dual_band<false, true, 3, 5>;
  return first ? dual_band_shape<true, false>(shape) : dual_band_shape<false, false>(shape);
}
inline PrimalBandFn primal_band_kernel(bool last, bool write_d, int shape) {
  if (last) return write_d ? k_primal_band<true, true, 4, 4> : k_primal_band<true, false, 4, 4>;
  switch (shape) {
    case 1: return k_primal_band<false, false, 3, 6>;
    case 2: return k_primal_band<false, false, 2, 6>;
    case 3: return k_primal_band<false, false, 2, 8>;
    default: return k_primal_band<false, false, 4, 4>;
  }
}
template <bool kFirst>
DualStagedFn dual_staged_shape(int shape) {
  switch (shape) {
    case 5: return k_dual_band_staged<kFirst, 10, 5>;
    case 6: return k_dual_band_staged<kFirst, 12, 4>;
    case 7: return k_dual_band_staged<kFirst, 6, 5>;
    default: return k_dual_band_staged<kFirst, 8, 5>;
  }
}
inline DualStagedFn dual_staged_kernel(bool first, int shape) {
  return first ? dual_staged_shape<true>(shape) : dual_staged_shape<false>(shape);
}
inline PrimalStagedFn primal_staged_kernel(int shape) {
  switch (shape) {
    case 5: return k_primal_band_staged<10, 5>;
    case 6: return k_primal_band_staged<12, 4>;
    case 7: return k_primal_band_staged<6, 5>;
    default: return k_primal_band_staged<8, 5>;
  }
}

}  // namespace
