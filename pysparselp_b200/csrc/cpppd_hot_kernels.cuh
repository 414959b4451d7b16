// The two hot kernels: k_primal (:198-228) and k_dual (:231-240, :333-341).
// Part of libcpppd (single translation unit, included by cpppd.cu).
#pragma once

namespace {

// ------------------------------------------------------------------------------------------
// halo exchange fused into the hot kernels (multi-GPU, peer memory over NVLink)
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#else  // host build of the sources for the CPU emulator (tests/emul): ranks are OS threads
inline unsigned long long ld_acquire_sys(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release_sys(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#endif

// Spin until *flag >= want.  A neighbour that never arrives (a crashed rank, a protocol error) must not
// hang the GPU: after st->timeout_ns the wait gives up, records it in st->timed_out (the host turns that
// into CPPPD_ERR_COMM at its next synchronisation) and the kernel runs on with whatever the ghosts hold.
__device__ __forceinline__ void wait_for_stamp(const unsigned long long *flag, unsigned long long want, SyncState *st,
                                               unsigned sleep_ns) {
  if (ld_acquire_sys(flag) >= want) return;
  const unsigned long long t0 = global_timer_ns(), limit = st->timeout_ns;
  while (ld_acquire_sys(flag) < want) {
    __nanosleep(sleep_ns);
    if (limit && global_timer_ns() - t0 > limit) {
      st->timed_out = 1u;
      break;
    }
  }
}

// Device-resident description of the halo a kernel produces (kind_out) and consumes (kind_in).
// role[s] of a slice: bit 0 = some row of the slice is sent to a neighbour, bit 1 = some row reads a
// ghost entry.  Warps of such slices first wait for the neighbours' stamps of the consumed halo
// (which also tells that the neighbours are done reading the ghosts this kernel is about to
// overwrite on their side), compute, and store sent values straight into the neighbours' ghost
// slots.  The last CTA of the grid publishes the new stamp to the neighbours.
struct FusedComm {
  const unsigned char *role;
  const int32_t *send_idx;            // local row ids, ascending inside each peer segment
  int64_t off[kMaxWorld], cnt[kMaxWorld], dst_base[kMaxWorld];
  double *peer_vec[kMaxWorld];        // neighbours' vectors (ghost tails are written)
  unsigned long long *peer_flags[kMaxWorld];
  const unsigned long long *my_flags; // 2 * world stamps written by the neighbours
  SyncState *st;
  int world, me, kind_out, kind_in;
  unsigned long long send_mask, recv_mask_in;
};

__device__ __forceinline__ void comm_wait(const FusedComm *cm, int lane) {
  // k_dual(k) consumes the xbar halo of k_primal(k) (stamp k+1); k_primal(k) consumes the y halo of
  // k_dual(k-1) (stamp k), so the very first primal kernel waits for nothing
  const unsigned long long want = cm->st->wait_stamp[cm->kind_in] + (cm->kind_in == 0 ? 1 : 0);
  for (int t = lane; t < cm->world; t += 32)
    if ((cm->recv_mask_in >> t) & 1ull)
      wait_for_stamp(cm->my_flags + cm->kind_in * cm->world + t, want, cm->st, 100);
  __syncwarp();
}

__device__ __forceinline__ void comm_push(const FusedComm *cm, int32_t row, double value) {
  for (int t = 0; t < cm->world; ++t) {
    if (!((cm->send_mask >> t) & 1ull)) continue;
    const int32_t *seg = cm->send_idx + cm->off[t];
    int64_t lo = 0, hi = cm->cnt[t] - 1;
    while (lo <= hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int32_t v = seg[mid];
      if (v == row) {
        cm->peer_vec[t][cm->dst_base[t] + mid] = value;
        break;
      }
      if (v < row) lo = mid + 1; else hi = mid - 1;
    }
  }
}

// every thread of the CTA calls this once at the end of the kernel
__device__ __forceinline__ void comm_finish(const FusedComm *cm) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x != 0) return;
  SyncState *st = cm->st;
  const unsigned int ticket = atomicAdd(&st->ticket[cm->kind_out], 1u);
  if (ticket != gridDim.x - 1) return;
  __threadfence_system();
  st->ticket[cm->kind_out] = 0;
  const unsigned long long stamp = st->push_stamp[cm->kind_out] + 1;
  st->push_stamp[cm->kind_out] = stamp;
  for (int t = 0; t < cm->world; ++t)
    if ((cm->send_mask >> t) & 1ull) st_release_sys(cm->peer_flags[t] + cm->kind_out * cm->world + cm->me, stamp);
  if (cm->recv_mask_in) st->wait_stamp[cm->kind_in] += 1;
}

// ------------------------------------------------------------------------------------------
// the two hot kernels
// ------------------------------------------------------------------------------------------
// Every variant of a kernel performs exactly the same floating point operations in exactly the same
// order (a row / column sum is accumulated sequentially in stored entry order), so the iterates do not
// depend on the variant.  What differs is how many independent loads a thread keeps in flight:
//   kChunk == 0 : the loop of the round-1 "v2" kernels measured on hardware (unroll 4 + scalar remainder);
//   kChunk >= 1 : entries are taken kChunk at a time — all index (and value) loads of a chunk are issued
//                 first, then all gathers, then the sequential accumulation — so a row of any width (also
//                 2 or 3, the Potts widths) keeps kChunk independent gathers in flight instead of one
//                 load-use chain per entry.
// kMinB is the CTAs/SM the variant is compiled for (register cap = 65536 / (256 * kMinB)).
// cpppd_create() times the variants on the actual operands and keeps the fastest (cpppd_host.cuh).

// How operands are loaded.
// LoadStream: one pass over operands far larger than the caches — matrix entries and vectors are read once
// (ld.global.cs, evict-first), gathers go through the read-only path (ld.global.nc): the gathered vector is not
// written by the kernel that gathers from it.
// LoadResident: the persistent kernel for tiny LPs (k_tiny_iterate) re-reads the whole LP every iteration from
// L1 / L2 and WRITES the vectors it gathers from between two barriers: ordinary coherent loads (ld.global.ca).
struct LoadStream {
  template <typename T> static __device__ __forceinline__ T entry(const T *p) { return __ldcs(p); }
  static __device__ __forceinline__ double gather(const double *p) { return __ldg(p); }
  static __device__ __forceinline__ double vec(const Vec &v, int64_t i) { return v.at(i); }
};
// LoadPeer: the kernels that carry the halo exchange themselves (kComm) gather ghost entries that a neighbour GPU
// stores over NVLink WHILE this kernel runs.  L1 is not coherent with those stores: a warp that needs no ghost (and
// therefore does not wait) may pull the 32-byte sector that holds the last owned entries AND the first ghosts into
// L1 before the neighbour's store lands, and a later ld.global.nc / .ca of the ghost on the same SM would hit that
// stale sector.  Gathers therefore go to L2, the point of coherence for peer stores (ld.global.cg).
struct LoadPeer {
  template <typename T> static __device__ __forceinline__ T entry(const T *p) { return __ldcs(p); }
  static __device__ __forceinline__ double gather(const double *p) { return __ldcg(p); }
  static __device__ __forceinline__ double vec(const Vec &v, int64_t i) { return v.at(i); }
};
struct LoadResident {
  template <typename T> static __device__ __forceinline__ T entry(const T *p) { return __ldca(p); }
  static __device__ __forceinline__ double gather(const double *p) { return __ldca(p); }
  static __device__ __forceinline__ double vec(const Vec &v, int64_t i) { return v.p ? __ldca(v.p + i) : v.c; }
};

// sum of column j of A times y, equality and inequality rows apart (:206, :216)
template <bool kDict, int kChunk, typename L = LoadStream>
__device__ __forceinline__ void primal_sums(const SellView &AT, const double *__restrict__ y, const double *sdict,
                                            int64_t p0, int64_t p1, int lane, double &s_eq, double &s_in) {
  const int32_t *ip = AT.idx + p0 + lane;
  const double *vp = AT.val + p0 + lane;
  const int width = (int)((p1 - p0) >> 5);
  const int32_t mask = AT.idx_mask;
  if (kChunk == 0) {
#pragma unroll 4
    for (int k = 0; k < width; ++k) {
      const int32_t r = L::entry(ip + k * kSlice);
      double a;
      if (kDict) a = sdict[(r >> AT.idx_bits) & AT.code_mask]; else a = L::entry(vp + k * kSlice);
      if (r >= 0) {
        const double t = __dmul_rn(a, L::gather(y + (r & mask)));
        if (r & kEqBit) s_eq = __dadd_rn(s_eq, t); else s_in = __dadd_rn(s_in, t);
      }
    }
  } else {
    constexpr int kC = kChunk > 0 ? kChunk : 1;
#pragma unroll 1
    for (int k0 = 0; k0 < width; k0 += kC) {
      int32_t r[kC];
      double a[kC], g[kC];
#pragma unroll
      for (int u = 0; u < kC; ++u) {
        const bool ok = k0 + u < width;
        r[u] = ok ? L::entry(ip + (k0 + u) * kSlice) : kPad;
        a[u] = (!kDict && ok) ? L::entry(vp + (k0 + u) * kSlice) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kC; ++u) g[u] = r[u] >= 0 ? L::gather(y + (r[u] & mask)) : 0.0;
#pragma unroll
      for (int u = 0; u < kC; ++u) {
        if (r[u] >= 0) {
          const double av = kDict ? sdict[(r[u] >> AT.idx_bits) & AT.code_mask] : a[u];
          const double t = __dmul_rn(av, g[u]);
          if (r[u] & kEqBit) s_eq = __dadd_rn(s_eq, t); else s_in = __dadd_rn(s_in, t);
        }
      }
    }
  }
}

// row i of A times xbar (:235, :240)
template <bool kDict, int kChunk, typename L = LoadStream>
__device__ __forceinline__ double dual_sum(const SellView &A, const double *__restrict__ xbar, const double *sdict,
                                           int64_t p0, int64_t p1, int lane) {
  const int32_t *ip = A.idx + p0 + lane;
  const double *vp = A.val + p0 + lane;
  const int width = (int)((p1 - p0) >> 5);
  const int32_t mask = A.idx_mask;
  double acc = 0.0;
  if (kChunk == 0) {
#pragma unroll 4
    for (int k = 0; k < width; ++k) {
      const int32_t jc = L::entry(ip + k * kSlice);
      double a;
      if (kDict) a = sdict[(jc >> A.idx_bits) & A.code_mask]; else a = L::entry(vp + k * kSlice);
      if (jc >= 0) acc = __dadd_rn(acc, __dmul_rn(a, L::gather(xbar + (jc & mask))));
    }
  } else {
    constexpr int kC = kChunk > 0 ? kChunk : 1;
#pragma unroll 1
    for (int k0 = 0; k0 < width; k0 += kC) {
      int32_t jc[kC];
      double a[kC], g[kC];
#pragma unroll
      for (int u = 0; u < kC; ++u) {
        const bool ok = k0 + u < width;
        jc[u] = ok ? L::entry(ip + (k0 + u) * kSlice) : kPad;
        a[u] = (!kDict && ok) ? L::entry(vp + (k0 + u) * kSlice) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kC; ++u) g[u] = jc[u] >= 0 ? L::gather(xbar + (jc[u] & mask)) : 0.0;
#pragma unroll
      for (int u = 0; u < kC; ++u) {
        if (jc[u] >= 0) {
          const double av = kDict ? sdict[(jc[u] >> A.idx_bits) & A.code_mask] : a[u];
          acc = __dadd_rn(acc, __dmul_rn(av, g[u]));
        }
      }
    }
  }
  return acc;
}

// ---- two rows per thread -------------------------------------------------------------------------------------
// A warp takes TWO consecutive slices (rows 64w + lane and 64w + 32 + lane) and walks them in lock step, kChunk
// entries of each at a time: twice the independent loads in flight per thread for the short rows of the Potts LP
// (3 entries per row, 2 per edge column), where one row per thread leaves a lane with a single short dependent
// chain.  Each row is still summed alone, sequentially, in stored order — same bits as every other variant.
template <bool kDict, int kChunk, typename L = LoadStream>
__device__ __forceinline__ void primal_sums2(const SellView &AT, const double *__restrict__ y, const double *sdict,
                                             int64_t pa0, int64_t pa1, int64_t pb0, int64_t pb1, int lane,
                                             double (&s_eq)[2], double (&s_in)[2]) {
  const int32_t *ip[2] = {AT.idx + pa0 + lane, AT.idx + pb0 + lane};
  const double *vp[2] = {AT.val + pa0 + lane, AT.val + pb0 + lane};
  const int width[2] = {(int)((pa1 - pa0) >> 5), (int)((pb1 - pb0) >> 5)};
  const int widest = width[0] > width[1] ? width[0] : width[1];
  const int32_t mask = AT.idx_mask;
  constexpr int kC = kChunk > 0 ? kChunk : 1;
#pragma unroll 1
  for (int k0 = 0; k0 < widest; k0 += kC) {
    int32_t r[2][kC];
    double a[2][kC], g[2][kC];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int u = 0; u < kC; ++u) {
        const bool ok = k0 + u < width[q];
        r[q][u] = ok ? L::entry(ip[q] + (k0 + u) * kSlice) : kPad;
        a[q][u] = (!kDict && ok) ? L::entry(vp[q] + (k0 + u) * kSlice) : 0.0;
      }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int u = 0; u < kC; ++u) g[q][u] = r[q][u] >= 0 ? L::gather(y + (r[q][u] & mask)) : 0.0;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int u = 0; u < kC; ++u) {
        if (r[q][u] >= 0) {
          const double av = kDict ? sdict[(r[q][u] >> AT.idx_bits) & AT.code_mask] : a[q][u];
          const double t = __dmul_rn(av, g[q][u]);
          if (r[q][u] & kEqBit) s_eq[q] = __dadd_rn(s_eq[q], t); else s_in[q] = __dadd_rn(s_in[q], t);
        }
      }
  }
}

template <bool kDict, int kChunk, typename L = LoadStream>
__device__ __forceinline__ void dual_sum2(const SellView &A, const double *__restrict__ xbar, const double *sdict,
                                          int64_t pa0, int64_t pa1, int64_t pb0, int64_t pb1, int lane,
                                          double (&acc)[2]) {
  const int32_t *ip[2] = {A.idx + pa0 + lane, A.idx + pb0 + lane};
  const double *vp[2] = {A.val + pa0 + lane, A.val + pb0 + lane};
  const int width[2] = {(int)((pa1 - pa0) >> 5), (int)((pb1 - pb0) >> 5)};
  const int widest = width[0] > width[1] ? width[0] : width[1];
  const int32_t mask = A.idx_mask;
  constexpr int kC = kChunk > 0 ? kChunk : 1;
#pragma unroll 1
  for (int k0 = 0; k0 < widest; k0 += kC) {
    int32_t jc[2][kC];
    double a[2][kC], g[2][kC];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int u = 0; u < kC; ++u) {
        const bool ok = k0 + u < width[q];
        jc[q][u] = ok ? L::entry(ip[q] + (k0 + u) * kSlice) : kPad;
        a[q][u] = (!kDict && ok) ? L::entry(vp[q] + (k0 + u) * kSlice) : 0.0;
      }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int u = 0; u < kC; ++u) g[q][u] = jc[q][u] >= 0 ? L::gather(xbar + (jc[q][u] & mask)) : 0.0;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int u = 0; u < kC; ++u) {
        if (jc[q][u] >= 0) {
          const double av = kDict ? sdict[(jc[q][u] >> A.idx_bits) & A.code_mask] : a[q][u];
          acc[q] = __dadd_rn(acc[q], __dmul_rn(av, g[q][u]));
        }
      }
  }
}

// k_primal for the two columns 64w + lane and 64w + 32 + lane of warp w
template <bool kWriteD, bool kDict, int kChunk, typename L = LoadStream>
__device__ __forceinline__ void primal_rows2(const SellView &AT, const double *__restrict__ y, const Vec &c, const Vec &T,
                                             const Vec &lb, const Vec &ub, double *__restrict__ x,
                                             double *__restrict__ xbar, double *__restrict__ d_out, int64_t n,
                                             int has_eq, int has_ineq, double theta, double one_plus_theta,
                                             const double *sdict, int64_t w) {
  const int lane = threadIdx.x & 31;
  const int64_t sa = 2 * w, sb = sa + 1;
  if (sa >= AT.nslices) return;
  int64_t p0[2] = {0, 0}, p1[2] = {0, 0};
  slice_range(AT, sa, p0[0], p1[0]);
  if (sb < AT.nslices) slice_range(AT, sb, p0[1], p1[1]);
  const int64_t j[2] = {sa * kSlice + lane, sb * kSlice + lane};
  const bool live[2] = {j[0] < n, j[1] < n};  // (a missing second slice starts at or behind n)
  double cj[2] = {0.0, 0.0}, tj[2] = {0.0, 0.0}, xo[2] = {0.0, 0.0};
#pragma unroll
  for (int q = 0; q < 2; ++q)
    if (live[q]) {
      cj[q] = L::vec(c, j[q]);
      tj[q] = L::vec(T, j[q]);
      xo[q] = L::entry(x + j[q]);
    }
  double s_eq[2] = {0.0, 0.0}, s_in[2] = {0.0, 0.0};
  primal_sums2<kDict, kChunk, L>(AT, y, sdict, p0[0], p1[0], p0[1], p1[1], lane, s_eq, s_in);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (!live[q]) continue;
    double d = cj[q];
    if (has_eq) d = __dadd_rn(d, s_eq[q]);
    if (has_ineq) d = __dadd_rn(d, s_in[q]);
    const double l = L::vec(lb, j[q]), u = L::vec(ub, j[q]);
    double x2 = __dsub_rn(xo[q], __dmul_rn(tj[q], d));
    x2 = (l > x2) ? l : x2;
    x2 = (u < x2) ? u : x2;
    xbar[j[q]] = __dsub_rn(__dmul_rn(one_plus_theta, x2), __dmul_rn(theta, xo[q]));
    x[j[q]] = x2;
    if (kWriteD) d_out[j[q]] = d;
  }
}

// k_dual for the two rows 64w + lane and 64w + 32 + lane of warp w
template <bool kDict, int kChunk, typename L = LoadStream>
__device__ __forceinline__ void dual_rows2(const SellView &A, const double *__restrict__ xbar, const Vec &b,
                                           const Vec &sigma, double *__restrict__ y, int64_t m, int64_t m_eq,
                                           const double *sdict, int64_t w) {
  const int lane = threadIdx.x & 31;
  const int64_t sa = 2 * w, sb = sa + 1;
  if (sa >= A.nslices) return;
  int64_t p0[2] = {0, 0}, p1[2] = {0, 0};
  slice_range(A, sa, p0[0], p1[0]);
  if (sb < A.nslices) slice_range(A, sb, p0[1], p1[1]);
  const int64_t i[2] = {sa * kSlice + lane, sb * kSlice + lane};
  const bool live[2] = {i[0] < m, i[1] < m};
  double bi[2] = {0.0, 0.0}, si[2] = {0.0, 0.0}, yi[2] = {0.0, 0.0};
#pragma unroll
  for (int q = 0; q < 2; ++q)
    if (live[q]) {
      bi[q] = L::vec(b, i[q]);
      si[q] = L::vec(sigma, i[q]);
      yi[q] = L::entry(y + i[q]);
    }
  double acc[2] = {0.0, 0.0};
  dual_sum2<kDict, kChunk, L>(A, xbar, sdict, p0[0], p1[0], p0[1], p1[1], lane, acc);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (!live[q]) continue;
    const double r = __dsub_rn(acc[q], bi[q]);
    double yn = __dadd_rn(yi[q], __dmul_rn(si[q], r));
    if (i[q] >= m_eq) yn = (yn < 0.0) ? 0.0 : yn;
    y[i[q]] = yn;
  }
}

// body of k_primal for one thread (column j of slice s)
template <bool kWriteD, bool kDict, int kChunk, bool kEarlyBounds, bool kComm, typename L = LoadStream>
__device__ __forceinline__ void primal_rows(const SellView &AT, const double *__restrict__ y, const Vec &c, const Vec &T,
                                            const Vec &lb, const Vec &ub, double *__restrict__ x,
                                            double *__restrict__ xbar, double *__restrict__ d_out, int64_t n,
                                            int has_eq, int has_ineq, double theta, double one_plus_theta,
                                            const FusedComm *__restrict__ cm, const double *sdict, int64_t j, int64_t s) {
  const int lane = threadIdx.x & 31;
  int role = 0;
  if (kComm) {
    role = cm->role[s];
    if (role) comm_wait(cm, lane);
  }
  int64_t p0, p1;
  slice_range(AT, s, p0, p1);
  const bool live = j < n;
  // loads that do not depend on the matrix are issued first: they are in flight together with the entries
  double cj = 0.0, tj = 0.0, xo = 0.0, l = 0.0, u = 0.0;
  if (live) {
    cj = L::vec(c, j);
    tj = L::vec(T, j);
    xo = L::entry(x + j);
    if (kEarlyBounds) {
      l = L::vec(lb, j);
      u = L::vec(ub, j);
    }
  }
  double s_eq = 0.0, s_in = 0.0;
  primal_sums<kDict, kChunk, L>(AT, y, sdict, p0, p1, lane, s_eq, s_in);
  if (!live) return;
  double d = cj;
  if (has_eq) d = __dadd_rn(d, s_eq);
  if (has_ineq) d = __dadd_rn(d, s_in);
  if (!kEarlyBounds) {
    l = L::vec(lb, j);
    u = L::vec(ub, j);
  }
  double x2 = __dsub_rn(xo, __dmul_rn(tj, d));
  x2 = (l > x2) ? l : x2;  // np.maximum(x2, lb)  (NaN in x2 propagates)
  x2 = (u < x2) ? u : x2;  // np.minimum(x2, ub)
  const double xb = __dsub_rn(__dmul_rn(one_plus_theta, x2), __dmul_rn(theta, xo));
  xbar[j] = xb;
  x[j] = x2;
  if (kWriteD) d_out[j] = d;
  if (kComm && (role & 1)) comm_push(cm, (int32_t)j, xb);
}

// Primal half-iteration (:198-228).  Thread j owns column j of A (row j of A^T); matrix entries are read
// once (ld.global.cs).
// kDict: entries are single 32-bit words [pad][eq][code][index]; values come from a <= 256 entry
// dictionary staged in shared memory.
// kComm: the halo exchange is done by the kernel itself (see FusedComm); cm is not touched otherwise.
// kPersist: the grid is a few CTAs per SM and every warp strides over the slices (no CTA churn: with two or three
// entries per row a CTA of the one-slice-per-warp grid lives for a microsecond or two)
template <bool kWriteD, bool kDict, int kChunk, int kMinB, bool kComm, int kRows = 1, bool kPersist = false>
__global__ void __launch_bounds__(kBlock, kMinB)
k_primal(SellView AT, const double *__restrict__ y, Vec c, Vec T, Vec lb, Vec ub, double *__restrict__ x,
         double *__restrict__ xbar, double *__restrict__ d_out, int64_t n, int has_eq, int has_ineq,
         double theta, double one_plus_theta, const FusedComm *__restrict__ cm) {
  __shared__ double sdict[kDict ? 256 : 1];
  if (kDict) {
    if ((int)threadIdx.x < AT.ndict) sdict[threadIdx.x] = AT.dict[threadIdx.x];
    __syncthreads();
  }
  const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t s = j >> 5;
  if constexpr (kRows == 2) {  // warp s takes slices 2s and 2s + 1 (the grid covers half as many warps)
    primal_rows2<kWriteD, kDict, kChunk>(AT, y, c, T, lb, ub, x, xbar, d_out, n, has_eq, has_ineq, theta, one_plus_theta,
                                         sdict, s);
  } else if constexpr (kPersist) {
    for (int64_t jj = j; (jj >> 5) < AT.nslices; jj += (int64_t)gridDim.x * kBlock)
      primal_rows<kWriteD, kDict, kChunk, false, false, LoadStream>(AT, y, c, T, lb, ub, x, xbar, d_out, n, has_eq, has_ineq, theta,
                                                                    one_plus_theta, cm, sdict, jj, jj >> 5);
  } else if (s < AT.nslices)
    primal_rows<kWriteD, kDict, kChunk, (kChunk > 0 && kMinB <= 6), kComm, typename std::conditional<kComm, LoadPeer, LoadStream>::type>(
        AT, y, c, T, lb, ub, x, xbar, d_out, n, has_eq, has_ineq, theta, one_plus_theta, cm, sdict, j, s);
  if (kComm) comm_finish(cm);  // every thread of the CTA gets here (no early return above)
}

// body of k_dual for one thread (row i of slice s)
template <bool kDict, int kChunk, bool kComm, typename L = LoadStream>
__device__ __forceinline__ void dual_rows(const SellView &A, const double *__restrict__ xbar, const Vec &b,
                                          const Vec &sigma, double *__restrict__ y, int64_t m, int64_t m_eq,
                                          const FusedComm *__restrict__ cm, const double *sdict, int64_t i, int64_t s) {
  const int lane = threadIdx.x & 31;
  int role = 0;
  if (kComm) {
    role = cm->role[s];
    if (role) comm_wait(cm, lane);
  }
  int64_t p0, p1;
  slice_range(A, s, p0, p1);
  const bool live = i < m;
  double bi = 0.0, si = 0.0, yi = 0.0;
  if (live) {
    bi = L::vec(b, i);
    si = L::vec(sigma, i);
    yi = L::entry(y + i);
  }
  const double acc = dual_sum<kDict, kChunk, L>(A, xbar, sdict, p0, p1, lane);
  if (!live) return;
  const double r = __dsub_rn(acc, bi);
  double yn = __dadd_rn(yi, __dmul_rn(si, r));
  if (i >= m_eq) yn = (yn < 0.0) ? 0.0 : yn;  // np.maximum(y_ineq, 0): NaN stays NaN, -0.0 stays
  y[i] = yn;
  if (kComm && (role & 1)) comm_push(cm, (int32_t)i, yn);
}

// Dual half-iteration (:231-240, :333-341).  Thread i owns row i of A.
template <bool kDict, int kChunk, int kMinB, bool kComm, int kRows = 1, bool kPersist = false>
__global__ void __launch_bounds__(kBlock, kMinB)
k_dual(SellView A, const double *__restrict__ xbar, Vec b, Vec sigma, double *__restrict__ y, int64_t m,
       int64_t m_eq, const FusedComm *__restrict__ cm) {
  __shared__ double sdict[kDict ? 256 : 1];
  if (kDict) {
    if ((int)threadIdx.x < A.ndict) sdict[threadIdx.x] = A.dict[threadIdx.x];
    __syncthreads();
  }
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t s = i >> 5;
  if constexpr (kRows == 2) {
    dual_rows2<kDict, kChunk>(A, xbar, b, sigma, y, m, m_eq, sdict, s);
  } else if constexpr (kPersist) {
    for (int64_t ii = i; (ii >> 5) < A.nslices; ii += (int64_t)gridDim.x * kBlock)
      dual_rows<kDict, kChunk, false, LoadStream>(A, xbar, b, sigma, y, m, m_eq, cm, sdict, ii, ii >> 5);
  } else if (s < A.nslices)
    dual_rows<kDict, kChunk, kComm, typename std::conditional<kComm, LoadPeer, LoadStream>::type>(A, xbar, b, sigma, y, m, m_eq, cm,
                                                                                               sdict, i, s);
  if (kComm) comm_finish(cm);
}

// Tiny LPs (SC105: 103 x 105; everything fits the L1 of one SM): a whole iteration is two kernels of a few
// microseconds of launch latency and a few hundred nanoseconds of work.  k_tiny_iterate runs `iters` complete
// iterations in ONE CTA — primal half, barrier, dual half, barrier — with the threads striding over the slices.
// Same per-row code as k_primal / k_dual (variant 1), so the iterates are the same bits; loads are ordinary
// coherent ones (LoadResident) because y and xbar are rewritten between the barriers.
constexpr int kTinyBlock = 1024;
template <bool kDict>
__global__ void __launch_bounds__(kTinyBlock, 1)
k_tiny_iterate(SellView AT, SellView A, Vec c, Vec T, Vec lb, Vec ub, Vec b, Vec sigma, double *x, double *xbar, double *y,
               int64_t n, int64_t m, int64_t m_eq, int has_eq, int has_ineq, double theta, double one_plus_theta,
               int64_t iters) {
  __shared__ double sdict[kDict ? 256 : 1];
  if (kDict) {  // (the CTA may have fewer threads than the dictionary has entries)
    for (int t = threadIdx.x; t < AT.ndict; t += blockDim.x) sdict[t] = AT.dict[t];
    __syncthreads();
  }
  const int64_t cols = AT.nslices * kSlice, rows = A.nslices * kSlice;
  for (int64_t it = 0; it < iters; ++it) {
    for (int64_t j = threadIdx.x; j < cols; j += blockDim.x)  // blockDim.x is a multiple of 32: lane == j % 32
      primal_rows<false, kDict, 0, false, false, LoadResident>(AT, y, c, T, lb, ub, x, xbar, nullptr, n, has_eq, has_ineq,
                                                               theta, one_plus_theta, nullptr, sdict, j, j >> 5);
    __syncthreads();
    for (int64_t i = threadIdx.x; i < rows; i += blockDim.x)
      dual_rows<kDict, 0, false, LoadResident>(A, xbar, b, sigma, y, m, m_eq, nullptr, sdict, i, i >> 5);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// variant table
// ------------------------------------------------------------------------------------------
struct HotVariant {
  int chunk, min_blocks, rows;  // rows: slices (rows per lane) a warp walks in lock step; the grid shrinks accordingly
  const char *name;
  int persist;                  // > 0: grid of sm_count * persist CTAs, warps stride over the slices
};
// index 0 is the variant every other one is measured against (and the one used without tuning)
constexpr int kNumVariants = 9;
constexpr HotVariant kVariants[kNumVariants] = {
    {0, 8, 1, "loop-unroll4/8cta", 0}, {2, 8, 1, "chunk2/8cta", 0},        {4, 6, 1, "chunk4/6cta", 0},
    {4, 4, 1, "chunk4/4cta", 0},       {8, 4, 1, "chunk8/4cta", 0},        {4, 3, 2, "rows2-chunk4/3cta", 0},
    {2, 4, 2, "rows2-chunk2/4cta", 0}, {2, 8, 1, "stride-chunk2/8cta", 8}, {4, 6, 1, "stride-chunk4/6cta", 6}};

using PrimalFn = void (*)(SellView, const double *, Vec, Vec, Vec, Vec, double *, double *, double *, int64_t, int, int,
                          double, double, const FusedComm *);
using DualFn = void (*)(SellView, const double *, Vec, Vec, double *, int64_t, int64_t, const FusedComm *);

template <bool kWriteD, bool kDict>
PrimalFn primal_variant(int v) {
  switch (v) {
    case 1: return k_primal<kWriteD, kDict, 2, 8, false>;
    case 2: return k_primal<kWriteD, kDict, 4, 6, false>;
    case 3: return k_primal<kWriteD, kDict, 4, 4, false>;
    case 4: return k_primal<kWriteD, kDict, 8, 4, false>;
    case 5: return k_primal<kWriteD, kDict, 4, 3, false, 2>;
    case 6: return k_primal<kWriteD, kDict, 2, 4, false, 2>;
    case 7: return k_primal<kWriteD, kDict, 2, 8, false, 1, true>;
    case 8: return k_primal<kWriteD, kDict, 4, 6, false, 1, true>;
    default: return k_primal<kWriteD, kDict, 0, 8, false>;
  }
}
template <bool kDict>
DualFn dual_variant(int v) {
  switch (v) {
    case 1: return k_dual<kDict, 2, 8, false>;
    case 2: return k_dual<kDict, 4, 6, false>;
    case 3: return k_dual<kDict, 4, 4, false>;
    case 4: return k_dual<kDict, 8, 4, false>;
    case 5: return k_dual<kDict, 4, 3, false, 2>;
    case 6: return k_dual<kDict, 2, 4, false, 2>;
    case 7: return k_dual<kDict, 2, 8, false, 1, true>;
    case 8: return k_dual<kDict, 4, 6, false, 1, true>;
    default: return k_dual<kDict, 0, 8, false>;
  }
}
inline PrimalFn primal_kernel(bool write_d, bool dict, int v) {
  if (write_d) return dict ? primal_variant<true, true>(v) : primal_variant<true, false>(v);
  return dict ? primal_variant<false, true>(v) : primal_variant<false, false>(v);
}
inline DualFn dual_kernel(bool dict, int v) { return dict ? dual_variant<true>(v) : dual_variant<false>(v); }
// kernels that carry the halo exchange themselves (CPPPD_FLAG_FUSED_HALO): variant 0 only
inline PrimalFn primal_kernel_fused(bool write_d, bool dict) {
  if (write_d) return dict ? k_primal<true, true, 0, 8, true> : k_primal<true, false, 0, 8, true>;
  return dict ? k_primal<false, true, 0, 8, true> : k_primal<false, false, 0, 8, true>;
}
inline DualFn dual_kernel_fused(bool dict) { return dict ? k_dual<true, 0, 8, true> : k_dual<false, 0, 8, true>; }

}  // namespace
