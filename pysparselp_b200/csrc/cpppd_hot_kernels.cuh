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
      while (ld_acquire_sys(cm->my_flags + cm->kind_in * cm->world + t) < want) __nanosleep(100);
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
// body of k_primal for one thread (column j of slice s)
template <bool kWriteD, bool kDict, int kChunk>
__device__ __forceinline__ void primal_rows(const SellView &AT, const double *__restrict__ y, const Vec &c, const Vec &T,
                                            const Vec &lb, const Vec &ub, double *__restrict__ x,
                                            double *__restrict__ xbar, double *__restrict__ d_out, int64_t n,
                                            int has_eq, int has_ineq, double theta, double one_plus_theta,
                                            const FusedComm *__restrict__ cm, const double *sdict, int64_t j, int64_t s) {
  const int lane = threadIdx.x & 31;
  int role = 0;
  if (cm) {
    role = cm->role[s];
    if (role) comm_wait(cm, lane);
  }
  int64_t p0, p1;
  slice_range(AT, s, p0, p1);
  const bool live = j < n;
  double cj = 0.0, tj = 0.0, xo = 0.0;
  if (live) {
    cj = c.at(j);
    tj = T.at(j);
    xo = __ldcs(x + j);
  }
  double s_eq = 0.0, s_in = 0.0;
  {
    const int32_t *ip = AT.idx + p0 + lane;
    const double *vp = AT.val + p0 + lane;
    const int width = (int)((p1 - p0) >> 5);
    const int32_t mask = AT.idx_mask;
    // entries are taken kChunk at a time: all index (and value) loads of a chunk are issued first,
    // then all gathers, then the sequential accumulation — so a row of any width (also 2 or 3) keeps
    // kChunk independent gathers in flight instead of one load-use chain per entry
#pragma unroll 1
    for (int k0 = 0; k0 < width; k0 += kChunk) {
      int32_t r[kChunk];
      double a[kChunk], g[kChunk];
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        const bool ok = k0 + u < width;
        r[u] = ok ? __ldcs(ip + (k0 + u) * kSlice) : kPad;
        a[u] = (!kDict && ok) ? __ldcs(vp + (k0 + u) * kSlice) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kChunk; ++u) g[u] = r[u] >= 0 ? __ldg(y + (r[u] & mask)) : 0.0;
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        if (r[u] >= 0) {
          const double av = kDict ? sdict[(r[u] >> AT.idx_bits) & AT.code_mask] : a[u];
          const double t = __dmul_rn(av, g[u]);
          if (r[u] & kEqBit) s_eq = __dadd_rn(s_eq, t); else s_in = __dadd_rn(s_in, t);
        }
      }
    }
  }
  if (!live) return;
  double d = cj;
  if (has_eq) d = __dadd_rn(d, s_eq);
  if (has_ineq) d = __dadd_rn(d, s_in);
  const double l = lb.at(j), u = ub.at(j);
  double x2 = __dsub_rn(xo, __dmul_rn(tj, d));
  x2 = (l > x2) ? l : x2;  // np.maximum(x2, lb)  (NaN in x2 propagates)
  x2 = (u < x2) ? u : x2;  // np.minimum(x2, ub)
  const double xb = __dsub_rn(__dmul_rn(one_plus_theta, x2), __dmul_rn(theta, xo));
  xbar[j] = xb;
  x[j] = x2;
  if (kWriteD) d_out[j] = d;
  if (role & 1) comm_push(cm, (int32_t)j, xb);
}

// Primal half-iteration (:198-228).  Thread j owns column j of A (row j of A^T).
// Loads that do not depend on the matrix (c, T, x) are issued first so that they are in flight
// together with the slice entries; matrix entries are read once (ld.global.cs).
// kDict: entries are single 32-bit words [pad][eq][code][index]; values come from a <= 256 entry
// dictionary staged in shared memory.
template <bool kWriteD, bool kDict, int kChunk>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
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
  if (s < AT.nslices) primal_rows<kWriteD, kDict, kChunk>(AT, y, c, T, lb, ub, x, xbar, d_out, n, has_eq, has_ineq, theta,
                                                          one_plus_theta, cm, sdict, j, s);
  if (cm) comm_finish(cm);  // every thread of the CTA gets here (no early return above)
}

// body of k_dual for one thread (row i of slice s)
template <bool kDict, int kChunk>
__device__ __forceinline__ void dual_rows(const SellView &A, const double *__restrict__ xbar, const Vec &b,
                                          const Vec &sigma, double *__restrict__ y, int64_t m, int64_t m_eq,
                                          const FusedComm *__restrict__ cm, const double *sdict, int64_t i, int64_t s) {
  const int lane = threadIdx.x & 31;
  int role = 0;
  if (cm) {
    role = cm->role[s];
    if (role) comm_wait(cm, lane);
  }
  int64_t p0, p1;
  slice_range(A, s, p0, p1);
  const bool live = i < m;
  double bi = 0.0, si = 0.0, yi = 0.0;
  if (live) {
    bi = b.at(i);
    si = sigma.at(i);
    yi = __ldcs(y + i);
  }
  double acc = 0.0;
  {
    const int32_t *ip = A.idx + p0 + lane;
    const double *vp = A.val + p0 + lane;
    const int width = (int)((p1 - p0) >> 5);
    const int32_t mask = A.idx_mask;
#pragma unroll 1
    for (int k0 = 0; k0 < width; k0 += kChunk) {  // see k_primal: loads of a chunk first, then gathers, then sums
      int32_t jc[kChunk];
      double a[kChunk], g[kChunk];
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        const bool ok = k0 + u < width;
        jc[u] = ok ? __ldcs(ip + (k0 + u) * kSlice) : kPad;
        a[u] = (!kDict && ok) ? __ldcs(vp + (k0 + u) * kSlice) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kChunk; ++u) g[u] = jc[u] >= 0 ? __ldg(xbar + (jc[u] & mask)) : 0.0;
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        if (jc[u] >= 0) {
          const double av = kDict ? sdict[(jc[u] >> A.idx_bits) & A.code_mask] : a[u];
          acc = __dadd_rn(acc, __dmul_rn(av, g[u]));
        }
      }
    }
  }
  if (!live) return;
  const double r = __dsub_rn(acc, bi);
  double yn = __dadd_rn(yi, __dmul_rn(si, r));
  if (i >= m_eq) yn = (yn < 0.0) ? 0.0 : yn;  // np.maximum(y_ineq, 0): NaN stays NaN, -0.0 stays
  y[i] = yn;
  if (role & 1) comm_push(cm, (int32_t)i, yn);
}

// Dual half-iteration (:231-240, :333-341).  Thread i owns row i of A.
template <bool kDict, int kChunk>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
k_dual(SellView A, const double *__restrict__ xbar, Vec b, Vec sigma, double *__restrict__ y, int64_t m,
       int64_t m_eq, const FusedComm *__restrict__ cm) {
  __shared__ double sdict[kDict ? 256 : 1];
  if (kDict) {
    if ((int)threadIdx.x < A.ndict) sdict[threadIdx.x] = A.dict[threadIdx.x];
    __syncthreads();
  }
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t s = i >> 5;
  if (s < A.nslices) dual_rows<kDict, kChunk>(A, xbar, b, sigma, y, m, m_eq, cm, sdict, i, s);
  if (cm) comm_finish(cm);
}

}  // namespace
