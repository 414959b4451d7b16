// Small LPs that exceed one SM: all k iterations in ONE launch of ONE thread-block cluster (<= 16 CTAs).
// Part of libcpppd (single translation unit, included by cpppd.cu after cpppd_hot_kernels.cuh).
//
// The Potts 50x50 LP of the reference's regression test (tests/test_pott_segmentation.py: 7 400 x 9 800, 29 400
// entries; BASELINE configs[0]) is bound by launch latency on the CUDA-graph path — two graph nodes of ~3.5 us around
// well under a microsecond of work — and too large for the one-CTA kernel (k_tiny_iterate).  Here every CTA of the
// cluster owns a contiguous range of slices of A^T (columns) and of A (rows) and keeps, for the whole launch, in its
// own shared memory: its entries of both operands, c / T / lb / ub / x / xbar of its columns, b / sigma / y of its
// rows.  A gather of y[i] or xbar[j] is a load from the OWNER's shared memory (distributed shared memory: the index
// words are translated once, at staging, into [owner CTA][offset]); the two halves of an iteration are separated by
// the hardware cluster barrier (barrier.cluster arrive.release / wait.acquire) instead of a kernel boundary.
// Nothing touches L2 / HBM between staging and the final write-back of x, xbar, y.
//
// Arithmetic: per row / column the same operations in the same order as k_primal / k_dual variant 0 (sequential sum
// in stored entry order, equality / inequality parts apart, the same epilogues) — the iterates are the same bits;
// tests/test_gpu_parity.py compares them with the goldens minted from the reference.
//
// Not compiled for the CPU emulation (tests/emul runs one CTA at a time): there the LP stays on the graph path.
#pragma once

#ifdef __CUDACC__
#include <cooperative_groups.h>
#endif

namespace {

constexpr int kClusterBlock = 1024;
constexpr int kClusterMaxCtas = 16;
constexpr int64_t kClusterMaxEntries = 65536;  // padded entries per operand: beyond, a half-iteration is no longer
                                               // latency-bound and the streaming kernels are the better shape
constexpr int32_t kClOffMask = 0x03ffffff;     // translated entry word: [pad:1][eq:1][owner:4][offset:26]
constexpr int kClOwnerShift = 26;

// layout of a CTA's dynamic shared memory (identical in every CTA: a remote address is the local one, mapped)
struct ClusterSmem {
  double *val_at, *val_a, *c, *T, *lb, *ub, *x, *xbar, *b, *sigma, *y;
  int32_t *w_at, *w_a, *sp_at, *sp_a;
};

inline __host__ __device__ size_t cluster_smem_bytes(int spc_at, int spc_a, int ent_at, int ent_a) {
  const size_t cols = (size_t)spc_at * 32, rows = (size_t)spc_a * 32;
  return 8 * ((size_t)ent_at + ent_a + 6 * cols + 3 * rows) + 4 * ((size_t)ent_at + ent_a + spc_at + spc_a + 2) + 16;
}

#ifdef __CUDACC__
namespace cg = cooperative_groups;

__device__ __forceinline__ ClusterSmem cluster_carve(unsigned char *base, int spc_at, int spc_a, int ent_at, int ent_a) {
  ClusterSmem S;
  const size_t cols = (size_t)spc_at * 32, rows = (size_t)spc_a * 32;
  double *d = reinterpret_cast<double *>(base);
  S.val_at = d; d += ent_at;
  S.val_a = d; d += ent_a;
  S.c = d; d += cols;
  S.T = d; d += cols;
  S.lb = d; d += cols;
  S.ub = d; d += cols;
  S.x = d; d += cols;
  S.xbar = d; d += cols;
  S.b = d; d += rows;
  S.sigma = d; d += rows;
  S.y = d; d += rows;
  int32_t *w = reinterpret_cast<int32_t *>(d);
  S.w_at = w; w += ent_at;
  S.w_a = w; w += ent_a;
  S.sp_at = w; w += spc_at + 1;
  S.sp_a = w;
  return S;
}

// entries of the slices [s_lo, s_hi) of S into shared memory: value, and the index word translated to
// [pad][eq][owner CTA of the gathered element][offset inside the owner's vector]
__device__ __forceinline__ void cluster_stage(const SellView &S, int64_t s_lo, int64_t s_hi, int spc_other, double *val,
                                              int32_t *w, int32_t *sp) {
  int64_t e_lo, e_hi, tmp;
  slice_range(S, s_lo, e_lo, tmp);
  if (s_hi > s_lo) slice_range(S, s_hi - 1, tmp, e_hi); else e_hi = e_lo;
  for (int64_t s = s_lo + threadIdx.x; s <= s_hi; s += blockDim.x) {
    int64_t p0, p1;
    if (s < s_hi) slice_range(S, s, p0, p1); else p0 = e_hi;
    sp[s - s_lo] = (int32_t)(p0 - e_lo);
  }
  const int32_t per_owner = spc_other * 32;
  for (int64_t e = e_lo + threadIdx.x; e < e_hi; e += blockDim.x) {
    const int32_t word = S.idx[e];
    int32_t out = kPad;
    double a = 0.0;
    if (word >= 0) {
      a = entry_value(S, e, word);
      const int32_t i = word & S.idx_mask;
      const int32_t owner = i / per_owner;
      out = (word & kEqBit) | (owner << kClOwnerShift) | (i - owner * per_owner);
    }
    val[e - e_lo] = a;
    w[e - e_lo] = out;
  }
}

// The barrier between the two halves of an iteration.  What has to be ordered is: my st.shared of xbar / y, then the
// other CTAs' loads of it from my shared memory after the barrier.  barrier.cluster.arrive.release compiles to
// MEMBAR.ALL.GPU + UCGABAR_ARV (cuobjdump): a GPU-scope fence, twice per iteration, for data that never leaves shared
// memory — measured 3.3 us per iteration of the Potts 50x50 LP.  kStrict == 0 uses membar.cta + the RELAXED arrive
// instead: shared memory has one physical copy (no cache in front of it), so once the store is performed at CTA
// scope a later load through the cluster network reads it; the wait keeps its acquire form, the loads after it are
// issued in program order.  kStrict == 1 (CPPPD_CLUSTER_STRICT=1) keeps release / acquire.
template <int kStrict>
__device__ __forceinline__ void cluster_barrier() {
  if (kStrict) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  } else {
    __threadfence_block();
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n" ::: "memory");
  }
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// element `word & kClOffMask` of the owner's copy of `vec`: own shared memory when this CTA is the owner
__device__ __forceinline__ double cluster_gather(cg::cluster_group &cluster, double *vec, int me, int32_t word) {
  const int owner = (word >> kClOwnerShift) & 15;
  const int off = word & kClOffMask;
  return owner == me ? vec[off] : cluster.map_shared_rank(vec, owner)[off];
}

template <int kStrict>
__global__ void __launch_bounds__(kClusterBlock, 1)
k_cluster_iterate(SellView AT, SellView A, Vec c, Vec T, Vec lb, Vec ub, Vec b, Vec sigma, double *x, double *xbar, double *y,
                  int64_t n, int64_t m, int64_t m_eq, int has_eq, int has_ineq, double theta, double one_plus_theta,
                  int64_t iters, int spc_at, int spc_a, int ent_at, int ent_a) {
  extern __shared__ __align__(16) unsigned char cluster_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int me = (int)cluster.block_rank();
  const ClusterSmem S = cluster_carve(cluster_raw, spc_at, spc_a, ent_at, ent_a);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // my slices
  const int64_t c_lo = min((int64_t)me * spc_at, AT.nslices), c_hi = min(c_lo + spc_at, AT.nslices);
  const int64_t r_lo = min((int64_t)me * spc_a, A.nslices), r_hi = min(r_lo + spc_a, A.nslices);
  const int ncs = (int)(c_hi - c_lo), nrs = (int)(r_hi - r_lo);
  // ---- staging
  cluster_stage(AT, c_lo, c_hi, spc_a, S.val_at, S.w_at, S.sp_at);   // A^T gathers y: owners by row slices
  cluster_stage(A, r_lo, r_hi, spc_at, S.val_a, S.w_a, S.sp_a);      // A gathers xbar: owners by column slices
  for (int t = threadIdx.x; t < spc_at * 32; t += blockDim.x) {
    const int64_t j = c_lo * 32 + t;
    const bool live = t < ncs * 32 && j < n;
    S.c[t] = live ? (c.p ? c.p[j] : c.c) : 0.0;
    S.T[t] = live ? (T.p ? T.p[j] : T.c) : 0.0;
    S.lb[t] = live ? (lb.p ? lb.p[j] : lb.c) : 0.0;
    S.ub[t] = live ? (ub.p ? ub.p[j] : ub.c) : 0.0;
    S.x[t] = live ? x[j] : 0.0;
    S.xbar[t] = live ? xbar[j] : 0.0;
  }
  for (int t = threadIdx.x; t < spc_a * 32; t += blockDim.x) {
    const int64_t i = r_lo * 32 + t;
    const bool live = t < nrs * 32 && i < m;
    S.b[t] = live ? (b.p ? b.p[i] : b.c) : 0.0;
    S.sigma[t] = live ? (sigma.p ? sigma.p[i] : sigma.c) : 0.0;
    S.y[t] = live ? y[i] : 0.0;
  }
  cluster.sync();
  // gathers in flight per thread: a whole pixel column of the Potts LP (8 entries) / a whole row (3) in one round trip
  constexpr int kCP = 8, kCD = 4;
  for (int64_t it = 0; it < iters; ++it) {
    // ---- primal half (:198-228): column sums of A against y (remote shared memory), fused epilogue
    for (int sl = warp; sl < ncs; sl += nwarps) {
      const int q0 = S.sp_at[sl] + lane, width = (S.sp_at[sl + 1] - S.sp_at[sl]) >> 5;
      double s_eq = 0.0, s_in = 0.0;
#pragma unroll 1
      for (int k0 = 0; k0 < width; k0 += kCP) {
        int32_t w[kCP];
        double g[kCP];
#pragma unroll
        for (int u = 0; u < kCP; ++u) w[u] = k0 + u < width ? S.w_at[q0 + (k0 + u) * 32] : kPad;
#pragma unroll
        for (int u = 0; u < kCP; ++u) g[u] = w[u] >= 0 ? cluster_gather(cluster, S.y, me, w[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < kCP; ++u)
          if (w[u] >= 0) {
            const double t = __dmul_rn(S.val_at[q0 + (k0 + u) * 32], g[u]);
            if (w[u] & kEqBit) s_eq = __dadd_rn(s_eq, t); else s_in = __dadd_rn(s_in, t);
          }
      }
      const int jl = sl * 32 + lane;
      if (c_lo * 32 + jl < n) {
        const double xo = S.x[jl];
        double d = S.c[jl];
        if (has_eq) d = __dadd_rn(d, s_eq);
        if (has_ineq) d = __dadd_rn(d, s_in);
        const double l = S.lb[jl], u = S.ub[jl];
        double x2 = __dsub_rn(xo, __dmul_rn(S.T[jl], d));
        x2 = (l > x2) ? l : x2;
        x2 = (u < x2) ? u : x2;
        S.xbar[jl] = __dsub_rn(__dmul_rn(one_plus_theta, x2), __dmul_rn(theta, xo));
        S.x[jl] = x2;
      }
    }
    cluster_barrier<kStrict>();
    // ---- dual half (:231-240, :333-341): row sums of A against xbar, fused epilogue
    for (int sl = warp; sl < nrs; sl += nwarps) {
      const int q0 = S.sp_a[sl] + lane, width = (S.sp_a[sl + 1] - S.sp_a[sl]) >> 5;
      double acc = 0.0;
#pragma unroll 1
      for (int k0 = 0; k0 < width; k0 += kCD) {
        int32_t w[kCD];
        double a[kCD], g[kCD];
#pragma unroll
        for (int u = 0; u < kCD; ++u) w[u] = k0 + u < width ? S.w_a[q0 + (k0 + u) * 32] : kPad;
#pragma unroll
        for (int u = 0; u < kCD; ++u) g[u] = w[u] >= 0 ? cluster_gather(cluster, S.xbar, me, w[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < kCD; ++u) a[u] = k0 + u < width ? S.val_a[q0 + (k0 + u) * 32] : 0.0;
#pragma unroll
        for (int u = 0; u < kCD; ++u)
          if (w[u] >= 0) acc = __dadd_rn(acc, __dmul_rn(a[u], g[u]));
      }
      const int il = sl * 32 + lane;
      const int64_t i = r_lo * 32 + il;
      if (i < m) {
        const double r = __dsub_rn(acc, S.b[il]);
        double yn = __dadd_rn(S.y[il], __dmul_rn(S.sigma[il], r));
        if (i >= m_eq) yn = (yn < 0.0) ? 0.0 : yn;
        S.y[il] = yn;
      }
    }
    cluster_barrier<kStrict>();
  }
  // ---- write-back
  for (int t = threadIdx.x; t < ncs * 32; t += blockDim.x) {
    const int64_t j = c_lo * 32 + t;
    if (j < n) {
      x[j] = S.x[t];
      xbar[j] = S.xbar[t];
    }
  }
  for (int t = threadIdx.x; t < nrs * 32; t += blockDim.x) {
    const int64_t i = r_lo * 32 + t;
    if (i < m) y[i] = S.y[t];
  }
}
#endif  // __CUDACC__

}  // namespace
