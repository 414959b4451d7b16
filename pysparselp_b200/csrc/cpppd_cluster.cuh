// Small LPs that exceed one SM: all k iterations in ONE launch of ONE thread-block cluster (<= 16 CTAs).
// Part of libcpppd (single translation unit, included by cpppd.cu after cpppd_hot_kernels.cuh).
//
// The Potts 50x50 LP of the reference's regression test (tests/test_pott_segmentation.py: 7 400 x 9 800, 29 400
// entries; BASELINE configs[0]) is bound by launch latency on the CUDA-graph path — two graph nodes of ~3.5 us around
// well under a microsecond of work — and too large for the one-CTA kernel (k_tiny_iterate).  Here every CTA of the
// cluster owns a contiguous range of slices of A^T (columns) and of A (rows) and keeps, for the whole launch, in its
// own shared memory: its entries of both operands, c / T / lb / ub / x / xbar of its columns, b / sigma / y of its
// rows.  A gather of y[i] or xbar[j] is a load from the OWNER's shared memory (distributed shared memory: the index
// words are translated once, at staging, into shared::cluster addresses); the two halves of an iteration are separated by
// the hardware cluster barrier (barrier.cluster arrive.release / wait.acquire) instead of a kernel boundary.
// Nothing touches L2 / HBM between staging and the final write-back of x, xbar, y.
//
// Arithmetic: per row / column the same operations in the same order as k_primal / k_dual variant 0 (sequential sum
// in stored entry order, equality / inequality parts apart, the same epilogues) — the iterates are the same bits;
// tests/test_gpu_parity.py compares them with the goldens minted from the reference.
//
// CPU emulation (tests/emul runs one CTA at a time): the staging and the per-row code below are compiled for the host
// as they are and driven phase by phase — one emulated launch per half-iteration instead of the cluster barrier, a
// host buffer per CTA instead of its shared memory (see the `#else` branches); what only hardware shows (the barrier,
// distributed shared memory, the one-pass register form of the kernel) stays with the -m gpu tests.
#pragma once

#ifdef __CUDACC__
#include <cooperative_groups.h>
#endif

namespace {

constexpr int kClusterBlock = 1024;
constexpr int kClusterMaxCtas = 16;
constexpr int64_t kClusterMaxEntries = 131072;  // padded entries per operand: beyond, a half-iteration is no longer
                                               // latency-bound and the streaming kernels are the better shape
constexpr int kClusterDefaultMode = 0;         // see cluster_barrier()
constexpr int kClC = 4;                        // entries per chunk: gathers in flight per thread

// layout of a CTA's dynamic shared memory (identical in every CTA: a remote address is the local one, mapped)
struct ClusterSmem {
  double *val_at, *val_a, *c, *T, *lb, *ub, *x, *xbar, *b, *sigma, *y, *zero;
  uint32_t *w_at, *w_a;
  int32_t *sp_at, *sp_a;
};

inline __host__ __device__ size_t cluster_smem_bytes(int spc_at, int spc_a, int ent_at, int ent_a) {
  const size_t cols = (size_t)spc_at * 32, rows = (size_t)spc_a * 32;
  return 8 * ((size_t)ent_at + ent_a + 6 * cols + 3 * rows + 2) + 4 * ((size_t)ent_at + ent_a + spc_at + spc_a + 2) + 16;
}

#ifdef __CUDACC__
namespace cg = cooperative_groups;
#endif

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
  S.zero = d; d += 2;
  uint32_t *w = reinterpret_cast<uint32_t *>(d);
  S.w_at = w; w += ent_at;
  S.w_a = w; w += ent_a;
  int32_t *sp = reinterpret_cast<int32_t *>(w);
  S.sp_at = sp; sp += spc_at + 1;
  S.sp_a = sp;
  return S;
}

#ifdef __CUDACC__
// 32-bit shared::cluster address of `local` (an address in this CTA's shared window) inside CTA `rank`
__device__ __forceinline__ uint32_t cluster_map(uint32_t local, int rank) {
  uint32_t out;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(local), "r"(rank));
  return out;
}
// generic address of a shared::cluster address: (shared window base, constant for the kernel) | 32-bit address
__device__ __forceinline__ uint64_t cluster_generic_base(uint32_t some_cluster_addr) {
  uint64_t gen;
  asm("cvta.shared::cluster.u64 %0, %1;" : "=l"(gen) : "l"((uint64_t)some_cluster_addr));
  return gen & 0xffffffff00000000ull;
}
__device__ __forceinline__ double cluster_load(uint64_t base, uint32_t addr) {
  double v;
  asm volatile("ld.f64 %0, [%1];" : "=d"(v) : "l"(base | addr) : "memory");
  return v;
}

#else
// ---- CPU emulation (tests/emul): one host buffer per CTA stands for its shared memory; a shared::cluster address is
//      [CTA rank : 6][offset inside the buffer : 26], the "generic base" is 0, mapa / cvta / ld are arithmetic on that.
inline thread_local unsigned char *g_emul_cluster_smem[kClusterMaxCtas] = {};
inline size_t __cvta_generic_to_shared(const void *p) {
  return (size_t)(static_cast<const unsigned char *>(p) - g_emul_cluster_smem[blockIdx.x]);
}
inline uint32_t cluster_map(uint32_t local, int rank) { return ((uint32_t)rank << 26) | local; }
inline uint64_t cluster_generic_base(uint32_t) { return 0; }
inline double cluster_load(uint64_t, uint32_t addr) {
  return *reinterpret_cast<const double *>(g_emul_cluster_smem[addr >> 26] + (addr & 0x03fffff8u));
}
#endif

// Entries of the slices [s_lo, s_hi) of S into shared memory: the value, and the index word translated ONCE into
// the shared::cluster address of the gathered element inside its owner CTA (mapa) — a gather in the loop is then
// one load, no owner / offset arithmetic.  The addresses are 8-byte aligned: bit 0 carries the equality flag.
// Every slice is staged with its width rounded up to a whole number of chunks of kClC entries, and every padding
// entry (the SELL padding of shorter rows and the chunk padding) becomes value 0.0 x the address of a zero in this
// CTA's own shared memory: the loops then need no bounds or validity tests, and the iterates keep their bits — a
// partial sum starts at +0.0 and can never become -0.0, so adding the +0.0 of a padding entry changes nothing.
__device__ __forceinline__ void cluster_stage(const SellView &S, int64_t s_lo, int64_t s_hi, int spc_other,
                                              const double *gathered_vec, const double *zero, int me, double *val,
                                              uint32_t *w, int32_t *sp) {
  if (threadIdx.x == 0) {  // local offsets of the padded slices (at most a few dozen per CTA)
    int32_t off = 0;
    for (int64_t s = s_lo; s < s_hi; ++s) {
      int64_t p0, p1;
      slice_range(S, s, p0, p1);
      sp[s - s_lo] = off;
      off += (int32_t)(((p1 - p0) / 32 + kClC - 1) / kClC * kClC * 32);
    }
    sp[s_hi - s_lo] = off;
  }
  __syncthreads();
  const int32_t per_owner = spc_other * 32;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(gathered_vec);
  const uint32_t pad_word = cluster_map((uint32_t)__cvta_generic_to_shared(zero), me);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t s = s_lo + warp; s < s_hi; s += nwarps) {
    int64_t p0, p1;
    slice_range(S, s, p0, p1);
    const int width = (int)((p1 - p0) >> 5), padded = (sp[s - s_lo + 1] - sp[s - s_lo]) >> 5;
    for (int k = 0; k < padded; ++k) {
      uint32_t out = pad_word;
      double a = 0.0;
      if (k < width) {
        const int64_t e = p0 + (int64_t)k * 32 + lane;
        const int32_t word = S.idx[e];
        if (word >= 0) {
          a = entry_value(S, e, word);
          const int32_t i = word & S.idx_mask;
          const int32_t owner = i / per_owner;
          out = cluster_map(base + 8u * (uint32_t)(i - owner * per_owner), owner) | ((word & kEqBit) ? 1u : 0u);
        }
      }
      val[sp[s - s_lo] + k * 32 + lane] = a;
      w[sp[s - s_lo] + k * 32 + lane] = out;
    }
  }
}

#ifdef __CUDACC__
// The barrier between the two halves of an iteration.  What has to be ordered is: my st.shared of xbar / y, then the
// other CTAs' loads of it from my shared memory after the barrier.  barrier.cluster.arrive.release compiles to
// MEMBAR.ALL.GPU + UCGABAR_ARV (cuobjdump): a GPU-scope fence, twice per iteration, for data that never leaves shared
// memory (measured: 10 % of the iteration time of the Potts 50x50 LP).  kStrict == 0 uses membar.cta + the RELAXED
// arrive instead: shared memory has one physical copy (no cache in front of it), so once the store is performed at
// CTA scope a later load through the cluster network reads it; the wait keeps its acquire form, the loads after it
// are issued in program order.  kStrict == 1 (CPPPD_CLUSTER_MODE=1) keeps release / acquire.
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

#endif

// per-column / per-row constants and state of one thread
struct ClusterCol { double c, T, lb, ub, x; int q0, width; bool live; };   // width: padded, a multiple of kClC
struct ClusterRow { double b, sigma, y; int q0, width; bool live, ineq; };

// primal half for one column (:198-228): same operations, same order as primal_rows / primal_sums<_, 0>
// (kEq false: the LP has no equality row, no entry carries the flag)
template <bool kEq>
__device__ __forceinline__ void cluster_primal(const ClusterSmem &S, uint64_t gbase, ClusterCol &C, int jl, int has_eq,
                                               int has_ineq, double theta, double one_plus_theta) {
  double s_eq = 0.0, s_in = 0.0;
#pragma unroll 1
  for (int k0 = 0; k0 < C.width; k0 += kClC) {
    uint32_t w[kClC];
    double a[kClC], g[kClC];
#pragma unroll
    for (int u = 0; u < kClC; ++u) w[u] = S.w_at[C.q0 + (k0 + u) * 32];
#pragma unroll
    for (int u = 0; u < kClC; ++u) g[u] = cluster_load(gbase, kEq ? (w[u] & ~1u) : w[u]);
#pragma unroll
    for (int u = 0; u < kClC; ++u) a[u] = S.val_at[C.q0 + (k0 + u) * 32];
#pragma unroll
    for (int u = 0; u < kClC; ++u) {
      const double t = __dmul_rn(a[u], g[u]);
      if (kEq && (w[u] & 1u)) s_eq = __dadd_rn(s_eq, t); else s_in = __dadd_rn(s_in, t);
    }
  }
  if (!C.live) return;
  double d = C.c;
  if (has_eq) d = __dadd_rn(d, s_eq);
  if (has_ineq) d = __dadd_rn(d, s_in);
  double x2 = __dsub_rn(C.x, __dmul_rn(C.T, d));
  x2 = (C.lb > x2) ? C.lb : x2;
  x2 = (C.ub < x2) ? C.ub : x2;
  S.xbar[jl] = __dsub_rn(__dmul_rn(one_plus_theta, x2), __dmul_rn(theta, C.x));
  C.x = x2;
}

// dual half for one row (:231-240, :333-341): same operations, same order as dual_rows / dual_sum<_, 0>
__device__ __forceinline__ void cluster_dual(const ClusterSmem &S, uint64_t gbase, ClusterRow &R, int il) {
  double acc = 0.0;
#pragma unroll 1
  for (int k0 = 0; k0 < R.width; k0 += kClC) {
    uint32_t w[kClC];
    double a[kClC], g[kClC];
#pragma unroll
    for (int u = 0; u < kClC; ++u) w[u] = S.w_a[R.q0 + (k0 + u) * 32];
#pragma unroll
    for (int u = 0; u < kClC; ++u) g[u] = cluster_load(gbase, w[u]);
#pragma unroll
    for (int u = 0; u < kClC; ++u) a[u] = S.val_a[R.q0 + (k0 + u) * 32];
#pragma unroll
    for (int u = 0; u < kClC; ++u) acc = __dadd_rn(acc, __dmul_rn(a[u], g[u]));
  }
  if (!R.live) return;
  const double r = __dsub_rn(acc, R.b);
  double yn = __dadd_rn(R.y, __dmul_rn(R.sigma, r));
  if (R.ineq) yn = (yn < 0.0) ? 0.0 : yn;
  R.y = yn;
  S.y[il] = yn;
}

__device__ __forceinline__ ClusterCol cluster_col(const ClusterSmem &S, int sl, int lane, bool live) {
  const int jl = sl * 32 + lane;
  return ClusterCol{S.c[jl], S.T[jl], S.lb[jl], S.ub[jl], S.x[jl], S.sp_at[sl] + lane, (S.sp_at[sl + 1] - S.sp_at[sl]) >> 5, live};
}
__device__ __forceinline__ ClusterRow cluster_row(const ClusterSmem &S, int sl, int lane, bool live, bool ineq) {
  const int il = sl * 32 + lane;
  return ClusterRow{S.b[il], S.sigma[il], S.y[il], S.sp_a[sl] + lane, (S.sp_a[sl + 1] - S.sp_a[sl]) >> 5, live, ineq};
}

#ifdef __CUDACC__
// kOnePass: every CTA has at most one slice of A^T and one of A per warp — a thread keeps its column and its row
// (constants, x, y, entry offsets) in registers for the whole launch.  Otherwise the warps stride over the slices and
// reload them from shared memory every iteration.
template <int kStrict, bool kOnePass>
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
  if (threadIdx.x < 2) S.zero[threadIdx.x] = 0.0;
  cluster_stage(AT, c_lo, c_hi, spc_a, S.y, S.zero, me, S.val_at, S.w_at, S.sp_at);    // A^T gathers y: owners by row slices
  cluster_stage(A, r_lo, r_hi, spc_at, S.xbar, S.zero, me, S.val_a, S.w_a, S.sp_a);    // A gathers xbar: owners by column slices
  const uint64_t gbase = cluster_generic_base(cluster_map((uint32_t)__cvta_generic_to_shared(S.zero), me));
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
  if (kOnePass) {
    const bool has_col = warp < ncs, has_row = warp < nrs;
    ClusterCol C = cluster_col(S, has_col ? warp : 0, lane, has_col && c_lo * 32 + warp * 32 + lane < n);
    ClusterRow R = cluster_row(S, has_row ? warp : 0, lane, has_row && r_lo * 32 + warp * 32 + lane < m,
                               r_lo * 32 + warp * 32 + lane >= m_eq);
    if (!has_col) C.width = 0;
    if (!has_row) R.width = 0;
    for (int64_t it = 0; it < iters; ++it) {
      if (has_eq) cluster_primal<true>(S, gbase, C, warp * 32 + lane, has_eq, has_ineq, theta, one_plus_theta);
      else cluster_primal<false>(S, gbase, C, warp * 32 + lane, has_eq, has_ineq, theta, one_plus_theta);
      cluster_barrier<kStrict>();
      cluster_dual(S, gbase, R, warp * 32 + lane);
      cluster_barrier<kStrict>();
    }
    if (C.live) S.x[warp * 32 + lane] = C.x;
  } else {
    for (int64_t it = 0; it < iters; ++it) {
      for (int sl = warp; sl < ncs; sl += nwarps) {
        ClusterCol C = cluster_col(S, sl, lane, c_lo * 32 + sl * 32 + lane < n);
        if (has_eq) cluster_primal<true>(S, gbase, C, sl * 32 + lane, has_eq, has_ineq, theta, one_plus_theta);
        else cluster_primal<false>(S, gbase, C, sl * 32 + lane, has_eq, has_ineq, theta, one_plus_theta);
        if (C.live) S.x[sl * 32 + lane] = C.x;
      }
      cluster_barrier<kStrict>();
      for (int sl = warp; sl < nrs; sl += nwarps) {
        ClusterRow R = cluster_row(S, sl, lane, r_lo * 32 + sl * 32 + lane < m, r_lo * 32 + sl * 32 + lane >= m_eq);
        cluster_dual(S, gbase, R, sl * 32 + lane);
      }
      cluster_barrier<kStrict>();
    }
  }
  __syncthreads();
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
#else  // ---- CPU emulation of the cluster kernel: the same staging and per-row code, one emulated launch per phase
// (the kernel boundary stands for the cluster barrier; CTAs run one after the other, a remote gather is a read of the
// other CTA's host buffer).  Per-thread state is reloaded from "shared memory" in every phase (the kOnePass = false form).
struct EmulClusterArgs {
  SellView AT, A;
  Vec c, T, lb, ub, b, sigma;
  double *x, *xbar, *y;
  int64_t n, m, m_eq;
  int has_eq, has_ineq;
  double theta, one_plus_theta;
  int spc_at, spc_a, ent_at, ent_a;
};
struct EmulClusterCta {
  ClusterSmem S;
  int me, lane, warp, nwarps, ncs, nrs;
  int64_t c_lo, r_lo;
};
inline EmulClusterCta emul_cluster_cta(const EmulClusterArgs &a) {
  EmulClusterCta q;
  q.me = (int)blockIdx.x;
  q.S = cluster_carve(g_emul_cluster_smem[q.me], a.spc_at, a.spc_a, a.ent_at, a.ent_a);
  q.lane = threadIdx.x & 31;
  q.warp = threadIdx.x >> 5;
  q.nwarps = blockDim.x >> 5;
  q.c_lo = std::min((int64_t)q.me * a.spc_at, a.AT.nslices);
  q.r_lo = std::min((int64_t)q.me * a.spc_a, a.A.nslices);
  q.ncs = (int)(std::min(q.c_lo + a.spc_at, a.AT.nslices) - q.c_lo);
  q.nrs = (int)(std::min(q.r_lo + a.spc_a, a.A.nslices) - q.r_lo);
  return q;
}
__global__ void k_emul_cluster_stage(EmulClusterArgs a) {
  const EmulClusterCta q = emul_cluster_cta(a);
  const ClusterSmem &S = q.S;
  if (threadIdx.x < 2) S.zero[threadIdx.x] = 0.0;
  cluster_stage(a.AT, q.c_lo, q.c_lo + q.ncs, a.spc_a, S.y, S.zero, q.me, S.val_at, S.w_at, S.sp_at);
  cluster_stage(a.A, q.r_lo, q.r_lo + q.nrs, a.spc_at, S.xbar, S.zero, q.me, S.val_a, S.w_a, S.sp_a);
  for (int t = threadIdx.x; t < a.spc_at * 32; t += blockDim.x) {
    const int64_t j = q.c_lo * 32 + t;
    const bool live = t < q.ncs * 32 && j < a.n;
    S.c[t] = live ? (a.c.p ? a.c.p[j] : a.c.c) : 0.0;
    S.T[t] = live ? (a.T.p ? a.T.p[j] : a.T.c) : 0.0;
    S.lb[t] = live ? (a.lb.p ? a.lb.p[j] : a.lb.c) : 0.0;
    S.ub[t] = live ? (a.ub.p ? a.ub.p[j] : a.ub.c) : 0.0;
    S.x[t] = live ? a.x[j] : 0.0;
    S.xbar[t] = live ? a.xbar[j] : 0.0;
  }
  for (int t = threadIdx.x; t < a.spc_a * 32; t += blockDim.x) {
    const int64_t i = q.r_lo * 32 + t;
    const bool live = t < q.nrs * 32 && i < a.m;
    S.b[t] = live ? (a.b.p ? a.b.p[i] : a.b.c) : 0.0;
    S.sigma[t] = live ? (a.sigma.p ? a.sigma.p[i] : a.sigma.c) : 0.0;
    S.y[t] = live ? a.y[i] : 0.0;
  }
}
__global__ void k_emul_cluster_primal(EmulClusterArgs a) {
  const EmulClusterCta q = emul_cluster_cta(a);
  const uint64_t gbase = cluster_generic_base(0);
  for (int sl = q.warp; sl < q.ncs; sl += q.nwarps) {
    ClusterCol C = cluster_col(q.S, sl, q.lane, q.c_lo * 32 + sl * 32 + q.lane < a.n);
    if (a.has_eq) cluster_primal<true>(q.S, gbase, C, sl * 32 + q.lane, a.has_eq, a.has_ineq, a.theta, a.one_plus_theta);
    else cluster_primal<false>(q.S, gbase, C, sl * 32 + q.lane, a.has_eq, a.has_ineq, a.theta, a.one_plus_theta);
    if (C.live) q.S.x[sl * 32 + q.lane] = C.x;
  }
}
__global__ void k_emul_cluster_dual(EmulClusterArgs a) {
  const EmulClusterCta q = emul_cluster_cta(a);
  const uint64_t gbase = cluster_generic_base(0);
  for (int sl = q.warp; sl < q.nrs; sl += q.nwarps) {
    ClusterRow R = cluster_row(q.S, sl, q.lane, q.r_lo * 32 + sl * 32 + q.lane < a.m, q.r_lo * 32 + sl * 32 + q.lane >= a.m_eq);
    cluster_dual(q.S, gbase, R, sl * 32 + q.lane);
  }
}
__global__ void k_emul_cluster_writeback(EmulClusterArgs a) {
  const EmulClusterCta q = emul_cluster_cta(a);
  for (int t = threadIdx.x; t < q.ncs * 32; t += blockDim.x) {
    const int64_t j = q.c_lo * 32 + t;
    if (j < a.n) {
      a.x[j] = q.S.x[t];
      a.xbar[j] = q.S.xbar[t];
    }
  }
  for (int t = threadIdx.x; t < q.nrs * 32; t += blockDim.x) {
    const int64_t i = q.r_lo * 32 + t;
    if (i < a.m) a.y[i] = q.S.y[t];
  }
}
#endif  // __CUDACC__

}  // namespace
