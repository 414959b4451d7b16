// Stats block kernels (:248-291) and the halo push / wait / pack kernels.
// Part of libcpppd (single translation unit, included by cpppd.cu).
#pragma once

namespace {

// ------------------------------------------------------------------------------------------
// stats block (:248-291)
// ------------------------------------------------------------------------------------------
// Column pass: c.x, c.x4, c.xr, #(xbar == 0), max(lb - x, x - ub); turns the d buffer into x4 in place
// and (force_integer) stores xr into xr_out.
__global__ void __launch_bounds__(kBlock)
k_stats_cols(Vec c, const double *__restrict__ x, const double *__restrict__ xbar, Vec lb, Vec ub,
             double *__restrict__ d_x4, double *__restrict__ xr_out, int64_t n, int force_integer,
             double *__restrict__ part) {
  double v[kColQ] = {0.0, 0.0, 0.0, 0.0, -INFINITY};
  for (int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x; j < n; j += (int64_t)gridDim.x * kBlock) {
    const double cj = c.at(j), xj = x[j];
    const double lj = lb.at(j), uj = ub.at(j);
    const double x4 = d_x4[j] < 0.0 ? uj : lj;  // x4 = lb; x4[d < 0] = ub[d < 0]  (:260-261)
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
    // bound part of SparseLP.max_constraint_violation (reference SparseLP.py:189-190)
    v[4] = nan_max(v[4], nan_max(__dsub_rn(lj, xj), __dsub_rn(xj, uj)));
  }
  block_reduce_write<kColQ>(v, 0x10u, part + (int64_t)blockIdx.x * kColQ);
}

// Row pass: A x, A x4, A xbar, A xr per row -> energy terms and violation maxima.
__global__ void __launch_bounds__(kBlock)
k_stats_rows(SellView A, const double *__restrict__ x, const double *__restrict__ x4,
             const double *__restrict__ xbar, const double *__restrict__ xr, Vec b,
             const double *__restrict__ y, int64_t m, int64_t m_eq, int force_integer,
             const double *__restrict__ row_off, double *__restrict__ part) {
  const double ninf = -INFINITY;
  double v[kRowQ] = {0.0, 0.0, 0.0, 0.0, ninf, ninf, ninf, ninf, ninf};
  const int lane = threadIdx.x & 31;
  for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; (i >> 5) < A.nslices;
       i += (int64_t)gridDim.x * kBlock) {
    const int64_t s = i >> 5;
    int64_t p0, p1;
    slice_range(A, s, p0, p1);
    double ax = 0.0, ax4 = 0.0, axb = 0.0, axr = 0.0;
    for (int64_t p = p0 + lane; p < p1; p += kSlice) {
      const int32_t jr = A.idx[p];
      if (jr >= 0) {
        const int32_t jc = jr & A.idx_mask;
        const double a = entry_value(A, p, jr);
        ax = __dadd_rn(ax, __dmul_rn(a, x[jc]));
        ax4 = __dadd_rn(ax4, __dmul_rn(a, x4[jc]));
        if (i < m_eq) axb = __dadd_rn(axb, __dmul_rn(a, xbar[jc]));
        if (force_integer) axr = __dadd_rn(axr, __dmul_rn(a, xr[jc]));
      }
    }
    if (i < m) {
      if (!force_integer) axr = ax;
      const double bi = b.at(i), yi = y[i];
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
      // residual of the caller's FULL LP (cpppd_set_row_offsets: variables eliminated before the solve shift every
      // row by a constant), for SparseLP.max_constraint_violation (reference SparseLP.py:186-204)
      const double full = __dsub_rn(__dsub_rn(axr, bi), row_off ? row_off[i] : 0.0);
      if (i < m_eq) v[7] = nan_max(v[7], fabs(full)); else v[8] = nan_max(v[8], full);
    }
  }
  block_reduce_write<kRowQ>(v, 0x1F0u, part + (int64_t)blockIdx.x * kRowQ);
}

// Ground-truth pass (solve()'s distance curves, reference SparseLP.py:1074-1082): for the entries of the
// ground truth this rank owns, sum |gt - x[idx]| and sum |gt - round(x[idx])|.
__global__ void __launch_bounds__(kBlock)
k_stats_gt(const int32_t *__restrict__ gt_idx, const double *__restrict__ gt_val, int64_t count,
           const double *__restrict__ x, double *__restrict__ part) {
  double v[kGtQ] = {0.0, 0.0};
  for (int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x; k < count; k += (int64_t)gridDim.x * kBlock) {
    const double xv = x[gt_idx[k]], g = gt_val[k];
    v[0] = __dadd_rn(v[0], fabs(__dsub_rn(g, xv)));
    v[1] = __dadd_rn(v[1], fabs(__dsub_rn(g, rint(xv))));
  }
  block_reduce_write<kGtQ>(v, 0u, part + (int64_t)blockIdx.x * kGtQ);
}

// One CTA: fold this rank's per-CTA partials in a fixed order into kStatQ numbers.
__global__ void __launch_bounds__(kBlock)
k_stats_local(const double *__restrict__ colpart, int nbc, const double *__restrict__ rowpart, int nbr,
              const double *__restrict__ gtpart, int nbg, double *__restrict__ out) {
  double cv[kColQ] = {0.0, 0.0, 0.0, 0.0, -INFINITY};
  double gv[kGtQ] = {0.0, 0.0};
  for (int bi = threadIdx.x; bi < nbg; bi += kBlock)
#pragma unroll
    for (int q = 0; q < kGtQ; ++q) gv[q] = __dadd_rn(gv[q], gtpart[(int64_t)bi * kGtQ + q]);
  const double ninf = -INFINITY;
  double rv[kRowQ] = {0.0, 0.0, 0.0, 0.0, ninf, ninf, ninf, ninf, ninf};
  for (int bi = threadIdx.x; bi < nbc; bi += kBlock) {
#pragma unroll
    for (int q = 0; q < 4; ++q) cv[q] = __dadd_rn(cv[q], colpart[(int64_t)bi * kColQ + q]);
    cv[4] = nan_max(cv[4], colpart[(int64_t)bi * kColQ + 4]);
  }
  for (int bi = threadIdx.x; bi < nbr; bi += kBlock) {
#pragma unroll
    for (int q = 0; q < 4; ++q) rv[q] = __dadd_rn(rv[q], rowpart[(int64_t)bi * kRowQ + q]);
#pragma unroll
    for (int q = 4; q < kRowQ; ++q) rv[q] = nan_max(rv[q], rowpart[(int64_t)bi * kRowQ + q]);
  }
  __shared__ double fin[kStatQ];
  block_reduce_write<kColQ>(cv, 0x10u, fin);
  __syncthreads();
  block_reduce_write<kRowQ>(rv, 0x1F0u, fin + kColQ);
  __syncthreads();
  block_reduce_write<kGtQ>(gv, 0u, fin + kColQ + kRowQ);
  __syncthreads();
  if (threadIdx.x < kStatQ) out[threadIdx.x] = fin[threadIdx.x];
}

// One thread: fold the ranks' numbers in rank order, then apply :248-291's scalar logic.
__global__ void k_stats_final(const double *__restrict__ all, int world, int64_t n_glob, int has_eq, int has_ineq,
                              int64_t niter, int64_t gt_count, StatsDev *out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double fin[kStatQ];
  for (int q = 0; q < kStatQ; ++q) fin[q] = all[q];
  for (int r = 1; r < world; ++r)
    for (int q = 0; q < kStatQ; ++q) {
      const double v = all[r * kStatQ + q];
      fin[q] = stat_is_max(q) ? nan_max(fin[q], v) : __dadd_rn(fin[q], v);
    }
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
  s.max_violated_equality_full = has_eq ? fin[kColQ + 7] : 0.0;
  s.max_violated_inequality_full = fin[kColQ + 8];
  s.energy_rounded = fin[2];
  s.frac_zero_xbar = n_glob > 0 ? fin[3] / (double)n_glob : 0.0;
  s.max_bound_violation = fin[4];
  s.distance_to_ground_truth = gt_count > 0 ? fin[kColQ + kRowQ + 0] / (double)gt_count : 0.0;
  s.distance_to_ground_truth_rounded = gt_count > 0 ? fin[kColQ + kRowQ + 1] / (double)gt_count : 0.0;
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

// Halo push over peer memory: entry k of the send list goes to peer push_peer[k], element
// push_dst[k] of that peer's vector (its ghost slot).  The last CTA to finish raises, on every
// neighbour it sent to, the flag [kind * world + me] to the new stamp (release at system scope after
// every CTA fenced its stores).
// Tail of a push kernel, run by thread 0 of every CTA after the CTA's stores were fenced: the LAST CTA to arrive
// publishes the new stamp to the neighbours and then — the wait of this exchange, merged into the same launch: one
// kernel less per half-iteration, which counts when a rank's kernels take 100 us — spins until the neighbours' stamps
// of this exchange arrived (`my_flags` != nullptr; with a time-out, see wait_for_stamp).
__device__ __forceinline__ void push_finish(const PeerPtrs &P, int kind, int world, int me, unsigned long long send_mask,
                                            const unsigned long long *my_flags, unsigned long long recv_mask, SyncState *st) {
  const unsigned int ticket = atomicAdd(&st->ticket[kind], 1u);
  if (ticket != gridDim.x - 1) return;
  __threadfence_system();
  st->ticket[kind] = 0;
  const unsigned long long stamp = st->push_stamp[kind] + 1;
  st->push_stamp[kind] = stamp;
  for (int t = 0; t < world; ++t)
    if ((send_mask >> t) & 1ull) st_release_sys(P.flags[t] + kind * world + me, stamp);
  if (my_flags && recv_mask) {
    const unsigned long long want = st->wait_stamp[kind] + 1;
    for (int t = 0; t < world; ++t)
      if ((recv_mask >> t) & 1ull) wait_for_stamp(my_flags + kind * world + t, want, st, 200);
    st->wait_stamp[kind] = want;
  }
}

__global__ void __launch_bounds__(kBlock)
k_push(const double *__restrict__ vec, const int32_t *__restrict__ src, const int64_t *__restrict__ dst,
       const int32_t *__restrict__ peer, int64_t count, PeerPtrs P, int kind, int world, int me,
       unsigned long long send_mask, const unsigned long long *__restrict__ my_flags, unsigned long long recv_mask,
       SyncState *st) {
  const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (k < count) P.vec[peer[k]][dst[k]] = vec[src[k]];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x != 0) return;
  push_finish(P, kind, world, me, send_mask, my_flags, recv_mask, st);
}

// Dense halo (patterns without locality: every peer needs every owned entry, and its ghost slots for this rank are
// one contiguous run in this rank's order): no index lists, the owned part of the vector is streamed once and
// stored to every peer, coalesced.  Same stamp protocol as k_push.
struct DenseDst {
  int64_t base[kMaxWorld];  // first ghost slot of this rank's entries in peer t's vector
};
__global__ void __launch_bounds__(kBlock)
k_push_dense(const double *__restrict__ vec, int64_t owned, PeerPtrs P, DenseDst D, int kind, int world, int me,
             unsigned long long send_mask, const unsigned long long *__restrict__ my_flags, unsigned long long recv_mask,
             SyncState *st) {
  for (int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x; k < owned; k += (int64_t)gridDim.x * kBlock) {
    const double v = vec[k];
    for (int t = 0; t < world; ++t)
      if ((send_mask >> t) & 1ull) P.vec[t][D.base[t] + k] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x != 0) return;
  push_finish(P, kind, world, me, send_mask, my_flags, recv_mask, st);
}

// Wait until every neighbour this rank receives from has pushed its halo for this exchange (a rank that receives
// without sending anything; otherwise the wait is the tail of the push kernel).
__global__ void k_wait(const unsigned long long *__restrict__ flags, int kind, int world,
                       unsigned long long recv_mask, SyncState *st) {
  const int t = threadIdx.x;
  const unsigned long long want = st->wait_stamp[kind] + 1;
  if (t < world && ((recv_mask >> t) & 1ull)) wait_for_stamp(flags + kind * world + t, want, st, 200);
  __syncthreads();
  if (t == 0) st->wait_stamp[kind] = want;
}

// role byte of every slice (see FusedComm): bit 0 = holds a row that is sent, bit 1 = reads a ghost
__global__ void k_slice_roles(SellView S, int64_t owned_other, const int32_t *__restrict__ send_idx,
                              int64_t send_total, unsigned int *__restrict__ role32) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < send_total) atomicOr(role32 + (send_idx[t] >> 5), 1u);
  const int64_t s = t >> 5;
  if (s >= S.nslices) return;
  const int lane = threadIdx.x & 31;
  int64_t p0, p1;
  slice_range(S, s, p0, p1);
  bool ghost = false;
  for (int64_t p = p0 + lane; p < p1; p += kSlice) {
    const int32_t w = S.idx[p];
    if (w >= 0 && (w & S.idx_mask) >= owned_other) ghost = true;
  }
  if (__any_sync(0xffffffffu, ghost) && lane == 0) atomicOr(role32 + s, 2u);
}

__global__ void k_narrow_roles(const unsigned int *__restrict__ role32, int64_t count, unsigned char *__restrict__ role) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < count) role[s] = (unsigned char)role32[s];
}

// halo staging: buf[k] = vec[idx[k]]
__global__ void k_pack(const double *__restrict__ vec, const int32_t *__restrict__ idx, int64_t count,
                       double *__restrict__ buf) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) buf[k] = vec[idx[k]];
}

}  // namespace
