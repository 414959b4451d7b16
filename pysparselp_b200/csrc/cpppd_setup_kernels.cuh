// Setup kernels: validation, locality keys, partition, local matrices, SELL-32 fill, transpose,
// Part of libcpppd (single translation unit, included by cpppd.cu).
// preconditioners (reference pysparselp/ChambollePockPPD.py:122-179).
#pragma once

namespace {

// ------------------------------------------------------------------------------------------
// setup kernels: validation, locality keys, partition, local matrices, SELL-32, transpose
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

__global__ void k_fill_i32(int32_t *p, int64_t count, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) p[i] = v;
}

// row_key[i] = min column index of row i (n for an empty row)
__global__ void k_row_key(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices, int64_t m,
                          int32_t n, int32_t *__restrict__ row_key) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  int32_t k = n;
  for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) k = min(k, indices[e]);
  row_key[i] = k;
}

// col_key[j] = min row_key over the rows that hit column j; col_len[j] = entries of column j
__global__ void k_col_key(const int32_t *__restrict__ indices, const uint32_t *__restrict__ row_of, int64_t nnz,
                          const int32_t *__restrict__ row_key, int32_t *__restrict__ col_key,
                          int32_t *__restrict__ col_len) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int32_t j = indices[e];
  atomicMin(col_key + j, row_key[row_of[e]]);
  atomicAdd(col_len + j, 1);
}

// work[bucket] += entries (rows: their length; columns: their length)
__global__ void k_bucket_work(const int32_t *__restrict__ key, const int64_t *__restrict__ rowptr,
                              const int32_t *__restrict__ len32, int64_t count, int32_t granule,
                              unsigned long long *__restrict__ work) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  unsigned long long w = rowptr ? (unsigned long long)(rowptr[i + 1] - rowptr[i]) : (unsigned long long)len32[i];
  if (w) atomicAdd(work + key[i] / granule, w);
}

// sort key of a row / column: (owner, [is_ineq,] bucket); also counts per owner.
// The owner is the owner of the bucket, or — balanced split, prefix != nullptr — decided per row / column from
// the entries in front of it in original order: min(world - 1, prefix[i] * world / total).
__global__ void k_sort_keys(const int32_t *__restrict__ key, int64_t count, int32_t granule,
                            const int32_t *__restrict__ owner_of_bucket, int64_t m_eq, int is_rows,
                            const int64_t *__restrict__ rowptr, const int32_t *__restrict__ len32,
                            const int64_t *__restrict__ prefix, int64_t total, int world, int keep_order,
                            uint64_t *__restrict__ out_key, uint32_t *__restrict__ out_id,
                            int32_t *__restrict__ count_per_owner, int32_t *__restrict__ eq_per_owner) {
  // per-owner counts: one shared-memory histogram per CTA, then one global atomic per CTA and owner (a global
  // atomicAdd per row on `world` addresses serialised 117 M updates of the 4096^2 Potts LP on 8 counters)
  __shared__ int32_t hist[2 * kMaxWorld];
  for (int t = threadIdx.x; t < 2 * kMaxWorld; t += blockDim.x) hist[t] = 0;
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    int32_t q = key[i] / granule;
    int32_t o = owner_of_bucket[q];
    if (prefix) o = (int32_t)min((int64_t)world - 1, prefix[i] * world / total);
    // inside a bucket, rows / columns of equal length sit together (SELL sigma-sorting: slices of
    // 32 neighbours then have nearly equal widths and little padding)
    int64_t len = rowptr ? rowptr[i + 1] - rowptr[i] : (int64_t)len32[i];
    uint64_t len12 = (uint64_t)(len > 4095 ? 4095 : len);
    uint64_t major = is_rows ? (uint64_t)o * 2 + (i >= m_eq ? 1 : 0) : (uint64_t)o;
    // keep_order (banded operands on several GPUs): original order inside (owner, kind) — the sort is stable
    out_key[i] = keep_order ? (major << 44) : ((major << 44) | ((uint64_t)(uint32_t)q << 12) | len12);
    out_id[i] = (uint32_t)i;
    atomicAdd(hist + o, 1);
    if (is_rows && i < m_eq) atomicAdd(hist + kMaxWorld + o, 1);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < world; t += blockDim.x) {
    if (hist[t]) atomicAdd(count_per_owner + t, hist[t]);
    if (eq_per_owner && hist[kMaxWorld + t]) atomicAdd(eq_per_owner + t, hist[kMaxWorld + t]);
  }
}

__global__ void k_col_len(const int32_t *__restrict__ indices, int64_t nnz, int32_t *__restrict__ col_len) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) atomicAdd(col_len + indices[e], 1);
}

// total[0] += 32 * (longest row of each slice of 32 consecutive rows)
__global__ void k_padded_total(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ len32, int64_t nrows,
                               unsigned long long *__restrict__ total) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int lane = threadIdx.x & 31;
  if (r - lane >= nrows) return;
  long long len = 0;
  if (r < nrows) len = rowptr ? rowptr[r + 1] - rowptr[r] : (long long)len32[r];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if (lane == 0 && len) atomicAdd(total, (unsigned long long)len * kSlice);
}

__global__ void k_invert(const uint32_t *__restrict__ order, int64_t count, int32_t *__restrict__ pos) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < count) pos[order[p]] = (int32_t)p;
}

// ghost marking: a column position is a ghost of this rank when one of this rank's rows hits a
// column owned elsewhere; a row position is a ghost when it hits one of this rank's columns.
__global__ void k_mark_ghosts(const int32_t *__restrict__ indices, const uint32_t *__restrict__ row_of, int64_t nnz,
                              const int32_t *__restrict__ row_pos, const int32_t *__restrict__ col_pos, int32_t rs,
                              int32_t re, int32_t cs, int32_t ce, int32_t *__restrict__ gcol_flag,
                              int32_t *__restrict__ grow_flag) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int32_t rp = row_pos[row_of[e]], cp = col_pos[indices[e]];
  bool row_mine = rp >= rs && rp < re, col_mine = cp >= cs && cp < ce;
  if (row_mine && !col_mine) gcol_flag[cp] = 1;
  if (col_mine && !row_mine) grow_flag[rp] = 1;
}

// What this rank must send to its peers: its columns hit by their rows, its rows hitting their columns — for all
// peers in ONE pass over the entries: bit t of sendx_mask[j] / sendy_mask[i] is set when peer t needs
// this rank's column j / row i (world <= 64).  The owner of a position is found in the rank boundaries (N + 1 each).
struct RankStarts {
  int32_t row[kMaxWorld + 1], col[kMaxWorld + 1];
};
__device__ __forceinline__ int owner_of(const int32_t *start, int world, int32_t pos) {
  int lo = 0, hi = world - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (start[mid] <= pos) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__global__ void k_mark_send_masks(const int32_t *__restrict__ indices, const uint32_t *__restrict__ row_of, int64_t nnz,
                                  const int32_t *__restrict__ row_pos, const int32_t *__restrict__ col_pos, int32_t rs,
                                  int32_t re, int32_t cs, int32_t ce, RankStarts starts, int world,
                                  unsigned long long *__restrict__ sendx_mask, unsigned long long *__restrict__ sendy_mask) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int32_t rp = row_pos[row_of[e]], cp = col_pos[indices[e]];
  const bool row_mine = rp >= rs && rp < re, col_mine = cp >= cs && cp < ce;
  if (col_mine && !row_mine) {
    const unsigned long long bit = 1ull << owner_of(starts.row, world, rp);
    if (!(sendx_mask[cp - cs] & bit)) atomicOr(sendx_mask + (cp - cs), bit);
  }
  if (row_mine && !col_mine) {
    const unsigned long long bit = 1ull << owner_of(starts.col, world, cp);
    if (!(sendy_mask[rp - rs] & bit)) atomicOr(sendy_mask + (rp - rs), bit);
  }
}
// dense halo: every position outside [lo, hi) is a ghost / every owned entry goes to every peer
__global__ void k_flag_outside(int32_t *__restrict__ flag, int64_t count, int32_t lo, int32_t hi) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) flag[i] = (i < lo || i >= hi) ? 1 : 0;
}
__global__ void k_fill_u64(unsigned long long *__restrict__ p, int64_t count, unsigned long long v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) p[i] = v;
}
__global__ void k_flag_from_mask(const unsigned long long *__restrict__ mask, int64_t count, int peer, int32_t *__restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) flag[i] = (int32_t)((mask[i] >> peer) & 1ull);
}

// out[base + scan[i]] = value(i) for flagged i
__global__ void k_compact(const int32_t *__restrict__ flag, const int32_t *__restrict__ scan, int64_t count,
                          const uint32_t *__restrict__ map, int32_t add, int32_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count || !flag[i]) return;
  out[scan[i]] = map ? (int32_t)map[i] : (int32_t)i + add;
}

__global__ void k_copy_u32_i32(const uint32_t *__restrict__ in, int64_t count, int32_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (int32_t)in[i];
}

// lengths of this rank's rows in local order
__global__ void k_local_row_len(const uint32_t *__restrict__ row_order, int32_t rs, int64_t mloc,
                                const int64_t *__restrict__ rowptr, int64_t *__restrict__ len) {
  int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (li > mloc) return;
  if (li == mloc) { len[li] = 0; return; }
  uint32_t old = row_order[rs + li];
  len[li] = rowptr[old + 1] - rowptr[old];
}

// this rank's rows of A in local numbering (entry order inside a row untouched)
__global__ void k_local_rows_fill(const uint32_t *__restrict__ row_order, int32_t rs, int64_t mloc,
                                  const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices,
                                  const double *__restrict__ values, const int32_t *__restrict__ col_pos, int32_t cs,
                                  int32_t ce, const int32_t *__restrict__ gcol_scan,
                                  const int64_t *__restrict__ lrowptr, int32_t *__restrict__ out_idx,
                                  double *__restrict__ out_val) {
  int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= mloc) return;
  uint32_t old = row_order[rs + li];
  int64_t src = rowptr[old], len = rowptr[old + 1] - src, dst = lrowptr[li];
  for (int64_t k = 0; k < len; ++k) {
    int32_t cp = col_pos[indices[src + k]];
    out_idx[dst + k] = (cp >= cs && cp < ce) ? cp - cs : (ce - cs) + gcol_scan[cp];
    out_val[dst + k] = values[src + k];
  }
}

__global__ void k_entry_col_pos(const int32_t *__restrict__ indices, const int32_t *__restrict__ col_pos,
                                int64_t nnz, uint32_t *__restrict__ out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) out[e] = (uint32_t)col_pos[indices[e]];
}

// first sorted position whose key is >= j, for j in [j0, j0 + count]
__global__ void k_lower_bounds(const uint32_t *__restrict__ sorted, int64_t nnz, int64_t j0, int64_t count,
                               int64_t *__restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > count) return;
  int64_t j = j0 + t, lo = 0, hi = nnz;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)sorted[mid] < j) lo = mid + 1; else hi = mid;
  }
  out[t] = lo;
}

// this rank's columns of A as rows of A^T: entries in original row order, local row numbering,
// equality rows tagged with kEqBit
__global__ void k_local_cols_fill(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ row_of,
                                  const double *__restrict__ values, int64_t first, int64_t count,
                                  const int32_t *__restrict__ row_pos, int32_t rs, int32_t re,
                                  const int32_t *__restrict__ grow_scan, int64_t m_eq_glob,
                                  int32_t *__restrict__ out_idx, double *__restrict__ out_val) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  uint32_t e = perm[first + t];
  uint32_t old = row_of[e];
  int32_t rp = row_pos ? row_pos[old] : (int32_t)old;
  int32_t local = (rp >= rs && rp < re) ? rp - rs : (re - rs) + grow_scan[rp];
  out_idx[t] = local | ((int64_t)old < m_eq_glob ? kEqBit : 0);
  out_val[t] = values[e];
}

__global__ void k_subtract_base(int64_t *p, int64_t count, int64_t base) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) p[i] -= base;
}

// dst[i] = src[map[i]]
__global__ void k_gather_f64(const double *__restrict__ src, const int32_t *__restrict__ map, int64_t count,
                             double *__restrict__ dst) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = src[map[i]];
}

// dst[map[i]] = src[i]
__global__ void k_scatter_f64(const double *__restrict__ src, const int32_t *__restrict__ map, int64_t count,
                              double *__restrict__ dst) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[map[i]] = src[i];
}

// diag_t (:122-153): thread per column of A, sequential over the column in row order,
// equality and inequality parts accumulated separately then  (0 + s_eq) + s_ineq.
__global__ void k_precond_cols(SellView AT, int64_t n, int has_eq, int has_ineq, double power,
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
      double t = __dmul_rn(abs_pow(entry_value(AT, p, r), power), 1.0);
      if (r & kEqBit) s_eq = __dadd_rn(s_eq, t); else s_in = __dadd_rn(s_in, t);
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
    const int32_t w = A.idx[p];
    if (w >= 0) acc = __dadd_rn(acc, __dmul_rn(abs_pow(entry_value(A, p, w), power), 1.0));
  }
  if (i < m) {
    if (acc == 0.0) acc = 1.0;
    sigma[i] = __ddiv_rn(1.0, acc);
  }
}

}  // namespace
