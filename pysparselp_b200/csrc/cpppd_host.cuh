// Host side: SELL build, dictionary detection, partition + halos, peer-memory setup, launches.
// Part of libcpppd (single translation unit, included by cpppd.cu).
#pragma once

namespace {

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
SellView view(const Sell &s) {
  const int code_bits = s.dict ? 30 - s.idx_bits : 0;
  return SellView{s.slice_ptr, s.idx, s.val, s.nrows, s.nslices, s.uniform_width, s.dict,
                  s.dict ? (int32_t)((1u << s.idx_bits) - 1) : kIdxMask, s.idx_bits, (int32_t)((1u << code_bits) - 1), s.ndict};
}

template <typename T>
int exclusive_scan(cpppd_solver *h, const T *in, T *out, int64_t count) {
  size_t bytes = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, count, h->stream));
  void *tmp = dev_alloc(h, bytes, false);
  if (!tmp) return fail(h, CPPPD_ERR_NOMEM, "scan workspace allocation failed");
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, count, h->stream);
  cudaStreamSynchronize(h->stream);
  dev_free(h, tmp);
  CK(e);
  return 0;
}

template <typename K>
int sort_pairs(cpppd_solver *h, cub::DoubleBuffer<K> &keys, cub::DoubleBuffer<uint32_t> &vals, int64_t count,
               int end_bit) {
  if (count == 0) return 0;
  size_t bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, vals, count, 0, end_bit, h->stream));
  void *tmp = dev_alloc(h, bytes, false);
  if (!tmp) return fail(h, CPPPD_ERR_NOMEM, "sort workspace allocation failed");
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, bytes, keys, vals, count, 0, end_bit, h->stream);
  cudaStreamSynchronize(h->stream);
  dev_free(h, tmp);
  CK(e);
  return 0;
}

int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b)) ++b;
  return b;
}

// CSR (device, int64 rowptr) -> SELL-32 (device)
int build_sell(cpppd_solver *h, const int64_t *rowptr, const int32_t *indices, const double *values, int64_t nrows,
               Sell *out) {
  out->nrows = nrows;
  out->nslices = (nrows + kSlice - 1) / kSlice;
  const int64_t ns = out->nslices;
  Scratch tmp(h);
  int64_t *extent = nullptr, *mm = nullptr;
  if (int rc = tmp.get(&extent, ns + 1)) return rc;
  if (int rc = tmp.get(&mm, 2)) return rc;
  if (int rc = alloc_array(h, &out->slice_ptr, ns + 1)) return rc;
  CK(cudaMemsetAsync(extent, 0, sizeof(int64_t) * (ns + 1), h->stream));
  if (ns) k_slice_extent<<<grid_for(ns * 32), kBlock, 0, h->stream>>>(rowptr, nrows, ns, extent);
  if (int rc = exclusive_scan(h, extent, out->slice_ptr, ns + 1)) return rc;
  CK(cudaMemcpyAsync(&out->padded, out->slice_ptr + ns, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  int64_t host_mm[2] = {0, 1};
  if (ns) {  // uniform slice width <=> min extent == max extent
    size_t b1 = 0, b2 = 0;
    CK(cub::DeviceReduce::Min(nullptr, b1, extent, mm, ns, h->stream));
    CK(cub::DeviceReduce::Max(nullptr, b2, extent, mm + 1, ns, h->stream));
    char *t2 = nullptr;
    if (int rc = tmp.get(&t2, (int64_t)std::max(b1, b2))) return rc;
    CK(cub::DeviceReduce::Min(t2, b1, extent, mm, ns, h->stream));
    CK(cub::DeviceReduce::Max(t2, b2, extent, mm + 1, ns, h->stream));
    CK(cudaMemcpyAsync(host_mm, mm, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  out->uniform_width = (ns && host_mm[0] == host_mm[1]) ? host_mm[0] / kSlice : -1;
  if (int rc = alloc_array(h, &out->idx, out->padded)) return rc;
  if (!out->dict)
    if (int rc = alloc_array(h, &out->val, out->padded)) return rc;
  if (ns) k_fill_sell<<<grid_for(ns * 32), kBlock, 0, h->stream>>>(rowptr, indices, values, nrows, ns, out->slice_ptr,
                                                                  out->idx, out->val, out->dict ? h->dict : nullptr,
                                                                  h->ndict, out->idx_bits);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// elements of the gathered vector per window: cpppd_problem.band_window, CPPPD_BAND_WINDOW_MB, or 56 MB — inside the
// L2-hit plateau tools/probe/gather_probe.cu found (<= 64 MB with the matrix streaming past it) and the best of the
// 24 .. 80 MB sweep on the 20M x 40M random LP (profiles/r02_random_lp.md)
int64_t band_window_elems(const cpppd_solver *h) {
  if (h->band_window > 0) return h->band_window;
  if (const char *env = getenv("CPPPD_BAND_WINDOW"))
    if (atoll(env) > 0) return atoll(env);
  double mb = 56.0;
  if (const char *env = getenv("CPPPD_BAND_WINDOW_MB")) mb = atof(env);
  return std::max<int64_t>(32, (int64_t)(mb * 1048576.0 / 8.0));
}

// sampled sectors per gather of a thread-per-row kernel on this CSR (1.0: no two lanes ever share a sector)
int band_locality(cpppd_solver *h, const int64_t *rowptr, const int32_t *indices, int64_t nrows, double *out) {
  *out = 0.0;
  if (nrows == 0) return 0;
  Scratch tmp(h);
  unsigned long long *acc = nullptr, host[2] = {0, 0};
  if (int rc = tmp.get(&acc, 2)) return rc;
  CK(cudaMemsetAsync(acc, 0, 2 * sizeof(unsigned long long), h->stream));
  const int64_t warps = (nrows + 31) / 32, sampled = std::min<int64_t>(warps, 8192), stride = std::max<int64_t>(1, warps / sampled);
  k_band_locality<<<grid_for(sampled * 32), kBlock, 0, h->stream>>>(rowptr, indices, nrows, stride, acc);
  CK(cudaMemcpyAsync(host, acc, sizeof host, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (host[1]) *out = (double)host[0] / (double)host[1];
  return 0;
}

// CSR (device, int64 rowptr; A^T entries may carry kEqBit) -> window-major copy (cpppd_banded.cuh).
// vec_len: length of the gathered vector; split: its first element that belongs to an inequality row (A^T) or 0.
// Leaves out->built false (and frees nothing it did not allocate persistently) when the operand does not qualify.
// orig: local -> original ids of the gathered vector (nullptr on one GPU: identity); vec_len and split then count
// ORIGINAL ids (the whole LP), because that is the order in which the caller's rows are sorted.
int build_band(cpppd_solver *h, const int64_t *rowptr, const int32_t *indices, const double *values, int64_t nrows,
               int64_t nnz, int64_t vec_len, int64_t split, bool forced, const int32_t *orig, Band *out) {
  if (nrows == 0 || nnz == 0 || vec_len == 0) return 0;
  cudaStream_t st = h->stream;
  const int64_t target = band_window_elems(h);
  if (int rc = band_locality(h, rowptr, indices, nrows, &out->sectors_per_gather)) return rc;
  // worth it when the gathers of a row-streaming pass cannot stay in L2 on their own: a vector of more than two
  // windows gathered without locality
  if (!forced && !(vec_len > 2 * target && out->sectors_per_gather > 0.6)) return 0;
  BandGeometry geo{split, 1, 1, 0, 0};
  if (split > 0) {
    geo.eq_windows = (int)((split + target - 1) / target);
    geo.eq_elems = (split + geo.eq_windows - 1) / geo.eq_windows;
  }
  int in_windows = 0;
  if (vec_len > split) {
    in_windows = (int)((vec_len - split + target - 1) / target);
    geo.in_elems = (vec_len - split + in_windows - 1) / in_windows;
  }
  geo.windows = geo.eq_windows + in_windows;
  if (geo.windows > 4096) return 0;
  const int64_t rows_pad = (nrows + kBandTile - 1) / kBandTile * kBandTile, ntiles = rows_pad / kBandTile,
                cells = (int64_t)geo.windows * ntiles;
  Scratch tmp(h);
  uint32_t *total = nullptr;
  int *flag = nullptr;
  if (int rc = tmp.get(&total, cells + 1)) return rc;
  if (int rc = tmp.get(&flag, 1)) return rc;
  unsigned char *cnt = static_cast<unsigned char *>(dev_alloc(h, (size_t)geo.windows * rows_pad, false));
  if (!cnt) return fail(h, CPPPD_ERR_NOMEM, "device allocation of the window counts failed");
  CK(cudaMemsetAsync(cnt, 0, (size_t)geo.windows * rows_pad, st));
  CK(cudaMemsetAsync(flag, 0, sizeof(int), st));
  CK(cudaMemsetAsync(total + cells, 0, sizeof(uint32_t), st));
  k_band_count<<<grid_for(nrows), kBlock, 0, st>>>(rowptr, indices, nrows, rows_pad, geo, orig, cnt, flag);
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (bad) {  // a row visits its windows out of order (banding would change its summation order) or is too dense
    dev_free(h, cnt);
    return 0;
  }
  h->owned.push_back(cnt);
  h->device_bytes += (int64_t)geo.windows * rows_pad;
  out->cnt = cnt;
  k_band_tile_totals<<<grid_for(cells), kBlock, 0, st>>>(cnt, cells, total);
  if (int rc = alloc_array(h, &out->tile_base, cells + 1)) return rc;
  if (int rc = exclusive_scan(h, total, out->tile_base, cells + 1)) return rc;
  tmp.release(total);
  if (int rc = alloc_array(h, &out->idx, nnz + 4)) return rc;  // (+ slack: the bulk copies move whole 16-byte units)
  if (int rc = alloc_array(h, &out->val, nnz + 2)) return rc;
  CK(cudaMemsetAsync(out->idx + nnz, 0, 4 * sizeof(int32_t), st));
  CK(cudaMemsetAsync(out->val + nnz, 0, 2 * sizeof(double), st));
  if (int rc = alloc_array(h, &out->carry, nrows)) return rc;
  if (geo.eq_windows && in_windows)
    if (int rc = alloc_array(h, &out->carry_eq, nrows)) return rc;
  k_band_fill<<<grid_for(nrows), kBlock, 0, st>>>(rowptr, indices, values, nrows, rows_pad, geo.windows, cnt, out->tile_base,
                                                    out->idx, out->val);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  out->rows_pad = rows_pad;
  out->geo = geo;
  out->nrows = nrows;
  out->nnz = nnz;
  out->win_bytes = 8 * std::max(geo.eq_windows ? geo.eq_elems : 0, in_windows ? geo.in_elems : 0);
  out->built = true;
  return 0;
}

// Host -> device copy of an array that EVERY rank of a distributed solve holds (the LP is passed whole to each rank):
// a rank copies only its 1/N slice over PCIe and the slices are all-gathered over NVLink — the eight ranks of a node
// otherwise pull N copies of the LP through the host's memory system at the same time (measured on the 4.4 GB Potts
// LP: 54 / 83 / 110-130 ms per rank at N = 2 / 4 / 8).  The destination must hold shared_upload_count() elements.
// CPPPD_FULL_UPLOAD=1 keeps the plain copy.
inline bool shared_upload_on(const cpppd_solver *h) {
  static const bool full = [] { const char *e = getenv("CPPPD_FULL_UPLOAD"); return e && atoi(e) != 0; }();
  return h->world > 1 && h->comm && !full;
}
inline int64_t shared_upload_chunk(const cpppd_solver *h, int64_t bytes) {  // bytes per rank, a multiple of 256
  const int64_t per = (bytes + h->world - 1) / h->world;
  return (per + 255) / 256 * 256;
}
inline int64_t shared_upload_count(const cpppd_solver *h, int64_t count, size_t elem) {
  if (!shared_upload_on(h) || count == 0) return count;
  return (shared_upload_chunk(h, count * (int64_t)elem) * h->world + (int64_t)elem - 1) / (int64_t)elem;
}
int upload_shared(cpppd_solver *h, void *dst, const void *src, int64_t count, size_t elem) {
  if (count == 0) return 0;
  const int64_t bytes = count * (int64_t)elem;
  if (!shared_upload_on(h)) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return 0;
  }
  const int64_t chunk = shared_upload_chunk(h, bytes);
  const int64_t lo = std::min(bytes, chunk * h->rank), hi = std::min(bytes, lo + chunk);
  if (hi > lo)
    CK(cudaMemcpyAsync((char *)dst + lo, (const char *)src + lo, hi - lo, cudaMemcpyHostToDevice, h->stream));
  NK(g_nccl.AllGather((const char *)dst + chunk * h->rank, dst, (size_t)chunk, ncclInt8, h->comm, h->stream));
  return 0;
}

int upload_f64(cpppd_solver *h, double *dst, const double *src, int64_t count) {
  if (count == 0) return 0;
  CK(cudaMemcpyAsync(dst, src, sizeof(double) * count, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

int read_i32(cpppd_solver *h, const int32_t *dev, int32_t *host, int64_t count) {
  CK(cudaMemcpyAsync(host, dev, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// Cut the rows longer than h->long_threshold out of a CSR (device arrays, int64 row pointers): see
// cpppd_long_rows.cuh.  When there are none, *new_rowptr stays nullptr and nothing is allocated; otherwise
// the three new_* arrays (scratch) describe the CSR to hand to build_sell and L owns the long rows.
int split_long_rows(cpppd_solver *h, Scratch &tmp, const int64_t *rowptr, const int32_t *indices, const double *values,
                    int64_t nrows, int virt, int64_t tail_base, LongRows *L, int64_t **new_rowptr, int32_t **new_idx,
                    double **new_val) {
  *new_rowptr = nullptr;
  *new_idx = nullptr;
  *new_val = nullptr;
  L->virt = virt;
  L->tail_base = tail_base;
  if (h->long_threshold < 0 || nrows == 0) return 0;
  cudaStream_t st = h->stream;
  int32_t *flag = nullptr, *slot = nullptr;
  if (int rc = tmp.get(&flag, nrows + 1)) return rc;
  if (int rc = tmp.get(&slot, nrows + 1)) return rc;
  k_long_flag<<<grid_for(nrows + 1), kBlock, 0, st>>>(rowptr, nrows, h->long_threshold, flag);
  if (int rc = exclusive_scan(h, flag, slot, nrows + 1)) return rc;
  int32_t count = 0;
  if (int rc = read_i32(h, slot + nrows, &count, 1)) return rc;
  if (count == 0) {
    tmp.release(flag);
    tmp.release(slot);
    return 0;
  }
  // the long rows themselves: ids and lengths to the host (few), offsets and segments back
  L->count = count;
  int64_t *len_dev = nullptr;
  if (int rc = alloc_array(h, &L->row, count)) return rc;
  if (int rc = tmp.get(&len_dev, count)) return rc;
  k_long_list<<<grid_for(nrows), kBlock, 0, st>>>(rowptr, flag, slot, nrows, L->row, len_dev);
  std::vector<int64_t> len(count), ptr(count + 1, 0), seg_ptr(count + 1, 0);
  CK(cudaMemcpyAsync(len.data(), len_dev, sizeof(int64_t) * count, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  int seg_len = kLongSeg;
  if (const char *env = getenv("CPPPD_LONG_SEG"))
    if (atoi(env) >= 256) seg_len = atoi(env);
  L->seg_len = seg_len;
  for (int32_t r = 0; r < count; ++r) {
    ptr[r + 1] = ptr[r] + len[r];
    seg_ptr[r + 1] = seg_ptr[r] + (len[r] + seg_len - 1) / seg_len;
  }
  L->nnz = ptr[count];
  L->nseg = seg_ptr[count];
  if (L->nseg >= (int64_t)INT32_MAX) return fail(h, CPPPD_ERR_INVALID, "too many long-row segments");
  std::vector<int32_t> seg_row(L->nseg);
  for (int32_t r = 0; r < count; ++r)
    for (int64_t q = seg_ptr[r]; q < seg_ptr[r + 1]; ++q) seg_row[q] = r;
  // launch order: row-major, as stored; CPPPD_LONG_ORDER=1: by position inside the row, then by row (the CTAs resident
  // at one time then gather from about the same window of the vector — measured on the L1-SVM LP: no gain, 0.83 vs
  // 0.81 ms; what the kernel lacked was the software pipeline, profiles/r02p_long_sweep.jsonl)
  std::vector<int32_t> seg_order(L->nseg);
  for (int64_t q = 0; q < L->nseg; ++q) seg_order[q] = (int32_t)q;
  const char *order_env = getenv("CPPPD_LONG_ORDER"), *shape_env = getenv("CPPPD_LONG_SHAPE");
  const bool k_major = order_env && atoi(order_env) != 0;
  L->shape = shape_env ? atoi(shape_env) : 0;
  if (k_major)
    std::stable_sort(seg_order.begin(), seg_order.end(), [&](int32_t a, int32_t b) {
      return a - seg_ptr[seg_row[a]] < b - seg_ptr[seg_row[b]];
    });
  if (int rc = alloc_array(h, &L->ptr, count + 1)) return rc;
  if (int rc = alloc_array(h, &L->seg_ptr, count + 1)) return rc;
  if (int rc = alloc_array(h, &L->seg_row, L->nseg)) return rc;
  if (int rc = alloc_array(h, &L->seg_order, L->nseg)) return rc;
  if (int rc = alloc_array(h, &L->idx, L->nnz)) return rc;
  if (int rc = alloc_array(h, &L->val, L->nnz)) return rc;
  if (int rc = alloc_array(h, &L->partial, 2 * L->nseg)) return rc;
  CK(cudaMemcpyAsync(L->ptr, ptr.data(), sizeof(int64_t) * (count + 1), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(L->seg_ptr, seg_ptr.data(), sizeof(int64_t) * (count + 1), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(L->seg_row, seg_row.data(), sizeof(int32_t) * L->nseg, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(L->seg_order, seg_order.data(), sizeof(int32_t) * L->nseg, cudaMemcpyHostToDevice, st));
  k_long_copy<<<count, kBlock, 0, st>>>(rowptr, indices, values, L->row, L->ptr, L->idx, L->val);
  // the CSR that goes into SELL: short rows as they are, long rows reduced to their virtual entries
  int64_t *newlen = nullptr;
  if (int rc = tmp.get(&newlen, nrows + 1)) return rc;
  if (int rc = tmp.get(new_rowptr, nrows + 1)) return rc;
  k_long_newlen<<<grid_for(nrows + 1), kBlock, 0, st>>>(rowptr, flag, nrows, virt, newlen);
  if (int rc = exclusive_scan(h, newlen, *new_rowptr, nrows + 1)) return rc;
  int64_t new_nnz = 0;
  CK(cudaMemcpyAsync(&new_nnz, *new_rowptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (int rc = tmp.get(new_idx, new_nnz)) return rc;
  if (int rc = tmp.get(new_val, new_nnz)) return rc;
  k_long_rewrite<<<grid_for(nrows), kBlock, 0, st>>>(rowptr, indices, values, flag, slot, *new_rowptr, nrows, virt, tail_base,
                                                     *new_idx, *new_val);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  tmp.release(flag);
  tmp.release(slot);
  tmp.release(len_dev);
  tmp.release(newlen);
  return 0;
}

// k_long_partial in the compiled shape CPPPD_LONG_SHAPE asks for (default 0; all shapes give the same bits)
void launch_long_partial(cpppd_solver *h, const LongRows &L, const double *vec, double power) {
  const int shape = L.shape;
  const int grid = (int)L.nseg;
#define CPPPD_LONG_ARGS L.ptr, L.seg_ptr, L.seg_row, L.seg_order, L.seg_len, L.idx, L.val, vec, power, L.partial
  if (!vec)
    k_long_partial<2, 4, false><<<grid, kBlock, 0, h->stream>>>(CPPPD_LONG_ARGS);
  else if (shape == 1)
    k_long_partial<4, 5, true><<<grid, kBlock, 0, h->stream>>>(CPPPD_LONG_ARGS);
  else if (shape == 2)
    k_long_partial<8, 3, true><<<grid, kBlock, 0, h->stream>>>(CPPPD_LONG_ARGS);
  else
    k_long_partial<6, 4, true><<<grid, kBlock, 0, h->stream>>>(CPPPD_LONG_ARGS);
#undef CPPPD_LONG_ARGS
}

// sums of the long rows of L against `vec`, stored behind the ghosts of `out_vec` (its tail)
int long_pass(cpppd_solver *h, const LongRows &L, const double *vec, double *out_vec) {
  if (L.count == 0) return 0;
  launch_long_partial(h, L, vec, 0.0);
  k_long_finish<<<grid_for(L.count * 32), kBlock, 0, h->stream>>>(L.seg_ptr, L.row, L.count, L.partial,
                                                                 L.virt == 2 ? kLongSumsAT : kLongSumsA, 1, 1,
                                                                 out_vec + L.tail_base);
  return 0;
}

// preconditioner entries of the long rows (the SELL pass saw only their virtual entries)
int long_precond(cpppd_solver *h, const LongRows &L, double power, double *out) {
  if (L.count == 0) return 0;
  const int has_eq = h->m_eq_glob > 0, has_ineq = h->m_ineq_glob > 0;
  launch_long_partial(h, L, nullptr, power);
  k_long_finish<<<grid_for(L.count * 32), kBlock, 0, h->stream>>>(L.seg_ptr, L.row, L.count, L.partial,
                                                                 L.virt == 2 ? kLongPrecondT : kLongPrecondSigma, has_eq,
                                                                 has_ineq, out);
  return 0;
}

int64_t default_granule(int64_t n) {
  int64_t g = 32;
  while (g < (n >> 14)) g *= 2;
  return g;
}

// gather a full-length host vector into this rank's local layout (owned + ghosts)
int upload_local(cpppd_solver *h, Scratch &tmp, const double *host_full, int64_t full_count, const int32_t *map,
                 int64_t local_count, double *dst) {
  if (h->identity_layout) return upload_f64(h, dst, host_full, local_count);
  double *full = nullptr;
  if (int rc = tmp.get(&full, shared_upload_count(h, full_count, sizeof(double)))) return rc;
  if (int rc = upload_shared(h, full, host_full, full_count, sizeof(double))) return rc;
  if (local_count) k_gather_f64<<<grid_for(local_count), kBlock, 0, h->stream>>>(full, map, local_count, dst);
  CK(cudaStreamSynchronize(h->stream));
  tmp.release(full);
  return 0;
}

__global__ void k_sample_bits(const double *__restrict__ values, int64_t nnz, int64_t stride, int64_t count,
                              unsigned long long *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (unsigned long long)__double_as_longlong(values[min(i * stride, nnz - 1)]);
}

// CPPPD_FLAG_VALUE_DICT: if the matrix takes at most 256 distinct values (bit patterns), keep them,
// sorted, in h->dict.  Candidates come from a strided sample; a full pass then proves that every
// entry is covered (otherwise the dictionary is dropped and the generic format is used).
int detect_dictionary(cpppd_solver *h, Scratch &tmp, const double *values, int64_t nnz) {
  cudaStream_t st = h->stream;
  const int64_t count = std::min<int64_t>(nnz, 1 << 20), stride = std::max<int64_t>(1, nnz / count);
  unsigned long long *a = nullptr, *b = nullptr, *uniq = nullptr;
  int *num = nullptr, *flag = nullptr;
  if (int rc = tmp.get(&a, count)) return rc;
  if (int rc = tmp.get(&b, count)) return rc;
  if (int rc = tmp.get(&uniq, count)) return rc;
  if (int rc = tmp.get(&num, 1)) return rc;
  if (int rc = tmp.get(&flag, 1)) return rc;
  k_sample_bits<<<grid_for(count), kBlock, 0, st>>>(values, nnz, stride, count, a);
  size_t b1 = 0, b2 = 0;
  CK(cub::DeviceRadixSort::SortKeys(nullptr, b1, a, b, count, 0, 64, st));
  CK(cub::DeviceSelect::Unique(nullptr, b2, b, uniq, num, count, st));
  char *ws = nullptr;
  if (int rc = tmp.get(&ws, (int64_t)std::max(b1, b2))) return rc;
  CK(cub::DeviceRadixSort::SortKeys(ws, b1, a, b, count, 0, 64, st));
  CK(cub::DeviceSelect::Unique(ws, b2, b, uniq, num, count, st));
  int num_h = 0;
  CK(cudaMemcpyAsync(&num_h, num, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  std::vector<unsigned long long> words;
  if (num_h >= 1 && num_h <= 256) {
    words.resize(num_h);
    CK(cudaMemcpyAsync(words.data(), uniq, sizeof(unsigned long long) * num_h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h->long_threshold >= 0) {  // the virtual entries of long rows (cpppd_long_rows.cuh) carry the value 1.0
      const double one = 1.0;
      unsigned long long one_bits;
      memcpy(&one_bits, &one, sizeof one_bits);
      auto pos = std::lower_bound(words.begin(), words.end(), one_bits);
      if (pos == words.end() || *pos != one_bits) words.insert(pos, one_bits);
      num_h = (int)words.size();
    }
  }
  if (num_h >= 1 && num_h <= 256) {
    if (int rc = alloc_array(h, &h->dict, 256)) return rc;
    CK(cudaMemsetAsync(h->dict, 0, sizeof(unsigned long long) * 256, st));
    CK(cudaMemcpyAsync(h->dict, words.data(), sizeof(unsigned long long) * num_h, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(flag, 0, sizeof(int), st));
    k_dict_check<<<std::min(grid_for(nnz), h->sm_count * 16), kBlock, 0, st>>>(values, nnz, h->dict, num_h, flag);
    int miss = 0;
    CK(cudaMemcpyAsync(&miss, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (miss) h->dict = nullptr; else h->ndict = num_h;
  }
  tmp.release(a); tmp.release(b); tmp.release(uniq); tmp.release(num); tmp.release(flag); tmp.release(ws);
  return 0;
}

int setup_fused(cpppd_solver *h, const std::vector<int64_t> &dst_base_x, const std::vector<int64_t> &dst_base_y);
int tune_kernels(cpppd_solver *h);

// how long a halo wait may spin before it gives up (CPPPD_HALO_TIMEOUT_S, default 60 s, 0 = for ever)
unsigned long long halo_timeout_ns() {
  double seconds = 60.0;
  if (const char *env = getenv("CPPPD_HALO_TIMEOUT_S")) seconds = atof(env);
  return seconds > 0 ? (unsigned long long)(seconds * 1e9) : 0ull;
}

// after a synchronisation: did a halo wait give up?
int check_halo_timeout(cpppd_solver *h) {
  if (!h->p2p.active) return 0;
  unsigned int flag = 0;
  CK(cudaMemcpyAsync(&flag, &h->p2p.state->timed_out, sizeof flag, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (flag) return fail(h, CPPPD_ERR_COMM, "a halo wait timed out: a neighbour rank did not deliver its halo (CPPPD_HALO_TIMEOUT_S)");
  return 0;
}

// Unmap the peers' buffers and free this rank's (communicator tear-down, or before the pool grows).  CUDA leaves
// cudaFree of an exported region undefined while an importer still has it open: `between` runs after all imports of
// this rank are closed and before its own exports are freed — the pool rebuild passes a collective there, so that no
// rank frees what a slower peer still maps (at tear-down there is nobody left to wait for).
template <typename Between>
int pool_release(PeerPool &pl, Between between) {
  if (!pl.valid) return between();
  for (int t = 0; t < kMaxWorld; ++t) {
    for (void *p : {(void *)pl.peer_x[t], (void *)pl.peer_y[t], (void *)pl.peer_flags[t]})
      if (p) cudaIpcCloseMemHandle(p);
    pl.peer_x[t] = pl.peer_y[t] = nullptr;
    pl.peer_flags[t] = nullptr;
  }
  const int rc = between();
  for (void *p : {(void *)pl.xbar, (void *)pl.y, (void *)pl.flags, (void *)pl.state})
    if (p) cudaFree(p);
  cudaGetLastError();
  pl = PeerPool();
  return rc;
}
inline void pool_release(PeerPool &pl) { pool_release(pl, [] { return 0; }); }

// The two peer-written vectors of this solve, from the communicator's pool when every rank can take them there
// (collective: every rank of the solve calls this with its own lengths).  The pool is (re)built — cudaMalloc, export,
// all-gather of the handles, cudaIpcOpenMemHandle of every peer — when some rank needs more than it holds; a second
// solver alive on the same communicator, or a solver with a communicator of its own, gets private buffers as before
// (h->pooled stays false, setup() allocates).  CPPPD_NO_P2P_POOL=1 switches the pool off.
struct PoolHandles {
  cudaIpcMemHandle_t xbar, y, flags;
};
int pool_acquire(cpppd_solver *h, int64_t x_len, int64_t y_len) {
  static const bool off = [] { const char *e = getenv("CPPPD_NO_P2P_POOL"); return e && atoi(e) != 0; }();
  h->pooled = false;
  if (off || !h->shared) return 0;
  const int N = h->world, me = h->rank;
  PeerPool &pl = h->shared->pool;
  cudaStream_t st = h->stream;
  const size_t need_x = sizeof(double) * (size_t)std::max<int64_t>(x_len, 1), need_y = sizeof(double) * (size_t)std::max<int64_t>(y_len, 1);
  Scratch tmp(h);
  signed char *votes = nullptr;  // [0]: ranks whose pool is busy, [1]: ranks that cannot reuse theirs
  if (int rc = tmp.get(&votes, 2)) return rc;
  signed char mine[2] = {(signed char)(pl.busy ? 1 : 0),
                         (signed char)((pl.valid && pl.world == N && pl.cap_x >= need_x && pl.cap_y >= need_y) ? 0 : 1)};
  CK(cudaMemcpyAsync(votes, mine, 2, cudaMemcpyHostToDevice, st));
  NK(g_nccl.AllReduce(votes, votes, 2, ncclInt8, ncclSum, h->comm, st));
  CK(cudaMemcpyAsync(mine, votes, 2, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (mine[0]) {  // another solver of this communicator is alive: private buffers
    if (getenv("CPPPD_POOL_TRACE")) fprintf(stderr, "[cpppd pool rank %d] busy: private buffers\n", me);
    return 0;
  }
  if (mine[1]) {
    // every rank agreed to rebuild (the all-reduce above is the rendezvous: no solver is alive, nobody stores into
    // the old buffers any more); a second one separates "every rank has closed its imports" from "free the exports"
    if (int rc = pool_release(pl, [&]() -> int {
          NK(g_nccl.AllReduce(votes, votes, 1, ncclInt8, ncclSum, h->comm, st));
          CK(cudaStreamSynchronize(st));
          return 0;
        }))
      return rc;
    size_t round = (size_t)2 << 20;  // capacity granule (CPPPD_POOL_GRANULE: tests make it small to see the pool grow)
    if (const char *e = getenv("CPPPD_POOL_GRANULE"))
      if (atoll(e) >= 8) round = (size_t)atoll(e);
    pl.cap_x = (need_x + round - 1) / round * round;
    pl.cap_y = (need_y + round - 1) / round * round;
    pl.world = N;
    pl.valid = true;  // (from here on pool_release undoes whatever exists)
    CK(cudaMalloc(&pl.xbar, pl.cap_x));
    CK(cudaMalloc(&pl.y, pl.cap_y));
    CK(cudaMalloc(&pl.flags, sizeof(unsigned long long) * 2 * kMaxWorld));
    CK(cudaMalloc(&pl.state, sizeof(SyncState)));
    PoolHandles own;
    memset(&own, 0, sizeof own);
    CK(cudaIpcGetMemHandle(&own.xbar, pl.xbar));
    CK(cudaIpcGetMemHandle(&own.y, pl.y));
    CK(cudaIpcGetMemHandle(&own.flags, pl.flags));
    char *send = nullptr, *recv = nullptr;
    if (int rc = tmp.get(&send, (int64_t)sizeof(PoolHandles))) return rc;
    if (int rc = tmp.get(&recv, (int64_t)sizeof(PoolHandles) * N)) return rc;
    CK(cudaMemcpyAsync(send, &own, sizeof own, cudaMemcpyHostToDevice, st));
    NK(g_nccl.AllGather(send, recv, sizeof(PoolHandles), ncclInt8, h->comm, st));
    std::vector<PoolHandles> all(N);
    CK(cudaMemcpyAsync(all.data(), recv, sizeof(PoolHandles) * N, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int failed = 0;
    for (int t = 0; t < N && !failed; ++t) {
      if (t == me) continue;
      void *px = nullptr, *py = nullptr, *pf = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&px, all[t].xbar, cudaIpcMemLazyEnablePeerAccess);
      if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&py, all[t].y, cudaIpcMemLazyEnablePeerAccess);
      if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&pf, all[t].flags, cudaIpcMemLazyEnablePeerAccess);
      pl.peer_x[t] = (double *)px;
      pl.peer_y[t] = (double *)py;
      pl.peer_flags[t] = (unsigned long long *)pf;
      if (e != cudaSuccess) {
        cudaGetLastError();
        failed = 1;
        fail(h, CPPPD_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s (use CPPPD_FLAG_NO_P2P for the NCCL path)", t,
             cudaGetErrorString(e));
      }
    }
    // success is agreed on collectively (a rank that returned alone would leave its peers in a later collective)
    signed char any = (signed char)failed;
    CK(cudaMemcpyAsync(votes, &any, 1, cudaMemcpyHostToDevice, st));
    NK(g_nccl.AllReduce(votes, votes, 1, ncclInt8, ncclSum, h->comm, st));
    CK(cudaMemcpyAsync(&any, votes, 1, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (any) {
      pool_release(pl);
      if (!failed) fail(h, CPPPD_ERR_COMM, "a peer rank could not map this rank's halo buffers (cudaIpcOpenMemHandle); "
                                           "use CPPPD_FLAG_NO_P2P for the NCCL path");
      return CPPPD_ERR_COMM;
    }
  }
  if (getenv("CPPPD_POOL_TRACE"))
    fprintf(stderr, "[cpppd pool rank %d] %s (%zu + %zu bytes needed, %zu + %zu held)\n", me, mine[1] ? "built" : "reused",
            need_x, need_y, pl.cap_x, pl.cap_y);
  pl.busy = true;
  h->pooled = true;
  h->xbar = pl.xbar;
  h->y = pl.y;
  return 0;
}

// What every rank publishes so that its neighbours can write into its ghost slots.
struct PeerRecord {
  cudaIpcMemHandle_t xbar, y, flags;
  int64_t owned_x, owned_y;
  int64_t recv_off_x[kMaxWorld], recv_off_y[kMaxWorld];
};

int setup_p2p(cpppd_solver *h) {
  const int N = h->world, me = h->rank;
  P2P &pp = h->p2p;
  cudaStream_t st = h->stream;
  if (h->pooled) {
    pp.flags = h->shared->pool.flags;
    pp.state = h->shared->pool.state;
  } else {
    CK(cudaMalloc(&pp.flags, sizeof(unsigned long long) * 2 * N));
    pp.own.push_back(pp.flags);
    CK(cudaMalloc(&pp.state, sizeof(SyncState)));
    pp.own.push_back(pp.state);
  }
  CK(cudaMemsetAsync(pp.flags, 0, sizeof(unsigned long long) * 2 * N, st));
  {
    SyncState init;
    memset(&init, 0, sizeof init);
    init.timeout_ns = halo_timeout_ns();
    CK(cudaMemcpyAsync(pp.state, &init, sizeof init, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
  }
  PeerRecord mine;
  memset(&mine, 0, sizeof mine);
  if (!h->pooled) {
    CK(cudaIpcGetMemHandle(&mine.xbar, h->xbar));
    CK(cudaIpcGetMemHandle(&mine.y, h->y));
    CK(cudaIpcGetMemHandle(&mine.flags, pp.flags));
  }
  mine.owned_x = h->hx.owned;
  mine.owned_y = h->hy.owned;
  for (int t = 0; t < N; ++t) {
    mine.recv_off_x[t] = h->hx.recv_off[t];
    mine.recv_off_y[t] = h->hy.recv_off[t];
  }
  // all-gather the records (NCCL, setup only)
  Scratch tmp(h);
  char *send = nullptr, *recv = nullptr;
  if (int rc = tmp.get(&send, (int64_t)sizeof(PeerRecord))) return rc;
  if (int rc = tmp.get(&recv, (int64_t)sizeof(PeerRecord) * N)) return rc;
  CK(cudaMemcpyAsync(send, &mine, sizeof mine, cudaMemcpyHostToDevice, st));
  NK(g_nccl.AllGather(send, recv, sizeof(PeerRecord), ncclInt8, h->comm, st));
  std::vector<PeerRecord> all(N);
  CK(cudaMemcpyAsync(all.data(), recv, sizeof(PeerRecord) * N, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  // map the neighbours' vectors
  int ipc_failed = 0;
  memset(pp.ptrs, 0, sizeof pp.ptrs);
  for (int t = 0; t < N; ++t) {
    if (t == me) continue;
    const bool nb = h->hx.send_count[t] || h->hx.recv_count[t] || h->hy.send_count[t] || h->hy.recv_count[t];
    if (!nb) continue;
    if (h->pooled) {  // mapped when the pool was built
      pp.ptrs[0].vec[t] = h->shared->pool.peer_x[t];
      pp.ptrs[1].vec[t] = h->shared->pool.peer_y[t];
      pp.ptrs[0].flags[t] = pp.ptrs[1].flags[t] = h->shared->pool.peer_flags[t];
      continue;
    }
    void *px = nullptr, *py = nullptr, *pf = nullptr;
    cudaError_t e1 = cudaIpcOpenMemHandle(&px, all[t].xbar, cudaIpcMemLazyEnablePeerAccess);
    cudaError_t e2 = e1 == cudaSuccess ? cudaIpcOpenMemHandle(&py, all[t].y, cudaIpcMemLazyEnablePeerAccess) : e1;
    cudaError_t e3 = e2 == cudaSuccess ? cudaIpcOpenMemHandle(&pf, all[t].flags, cudaIpcMemLazyEnablePeerAccess) : e2;
    if (e3 != cudaSuccess) {
      cudaGetLastError();
      ipc_failed = 1;
      fail(h, CPPPD_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s (use CPPPD_FLAG_NO_P2P for the NCCL path)", t,
           cudaGetErrorString(e3));
      for (void *p : {px, py}) if (p) pp.opened.push_back(p);
      break;
    }
    pp.opened.insert(pp.opened.end(), {px, py, pf});
    pp.ptrs[0].vec[t] = (double *)px;
    pp.ptrs[1].vec[t] = (double *)py;
    pp.ptrs[0].flags[t] = pp.ptrs[1].flags[t] = (unsigned long long *)pf;
  }
  {  // success is agreed on collectively: a rank that returned alone would leave its peers in the barrier below for ever
    signed char any_failed = (signed char)ipc_failed;  // (world <= 64: the sum of the flags fits a byte)
    CK(cudaMemcpyAsync(send, &any_failed, 1, cudaMemcpyHostToDevice, st));
    NK(g_nccl.AllReduce(send, send, 1, ncclInt8, ncclSum, h->comm, st));
    CK(cudaMemcpyAsync(&any_failed, send, 1, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (any_failed) {
      if (!ipc_failed) fail(h, CPPPD_ERR_COMM, "a peer rank could not map this rank's halo buffers (cudaIpcOpenMemHandle); "
                                              "use CPPPD_FLAG_NO_P2P for the NCCL path");
      return CPPPD_ERR_COMM;
    }
  }
  // per-entry destinations of the two send lists
  std::vector<int64_t> dst_base[2] = {std::vector<int64_t>(N, 0), std::vector<int64_t>(N, 0)};
  for (int kind = 0; kind < 2; ++kind) {
    Halo &H = kind ? h->hy : h->hx;
    std::vector<int32_t> peer(H.send_total);
    std::vector<int64_t> dst(H.send_total);
    for (int t = 0; t < N; ++t) {
      if (H.send_count[t]) pp.send_mask[kind] |= 1ull << t;
      if (H.recv_count[t]) pp.recv_mask[kind] |= 1ull << t;
      const int64_t base = (kind ? all[t].owned_y : all[t].owned_x) + (kind ? all[t].recv_off_y[me] : all[t].recv_off_x[me]);
      dst_base[kind][t] = base;
      for (int64_t k = 0; k < H.send_count[t]; ++k) {
        peer[H.send_off[t] + k] = t;
        dst[H.send_off[t] + k] = base + k;
      }
    }
    pp.dense[kind] = h->dense_halo && H.owned > 0;
    for (int t = 0; t < N; ++t) {
      pp.dense_dst[kind][t] = dst_base[kind][t];
      if (t != me && H.send_count[t] != H.owned) pp.dense[kind] = false;
    }
    if (int rc = alloc_array(h, &pp.push_peer[kind], H.send_total)) return rc;
    if (int rc = alloc_array(h, &pp.push_dst[kind], H.send_total)) return rc;
    if (H.send_total) {
      CK(cudaMemcpy(pp.push_peer[kind], peer.data(), sizeof(int32_t) * H.send_total, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(pp.push_dst[kind], dst.data(), sizeof(int64_t) * H.send_total, cudaMemcpyHostToDevice));
    }
  }
  // (the fused kernels wait for the halo themselves; the long-row pre-passes would read the ghosts too early)
  if ((h->flags & CPPPD_FLAG_FUSED_HALO) && h->longA.count == 0 && h->longAT.count == 0 && !h->bandA.built && !h->bandAT.built)
    if (int rc = setup_fused(h, dst_base[0], dst_base[1])) return rc;
  // nobody may push before every rank has initialised its vectors and flags
  NK(g_nccl.AllReduce(send, send, 1, ncclInt8, ncclSum, h->comm, st));
  CK(cudaStreamSynchronize(st));
  pp.active = true;
  return 0;
}

// xbar (kind 0) / y (kind 1) halo over peer memory: push mine, then wait for the neighbours'.
int exchange_p2p(cpppd_solver *h, int kind) {
  P2P &pp = h->p2p;
  Halo &H = kind ? h->hy : h->hx;
  const double *vec = kind ? h->y : h->xbar;
  // CPPPD_SPLIT_HALO_WAIT=1: the wait as a launch of its own (k_wait) instead of the tail of the push kernel
  static const bool split_wait = [] { const char *e = getenv("CPPPD_SPLIT_HALO_WAIT"); return e && atoi(e) != 0; }();
  const unsigned long long *wait_flags = split_wait ? nullptr : pp.flags;
  bool pushed = false;
  if (pp.dense[kind]) {
    DenseDst D;
    memcpy(D.base, pp.dense_dst[kind], sizeof D.base);
    // (a few CTAs per SM: every thread streams its share of the owned entries to all peers)
    const int grid = (int)std::min<int64_t>(grid_for(H.owned), (int64_t)h->sm_count * 16);
    k_push_dense<<<grid, kBlock, 0, h->stream>>>(vec, H.owned, pp.ptrs[kind], D, kind, h->world, h->rank, pp.send_mask[kind],
                                                 wait_flags, pp.recv_mask[kind], pp.state);
    pushed = true;
  } else if (H.send_total) {
    k_push<<<grid_for(H.send_total), kBlock, 0, h->stream>>>(vec, H.send_idx, pp.push_dst[kind], pp.push_peer[kind],
                                                            H.send_total, pp.ptrs[kind], kind, h->world, h->rank,
                                                            pp.send_mask[kind], wait_flags, pp.recv_mask[kind], pp.state);
    pushed = true;
  }
  if (pp.recv_mask[kind] && (!pushed || split_wait))
    k_wait<<<1, kMaxWorld, 0, h->stream>>>(pp.flags, kind, h->world, pp.recv_mask[kind], pp.state);
  return 0;
}

// device tables for the halo exchange fused into k_primal / k_dual
int setup_fused(cpppd_solver *h, const std::vector<int64_t> &dst_base_x, const std::vector<int64_t> &dst_base_y) {
  P2P &pp = h->p2p;
  cudaStream_t st = h->stream;
  const int N = h->world;
  Scratch tmp(h);
  for (int kind = 0; kind < 2; ++kind) {  // kind 0: k_primal over A^T produces xbar; kind 1: k_dual over A produces y
    const Sell &S = kind ? h->A : h->AT;
    const Halo &out = kind ? h->hy : h->hx;
    const int64_t owned_other = kind ? h->hx.owned : h->hy.owned;  // entry indices >= this are ghosts
    const int64_t ns = S.nslices;
    unsigned int *role32 = nullptr;
    unsigned char *role = nullptr;
    if (int rc = tmp.get(&role32, ns + 1)) return rc;
    if (int rc = alloc_array(h, &role, ns + 1)) return rc;
    CK(cudaMemsetAsync(role32, 0, sizeof(unsigned int) * (ns + 1), st));
    const int64_t items = std::max<int64_t>(ns * 32, out.send_total);
    if (items) k_slice_roles<<<grid_for(items), kBlock, 0, st>>>(view(S), owned_other, out.send_idx, out.send_total, role32);
    k_narrow_roles<<<grid_for(ns + 1), kBlock, 0, st>>>(role32, ns + 1, role);
    FusedComm cm;
    memset(&cm, 0, sizeof cm);
    cm.role = role;
    cm.send_idx = out.send_idx;
    const std::vector<int64_t> &dst_base = kind ? dst_base_y : dst_base_x;
    for (int t = 0; t < N; ++t) {
      cm.off[t] = out.send_off[t];
      cm.cnt[t] = out.send_count[t];
      cm.dst_base[t] = dst_base[t];
      cm.peer_vec[t] = pp.ptrs[kind].vec[t];
      cm.peer_flags[t] = pp.ptrs[kind].flags[t];
    }
    cm.my_flags = pp.flags;
    cm.st = pp.state;
    cm.world = N;
    cm.me = h->rank;
    cm.kind_out = kind;
    cm.kind_in = 1 - kind;
    cm.send_mask = pp.send_mask[kind];
    cm.recv_mask_in = pp.recv_mask[1 - kind];
    if (int rc = alloc_array(h, &pp.fused[kind], 1)) return rc;
    CK(cudaMemcpyAsync(pp.fused[kind], &cm, sizeof cm, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    tmp.release(role32);
  }
  pp.use_fused = true;
  return 0;
}

#ifdef __CUDACC__
inline auto cluster_kernel(bool strict, bool one_pass) {
  if (strict) return one_pass ? k_cluster_iterate<1, true> : k_cluster_iterate<1, false>;
  return one_pass ? k_cluster_iterate<0, true> : k_cluster_iterate<0, false>;
}
#endif

// Can one thread-block cluster carry this LP (cpppd_cluster.cuh)?  Fills h->cluster.
int plan_cluster(cpppd_solver *h) {
  h->cluster = ClusterPlan();
  static const bool off = [] { const char *e = getenv("CPPPD_NO_CLUSTER"); return e && atoi(e) != 0; }();
  if (off || h->A.nslices == 0 || h->AT.nslices == 0) return 0;
  if (h->A.padded > kClusterMaxEntries || h->AT.padded > kClusterMaxEntries) return 0;
  std::vector<int64_t> sp[2];
  const Sell *ops[2] = {&h->AT, &h->A};
  for (int k = 0; k < 2; ++k) {
    const Sell &S = *ops[k];
    sp[k].resize(S.nslices + 1);
    if (S.uniform_width >= 0) {
      for (int64_t q = 0; q <= S.nslices; ++q) sp[k][q] = q * S.uniform_width * 32;
    } else {
      CK(cudaMemcpyAsync(sp[k].data(), S.slice_ptr, sizeof(int64_t) * (S.nslices + 1), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
    }
  }
#ifdef __CUDACC__
  int max_smem = 0;
  CK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  auto kernel = cluster_kernel(false, false);
#else
  const int max_smem = 232448;  // (CPU emulation: the 227 KB per CTA of sm_100)
#endif
  for (int ctas : {kClusterMaxCtas, 8}) {
    ClusterPlan P;
    P.ctas = ctas;
    P.spc_at = (int)((h->AT.nslices + ctas - 1) / ctas);
    P.spc_a = (int)((h->A.nslices + ctas - 1) / ctas);
    int64_t ent[2] = {0, 0};
    for (int k = 0; k < 2; ++k) {
      const int64_t ns = ops[k]->nslices, spc = k ? P.spc_a : P.spc_at;
      for (int r = 0; r < ctas; ++r) {
        const int64_t lo = std::min<int64_t>(r * spc, ns), hi = std::min<int64_t>(lo + spc, ns);
        int64_t padded = 0;  // every slice is staged with its width rounded up to whole chunks of kClC entries
        for (int64_t q = lo; q < hi; ++q) padded += ((sp[k][q + 1] - sp[k][q]) / 32 + kClC - 1) / kClC * kClC * 32;
        ent[k] = std::max(ent[k], padded);
      }
    }
    P.ent_at = (int)ent[0];
    P.ent_a = (int)ent[1];
    P.smem = cluster_smem_bytes(P.spc_at, P.spc_a, P.ent_at, P.ent_a);
    if ((int64_t)P.smem > max_smem) continue;
#ifdef __CUDACC__
    bool attr_ok = true;
    for (auto fn : {cluster_kernel(false, false), cluster_kernel(false, true), cluster_kernel(true, false), cluster_kernel(true, true)})
      attr_ok = attr_ok && cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) == cudaSuccess &&
                (ctas <= 8 || cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess);
    if (!attr_ok) {
      cudaGetLastError();
      continue;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kClusterBlock);
    cfg.dynamicSmemBytes = P.smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = ctas;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg) != cudaSuccess || clusters < 1) {
      cudaGetLastError();
      continue;
    }
#endif
    P.on = true;
    h->cluster = P;
    break;
  }
  return 0;
}

// CPPPD_SETUP_TIMING=1: wall-clock of the phases of setup() on stderr (after a stream synchronisation each)
struct PhaseTimer {
  cudaStream_t st;
  int rank;
  bool on;
  std::chrono::steady_clock::time_point t0;
  PhaseTimer(cudaStream_t s, int r) : st(s), rank(r), on(getenv("CPPPD_SETUP_TIMING") && atoi(getenv("CPPPD_SETUP_TIMING"))),
                                      t0(std::chrono::steady_clock::now()) {}
  void mark(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(st);
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[cpppd setup rank %d] %-28s %8.1f ms\n", rank, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

int setup(cpppd_solver *h, const cpppd_problem *P) {
  const int64_t n = h->n_glob, m = h->m_glob, nnz = h->nnz_glob, m_eq = h->m_eq_glob;
  const int N = h->world, me = h->rank;
  cudaStream_t st = h->stream;
  PhaseTimer phase(st, me);
  Scratch tmp(h);
  // ---- CSR of the whole A on the device (temporary; every rank analyses the same pattern)
  int64_t *rowptr = nullptr;
  int32_t *indices = nullptr;
  double *values = nullptr;
  if (int rc = tmp.get(&rowptr, shared_upload_count(h, m + 1, sizeof(int64_t)))) return rc;
  if (int rc = tmp.get(&indices, shared_upload_count(h, nnz, sizeof(int32_t)))) return rc;
  if (int rc = tmp.get(&values, shared_upload_count(h, nnz, sizeof(double)))) return rc;
  if (P->indptr_bits == 64) {
    if (int rc = upload_shared(h, rowptr, P->indptr, m + 1, sizeof(int64_t))) return rc;
  } else {
    int32_t *tmp32 = nullptr;
    if (int rc = tmp.get(&tmp32, shared_upload_count(h, m + 1, sizeof(int32_t)))) return rc;
    if (int rc = upload_shared(h, tmp32, P->indptr, m + 1, sizeof(int32_t))) return rc;
    k_widen_indptr<<<grid_for(m + 1), kBlock, 0, st>>>(tmp32, rowptr, m + 1);
    CK(cudaStreamSynchronize(st));
    tmp.release(tmp32);
  }
  if (int rc = upload_shared(h, indices, P->indices, nnz, sizeof(int32_t))) return rc;
  if (int rc = upload_shared(h, values, P->values, nnz, sizeof(double))) return rc;
  {  // validation on the device: monotone row pointers, column indices in range
    int *flag = nullptr;
    if (int rc = tmp.get(&flag, 1)) return rc;
    CK(cudaMemsetAsync(flag, 0, sizeof(int), st));
    int64_t items = std::max<int64_t>(m, std::min<int64_t>(nnz, (int64_t)h->sm_count * 64 * kBlock));
    if (items) k_validate<<<grid_for(items), kBlock, 0, st>>>(rowptr, m, indices, nnz, n, flag);
    int host_flag = 0;
    CK(cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (host_flag & 1) return fail(h, CPPPD_ERR_INVALID, "indptr is not non-decreasing");
    if (host_flag & 2) return fail(h, CPPPD_ERR_INVALID, "column index outside [0, n)");
  }
  phase.mark("upload + validation");
  if ((h->flags & CPPPD_FLAG_VALUE_DICT) && nnz)
    if (int rc = detect_dictionary(h, tmp, values, nnz)) return rc;
  uint32_t *row_of = nullptr, *entry_id = nullptr;
  if (int rc = tmp.get(&row_of, nnz)) return rc;
  if (int rc = tmp.get(&entry_id, nnz)) return rc;
  if (nnz) k_row_of_entry<<<grid_for(nnz), kBlock, 0, st>>>(rowptr, m, nnz, row_of, entry_id);

  bool reorder = N > 1 || (h->flags & CPPPD_FLAG_REORDER);
  // Small LPs that may run in one thread-block cluster (cpppd_cluster.cuh) are renumbered by locality: a CTA of the
  // cluster then owns rows AND the columns they touch, most gathers stay in its own shared memory, and the cluster
  // network — which bounds the kernel otherwise — carries a fraction of them (Potts 50x50: 404 000 -> 794 000 it/s).
  {
    static const bool no_cluster = [] { const char *e = getenv("CPPPD_NO_CLUSTER"); return e && atoi(e) != 0; }();
    const bool forced = h->variant_request != 0 || (getenv("CPPPD_KERNEL_VARIANT") && atoi(getenv("CPPPD_KERNEL_VARIANT")));
    if (N == 1 && nnz > 0 && nnz <= kClusterMaxEntries && !no_cluster && !forced &&
        !(h->flags & (CPPPD_FLAG_NO_REORDER | CPPPD_FLAG_NO_TINY_PERSISTENT | CPPPD_FLAG_BANDED)))
      reorder = true;
  }
  // Banded operands (cpppd_banded.cuh) keep the caller's numbering: the window order along a row is what keeps
  // the sums bit-exact.  Candidates: forced by flag, or a pattern without locality over vectors of several windows.
  // With more than one GPU they need the balanced split in original order (decided below): windows are ranges of
  // original ids, and the local layout must keep such a range in a few contiguous pieces.
  const bool band_ok = (N > 1 || !reorder) && nnz > 0 && !(h->flags & (CPPPD_FLAG_NO_BANDED | CPPPD_FLAG_VALUE_DICT));
  const bool band_forced = band_ok && (h->flags & CPPPD_FLAG_BANDED);
  bool band_candidate = band_forced;
  if (band_ok && !band_forced && std::max(n, m) > 2 * band_window_elems(h)) {
    double spg = 0.0;
    if (int rc = band_locality(h, rowptr, indices, m, &spg)) return rc;
    band_candidate = spg > 0.6;
  }
  if (!reorder && nnz && !band_candidate && !(h->flags & CPPPD_FLAG_NO_REORDER)) {
    // keep the caller's numbering unless SELL-32 would pad it by more than 15 %: then renumber
    // (rows / columns of equal length are grouped inside locality buckets)
    int32_t *col_len = nullptr;
    unsigned long long *total = nullptr, total_h = 0;
    if (int rc = tmp.get(&col_len, n)) return rc;
    if (int rc = tmp.get(&total, 1)) return rc;
    CK(cudaMemsetAsync(col_len, 0, sizeof(int32_t) * std::max<int64_t>(n, 1), st));
    CK(cudaMemsetAsync(total, 0, sizeof(unsigned long long), st));
    k_col_len<<<grid_for(nnz), kBlock, 0, st>>>(indices, nnz, col_len);
    if (m) k_padded_total<<<grid_for(((m + 31) / 32) * 32), kBlock, 0, st>>>(rowptr, nullptr, m, total);
    if (n) k_padded_total<<<grid_for(((n + 31) / 32) * 32), kBlock, 0, st>>>(nullptr, col_len, n, total);
    CK(cudaMemcpyAsync(&total_h, total, sizeof total_h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    tmp.release(col_len);
    tmp.release(total);
    reorder = (double)total_h > 1.15 * 2.0 * (double)nnz;
  }
  h->identity_layout = !reorder;
  int32_t rs = 0, re = (int32_t)m, cs = 0, ce = (int32_t)n;
  int64_t n_ghost = 0, m_ghost = 0;
  uint32_t *row_order = nullptr, *col_order = nullptr;
  int32_t *row_pos = nullptr, *col_pos = nullptr, *gcol_scan = nullptr, *grow_scan = nullptr;
  std::vector<int32_t> row_start(N + 1, 0), col_start(N + 1, 0), eq_count(N, 0);
  h->hx = Halo();
  h->hy = Halo();
  for (Halo *H : {&h->hx, &h->hy}) {
    H->send_count.assign(N, 0);
    H->send_off.assign(N, 0);
    H->recv_count.assign(N, 0);
    H->recv_off.assign(N, 0);
  }

  if (reorder) {
    phase.mark("row of entry + layout choice");
    // ---- locality keys -> buckets -> owners (oracle/partition_oracle.py restates this block)
    const int64_t G = h->granule > 0 ? h->granule : default_granule(n);
    h->granule = G;
    const int64_t nb = n / G + 2;
    int32_t *row_key = nullptr, *col_key = nullptr, *col_len = nullptr, *owner_dev = nullptr, *counts = nullptr;
    unsigned long long *work = nullptr;  // [0, nb): entries of the rows of a bucket, [nb, 2 nb): of its columns
    if (int rc = tmp.get(&row_key, m)) return rc;
    if (int rc = tmp.get(&col_key, n)) return rc;
    if (int rc = tmp.get(&col_len, n)) return rc;
    if (int rc = tmp.get(&work, 2 * nb)) return rc;
    if (int rc = tmp.get(&owner_dev, nb)) return rc;
    if (int rc = tmp.get(&counts, 3 * (int64_t)N)) return rc;
    if (m) k_row_key<<<grid_for(m), kBlock, 0, st>>>(rowptr, indices, m, (int32_t)n, row_key);
    if (n) k_fill_i32<<<grid_for(n), kBlock, 0, st>>>(col_key, n, (int32_t)n);
    CK(cudaMemsetAsync(col_len, 0, sizeof(int32_t) * std::max<int64_t>(n, 1), st));
    if (nnz) k_col_key<<<grid_for(nnz), kBlock, 0, st>>>(indices, row_of, nnz, row_key, col_key, col_len);
    CK(cudaMemsetAsync(work, 0, sizeof(unsigned long long) * 2 * nb, st));
    if (m) k_bucket_work<<<grid_for(m), kBlock, 0, st>>>(row_key, rowptr, nullptr, m, (int32_t)G, work);
    if (n) k_bucket_work<<<grid_for(n), kBlock, 0, st>>>(col_key, nullptr, col_len, n, (int32_t)G, work + nb);
    std::vector<unsigned long long> work_h(2 * nb);
    CK(cudaMemcpyAsync(work_h.data(), work, sizeof(unsigned long long) * 2 * nb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<int32_t> owner_h(nb, 0);
    unsigned long long total = 0, before = 0;
    for (auto w : work_h) total += w;
    std::vector<unsigned long long> row_share(N, 0), col_share(N, 0);
    for (int64_t q = 0; q < nb; ++q) {
      owner_h[q] = total ? (int32_t)std::min<unsigned long long>(N - 1, (unsigned __int128)before * N / total) : 0;
      before += work_h[q] + work_h[nb + q];
      row_share[owner_h[q]] += work_h[q];
      col_share[owner_h[q]] += work_h[nb + q];
    }
    CK(cudaMemcpyAsync(owner_dev, owner_h.data(), sizeof(int32_t) * nb, cudaMemcpyHostToDevice, st));
    phase.mark("  keys + bucket work");
    // A pattern without locality leaves some rank with far more than its share of the row entries or of the
    // column entries (every kernel then waits for that rank): fall back to the balanced split, where the owner
    // of a row / column follows from the entries in front of it (oracle/partition_oracle.py restates this).
    unsigned long long worst = 0;
    for (int r = 0; r < N; ++r) worst = std::max(worst, std::max(row_share[r], col_share[r]));
    h->balanced_split = nnz > 0 && (unsigned __int128)2 * N * worst > (unsigned __int128)3 * (unsigned long long)nnz;
    // banded operands on several GPUs: only with the balanced split, whose rows / columns then keep their original
    // order inside a rank (no locality buckets, no length sorting: the banded kernels do not pad)
    if (N > 1 && !h->balanced_split) band_candidate = false;
    const int keep_order = N > 1 && band_candidate ? 1 : 0;
    int64_t *col_prefix = nullptr;
    if (h->balanced_split) {
      int64_t *col_len64 = nullptr;
      if (int rc = tmp.get(&col_len64, n + 1)) return rc;
      if (int rc = tmp.get(&col_prefix, n + 1)) return rc;
      CK(cudaMemsetAsync(col_len64, 0, sizeof(int64_t) * (n + 1), st));
      if (n) k_widen_indptr<<<grid_for(n), kBlock, 0, st>>>(col_len, col_len64, n);
      if (int rc = exclusive_scan(h, col_len64, col_prefix, n + 1)) return rc;
      tmp.release(col_len64);
    }
    // ---- local orders: rows by (owner, is_ineq, bucket, id), columns by (owner, bucket, id)
    uint64_t *rk_a = nullptr, *rk_b = nullptr, *ck_a = nullptr, *ck_b = nullptr;
    uint32_t *ro_a = nullptr, *ro_b = nullptr, *co_a = nullptr, *co_b = nullptr;
    if (int rc = tmp.get(&rk_a, m)) return rc;
    if (int rc = tmp.get(&rk_b, m)) return rc;
    if (int rc = tmp.get(&ro_a, m)) return rc;
    if (int rc = tmp.get(&ro_b, m)) return rc;
    if (int rc = tmp.get(&ck_a, n)) return rc;
    if (int rc = tmp.get(&ck_b, n)) return rc;
    if (int rc = tmp.get(&co_a, n)) return rc;
    if (int rc = tmp.get(&co_b, n)) return rc;
    CK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 3 * N, st));
    if (m) k_sort_keys<<<grid_for(m), kBlock, 0, st>>>(row_key, m, (int32_t)G, owner_dev, m_eq, 1, rowptr, nullptr,
                                                       h->balanced_split ? rowptr : nullptr, nnz, N, keep_order, rk_a, ro_a, counts, counts + N);
    if (n) k_sort_keys<<<grid_for(n), kBlock, 0, st>>>(col_key, n, (int32_t)G, owner_dev, 0, 0, nullptr, col_len,
                                                       col_prefix, nnz, N, keep_order, ck_a, co_a, counts + 2 * N, nullptr);
    phase.mark("  sort keys");
    const int end_bit = 44 + bits_for((uint64_t)2 * N + 1);
    cub::DoubleBuffer<uint64_t> rk(rk_a, rk_b), ck(ck_a, ck_b);
    cub::DoubleBuffer<uint32_t> rov(ro_a, ro_b), cov(co_a, co_b);
    if (int rc = sort_pairs(h, rk, rov, m, end_bit)) return rc;
    if (int rc = sort_pairs(h, ck, cov, n, end_bit)) return rc;
    row_order = rov.Current();
    col_order = cov.Current();
    phase.mark("  two radix sorts");
    std::vector<int32_t> counts_h(3 * N);
    if (int rc = read_i32(h, counts, counts_h.data(), 3 * N)) return rc;
    for (int r = 0; r < N; ++r) {
      row_start[r + 1] = row_start[r] + counts_h[r];
      eq_count[r] = counts_h[N + r];
      col_start[r + 1] = col_start[r] + counts_h[2 * N + r];
    }
    if (int rc = tmp.get(&row_pos, m)) return rc;
    if (int rc = tmp.get(&col_pos, n)) return rc;
    if (m) k_invert<<<grid_for(m), kBlock, 0, st>>>(row_order, m, row_pos);
    if (n) k_invert<<<grid_for(n), kBlock, 0, st>>>(col_order, n, col_pos);
    rs = row_start[me]; re = row_start[me + 1]; cs = col_start[me]; ce = col_start[me + 1];
    tmp.release(rk_a); tmp.release(rk_b); tmp.release(ck_a); tmp.release(ck_b);
    if (row_order == ro_a) tmp.release(ro_b); else tmp.release(ro_a);
    if (col_order == co_a) tmp.release(co_b); else tmp.release(co_a);
    tmp.release(row_key); tmp.release(col_key); tmp.release(col_len); tmp.release(work); tmp.release(col_prefix);
    phase.mark("partition + local orders");
    // ---- ghosts of this rank
    int32_t *gcol_flag = nullptr, *grow_flag = nullptr;
    if (int rc = tmp.get(&gcol_flag, n + 1)) return rc;
    if (int rc = tmp.get(&grow_flag, m + 1)) return rc;
    if (int rc = tmp.get(&gcol_scan, n + 1)) return rc;
    if (int rc = tmp.get(&grow_scan, m + 1)) return rc;
    CK(cudaMemsetAsync(gcol_flag, 0, sizeof(int32_t) * (n + 1), st));
    CK(cudaMemsetAsync(grow_flag, 0, sizeof(int32_t) * (m + 1), st));
    // Patterns without locality on several GPUs (balanced split in original order): a rank's rows touch nearly every
    // column and its columns nearly every row (random LP, 8 GPUs: 86 % / 63 %).  Every foreign column / row is then
    // kept as a ghost: the halo of a peer is the peer's whole slice in its own order — one contiguous, index-free
    // copy per peer (k_push_dense) instead of an indexed push of almost the same bytes.
    h->dense_halo = keep_order && !(h->flags & CPPPD_FLAG_NO_DENSE_HALO);
    if (N > 1 && h->dense_halo) {
      if (n) k_flag_outside<<<grid_for(n), kBlock, 0, st>>>(gcol_flag, n, cs, ce);
      if (m) k_flag_outside<<<grid_for(m), kBlock, 0, st>>>(grow_flag, m, rs, re);
    } else if (nnz && N > 1) {
      k_mark_ghosts<<<grid_for(nnz), kBlock, 0, st>>>(indices, row_of, nnz, row_pos, col_pos, rs, re, cs, ce, gcol_flag, grow_flag);
    }
    if (int rc = exclusive_scan(h, gcol_flag, gcol_scan, n + 1)) return rc;
    if (int rc = exclusive_scan(h, grow_flag, grow_scan, m + 1)) return rc;
    std::vector<int32_t> cb(N + 1), rb(N + 1);
    for (int r = 0; r <= N; ++r) {
      if (int rc = read_i32(h, gcol_scan + col_start[r], &cb[r], 1)) return rc;
      if (int rc = read_i32(h, grow_scan + row_start[r], &rb[r], 1)) return rc;
    }
    n_ghost = cb[N];
    m_ghost = rb[N];
    for (int r = 0; r < N; ++r) {
      h->hx.recv_count[r] = cb[r + 1] - cb[r];
      h->hx.recv_off[r] = cb[r];
      h->hy.recv_count[r] = rb[r + 1] - rb[r];
      h->hy.recv_off[r] = rb[r];
    }
    // ---- local -> original id maps (owned, then ghosts in exchange order)
    const int64_t nloc = ce - cs, mloc = re - rs;
    if (int rc = alloc_array(h, &h->col_old, nloc + n_ghost)) return rc;
    if (int rc = alloc_array(h, &h->row_old, mloc + m_ghost)) return rc;
    if (nloc) k_copy_u32_i32<<<grid_for(nloc), kBlock, 0, st>>>(col_order + cs, nloc, h->col_old);
    if (mloc) k_copy_u32_i32<<<grid_for(mloc), kBlock, 0, st>>>(row_order + rs, mloc, h->row_old);
    if (n_ghost) k_compact<<<grid_for(n), kBlock, 0, st>>>(gcol_flag, gcol_scan, n, col_order, 0, h->col_old + nloc);
    if (m_ghost) k_compact<<<grid_for(m), kBlock, 0, st>>>(grow_flag, grow_scan, m, row_order, 0, h->row_old + mloc);
    CK(cudaStreamSynchronize(st));
    tmp.release(gcol_flag);
    tmp.release(grow_flag);
    phase.mark("ghosts + id maps");
    // ---- what to send to every peer
    if (N > 1) {
      int32_t *sx_flag = nullptr, *sy_flag = nullptr, *sx_scan = nullptr, *sy_scan = nullptr;
      if (int rc = tmp.get(&sx_flag, nloc + 1)) return rc;
      if (int rc = tmp.get(&sy_flag, mloc + 1)) return rc;
      if (int rc = tmp.get(&sx_scan, nloc + 1)) return rc;
      if (int rc = tmp.get(&sy_scan, mloc + 1)) return rc;
      std::vector<std::vector<int32_t>> sx_lists(N), sy_lists(N);
      int32_t *list_dev = nullptr;
      if (int rc = tmp.get(&list_dev, std::max(nloc, mloc) + 1)) return rc;
      // one pass over the entries marks, per owned column / row, the set of peers that need it
      unsigned long long *sx_mask = nullptr, *sy_mask = nullptr;
      if (int rc = tmp.get(&sx_mask, nloc + 1)) return rc;
      if (int rc = tmp.get(&sy_mask, mloc + 1)) return rc;
      CK(cudaMemsetAsync(sx_mask, 0, sizeof(unsigned long long) * (nloc + 1), st));
      CK(cudaMemsetAsync(sy_mask, 0, sizeof(unsigned long long) * (mloc + 1), st));
      RankStarts starts;
      memset(&starts, 0, sizeof starts);
      for (int r = 0; r <= N; ++r) {
        starts.row[r] = row_start[r];
        starts.col[r] = col_start[r];
      }
      if (h->dense_halo) {
        const unsigned long long peers = (N >= 64 ? ~0ull : ((1ull << N) - 1ull)) & ~(1ull << me);
        if (nloc) k_fill_u64<<<grid_for(nloc), kBlock, 0, st>>>(sx_mask, nloc, peers);
        if (mloc) k_fill_u64<<<grid_for(mloc), kBlock, 0, st>>>(sy_mask, mloc, peers);
      } else if (nnz) {
        k_mark_send_masks<<<grid_for(nnz), kBlock, 0, st>>>(indices, row_of, nnz, row_pos, col_pos, rs, re, cs, ce, starts, N,
                                                           sx_mask, sy_mask);
      }
      for (int t = 0; t < N; ++t) {
        if (t == me) continue;
        CK(cudaMemsetAsync(sx_flag + nloc, 0, sizeof(int32_t), st));
        CK(cudaMemsetAsync(sy_flag + mloc, 0, sizeof(int32_t), st));
        if (nloc) k_flag_from_mask<<<grid_for(nloc), kBlock, 0, st>>>(sx_mask, nloc, t, sx_flag);
        if (mloc) k_flag_from_mask<<<grid_for(mloc), kBlock, 0, st>>>(sy_mask, mloc, t, sy_flag);
        if (int rc = exclusive_scan(h, sx_flag, sx_scan, nloc + 1)) return rc;
        if (int rc = exclusive_scan(h, sy_flag, sy_scan, mloc + 1)) return rc;
        int32_t cx = 0, cy = 0;
        if (int rc = read_i32(h, sx_scan + nloc, &cx, 1)) return rc;
        if (int rc = read_i32(h, sy_scan + mloc, &cy, 1)) return rc;
        if (cx) {
          k_compact<<<grid_for(nloc), kBlock, 0, st>>>(sx_flag, sx_scan, nloc, nullptr, 0, list_dev);
          sx_lists[t].resize(cx);
          if (int rc = read_i32(h, list_dev, sx_lists[t].data(), cx)) return rc;
        }
        if (cy) {
          k_compact<<<grid_for(mloc), kBlock, 0, st>>>(sy_flag, sy_scan, mloc, nullptr, 0, list_dev);
          sy_lists[t].resize(cy);
          if (int rc = read_i32(h, list_dev, sy_lists[t].data(), cy)) return rc;
        }
      }
      for (int pass = 0; pass < 2; ++pass) {
        Halo &H = pass ? h->hy : h->hx;
        auto &lists = pass ? sy_lists : sx_lists;
        std::vector<int32_t> flat;
        for (int t = 0; t < N; ++t) {
          H.send_off[t] = (int64_t)flat.size();
          H.send_count[t] = (int64_t)lists[t].size();
          flat.insert(flat.end(), lists[t].begin(), lists[t].end());
        }
        H.send_total = (int64_t)flat.size();
        if (int rc = alloc_array(h, &H.send_idx, H.send_total)) return rc;
        if (int rc = alloc_array(h, &H.send_buf, H.send_total)) return rc;
        if (H.send_total) CK(cudaMemcpy(H.send_idx, flat.data(), sizeof(int32_t) * flat.size(), cudaMemcpyHostToDevice));
      }
      tmp.release(sx_flag); tmp.release(sy_flag); tmp.release(sx_scan); tmp.release(sy_scan); tmp.release(list_dev);
      tmp.release(sx_mask); tmp.release(sy_mask);
    }
  }
  const int64_t nloc = ce - cs, mloc = re - rs;
  h->n = nloc;
  h->m = mloc;
  h->m_eq = reorder ? eq_count[me] : m_eq;
  h->hx.owned = nloc;
  h->hx.ghost = n_ghost;
  h->hy.owned = mloc;
  h->hy.ghost = m_ghost;

  if (h->dict) {  // an entry word must hold index + code below the eq / pad bits
    // gather indices reach behind the ghosts when long rows exist: at most nnz / threshold of them, with one
    // (A) or two (A^T) tail slots each
    const int64_t max_long = h->long_threshold < 0 ? 0 : nnz / std::max<int64_t>(h->long_threshold, 1) + 1;
    const int idx_bits = bits_for((uint64_t)std::max<int64_t>(
        std::max(nloc + n_ghost + max_long, mloc + m_ghost + 2 * max_long), 2) - 1);
    const int code_bits = bits_for((uint64_t)std::max(h->ndict, 2) - 1);
    if (idx_bits + code_bits <= 30) {
      for (Sell *S : {&h->A, &h->AT}) {
        S->dict = reinterpret_cast<const double *>(h->dict);
        S->idx_bits = idx_bits;
        S->ndict = h->ndict;
      }
    } else {
      h->dict = nullptr;  // (stays allocated, simply unused)
      h->ndict = 0;
    }
  }
  phase.mark("send lists");
  // ---- this rank's rows of A -> SELL-32
  int64_t *s_rowptr = nullptr;  // CSR without the long rows (see split_long_rows), when there are any
  int32_t *s_idx = nullptr;
  double *s_val = nullptr;
  if (!reorder) {
    h->nnz_rows = nnz;
    if (int rc = split_long_rows(h, tmp, rowptr, indices, values, m, 1, nloc + n_ghost, &h->longA, &s_rowptr, &s_idx, &s_val))
      return rc;
    if (s_rowptr) {
      if (int rc = build_sell(h, s_rowptr, s_idx, s_val, m, &h->A)) return rc;
      tmp.release(s_rowptr); tmp.release(s_idx); tmp.release(s_val);
    } else {
      if (int rc = build_sell(h, rowptr, indices, values, m, &h->A)) return rc;
      if (band_candidate)
        if (int rc = build_band(h, rowptr, indices, values, m, nnz, n, 0, band_forced, nullptr, &h->bandA)) return rc;
    }
  } else {
    int64_t *len = nullptr, *lrowptr = nullptr;
    if (int rc = tmp.get(&len, mloc + 1)) return rc;
    if (int rc = tmp.get(&lrowptr, mloc + 1)) return rc;
    k_local_row_len<<<grid_for(mloc + 1), kBlock, 0, st>>>(row_order, rs, mloc, rowptr, len);
    if (int rc = exclusive_scan(h, len, lrowptr, mloc + 1)) return rc;
    int64_t lnnz = 0;
    CK(cudaMemcpyAsync(&lnnz, lrowptr + mloc, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h->nnz_rows = lnnz;
    int32_t *lidx = nullptr;
    double *lval = nullptr;
    if (int rc = tmp.get(&lidx, lnnz)) return rc;
    if (int rc = tmp.get(&lval, lnnz)) return rc;
    if (mloc) k_local_rows_fill<<<grid_for(mloc), kBlock, 0, st>>>(row_order, rs, mloc, rowptr, indices, values, col_pos,
                                                                  cs, ce, gcol_scan, lrowptr, lidx, lval);
    if (int rc = split_long_rows(h, tmp, lrowptr, lidx, lval, mloc, 1, nloc + n_ghost, &h->longA, &s_rowptr, &s_idx, &s_val))
      return rc;
    if (s_rowptr) {
      if (int rc = build_sell(h, s_rowptr, s_idx, s_val, mloc, &h->A)) return rc;
      tmp.release(s_rowptr); tmp.release(s_idx); tmp.release(s_val);
    } else {
      if (int rc = build_sell(h, lrowptr, lidx, lval, mloc, &h->A)) return rc;
      if (band_candidate)
        if (int rc = build_band(h, lrowptr, lidx, lval, mloc, lnnz, n, 0, band_forced, h->col_old, &h->bandA)) return rc;
    }
    tmp.release(len); tmp.release(lrowptr); tmp.release(lidx); tmp.release(lval);
  }
  phase.mark("rows of A -> SELL (+ band)");
  // ---- this rank's columns of A as rows of A^T.  A stable radix sort of the entries (taken in CSR
  //      order) by column keeps, inside each column, the original row order — exactly the
  //      accumulation order of scipy's csc_matvec.
  {
    uint32_t *keys_a = nullptr, *keys_b = nullptr, *ids_b = nullptr;
    if (int rc = tmp.get(&keys_a, nnz)) return rc;
    if (int rc = tmp.get(&keys_b, nnz)) return rc;
    if (int rc = tmp.get(&ids_b, nnz)) return rc;
    if (nnz) {
      if (reorder) k_entry_col_pos<<<grid_for(nnz), kBlock, 0, st>>>(indices, col_pos, nnz, keys_a);
      else CK(cudaMemcpyAsync(keys_a, indices, sizeof(uint32_t) * nnz, cudaMemcpyDeviceToDevice, st));
    }
    cub::DoubleBuffer<uint32_t> keys(keys_a, keys_b), ids(entry_id, ids_b);
    if (int rc = sort_pairs(h, keys, ids, nnz, bits_for((uint64_t)std::max<int64_t>(n, 1)))) return rc;
    int64_t *lcolptr = nullptr;
    if (int rc = tmp.get(&lcolptr, nloc + 1)) return rc;
    k_lower_bounds<<<grid_for(nloc + 1), kBlock, 0, st>>>(keys.Current(), nnz, cs, nloc, lcolptr);
    int64_t first = 0, last = 0;
    CK(cudaMemcpyAsync(&first, lcolptr, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&last, lcolptr + nloc, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int64_t lnnz = last - first;
    h->nnz_cols = lnnz;
    if (first) k_subtract_base<<<grid_for(nloc + 1), kBlock, 0, st>>>(lcolptr, nloc + 1, first);
    const uint32_t *sorted_ids = ids.Current();
    if (keys.Current() == keys_a) tmp.release(keys_b); else tmp.release(keys_a);
    int32_t *t_idx = nullptr;
    double *t_val = nullptr;
    if (int rc = tmp.get(&t_idx, lnnz)) return rc;
    if (int rc = tmp.get(&t_val, lnnz)) return rc;
    if (lnnz) {
      if (reorder) {
        k_local_cols_fill<<<grid_for(lnnz), kBlock, 0, st>>>(sorted_ids, row_of, values, first, lnnz, row_pos, rs, re,
                                                             grow_scan, m_eq, t_idx, t_val);
      } else {
        // identity layout: row_pos / grow_scan do not exist; local row == original row
        k_local_cols_fill<<<grid_for(lnnz), kBlock, 0, st>>>(sorted_ids, row_of, values, first, lnnz, nullptr, 0,
                                                             (int32_t)m, nullptr, m_eq, t_idx, t_val);
      }
    }
    CK(cudaStreamSynchronize(st));
    tmp.release(keys_a); tmp.release(keys_b); tmp.release(ids_b); tmp.release(entry_id); tmp.release(row_of);
    tmp.release(indices); tmp.release(values); tmp.release(rowptr);
    if (int rc = split_long_rows(h, tmp, lcolptr, t_idx, t_val, nloc, 2, mloc + m_ghost, &h->longAT, &s_rowptr, &s_idx, &s_val))
      return rc;
    if (s_rowptr) {
      if (int rc = build_sell(h, s_rowptr, s_idx, s_val, nloc, &h->AT)) return rc;
      tmp.release(s_rowptr); tmp.release(s_idx); tmp.release(s_val);
    } else {
      if (int rc = build_sell(h, lcolptr, t_idx, t_val, nloc, &h->AT)) return rc;
      if (band_candidate)
        if (int rc = build_band(h, lcolptr, t_idx, t_val, nloc, lnnz, m, m_eq, band_forced, reorder ? h->row_old : nullptr,
                                &h->bandAT))
          return rc;
    }
    tmp.release(lcolptr); tmp.release(t_idx); tmp.release(t_val);
  }
  phase.mark("columns of A -> SELL (+ band)");
  // ---- vectors in local layout
  for (double **v : {&h->c, &h->T, &h->lb, &h->ub, &h->best})
    if (int rc = alloc_array(h, v, nloc)) return rc;
  const bool want_p2p = N > 1 && !(h->flags & CPPPD_FLAG_NO_P2P);
  // layout of a gathered vector: [owned | ghosts | sums of the long rows that gather from it]
  const int64_t x_len = nloc + n_ghost + h->longA.count, y_len = mloc + m_ghost + 2 * h->longAT.count;
  h->x_len = x_len;
  h->y_len = y_len;
  for (double **v : {&h->x, &h->dbuf})
    if (int rc = alloc_array(h, v, x_len)) return rc;
  for (double **v : {&h->b, &h->sigma})
    if (int rc = alloc_array(h, v, mloc)) return rc;
  if (want_p2p)    // ... from the communicator's pool of mapped peer memory when there is one
    if (int rc = pool_acquire(h, x_len, y_len)) return rc;
  if (h->pooled) {
    h->device_bytes += 8 * (x_len + y_len);
  } else if (want_p2p) {  // the two vectors with peer-written ghost tails: plain cudaMalloc, exportable by IPC
    CK(cudaMalloc(&h->xbar, sizeof(double) * std::max<int64_t>(x_len, 1)));
    h->p2p.own.push_back(h->xbar);
    CK(cudaMalloc(&h->y, sizeof(double) * std::max<int64_t>(y_len, 1)));
    h->p2p.own.push_back(h->y);
    h->device_bytes += 8 * (x_len + y_len);
  } else {
    if (int rc = alloc_array(h, &h->xbar, x_len)) return rc;
    if (int rc = alloc_array(h, &h->y, y_len)) return rc;
  }
  if (int rc = upload_local(h, tmp, P->c, n, h->col_old, nloc, h->c)) return rc;
  if (int rc = upload_local(h, tmp, P->lb, n, h->col_old, nloc, h->lb)) return rc;
  if (int rc = upload_local(h, tmp, P->ub, n, h->col_old, nloc, h->ub)) return rc;
  if (int rc = upload_local(h, tmp, P->b, m, h->row_old, mloc, h->b)) return rc;
  CK(cudaMemsetAsync(h->x, 0, sizeof(double) * std::max<int64_t>(x_len, 1), st));
  if (P->x0)
    if (int rc = upload_local(h, tmp, P->x0, n, h->col_old, nloc + n_ghost, h->x)) return rc;
  CK(cudaMemcpyAsync(h->xbar, h->x, sizeof(double) * x_len, cudaMemcpyDeviceToDevice, st));  // x3 = x (:190)
  CK(cudaMemsetAsync(h->dbuf, 0, sizeof(double) * std::max<int64_t>(x_len, 1), st));
  CK(cudaMemsetAsync(h->y, 0, sizeof(double) * std::max<int64_t>(y_len, 1), st));           // :166,:177
  // ---- preconditioners (:122-179): complete columns / rows are local, so no exchange is needed
  const int has_eq = h->m_eq_glob > 0, has_ineq = h->m_ineq_glob > 0;
  if (h->AT.nslices)
    k_precond_cols<<<grid_for(h->AT.nslices * 32), kBlock, 0, st>>>(view(h->AT), nloc, has_eq, has_ineq, 2.0 - h->alpha, h->T);
  if (h->A.nslices)
    k_precond_rows<<<grid_for(h->A.nslices * 32), kBlock, 0, st>>>(view(h->A), mloc, h->alpha, h->sigma);
  if (int rc = long_precond(h, h->longAT, 2.0 - h->alpha, h->T)) return rc;
  if (int rc = long_precond(h, h->longA, h->alpha, h->sigma)) return rc;
  h->vc = Vec{h->c, 0};
  h->vT = Vec{h->T, 0};
  h->vlb = Vec{h->lb, 0};
  h->vub = Vec{h->ub, 0};
  h->vb = Vec{h->b, 0};
  h->vsigma = Vec{h->sigma, 0};
  if (h->flags & CPPPD_FLAG_CONST_VECTORS) {
    struct { Vec *v; double *p; int64_t count; int bit; } cand[] = {
        {&h->vb, h->b, mloc, 0}, {&h->vsigma, h->sigma, mloc, 1}, {&h->vlb, h->lb, nloc, 2},
        {&h->vub, h->ub, nloc, 3}, {&h->vc, h->c, nloc, 4},      {&h->vT, h->T, nloc, 5}};
    int *flag = nullptr;
    if (int rc = tmp.get(&flag, 1)) return rc;
    for (auto &cd : cand) {
      if (cd.count == 0) continue;
      int host_flag = 0;
      CK(cudaMemsetAsync(flag, 0, sizeof(int), st));
      k_not_constant<<<std::min(grid_for(cd.count), h->sm_count * 8), kBlock, 0, st>>>(cd.p, cd.count, flag);
      CK(cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
      double first = 0;
      CK(cudaMemcpyAsync(&first, cd.p, sizeof(double), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (!host_flag) {
        *cd.v = Vec{nullptr, first};
        h->const_mask |= 1 << cd.bit;
      }
    }
  }
  phase.mark("vectors + preconditioners");
  // ---- stats plumbing
  h->stat_blocks_c = (int)std::max<int64_t>(1, std::min<int64_t>(grid_for(nloc), (int64_t)h->sm_count * 8));
  h->stat_blocks_r = (int)std::max<int64_t>(1, std::min<int64_t>(grid_for(h->A.nslices * 32), (int64_t)h->sm_count * 8));
  if (int rc = alloc_array(h, &h->colpart, (int64_t)h->stat_blocks_c * kColQ)) return rc;
  if (int rc = alloc_array(h, &h->rowpart, (int64_t)h->stat_blocks_r * kRowQ)) return rc;
  if (int rc = alloc_array(h, &h->stat_local, kStatQ)) return rc;
  if (int rc = alloc_array(h, &h->stat_all, (int64_t)kStatQ * N)) return rc;
  if (int rc = alloc_array(h, &h->stats_dev, 1)) return rc;
  k_init_stats<<<1, 1, 0, st>>>(h->stats_dev);
  CK(cudaMallocHost(&h->stats_host, sizeof(cpppd_stats)));
  memset(h->stats_host, 0, sizeof(cpppd_stats));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  // a single CTA can carry a whole iteration when both operands fit the L1 of one SM (see k_tiny_iterate)
  const bool variant_forced = h->variant_request != 0 || (getenv("CPPPD_KERNEL_VARIANT") && atoi(getenv("CPPPD_KERNEL_VARIANT")));
  h->tiny = !(h->flags & CPPPD_FLAG_NO_TINY_PERSISTENT) && !variant_forced && N == 1 && h->longA.count == 0 && h->longAT.count == 0 &&
            !h->bandA.built && !h->bandAT.built &&
            std::max(nloc, mloc) <= 4096 && h->A.padded + h->AT.padded <= 16384;
  // ... a cluster of up to 16 CTAs when they fit the shared memory of 16 SMs (see k_cluster_iterate).
  // The cluster kernel is the faster one whenever it can be launched (SC105, 103 x 105: 801 000 it/s in one CTA,
  // 1 116 000 in the cluster; Potts 24x24: 131 000 / 961 000 — profiles/r02_kernels.md): the one-CTA kernel remains for
  // devices / LPs where the cluster launch is refused and for an explicit CPPPD_FLAG_TINY_PERSISTENT.
  const bool was_tiny = h->tiny;
  if (!(h->flags & CPPPD_FLAG_TINY_PERSISTENT)) h->tiny = false;
  if (!h->tiny && !(h->flags & CPPPD_FLAG_NO_TINY_PERSISTENT) && !variant_forced && N == 1 && h->longA.count == 0 &&
      h->longAT.count == 0 && !h->bandA.built && !h->bandAT.built)
    if (int rc = plan_cluster(h)) return rc;
  if (was_tiny && !h->cluster.on) h->tiny = true;
  phase.mark("stats plumbing");
  if (int rc = tune_kernels(h)) return rc;
  phase.mark("kernel timing");
  if (want_p2p)
    if (int rc = setup_p2p(h)) return rc;
  phase.mark("peer-memory mapping");
  return 0;
}

// Refresh the ghost part of a distributed vector: every rank sends the owned entries its peers
// need and receives its ghosts straight into vec[owned ...].
int exchange(cpppd_solver *h, double *vec, Halo &H) {
  if (h->world == 1) return 0;
  if (H.send_total) k_pack<<<grid_for(H.send_total), kBlock, 0, h->stream>>>(vec, H.send_idx, H.send_total, H.send_buf);
  NK(g_nccl.GroupStart());
  for (int t = 0; t < h->world; ++t) {
    if (H.send_count[t]) NK(g_nccl.Send(H.send_buf + H.send_off[t], (size_t)H.send_count[t], ncclFloat64, t, h->comm, h->stream));
    if (H.recv_count[t]) NK(g_nccl.Recv(vec + H.owned + H.recv_off[t], (size_t)H.recv_count[t], ncclFloat64, t, h->comm, h->stream));
  }
  NK(g_nccl.GroupEnd());
  return 0;
}

// Primal half + xbar halo.  world > 1: k_push / k_wait kernels over peer memory (default), the kernel itself
// waits, pushes and signals (CPPPD_FLAG_FUSED_HALO), or NCCL send/recv (CPPPD_FLAG_NO_P2P).
// one launch per window of the gathered vector (cpppd_banded.cuh)
constexpr int kBandVariant = -2;  // `variant` argument of launch_primal / launch_dual: time the banded kernels
// the piece of the gathered vector that window w of B covers (one GPU: local ids are original ids); CPPPD_BAND_PREFETCH=0
// switches the L2 prefetch off
BandPrefetch band_window_range(const cpppd_solver *h, const Band &B, const double *vec, int64_t vec_len, int w) {
  static const bool enabled = [] { const char *e = getenv("CPPPD_BAND_PREFETCH"); return !e || atoi(e) != 0; }();
  if (!enabled || h->world > 1) return BandPrefetch{nullptr, 0};
  const BandGeometry &g = B.geo;
  const int64_t lo = w < g.eq_windows ? (int64_t)w * g.eq_elems : g.split + (int64_t)(w - g.eq_windows) * g.in_elems;
  const int64_t end = w < g.eq_windows ? g.split : vec_len;
  const int64_t hi = std::min(end, lo + (w < g.eq_windows ? g.eq_elems : g.in_elems));
  const int64_t first = lo & ~(int64_t)1;  // 16-byte aligned start
  return BandPrefetch{reinterpret_cast<const char *>(vec + first), std::max<int64_t>(0, (hi - first) * 8)};
}

int launch_primal_band(cpppd_solver *h, bool write_d) {
  const Band &B = h->bandAT;
  const int has_eq = h->m_eq_glob > 0, has_ineq = h->m_ineq_glob > 0;
  const int64_t ntiles = B.rows_pad / kBandTile;
  const int grid = (int)((ntiles + kBandWarps - 1) / kBandWarps);  // a warp per tile of 128 rows
  for (int w = 0; w < B.geo.windows; ++w) {
    const bool eq = w < B.geo.eq_windows;
    const int mode = ((w == 0 || w == B.geo.eq_windows) ? kBandStart : 0) | (eq ? kBandEq : 0);
    const unsigned char *cnt = B.cnt + (int64_t)w * B.rows_pad;
    const uint32_t *base = B.tile_base + (int64_t)w * ntiles;
    double *ceq = B.carry_eq ? B.carry_eq : B.carry;  // one kind of rows only: a single carry serves it
    const BandPrefetch pf = band_window_range(h, B, h->y, h->m, w);
    const bool last = w == B.geo.windows - 1;
    if (!last && kBandShapes[B.shape].staged) {
      PrimalStagedFn sfn = primal_staged_kernel(B.shape);
      sfn<<<grid, kBlock, 0, h->stream>>>(cnt, base, B.idx, B.val, h->y, eq ? ceq : B.carry, mode & kBandStart, h->n, ntiles, pf);
      continue;
    }
    PrimalBandFn fn = primal_band_kernel(last, write_d, B.shape);
    fn<<<grid, kBlock, 0, h->stream>>>(
        cnt, base, B.idx, B.val, h->y, ceq, B.carry, mode, h->vc, h->vT, h->vlb, h->vub, h->x, h->xbar, h->dbuf, h->n, ntiles,
        has_eq, has_ineq, h->theta, h->one_plus_theta, pf);
  }
  return 0;
}

int launch_dual_band(cpppd_solver *h) {
  const Band &B = h->bandA;
  const int W = B.geo.windows;
  const int64_t ntiles = B.rows_pad / kBandTile;
  const int grid = (int)((ntiles + kBandWarps - 1) / kBandWarps);
  for (int w = 0; w < W; ++w) {
    const unsigned char *cnt = B.cnt + (int64_t)w * B.rows_pad;
    const uint32_t *base = B.tile_base + (int64_t)w * ntiles;
    const bool first = w == 0, last = w == W - 1;
    const BandPrefetch pf = band_window_range(h, B, h->xbar, h->n, w);
    if (!last && kBandShapes[B.shape].staged) {
      DualStagedFn sfn = dual_staged_kernel(first, B.shape);
      sfn<<<grid, kBlock, 0, h->stream>>>(cnt, base, B.idx, B.val, h->xbar, B.carry, h->m, ntiles, pf);
      continue;
    }
    DualBandFn fn = dual_band_kernel(first, last, B.shape);
    fn<<<grid, kBlock, 0, h->stream>>>(cnt, base, B.idx, B.val, h->xbar, B.carry, h->vb, h->vsigma, h->y, h->m, ntiles, h->m_eq, pf);
  }
  return 0;
}

int launch_primal(cpppd_solver *h, bool write_d, int variant = -1) {
  P2P &pp = h->p2p;
  if (variant == kBandVariant || (variant == -1 && h->bandAT.in_use)) {
    if (int rc = launch_primal_band(h, write_d)) return rc;
    if (variant != -1) return 0;
    return pp.active ? exchange_p2p(h, 0) : exchange(h, h->xbar, h->hx);
  }
  const FusedComm *cm = pp.use_fused ? pp.fused[0] : nullptr;
  if (int rc = long_pass(h, h->longAT, h->y, h->y)) return rc;  // long columns: their A^T y into the tail of y
  if (h->AT.nslices) {
    const bool dict = h->AT.dict != nullptr;
    const int has_eq = h->m_eq_glob > 0, has_ineq = h->m_ineq_glob > 0;
    const int v = cm ? 0 : (variant >= 0 ? variant : h->primal_variant);
    PrimalFn fn = cm ? primal_kernel_fused(write_d, dict) : primal_kernel(write_d, dict, v);
    const int64_t warps = (h->AT.nslices + kVariants[v].rows - 1) / kVariants[v].rows;  // a warp walks `rows` slices
    int grid = grid_for(warps * 32);
    if (!cm && kVariants[v].persist) grid = std::min(grid, h->sm_count * kVariants[v].persist);
    fn<<<grid, kBlock, 0, h->stream>>>(view(h->AT), h->y, h->vc, h->vT, h->vlb, h->vub, h->x, h->xbar,
                                                              h->dbuf, h->n, has_eq, has_ineq, h->theta, h->one_plus_theta, cm);
  }
  if (cm || variant >= 0) return 0;
  return pp.active ? exchange_p2p(h, 0) : exchange(h, h->xbar, h->hx);
}

int launch_dual(cpppd_solver *h, int variant = -1) {
  P2P &pp = h->p2p;
  if (variant == kBandVariant || (variant == -1 && h->bandA.in_use)) {
    if (int rc = launch_dual_band(h)) return rc;
    if (variant != -1) return 0;
    return pp.active ? exchange_p2p(h, 1) : exchange(h, h->y, h->hy);
  }
  const FusedComm *cm = pp.use_fused ? pp.fused[1] : nullptr;
  if (int rc = long_pass(h, h->longA, h->xbar, h->xbar)) return rc;  // long rows: their A xbar into the tail of xbar
  if (h->A.nslices) {
    const bool dict = h->A.dict != nullptr;
    const int v = cm ? 0 : (variant >= 0 ? variant : h->dual_variant);
    DualFn fn = cm ? dual_kernel_fused(dict) : dual_kernel(dict, v);
    const int64_t warps = (h->A.nslices + kVariants[v].rows - 1) / kVariants[v].rows;
    int grid = grid_for(warps * 32);
    if (!cm && kVariants[v].persist) grid = std::min(grid, h->sm_count * kVariants[v].persist);
    fn<<<grid, kBlock, 0, h->stream>>>(view(h->A), h->xbar, h->vb, h->vsigma, h->y, h->m, h->m_eq, cm);
  }
  if (cm || variant >= 0) return 0;
  return pp.active ? exchange_p2p(h, 1) : exchange(h, h->y, h->hy);
}

// Process-wide memory of tune_kernels(): the timings depend on the device and on the shape of the two SELL
// operands (slices, stored entries, uniform widths, long rows, storage variant, numbering), not on the values.
using TuneKey = std::array<int64_t, 17>;
struct TuneChoice {
  int primal = 0, dual = 0;
  float ms[2][CPPPD_KERNEL_VARIANTS] = {};
  bool band_primal = false, band_dual = false;
  float band_ms[2] = {0.f, 0.f};
  int band_shape[2] = {0, 0};
  float band_shape_ms[2][8] = {};
};
std::map<TuneKey, TuneChoice> g_tune_cache;
std::mutex g_tune_mutex;

TuneKey tune_key(const cpppd_solver *h) {
  return TuneKey{h->device,          h->n,          h->m,           h->A.nslices,        h->A.padded,
                 h->A.uniform_width, h->AT.nslices, h->AT.padded,   h->AT.uniform_width, h->longA.nnz,
                 h->longAT.nnz,      h->A.dict ? h->ndict : 0, h->const_mask, h->hx.ghost + h->hy.ghost,
                 h->identity_layout ? 0 : 1 + h->granule,
                 h->bandA.built ? h->bandA.geo.windows : 0, h->bandAT.built ? h->bandAT.geo.windows : 0};
}

// Choose the kernel variants (called at the end of setup(), before any neighbour may write into this
// rank's vectors).  Forced by cpppd_problem.kernel_variant / CPPPD_KERNEL_VARIANT, or — for LPs large
// enough for the choice to matter — measured: every variant runs on the real operands (one untimed launch
// each, then two passes of two timed launches, CUDA events), the fastest wins, and variant 0 is only given
// up for a gain above 2 %.
// The iterates do not depend on the choice; the state (x, xbar, y) is put back afterwards.  A measured choice is
// remembered per process and operand shape (g_tune_cache; CPPPD_AUTOTUNE_CACHE=0 measures every time).
int tune_kernels(cpppd_solver *h) {
  // a banded operand is used unless the timing below finds the SELL kernel faster
  h->bandA.in_use = h->bandA.built;
  h->bandAT.in_use = h->bandAT.built;
  if (const char *env = getenv("CPPPD_BAND_SHAPE"))  // (all shapes give the same bits; the timing below picks one otherwise)
    if (atoi(env) >= 0 && atoi(env) < kNumBandShapes) h->bandA.shape = h->bandAT.shape = atoi(env);
  int request = h->variant_request;
  if (request == 0)
    if (const char *env = getenv("CPPPD_KERNEL_VARIANT")) request = atoi(env);
  if (request != 0) {
    const int p = request & 0xff, d = (request >> 8) & 0xff ? (request >> 8) & 0xff : p;
    if (p < 1 || p > kNumVariants || d < 1 || d > kNumVariants)
      return fail(h, CPPPD_ERR_INVALID, "kernel_variant %d: variants are 1 .. %d", request, kNumVariants);
    h->primal_variant = p - 1;
    h->dual_variant = d - 1;
    if (!(h->flags & CPPPD_FLAG_BANDED)) h->bandA.in_use = h->bandAT.in_use = false;  // the caller asked for a SELL variant
    return 0;
  }
  int64_t min_nnz = (int64_t)1 << 22;
  if (const char *env = getenv("CPPPD_AUTOTUNE_MIN_NNZ")) min_nnz = atoll(env);
  if ((h->flags & CPPPD_FLAG_NO_AUTOTUNE) || std::max(h->nnz_rows, h->nnz_cols) < min_nnz) return 0;
  // operands of the same shape were timed before in this process (a second solve of the same LP family):
  // reuse that choice instead of spending another ~50 launches
  const TuneKey key = tune_key(h);
  bool use_cache = true;
  if (const char *env = getenv("CPPPD_AUTOTUNE_CACHE")) use_cache = atoi(env) != 0;
  if (use_cache) {
    std::lock_guard<std::mutex> lock(g_tune_mutex);
    auto hit = g_tune_cache.find(key);
    if (hit != g_tune_cache.end()) {
      h->primal_variant = hit->second.primal;
      h->dual_variant = hit->second.dual;
      memcpy(h->variant_ms, hit->second.ms, sizeof h->variant_ms);
      if (!(h->flags & CPPPD_FLAG_BANDED)) {  // (forced: stays in use whatever an earlier timing said)
        h->bandAT.in_use = hit->second.band_primal;
        h->bandA.in_use = hit->second.band_dual;
      }
      h->bandAT.ms = hit->second.band_ms[0];
      h->bandA.ms = hit->second.band_ms[1];
      h->bandAT.shape = hit->second.band_shape[0];
      h->bandA.shape = hit->second.band_shape[1];
      memcpy(h->bandAT.shape_ms, hit->second.band_shape_ms[0], sizeof h->bandAT.shape_ms);
      memcpy(h->bandA.shape_ms, hit->second.band_shape_ms[1], sizeof h->bandA.shape_ms);
      h->autotuned = true;
      return 0;
    }
  }
  cudaStream_t st = h->stream;
  Scratch tmp(h);
  const int64_t nx = h->x_len, ny = h->y_len;
  double *x_saved = nullptr;
  if (int rc = tmp.get(&x_saved, nx)) return rc;
  CK(cudaMemcpyAsync(x_saved, h->x, sizeof(double) * nx, cudaMemcpyDeviceToDevice, st));
  Events tune_ev(2);  // (destroyed on every way out, also the early returns of CK)
  CK(tune_ev.create());
  const cudaEvent_t e0 = tune_ev[0], e1 = tune_ev[1];
  int rc = 0;
  for (int kind = 0; kind < 2 && !rc; ++kind) {
    // pass 0 runs every variant once untimed (caches, clocks); passes 1 and 2 time two launches of each variant
    // in turn, and a variant keeps its better pass — so no variant is judged on a cold or ramping GPU
    for (int pass = 0; pass < 3 && !rc; ++pass) {
      for (int v = 0; v < kNumVariants && !rc; ++v) {
        float ms = 0.f;
        if (pass > 0) CK(cudaEventRecord(e0, st));
        for (int rep = 0; rep < (pass > 0 ? 2 : 1) && !rc; ++rep)
          rc = kind == 0 ? launch_primal(h, false, v) : launch_dual(h, v);
        if (rc || pass == 0) continue;
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= 2;
        if (pass == 1 || ms < h->variant_ms[kind][v]) h->variant_ms[kind][v] = ms;
      }
    }
    if (rc) break;
    int best = 0;
    for (int v = 1; v < kNumVariants; ++v)
      if (h->variant_ms[kind][v] < h->variant_ms[kind][best]) best = v;
    if (best != 0 && h->variant_ms[kind][best] > 0.98f * h->variant_ms[kind][0]) best = 0;
    (kind == 0 ? h->primal_variant : h->dual_variant) = best;
    // the banded copy of the operand, when it was built: same protocol, against the best SELL variant
    Band &band = kind == 0 ? h->bandAT : h->bandA;
    if (band.built) {
      int forced_shape = -1;
      if (const char *env = getenv("CPPPD_BAND_SHAPE")) forced_shape = atoi(env);
      for (int pass = 0; pass < 3 && !rc; ++pass) {
        for (int shape = 0; shape < kNumBandShapes && !rc; ++shape) {
          if (forced_shape >= 0 && forced_shape < kNumBandShapes && shape != forced_shape) continue;
          band.shape = shape;
          float ms = 0.f;
          if (pass > 0) CK(cudaEventRecord(e0, st));
          for (int rep = 0; rep < (pass > 0 ? 2 : 1) && !rc; ++rep)
            rc = kind == 0 ? launch_primal(h, false, kBandVariant) : launch_dual(h, kBandVariant);
          if (rc || pass == 0) continue;
          CK(cudaEventRecord(e1, st));
          CK(cudaEventSynchronize(e1));
          CK(cudaEventElapsedTime(&ms, e0, e1));
          ms /= 2;
          if (pass == 1 || ms < band.shape_ms[shape]) band.shape_ms[shape] = ms;
        }
      }
      band.shape = forced_shape >= 0 && forced_shape < kNumBandShapes ? forced_shape : 0;
      for (int shape = 0; shape < kNumBandShapes; ++shape)
        if (band.shape_ms[shape] > 0.f && band.shape_ms[shape] < band.shape_ms[band.shape]) band.shape = shape;
      band.ms = band.shape_ms[band.shape];
      band.in_use = (h->flags & CPPPD_FLAG_BANDED) || (!rc && band.ms < 0.98f * h->variant_ms[kind][best]);
    }
  }
  if (rc) return rc;
  h->autotuned = true;
  if (use_cache) {
    TuneChoice choice;
    choice.primal = h->primal_variant;
    choice.dual = h->dual_variant;
    memcpy(choice.ms, h->variant_ms, sizeof choice.ms);
    choice.band_primal = h->bandAT.built && h->bandAT.ms > 0.f && h->bandAT.ms < 0.98f * h->variant_ms[0][h->primal_variant];
    choice.band_dual = h->bandA.built && h->bandA.ms > 0.f && h->bandA.ms < 0.98f * h->variant_ms[1][h->dual_variant];
    choice.band_ms[0] = h->bandAT.ms;
    choice.band_ms[1] = h->bandA.ms;
    choice.band_shape[0] = h->bandAT.shape;
    choice.band_shape[1] = h->bandA.shape;
    memcpy(choice.band_shape_ms[0], h->bandAT.shape_ms, sizeof h->bandAT.shape_ms);
    memcpy(choice.band_shape_ms[1], h->bandA.shape_ms, sizeof h->bandA.shape_ms);
    std::lock_guard<std::mutex> lock(g_tune_mutex);
    g_tune_cache[key] = choice;
  }
  // back to the initial state: x = x0, xbar = x (:190), y = 0 (:166,:177)
  CK(cudaMemcpyAsync(h->x, x_saved, sizeof(double) * nx, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(h->xbar, x_saved, sizeof(double) * nx, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemsetAsync(h->y, 0, sizeof(double) * std::max<int64_t>(ny, 1), st));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  return 0;
}

int get_graph(cpppd_solver *h, int64_t k, cudaGraphExec_t *out) {
  auto it = h->graphs.find(k);
  if (it != h->graphs.end()) {
    *out = it->second;
    return 0;
  }
  cudaGraph_t g = nullptr;
  CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  int rc = 0;
  for (int64_t i = 0; i < k && !rc; ++i) {
    rc = launch_primal(h, false);
    if (!rc) rc = launch_dual(h);
  }
  cudaError_t e = cudaStreamEndCapture(h->stream, &g);
  if (rc) return rc;
  CK(e);
  cudaGraphExec_t ge = nullptr;
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaGraphDestroy(g);
  h->graphs[k] = ge;
  *out = ge;
  return 0;
}

// tiny LPs with CPPPD_FLAG_TINY_PERSISTENT: all k iterations in one launch of one CTA (k_tiny_iterate)
int run_tiny(cpppd_solver *h, int64_t k) {
  const int64_t widest = std::max(h->AT.nslices, h->A.nslices) * kSlice;
  const int threads = (int)std::min<int64_t>(kTinyBlock, std::max<int64_t>(kSlice, widest));
  const int has_eq = h->m_eq_glob > 0, has_ineq = h->m_ineq_glob > 0;
  if (h->A.dict)
    k_tiny_iterate<true><<<1, threads, 0, h->stream>>>(view(h->AT), view(h->A), h->vc, h->vT, h->vlb, h->vub, h->vb, h->vsigma,
                                                      h->x, h->xbar, h->y, h->n, h->m, h->m_eq, has_eq, has_ineq, h->theta,
                                                      h->one_plus_theta, k);
  else
    k_tiny_iterate<false><<<1, threads, 0, h->stream>>>(view(h->AT), view(h->A), h->vc, h->vT, h->vlb, h->vub, h->vb, h->vsigma,
                                                       h->x, h->xbar, h->y, h->n, h->m, h->m_eq, has_eq, has_ineq, h->theta,
                                                       h->one_plus_theta, k);
  CK(cudaGetLastError());
  h->niter += k;
  return 0;
}

// small LPs: all k iterations in one launch of one thread-block cluster (k_cluster_iterate)
int run_cluster(cpppd_solver *h, int64_t k) {
#ifdef __CUDACC__
  const ClusterPlan &P = h->cluster;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(P.ctas);
  cfg.blockDim = dim3(kClusterBlock);
  cfg.dynamicSmemBytes = P.smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = P.ctas;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  const int has_eq = h->m_eq_glob > 0, has_ineq = h->m_ineq_glob > 0;
  const char *mode_env = getenv("CPPPD_CLUSTER_MODE");  // 0: relaxed barrier arrive, 1: release / acquire
  const int mode = mode_env ? atoi(mode_env) : kClusterDefaultMode;
  const bool one_pass = std::max(P.spc_at, P.spc_a) <= kClusterBlock / 32;
  auto kernel = cluster_kernel(mode != 0, one_pass);
  CK(cudaLaunchKernelEx(&cfg, kernel, view(h->AT), view(h->A), h->vc, h->vT, h->vlb, h->vub, h->vb, h->vsigma,
                        h->x, h->xbar, h->y, h->n, h->m, h->m_eq, has_eq, has_ineq, h->theta, h->one_plus_theta, k, P.spc_at,
                        P.spc_a, P.ent_at, P.ent_a));
  h->niter += k;
  return 0;
#else
  // CPU emulation: a host buffer per CTA, one emulated launch per phase (cpppd_cluster.cuh)
  const ClusterPlan &P = h->cluster;
  std::vector<std::vector<unsigned char>> smem(P.ctas, std::vector<unsigned char>(P.smem + 16, (unsigned char)0xA5));
  for (int r = 0; r < P.ctas; ++r) {
    unsigned char *base = smem[r].data();
    g_emul_cluster_smem[r] = base + ((16 - reinterpret_cast<uintptr_t>(base) % 16) % 16);
  }
  EmulClusterArgs a{view(h->AT), view(h->A), h->vc, h->vT, h->vlb, h->vub, h->vb, h->vsigma, h->x, h->xbar, h->y, h->n, h->m,
                    h->m_eq, h->m_eq_glob > 0, h->m_ineq_glob > 0, h->theta, h->one_plus_theta, P.spc_at, P.spc_a, P.ent_at,
                    P.ent_a};
  k_emul_cluster_stage<<<P.ctas, kClusterBlock, 0, h->stream>>>(a);
  for (int64_t it = 0; it < k; ++it) {
    k_emul_cluster_primal<<<P.ctas, kClusterBlock, 0, h->stream>>>(a);
    k_emul_cluster_dual<<<P.ctas, kClusterBlock, 0, h->stream>>>(a);
  }
  k_emul_cluster_writeback<<<P.ctas, kClusterBlock, 0, h->stream>>>(a);
  for (int r = 0; r < P.ctas; ++r) g_emul_cluster_smem[r] = nullptr;
  h->niter += k;
  return 0;
#endif
}

int run_iterations(cpppd_solver *h, int64_t k) {
  if (h->tiny && k > 0) return run_tiny(h, k);
  if (h->cluster.on && k > 0) return run_cluster(h, k);
  const bool use_graph = !(h->flags & CPPPD_FLAG_NO_GRAPH) && (h->world == 1 || h->p2p.active || (h->flags & CPPPD_FLAG_GRAPH_COMM));
  while (k > 0) {
    int64_t step = std::min<int64_t>(k, kGraphChunk);
    if (use_graph && step >= 2) {
      cudaGraphExec_t ge = nullptr;
      if (int rc = get_graph(h, step, &ge)) return rc;
      CK(cudaGraphLaunch(ge, h->stream));
    } else {
      for (int64_t i = 0; i < step; ++i) {
        if (int rc = launch_primal(h, false)) return rc;
        if (int rc = launch_dual(h)) return rc;
      }
      CK(cudaGetLastError());
    }
    k -= step;
    h->niter += step;
  }
  return 0;
}

}  // namespace
