// libcpppd — B200 (sm_100a) core of the Chambolle-Pock PPD LP solver.  C ABI in include/cpppd.h.
//
// What runs here is the reference's pysparselp/ChambollePockPPD.py:122-343, restated
// for the GPU (file:line citations refer to that file unless noted):
//   * operator storage : A = [A_eq; A_ineq] and A^T, each in SELL-32 (sliced ELLPACK, slice
//                        height = warp size).  One thread owns one row; lane l of a warp
//                        reads element k of its row at  base + 32*k + l, so every warp-level
//                        load of values / indices is one contiguous 256 B / 128 B segment.
//   * k_primal         : d = c + A^T y (:198-217) fused with x - T d, the box clip and the
//                        theta extrapolation (:220-228).  d never reaches memory except on
//                        stats iterations.
//   * k_dual           : r = A xbar - b (:231-240) fused with y + Sigma r and the projection
//                        of y_ineq on >= 0 (:333-341).  r never reaches memory.
//   * k_precond_*      : column / row abs-power sums -> diag_t, diag_sigma (:122-179).
//   * k_stats_*        : the stats block (:248-291) as warp-shuffle + block reductions with a
//                        deterministic two-level tree, the best-integer bookkeeping on device.
//   * multi-GPU        : owner-computes partition.  Every rank owns a set of rows (its y) and a
//                        set of columns (its x); it stores its rows of A and its columns of A (as
//                        rows of A^T) and keeps *ghost* copies of the few xbar / y entries owned by
//                        other ranks that its rows / columns touch.  One halo exchange of xbar and
//                        one of y per iteration replace the dense all-reduce of A^T y: no partial
//                        sums cross GPUs, so the iterates stay bit-identical to the single-GPU
//                        (and reference) ones.  The partition is a pure integer function of the
//                        sparsity pattern (oracle/partition_oracle.py restates it).
//
// Floating point: IEEE fp64, compiled with -fmad=false and written with explicit
// __dmul_rn/__dadd_rn so products and sums round exactly like the numpy/scipy code of the
// reference.  A row (column) sum is accumulated sequentially in the stored entry order from
// 0.0 — the same order as scipy's csr_matvec (csc_matvec) — so x, xbar, y, T and Sigma are
// bit-identical to the reference.  Only the scalar dot products of the stats block use a
// different (tree) order than numpy.dot.
#include "cpppd_types.cuh"
#include "cpppd_setup_kernels.cuh"
#include "cpppd_hot_kernels.cuh"
#include "cpppd_cluster.cuh"
#include "cpppd_stats_kernels.cuh"
#include "cpppd_host.cuh"


// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
namespace {

// assemble a distributed vector (local layout, owned part) into original order on the host
int fetch_vector(cpppd_solver *h, const double *local, bool is_col, double *host_dst) {
  const int64_t owned = is_col ? h->n : h->m, full = is_col ? h->n_glob : h->m_glob;
  if (h->identity_layout) {
    if (full) CK(cudaMemcpyAsync(host_dst, local, sizeof(double) * full, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
  }
  Scratch tmp(h);
  double *buf = nullptr;
  if (int rc = tmp.get(&buf, full)) return rc;
  CK(cudaMemsetAsync(buf, 0, sizeof(double) * std::max<int64_t>(full, 1), h->stream));
  if (owned) k_scatter_f64<<<grid_for(owned), kBlock, 0, h->stream>>>(local, is_col ? h->col_old : h->row_old, owned, buf);
  // every entry is owned by exactly one rank and the others contribute all-zero bits: summing the BIT PATTERNS as
  // 64-bit integers reproduces the owner's double exactly (an fp64 sum would turn -0.0 into +0.0 and quiet NaNs)
  if (h->world > 1 && full) NK(g_nccl.AllReduce(buf, buf, (size_t)full, ncclUint64, ncclSum, h->comm, h->stream));
  if (full) CK(cudaMemcpyAsync(host_dst, buf, sizeof(double) * full, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int vector_ptr(cpppd_solver *h, int32_t which, double **p, bool *is_col) {
  switch (which) {
    case CPPPD_VEC_X: *p = h->x; *is_col = true; return 0;
    case CPPPD_VEC_XBAR: *p = h->xbar; *is_col = true; return 0;
    case CPPPD_VEC_Y: *p = h->y; *is_col = false; return 0;
    case CPPPD_VEC_T: *p = h->T; *is_col = true; return 0;
    case CPPPD_VEC_SIGMA: *p = h->sigma; *is_col = false; return 0;
    case CPPPD_VEC_BEST_INTEGER: *p = h->best; *is_col = true; return 0;
    case CPPPD_VEC_D: *p = h->dbuf; *is_col = true; return 0;
    default: return fail(h, CPPPD_ERR_INVALID, "unknown vector id %d", which);
  }
}

}  // namespace

extern "C" {

int cpppd_abi_version(void) { return CPPPD_ABI_VERSION; }

const char *cpppd_last_error(cpppd_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int cpppd_comm_unique_id(void *out128) {
  cpppd_solver *h = nullptr;
  if (!out128) return fail(h, CPPPD_ERR_INVALID, "null output");
  if (const char *e = load_nccl()) return fail(h, CPPPD_ERR_COMM, "%s", e);
  ncclUniqueId id;
  NK(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof id);
  return 0;
}

int cpppd_comm_create(const void *id128, int32_t rank, int32_t world_size, int32_t device, cpppd_comm *out) {
  cpppd_solver *h = nullptr;
  if (!id128 || !out || world_size < 2 || rank < 0 || rank >= world_size) return fail(h, CPPPD_ERR_INVALID, "bad argument");
  *out = nullptr;
  if (const char *e = load_nccl()) return fail(h, CPPPD_ERR_COMM, "%s", e);
  CK(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  cpppd_comm_s *c = new cpppd_comm_s();
  c->rank = rank;
  c->world = world_size;
  c->device = device;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, world_size, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return fail(h, CPPPD_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
  }
  *out = c;
  return 0;
}

int cpppd_comm_destroy(cpppd_comm comm) {
  if (!comm) return 0;
  cudaSetDevice(comm->device);
  pool_release(comm->pool);
  if (comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm->comm);
  delete comm;
  return 0;
}

int cpppd_create(const cpppd_problem *P, cpppd_handle *out) {
  cpppd_solver *h = nullptr;
  if (!P || !out) return fail(h, CPPPD_ERR_INVALID, "null argument");
  *out = nullptr;
  if (P->abi_version != CPPPD_ABI_VERSION)
    return fail(h, CPPPD_ERR_INVALID, "ABI version mismatch: caller %d, library %d", P->abi_version, CPPPD_ABI_VERSION);
  if (P->n < 0 || P->m_eq < 0 || P->m_ineq < 0 || P->nnz < 0) return fail(h, CPPPD_ERR_INVALID, "negative size");
  if (P->n >= (int64_t)1 << 30 || P->m_eq + P->m_ineq >= (int64_t)1 << 30)
    return fail(h, CPPPD_ERR_INVALID, "n and m must be below 2^30 (30-bit local indices)");
  if (P->nnz >= (int64_t)1 << 32) return fail(h, CPPPD_ERR_INVALID, "nnz must be below 2^32");
  if (P->index_bits != 32) return fail(h, CPPPD_ERR_INVALID, "column indices must be int32 (narrow them on the host)");
  if (P->indptr_bits != 32 && P->indptr_bits != 64) return fail(h, CPPPD_ERR_INVALID, "indptr_bits must be 32 or 64");
  const int world = P->world_size <= 0 ? 1 : P->world_size;
  if (world > kMaxWorld || P->rank < 0 || P->rank >= world) return fail(h, CPPPD_ERR_INVALID, "bad rank / world_size");
  if (world > 1 && !P->comm_id && !P->comm)
    return fail(h, CPPPD_ERR_INVALID, "world_size > 1 needs comm_id (cpppd_comm_unique_id) or comm (cpppd_comm_create)");
  const int64_t m = P->m_eq + P->m_ineq;
  if (!P->indptr || (P->nnz && (!P->indices || !P->values)) || (P->n && (!P->c || !P->lb || !P->ub)) || (m && !P->b))
    return fail(h, CPPPD_ERR_INVALID, "null array pointer");
  {
    int64_t first = P->indptr_bits == 64 ? ((const int64_t *)P->indptr)[0] : ((const int32_t *)P->indptr)[0];
    int64_t last = P->indptr_bits == 64 ? ((const int64_t *)P->indptr)[m] : ((const int32_t *)P->indptr)[m];
    if (first != 0 || last != P->nnz) return fail(h, CPPPD_ERR_INVALID, "indptr[0] must be 0 and indptr[m] == nnz");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(h, CPPPD_ERR_NODEVICE, "no CUDA device available: this solver has no CPU path");
  }
  if (P->device < 0 || P->device >= ndev) return fail(h, CPPPD_ERR_INVALID, "device %d out of range (%d devices)", P->device, ndev);
  h = new cpppd_solver();
  h->device = P->device;
  h->n_glob = P->n;
  h->m_eq_glob = P->m_eq;
  h->m_ineq_glob = P->m_ineq;
  h->m_glob = m;
  h->nnz_glob = P->nnz;
  h->alpha = P->alpha;
  h->theta = P->theta;
  h->one_plus_theta = P->one_plus_theta;
  h->flags = P->flags;
  h->granule = P->partition_granule;
  h->variant_request = P->kernel_variant;
  h->long_threshold = P->long_row_threshold == 0 ? kLongDefault : P->long_row_threshold;
  h->band_window = P->band_window;
  h->rank = P->rank;
  h->world = world;
  h->alloc = P->alloc;
  h->free_fn = P->free;
  h->alloc_user = P->alloc_user;
  int rc = 0;
  do {
    if (cudaSetDevice(h->device) != cudaSuccess) {
      rc = fail(h, CPPPD_ERR_CUDA, "cudaSetDevice(%d) failed", h->device);
      break;
    }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    // experiment knob: bytes the L2 fetches from DRAM per missing sector (device-wide limit, default 64).  Random
    // 8-byte gathers that miss L2 otherwise pull two sectors for one (profiles/r02_random_lp.md)
    if (const char *env = getenv("CPPPD_L2_FETCH_GRANULARITY"))
      if (atoi(env) > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(env));
    if (P->stream) {
      h->stream = (cudaStream_t)P->stream;
    } else {
      if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        rc = fail(h, CPPPD_ERR_CUDA, "cudaStreamCreate failed");
        break;
      }
      h->own_stream = true;
    }
    if (world > 1 && P->comm) {
      cpppd_comm_s *shared = static_cast<cpppd_comm_s *>(P->comm);
      if (shared->world != world || shared->rank != h->rank || shared->device != h->device) {
        rc = fail(h, CPPPD_ERR_INVALID, "cpppd_problem.comm was created for another rank / world size / device");
        break;
      }
      h->comm = shared->comm;
      h->own_comm = false;
      h->shared = shared;
    } else if (world > 1) {
      if (const char *e = load_nccl()) {
        rc = fail(h, CPPPD_ERR_COMM, "%s", e);
        break;
      }
      ncclUniqueId id;
      memcpy(&id, P->comm_id, sizeof id);
      ncclResult_t r = g_nccl.CommInitRank(&h->comm, world, id, h->rank);
      if (r != ncclSuccess) {
        rc = fail(h, CPPPD_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        break;
      }
    }
    rc = setup(h, P);
  } while (0);
  if (rc) {
    g_create_error = h->err;
    cpppd_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

int cpppd_destroy(cpppd_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second);
  if (h->p2p.active && h->comm && !h->sticky) {
    // neighbours may still be storing into this rank's ghost slots: rendezvous before unmapping
    if (g_nccl.AllReduce(h->p2p.flags, h->p2p.flags, 1, ncclInt8, ncclSum, h->comm, h->stream) == ncclSuccess)
      cudaStreamSynchronize(h->stream);
  }
  for (void *p : h->p2p.opened) cudaIpcCloseMemHandle(p);
  for (void *p : h->p2p.own) cudaFree(p);
  if (h->pooled && h->shared) h->shared->pool.busy = false;  // (the buffers stay mapped for the next solve)
  if (h->comm && h->own_comm) g_nccl.CommDestroy(h->comm);
  for (void *p : h->owned) dev_free(h, p);
  if (h->stats_host) cudaFreeHost(h->stats_host);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
  return 0;
}

int cpppd_iterate(cpppd_handle h, int64_t k) {
  CHECK_HANDLE(h);
  if (k < 0) return fail(h, CPPPD_ERR_INVALID, "negative iteration count");
  if (h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "cpppd_dual_step must close the open iteration first");
  return run_iterations(h, k);
}

int cpppd_primal_step(cpppd_handle h, int32_t keep_d) {
  CHECK_HANDLE(h);
  if (h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "primal step already issued for this iteration");
  if (int rc = launch_primal(h, keep_d != 0)) return rc;
  h->mid_iteration = true;
  h->have_d = keep_d != 0;
  CK(cudaGetLastError());
  return 0;
}

int cpppd_stats_step(cpppd_handle h, int32_t force_integer) {
  CHECK_HANDLE(h);
  if (!h->mid_iteration || !h->have_d)
    return fail(h, CPPPD_ERR_STATE, "cpppd_stats_step needs a preceding cpppd_primal_step(keep_d=1)");
  const int has_eq = h->m_eq_glob > 0, has_ineq = h->m_ineq_glob > 0;
  // column pass turns dbuf into x4 in place; with force_integer it also materialises xr = rint(x)
  double *xr = h->x;
  if (force_integer) {
    if (!h->xr_scratch)
      if (int rc = alloc_array(h, &h->xr_scratch, h->x_len)) return rc;
    xr = h->xr_scratch;
  }
  k_stats_cols<<<h->stat_blocks_c, kBlock, 0, h->stream>>>(h->vc, h->x, h->xbar, h->vlb, h->vub, h->dbuf, xr, h->n,
                                                          force_integer, h->colpart);
  // the row pass reads x, x4 and xr at ghost columns too
  if (int rc = exchange(h, h->x, h->hx)) return rc;
  if (int rc = exchange(h, h->dbuf, h->hx)) return rc;
  if (force_integer)
    if (int rc = exchange(h, xr, h->hx)) return rc;
  // long rows of A: their sums against x, x4, xbar (and xr) go to the tails the row pass gathers from
  if (int rc = long_pass(h, h->longA, h->x, h->x)) return rc;
  if (int rc = long_pass(h, h->longA, h->dbuf, h->dbuf)) return rc;
  if (int rc = long_pass(h, h->longA, h->xbar, h->xbar)) return rc;
  if (force_integer)
    if (int rc = long_pass(h, h->longA, xr, xr)) return rc;
  k_stats_rows<<<h->stat_blocks_r, kBlock, 0, h->stream>>>(view(h->A), h->x, h->dbuf, h->xbar, xr, h->vb, h->y, h->m,
                                                          h->m_eq, force_integer, h->row_off, h->rowpart);
  if (h->gt_local)
    k_stats_gt<<<h->stat_blocks_g, kBlock, 0, h->stream>>>(h->gt_idx, h->gt_val, h->gt_local, h->x, h->gtpart);
  k_stats_local<<<1, kBlock, 0, h->stream>>>(h->colpart, h->stat_blocks_c, h->rowpart, h->stat_blocks_r, h->gtpart,
                                             h->gt_local ? h->stat_blocks_g : 0, h->stat_local);
  const double *all = h->stat_local;
  if (h->world > 1) {
    NK(g_nccl.AllGather(h->stat_local, h->stat_all, kStatQ, ncclFloat64, h->comm, h->stream));
    all = h->stat_all;
  }
  k_stats_final<<<1, 32, 0, h->stream>>>(all, h->world, h->n_glob, has_eq, has_ineq, h->niter, h->gt_total, h->stats_dev);
  k_snapshot_best<<<std::max(1, std::min(grid_for(h->n), h->sm_count * 8)), kBlock, 0, h->stream>>>(h->stats_dev, xr,
                                                                                                  h->best, h->n);
  CK(cudaMemcpyAsync(h->stats_host, h->stats_dev, sizeof(cpppd_stats), cudaMemcpyDeviceToHost, h->stream));
  h->stats_pending = true;
  h->have_d = false;  // dbuf now holds x4
  CK(cudaGetLastError());
  return 0;
}

int cpppd_dual_step(cpppd_handle h) {
  CHECK_HANDLE(h);
  if (!h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "no primal step open");
  if (int rc = launch_dual(h)) return rc;
  h->mid_iteration = false;
  h->niter += 1;
  CK(cudaGetLastError());
  return 0;
}

int cpppd_sync(cpppd_handle h) {
  CHECK_HANDLE(h);
  CK(cudaStreamSynchronize(h->stream));
  return check_halo_timeout(h);
}

int cpppd_read_stats(cpppd_handle h, cpppd_stats *out) {
  CHECK_HANDLE(h);
  if (!out) return fail(h, CPPPD_ERR_INVALID, "null output");
  if (!h->stats_pending) return fail(h, CPPPD_ERR_STATE, "no stats step has been issued");
  CK(cudaStreamSynchronize(h->stream));
  if (int rc = check_halo_timeout(h)) return rc;
  *out = *h->stats_host;
  return 0;
}

int cpppd_time_iterations(cpppd_handle h, int64_t k, float *elapsed_ms) {
  CHECK_HANDLE(h);
  if (!elapsed_ms || k < 0) return fail(h, CPPPD_ERR_INVALID, "bad argument");
  if (h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "cpppd_dual_step must close the open iteration first");
  Events ev(2);
  CK(ev.create());
  CK(cudaEventRecord(ev[0], h->stream));
  int rc = run_iterations(h, k);
  if (!rc) {
    CK(cudaEventRecord(ev[1], h->stream));
    CK(cudaEventSynchronize(ev[1]));
    CK(cudaEventElapsedTime(elapsed_ms, ev[0], ev[1]));
  }
  return rc ? rc : check_halo_timeout(h);
}

int cpppd_time_kernels(cpppd_handle h, int64_t k, float *primal_ms, float *dual_ms) {
  CHECK_HANDLE(h);
  if (!primal_ms || !dual_ms || k < 0 || k > 64) return fail(h, CPPPD_ERR_INVALID, "bad argument (k must be in [0, 64])");
  if (h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "cpppd_dual_step must close the open iteration first");
  // the halo exchanges (world > 1) fall inside the brackets of the kernel that produces the data
  Events ev(2 * k + 1);
  CK(ev.create());
  CK(cudaEventRecord(ev[0], h->stream));
  for (int64_t i = 0; i < k; ++i) {
    if (int rc = launch_primal(h, false)) return rc;
    CK(cudaEventRecord(ev[2 * i + 1], h->stream));
    if (int rc = launch_dual(h)) return rc;
    CK(cudaEventRecord(ev[2 * i + 2], h->stream));
  }
  h->niter += k;
  CK(cudaEventSynchronize(ev[2 * k]));
  *primal_ms = *dual_ms = 0.f;
  for (int64_t i = 0; i < k; ++i) {
    float a = 0.f, b = 0.f;
    CK(cudaEventElapsedTime(&a, ev[2 * i], ev[2 * i + 1]));
    CK(cudaEventElapsedTime(&b, ev[2 * i + 1], ev[2 * i + 2]));
    *primal_ms += a;
    *dual_ms += b;
  }
  return 0;
}

int cpppd_get_vector(cpppd_handle h, int32_t which, double *host_dst) {
  CHECK_HANDLE(h);
  double *p = nullptr;
  bool is_col = true;
  if (int rc = vector_ptr(h, which, &p, &is_col)) return rc;
  if (!host_dst) return fail(h, CPPPD_ERR_INVALID, "null destination");
  if (int rc = fetch_vector(h, p, is_col, host_dst)) return rc;
  return check_halo_timeout(h);
}

int cpppd_set_vector(cpppd_handle h, int32_t which, const double *host_src) {
  CHECK_HANDLE(h);
  if (which != CPPPD_VEC_X && which != CPPPD_VEC_XBAR && which != CPPPD_VEC_Y)
    return fail(h, CPPPD_ERR_INVALID, "only x, xbar and y can be set");
  double *p = nullptr;
  bool is_col = true;
  if (int rc = vector_ptr(h, which, &p, &is_col)) return rc;
  if (!host_src) return fail(h, CPPPD_ERR_INVALID, "null source");
  Scratch tmp(h);
  const int64_t local = is_col ? h->n + h->hx.ghost : h->m + h->hy.ghost;
  if (int rc = upload_local(h, tmp, host_src, is_col ? h->n_glob : h->m_glob, is_col ? h->col_old : h->row_old, local, p))
    return rc;
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int cpppd_set_ground_truth(cpppd_handle h, const int32_t *indices, const double *values, int64_t count) {
  CHECK_HANDLE(h);
  if (count < 0 || (count && (!indices || !values))) return fail(h, CPPPD_ERR_INVALID, "bad ground truth");
  h->gt_local = h->gt_total = 0;
  if (count == 0) return 0;
  // keep the entries whose column this rank owns, in local numbering
  std::vector<int32_t> local_of(h->n_glob, -1);
  if (h->identity_layout) {
    for (int64_t j = 0; j < h->n_glob; ++j) local_of[j] = (int32_t)j;
  } else {
    std::vector<int32_t> col_old(h->n);
    if (h->n) CK(cudaMemcpyAsync(col_old.data(), h->col_old, sizeof(int32_t) * h->n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int64_t lj = 0; lj < h->n; ++lj) local_of[col_old[lj]] = (int32_t)lj;
  }
  std::vector<int32_t> idx;
  std::vector<double> val;
  for (int64_t k = 0; k < count; ++k) {
    if (indices[k] < 0 || indices[k] >= h->n_glob) return fail(h, CPPPD_ERR_INVALID, "ground truth index out of range");
    const int32_t lj = local_of[indices[k]];
    if (lj >= 0) {
      idx.push_back(lj);
      val.push_back(values[k]);
    }
  }
  h->gt_total = count;
  h->gt_local = (int64_t)idx.size();
  if (h->gt_local) {
    h->stat_blocks_g = (int)std::max<int64_t>(1, std::min<int64_t>(grid_for(h->gt_local), (int64_t)h->sm_count * 8));
    if (int rc = alloc_array(h, &h->gt_idx, h->gt_local)) return rc;
    if (int rc = alloc_array(h, &h->gt_val, h->gt_local)) return rc;
    if (int rc = alloc_array(h, &h->gtpart, (int64_t)h->stat_blocks_g * kGtQ)) return rc;
    CK(cudaMemcpyAsync(h->gt_idx, idx.data(), sizeof(int32_t) * idx.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->gt_val, val.data(), sizeof(double) * val.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

int cpppd_set_row_offsets(cpppd_handle h, const double *offsets) {
  CHECK_HANDLE(h);
  if (!offsets) {
    h->row_off = nullptr;  // (the buffer, if any, stays allocated until the handle is destroyed)
    return 0;
  }
  double *buf = nullptr;
  if (int rc = alloc_array(h, &buf, h->m)) return rc;
  Scratch tmp(h);
  if (int rc = upload_local(h, tmp, offsets, h->m_glob, h->row_old, h->m, buf)) return rc;
  CK(cudaStreamSynchronize(h->stream));
  h->row_off = buf;
  return 0;
}

int cpppd_get_layout(cpppd_handle h, int32_t columns, int64_t *owned, int64_t *ghost, int32_t *ids) {
  CHECK_HANDLE(h);
  const Halo &H = columns ? h->hx : h->hy;
  if (owned) *owned = H.owned;
  if (ghost) *ghost = H.ghost;
  if (!ids) return 0;
  const int64_t count = H.owned + H.ghost;
  if (h->identity_layout) {
    for (int64_t i = 0; i < count; ++i) ids[i] = (int32_t)i;
    return 0;
  }
  if (count) CK(cudaMemcpyAsync(ids, columns ? h->col_old : h->row_old, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int cpppd_get_info(cpppd_handle h, cpppd_info *out) {
  CHECK_HANDLE(h);
  if (!out) return fail(h, CPPPD_ERR_INVALID, "null output");
  memset(out, 0, sizeof *out);
  out->n = h->n_glob;
  out->m_eq = h->m_eq_glob;
  out->m_ineq = h->m_ineq_glob;
  out->nnz = h->nnz_glob;
  out->a_padded_entries = h->A.padded;
  out->at_padded_entries = h->AT.padded;
  out->device_bytes = h->device_bytes;
  const int64_t P = h->nnz_glob >= ((int64_t)1 << 31) ? 8 : 4;
  out->bytes_per_iteration_algorithmic =
      2 * h->nnz_glob * 12 + P * (h->m_glob + 1) + P * (h->n_glob + 1) + 8 * (8 * h->n_glob + 5 * h->m_glob);
  const int64_t entry_bytes = h->A.dict ? 4 : 12;
  int64_t vec_bytes = 8 * (3 * h->n + 2 * h->m);  // x read+write, xbar write, y read+write
  vec_bytes += 8 * (h->n + h->hx.ghost + h->m + h->hy.ghost);  // xbar / y gathered once
  const int64_t per_elem[6] = {h->m, h->m, h->n, h->n, h->n, h->n};
  for (int bit = 0; bit < 6; ++bit)
    if (!((h->const_mask >> bit) & 1)) vec_bytes += 8 * per_elem[bit];
  out->balanced_split = h->balanced_split ? 1 : 0;
  out->dense_halo = h->dense_halo && h->world > 1 ? 1 : 0;
  out->tiny_persistent = h->tiny ? 1 : (h->cluster.on ? 2 : 0);
  out->long_rows = h->longA.count;
  out->long_cols = h->longAT.count;
  out->long_entries = h->longA.nnz + h->longAT.nnz;
  vec_bytes += 12 * (h->longA.nnz + h->longAT.nnz) + 8 * (h->longA.nnz + h->longAT.nnz);  // entries + their gathers
  // an operand streams either its SELL slices or — banded — its entries once, one count byte per row and one offset
  // per tile and window, and the carries (read + written between consecutive windows of a kind)
  int64_t op_bytes[2];
  const Sell *sell[2] = {&h->A, &h->AT};
  const Band *band[2] = {&h->bandA, &h->bandAT};
  for (int k = 0; k < 2; ++k) {
    const Band &B = *band[k];
    if (B.in_use) {
      const int kinds = (B.geo.eq_windows ? 1 : 0) + (B.geo.windows > B.geo.eq_windows ? 1 : 0);
      op_bytes[k] = 12 * B.nnz + (int64_t)B.geo.windows * (B.rows_pad + B.rows_pad / 32) + 16 * B.nrows * (B.geo.windows - kinds) +
                    (kinds == 2 ? 16 * B.nrows : 0);
    } else {
      op_bytes[k] = sell[k]->padded * entry_bytes + (sell[k]->uniform_width >= 0 ? 0 : 8 * (sell[k]->nslices + 1));
    }
    out->band_windows[k] = B.built ? B.geo.windows : 0;
    out->band_in_use[k] = B.in_use ? 1 : 0;
    out->band_ms[k] = B.ms;
    out->band_sectors_per_gather[k] = (float)B.sectors_per_gather;
    out->band_shape[k] = B.shape;
    memcpy(out->band_shape_ms[k], B.shape_ms, sizeof B.shape_ms);
    if (B.built) out->band_window_bytes = std::max(out->band_window_bytes, B.win_bytes);
  }
  out->bytes_per_iteration_actual = op_bytes[0] + op_bytes[1] + vec_bytes;
  out->value_bytes = h->A.dict ? 0 : 8;
  out->const_vector_mask = h->const_mask;
  out->sm_count = h->sm_count;
  out->world_size = h->world;
  out->rank = h->rank;
  out->n_local = h->n;
  out->m_local = h->m;
  out->m_eq_local = h->m_eq;
  out->n_ghost = h->hx.ghost;
  out->m_ghost = h->hy.ghost;
  out->nnz_local_rows = h->nnz_rows;
  out->nnz_local_cols = h->nnz_cols;
  out->halo_send_bytes_per_iteration = 8 * (h->hx.send_total + h->hy.send_total);
  out->partition_granule = h->granule;
  out->primal_variant = h->primal_variant + 1;
  out->dual_variant = h->dual_variant + 1;
  out->autotuned = h->autotuned ? 1 : 0;
  memcpy(out->variant_ms, h->variant_ms, sizeof out->variant_ms);
  return 0;
}

int64_t cpppd_iteration_count(cpppd_handle h) { return h ? h->niter : -1; }

}  // extern "C"
