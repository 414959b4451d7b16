// Types, constants, error / allocation plumbing and the NCCL loader.
// Part of libcpppd (single translation unit, included by cpppd.cu).
#pragma once

#include "../../include/cpppd.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cub/cub.cuh>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <climits>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "cpppd_device_types.cuh"
#include "cpppd_device.cuh"
#include "cpppd_long_rows.cuh"
#include "cpppd_banded.cuh"

namespace {

thread_local std::string g_create_error;

struct Sell {
  int64_t nrows = 0, nslices = 0, padded = 0;
  int64_t uniform_width = -1;    // >= 0 when every slice has this width (slice_ptr is then implicit)
  int64_t *slice_ptr = nullptr;  // nslices+1 element offsets
  int32_t *idx = nullptr;        // padded entries, kPad = padding
  double *val = nullptr;         // nullptr in dictionary mode
  // dictionary mode (CPPPD_FLAG_VALUE_DICT): the matrix takes <= 256 distinct values; an entry is one
  // 32-bit word  [pad:1][eq:1][code][index]  and its value is dict[code] (the exact original double)
  const double *dict = nullptr;
  int idx_bits = 30, ndict = 0;
};

// One thread-block cluster carries all iterations of a small LP (cpppd_cluster.cuh)
struct ClusterPlan {
  bool on = false;
  int ctas = 0;
  int spc_at = 0, spc_a = 0;  // slices per CTA of A^T / A
  int ent_at = 0, ent_a = 0;  // entries per CTA (max over the CTAs, multiples of 32)
  size_t smem = 0;
};

// Window-major copy of an operand for patterns without locality (cpppd_banded.cuh)
struct Band {
  bool built = false;   // the window-major arrays exist
  bool in_use = false;  // the half-iteration runs k_*_band (one launch per window) instead of the SELL kernel
  BandGeometry geo{0, 1, 1, 0, 0};
  int64_t nrows = 0, nnz = 0, win_bytes = 0;
  int64_t rows_pad = 0;             // nrows rounded up to whole tiles of kBandTile rows
  unsigned char *cnt = nullptr;     // windows * rows_pad : entries of a row in a window
  uint32_t *tile_base = nullptr;    // windows * (rows_pad / kBandTile) + 1 : first entry of a tile in a window (into idx / val)
  int32_t *idx = nullptr;
  double *val = nullptr;
  double *carry = nullptr, *carry_eq = nullptr;  // partial sums between windows (carry_eq: A^T with both row kinds)
  double sectors_per_gather = 0;                  // sampled locality of the operand (k_band_locality)
  float ms = 0.f;                                 // per half-iteration, measured by tune_kernels()
  int shape = 0;                                  // kBandShapes index of the windows before the last one
  float shape_ms[8] = {};
};

struct StatsDev {  // device-resident, copied verbatim into cpppd_stats
  cpppd_stats s;
};

// Halo of one distributed vector: which owned entries go to which peer, where ghosts land.
struct Halo {
  int64_t owned = 0, ghost = 0, send_total = 0;
  std::vector<int64_t> send_count, send_off, recv_count, recv_off;  // per peer rank
  int32_t *send_idx = nullptr;  // send_total local indices (owned part), grouped by peer
  double *send_buf = nullptr;   // send_total staging values
};

// Peer-memory halo exchange (world > 1): the ghost tails of xbar / y live in cudaMalloc'ed memory
// that every neighbour maps through CUDA IPC; a push kernel stores the halo values straight into the
// neighbours' ghost slots over NVLink and then raises a per-neighbour flag, a wait kernel spins on
// the local flags.  No staging buffer, no NCCL call, and the whole iteration is graph-capturable.
struct FusedComm;
struct PeerPtrs {
  double *vec[kMaxWorld];
  unsigned long long *flags[kMaxWorld];
};
struct P2P {
  bool active = false;
  PeerPtrs ptrs[2];                  // [0]: peers' xbar, [1]: peers' y (+ their flag arrays)
  unsigned long long *flags = nullptr;  // 2 * world stamps written by the peers
  SyncState *state = nullptr;
  int32_t *push_peer[2] = {nullptr, nullptr};
  int64_t *push_dst[2] = {nullptr, nullptr};
  unsigned long long send_mask[2] = {0, 0}, recv_mask[2] = {0, 0};
  bool dense[2] = {false, false};        // every peer takes all owned entries, contiguously (k_push_dense)
  int64_t dense_dst[2][kMaxWorld] = {};  // first ghost slot of this rank's entries in every peer's vector
  // halo exchange fused into k_primal ([0], produces xbar) and k_dual ([1], produces y)
  FusedComm *fused[2] = {nullptr, nullptr};
  bool use_fused = false;
  std::vector<void *> opened, own;
};

// NCCL is resolved at run time (dlopen) so that the library loads without it on one GPU.
struct NcclApi {
  void *dl = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

const char *load_nccl() {
  if (g_nccl.dl) return nullptr;
  const char *env = getenv("CPPPD_NCCL_LIB");
  void *dl = nullptr;
  if (env && *env) dl = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  if (!dl) dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!dl) dl = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!dl) return "cannot dlopen libnccl.so.2 (set CPPPD_NCCL_LIB)";
#define SYM(field, name)                                       \
  g_nccl.field = (decltype(g_nccl.field))dlsym(dl, name);      \
  if (!g_nccl.field) return "libnccl lacks symbol " name;
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(AllGather, "ncclAllGather")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.dl = dl;
  return nullptr;
}

}  // namespace

// Peer-memory buffers of a communicator, kept between solves (cpppd_host.cuh: pool_acquire).  A solve used to
// cudaMalloc xbar / y / the stamps, export them, map those of its neighbours (cudaIpcOpenMemHandle) and undo all of
// it in cpppd_destroy — measured on 2 GPUs: 0.12-0.62 s per end-to-end call in the unmap + cudaFree alone.
struct PeerPool {
  bool valid = false, busy = false;
  int world = 0;
  size_t cap_x = 0, cap_y = 0;  // bytes of this rank's two vectors
  double *xbar = nullptr, *y = nullptr;
  unsigned long long *flags = nullptr;
  SyncState *state = nullptr;
  double *peer_x[kMaxWorld] = {}, *peer_y[kMaxWorld] = {};
  unsigned long long *peer_flags[kMaxWorld] = {};
};

struct cpppd_comm_s {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  PeerPool pool;
};

struct cpppd_solver {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // global problem
  int64_t n_glob = 0, m_eq_glob = 0, m_ineq_glob = 0, m_glob = 0, nnz_glob = 0;
  // this rank's share (== global on one GPU)
  int64_t n = 0, m = 0, m_eq = 0, nnz_rows = 0, nnz_cols = 0;
  int rank = 0, world = 1;
  bool identity_layout = true;  // local index == original index (one GPU, no reordering)
  bool tiny = false;            // iterations run in k_tiny_iterate (one persistent CTA)
  ClusterPlan cluster;          // ... or in k_cluster_iterate (one persistent cluster of CTAs)
  bool balanced_split = false;  // ownership by prefix sums instead of locality buckets (see setup())
  bool dense_halo = false;      // patterns without locality: every rank keeps ghosts of ALL foreign columns / rows
  int32_t *col_old = nullptr;   // n + ghosts : original column id of a local column
  int32_t *row_old = nullptr;   // m + ghosts : original row id of a local row
  Halo hx, hy;                  // xbar-like vectors (columns) / y-like vectors (rows)
  ncclComm_t comm = nullptr;
  bool own_comm = true;         // false: borrowed from a cpppd_comm (cpppd_problem.comm)
  cpppd_comm_s *shared = nullptr;  // that cpppd_comm (owner of the peer-memory pool)
  bool pooled = false;          // xbar / y / stamps / peer mappings belong to shared->pool
  P2P p2p;
  double alpha = 1, theta = 1, one_plus_theta = 2;
  uint32_t flags = 0;
  int64_t granule = 0;
  cpppd_alloc_fn alloc = nullptr;
  cpppd_free_fn free_fn = nullptr;
  void *alloc_user = nullptr;
  std::vector<void *> owned;
  int64_t device_bytes = 0;
  Sell A, AT;
  Band bandA, bandAT;            // window-major copies for patterns without locality (cpppd_banded.cuh)
  int64_t band_window = 0;       // elements of the gathered vector per window (cpppd_problem.band_window)
  LongRows longA, longAT;        // rows of A / columns of A cut out of the SELL operands (cpppd_long_rows.cuh)
  int64_t long_threshold = 0;    // rows with more entries are long; < 0: never
  int64_t x_len = 0, y_len = 0;  // allocated length of x-like / y-like vectors (owned + ghosts + long-row tails)
  double *c = nullptr, *T = nullptr, *lb = nullptr, *ub = nullptr, *x = nullptr, *xbar = nullptr;
  double *b = nullptr, *sigma = nullptr, *y = nullptr, *dbuf = nullptr, *best = nullptr;
  double *row_off = nullptr;  // cpppd_set_row_offsets (m, local order) or nullptr
  Vec vc{nullptr, 0}, vT{nullptr, 0}, vlb{nullptr, 0}, vub{nullptr, 0}, vb{nullptr, 0}, vsigma{nullptr, 0};
  int const_mask = 0;               // bit0 b, bit1 sigma, bit2 lb, bit3 ub, bit4 c, bit5 T folded to scalars
  unsigned long long *dict = nullptr;  // sorted bit patterns of the distinct matrix values (dictionary mode)
  int ndict = 0;
  double *colpart = nullptr, *rowpart = nullptr, *xr_scratch = nullptr;
  double *stat_local = nullptr, *stat_all = nullptr;  // kStatQ / world*kStatQ
  int32_t *gt_idx = nullptr;    // local column ids of the ground-truth entries this rank owns
  double *gt_val = nullptr, *gtpart = nullptr;
  int64_t gt_local = 0, gt_total = 0;
  int stat_blocks_g = 0;
  int stat_blocks_c = 0, stat_blocks_r = 0;
  StatsDev *stats_dev = nullptr;
  cpppd_stats *stats_host = nullptr;
  int64_t niter = 0;
  bool mid_iteration = false;  // primal step issued, dual step pending
  bool stats_pending = false;
  bool have_d = false;
  int sm_count = 148;
  std::map<int64_t, cudaGraphExec_t> graphs;
  // kernel variants (0-based indices into kVariants): chosen at creation, see tune_kernels()
  int variant_request = 0;  // cpppd_problem.kernel_variant
  int primal_variant = 0, dual_variant = 0;
  bool autotuned = false;
  float variant_ms[2][CPPPD_KERNEL_VARIANTS] = {};
  std::string err;
  int sticky = 0;
};

namespace {

int fail(cpppd_solver *h, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) {
    h->err = buf;
    if (code == CPPPD_ERR_CUDA || code == CPPPD_ERR_COMM) h->sticky = code;
  }
  g_create_error = buf;
  return code;
}

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(h, CPPPD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                        \
  } while (0)

#define NK(call)                                                                                    \
  do {                                                                                              \
    ncclResult_t r_ = (call);                                                                       \
    if (r_ != ncclSuccess)                                                                          \
      return fail(h, CPPPD_ERR_COMM, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_),    \
                  __FILE__, __LINE__);                                                              \
  } while (0)

#define CHECK_HANDLE(h)                                  \
  do {                                                   \
    if (!(h)) return CPPPD_ERR_INVALID;                  \
    if ((h)->sticky) return (h)->sticky;                 \
    cudaSetDevice((h)->device);                          \
  } while (0)

void *dev_alloc(cpppd_solver *h, size_t bytes, bool persistent) {
  if (bytes == 0) bytes = 256;
  void *p = nullptr;
  if (h->alloc) {
    p = h->alloc(bytes, h->alloc_user);
  } else if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    p = nullptr;
  }
  if (p && persistent) {
    h->owned.push_back(p);
    h->device_bytes += (int64_t)bytes;
  }
  return p;
}

void dev_free(cpppd_solver *h, void *p) {
  if (!p) return;
  if (h->alloc) {
    if (h->free_fn) h->free_fn(p, h->alloc_user);
  } else {
    cudaFree(p);
  }
}

template <typename T>
int alloc_array(cpppd_solver *h, T **out, int64_t count, bool persistent = true) {
  *out = static_cast<T *>(dev_alloc(h, sizeof(T) * (size_t)std::max<int64_t>(count, 1), persistent));
  if (!*out) return fail(h, CPPPD_ERR_NOMEM, "device allocation of %lld bytes failed", (long long)(sizeof(T) * count));
  return 0;
}

// temporaries of setup(): freed on scope exit
struct Scratch {
  cpppd_solver *h;
  std::vector<void *> ptrs;
  explicit Scratch(cpppd_solver *h_) : h(h_) {}
  ~Scratch() { for (void *p : ptrs) dev_free(h, p); }
  template <typename T>
  int get(T **out, int64_t count) {
    int rc = alloc_array(h, out, count, false);
    if (!rc) ptrs.push_back(*out);
    return rc;
  }
  // frees now and NULLs the caller's variable: the allocator may hand the same address out again,
  // so a stale copy of the pointer must never reach release() a second time
  template <typename T>
  void release(T *&p) {
    if (!p) return;
    for (auto &q : ptrs)
      if (q == (void *)p) {
        dev_free(h, q);
        q = nullptr;
        break;
      }
    p = nullptr;
  }
};

// CUDA events that are destroyed on every way out of a function (the CK / NK macros return early)
struct Events {
  std::vector<cudaEvent_t> ev;
  explicit Events(size_t count) : ev(count, nullptr) {}
  ~Events() {
    for (cudaEvent_t e : ev)
      if (e) cudaEventDestroy(e);
  }
  cudaError_t create() {
    for (auto &e : ev)
      if (cudaError_t err = cudaEventCreate(&e)) return err;
    return cudaSuccess;
  }
  cudaEvent_t operator[](size_t i) const { return ev[i]; }
};

inline int grid_for(int64_t items, int block = kBlock) { return (int)std::max<int64_t>(1, (items + block - 1) / block); }

}  // namespace
