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
#include "../../include/cpppd.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <climits>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

constexpr int kSlice = 32;       // SELL slice height C (= warp size)
constexpr int kBlock = 256;      // threads per CTA for the streaming kernels (8 slices)
constexpr int kGraphChunk = 50;  // iterations captured per CUDA graph
constexpr int kColQ = 4;         // column-pass partial sums per CTA
constexpr int kRowQ = 7;         // row-pass partial sums per CTA
constexpr int kStatQ = kColQ + kRowQ;
constexpr int32_t kEqBit = 0x40000000;   // A^T entries: set when the source row is an equality
constexpr int32_t kIdxMask = 0x3fffffff;
// padding entry of a slice: negative, and its masked index is 0 so that a gather the compiler
// hoists above the `idx >= 0` test still reads a valid address
constexpr int32_t kPad = INT32_MIN;
constexpr int kMaxWorld = 64;

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

struct SellView {
  const int64_t *__restrict__ slice_ptr;
  const int32_t *__restrict__ idx;
  const double *__restrict__ val;
  int64_t nrows, nslices;
  int64_t uniform_width;  // -1: read slice_ptr
  const double *__restrict__ dict;
  int32_t idx_mask;       // low bits of an entry word that hold the gather index
  int32_t idx_bits, code_mask, ndict;
};

// a vector operand that may have been folded into a scalar (CPPPD_FLAG_CONST_VECTORS)
struct Vec {
  const double *p;
  double c;
  __device__ __forceinline__ double at(int64_t i) const { return p ? __ldcs(p + i) : c; }
};

// first / one-past-last element offset of slice s
__device__ __forceinline__ void slice_range(const SellView &S, int64_t s, int64_t &p0, int64_t &p1) {
  if (S.uniform_width >= 0) {
    p0 = s * S.uniform_width * 32;
    p1 = p0 + S.uniform_width * 32;
  } else {
    p0 = __ldg(S.slice_ptr + s);
    p1 = __ldg(S.slice_ptr + s + 1);
  }
}

// value of the entry stored at position p whose index word is w (non-hot kernels)
__device__ __forceinline__ double entry_value(const SellView &S, int64_t p, int32_t w) {
  return S.dict ? S.dict[(w >> S.idx_bits) & S.code_mask] : S.val[p];
}

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
struct PeerPtrs {
  double *vec[kMaxWorld];
  unsigned long long *flags[kMaxWorld];
};
struct SyncState {
  unsigned long long push_stamp[2];  // halos pushed so far        ([0] xbar, [1] y)
  unsigned long long wait_stamp[2];  // halos consumed so far
  unsigned int ticket[2];            // CTA arrival counter of k_push
};
struct P2P {
  bool active = false;
  PeerPtrs ptrs[2];                  // [0]: peers' xbar, [1]: peers' y (+ their flag arrays)
  unsigned long long *flags = nullptr;  // 2 * world stamps written by the peers
  SyncState *state = nullptr;
  int32_t *push_peer[2] = {nullptr, nullptr};
  int64_t *push_dst[2] = {nullptr, nullptr};
  unsigned long long send_mask[2] = {0, 0}, recv_mask[2] = {0, 0};
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
  int32_t *col_old = nullptr;   // n + ghosts : original column id of a local column
  int32_t *row_old = nullptr;   // m + ghosts : original row id of a local row
  Halo hx, hy;                  // xbar-like vectors (columns) / y-like vectors (rows)
  ncclComm_t comm = nullptr;
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
  double *c = nullptr, *T = nullptr, *lb = nullptr, *ub = nullptr, *x = nullptr, *xbar = nullptr;
  double *b = nullptr, *sigma = nullptr, *y = nullptr, *dbuf = nullptr, *best = nullptr;
  Vec vc{nullptr, 0}, vT{nullptr, 0}, vlb{nullptr, 0}, vub{nullptr, 0}, vb{nullptr, 0}, vsigma{nullptr, 0};
  int const_mask = 0;               // bit0 b, bit1 sigma, bit2 lb, bit3 ub, bit4 c, bit5 T folded to scalars
  unsigned long long *dict = nullptr;  // sorted bit patterns of the distinct matrix values (dictionary mode)
  int ndict = 0;
  double *colpart = nullptr, *rowpart = nullptr, *xr_scratch = nullptr;
  double *stat_local = nullptr, *stat_all = nullptr;  // kStatQ / world*kStatQ
  int stat_blocks_c = 0, stat_blocks_r = 0;
  StatsDev *stats_dev = nullptr;
  cpppd_stats *stats_host = nullptr;
  int64_t niter = 0;
  bool mid_iteration = false;  // primal step issued, dual step pending
  bool stats_pending = false;
  bool have_d = false;
  int sm_count = 148;
  std::map<int64_t, cudaGraphExec_t> graphs;
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

inline int grid_for(int64_t items, int block = kBlock) { return (int)std::max<int64_t>(1, (items + block - 1) / block); }

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double nan_max(double a, double b) {
  // numpy.max semantics: NaN wins
  if (a != a) return a;
  if (b != b) return b;
  return a > b ? a : b;
}

__device__ __forceinline__ double abs_pow(double a, double p) {
  // numpy: np.abs(data) ** p.  numpy special-cases the scalar exponents 1 and 2 (exact), so do we.
  double v = fabs(a);
  if (p == 1.0) return v;
  if (p == 2.0) return __dmul_rn(v, v);
  if (p == 0.0) return 1.0;
  return pow(v, p);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_nanmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = nan_max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Reduce Q per-thread values over the CTA; thread 0 writes them to out[0..Q).
// is_max bit q set -> NaN-propagating max, else sum.
template <int Q>
__device__ __forceinline__ void block_reduce_write(double (&v)[Q], unsigned is_max, double *out) {
  __shared__ double sh[Q][kBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    double r = (is_max >> q) & 1u ? warp_nanmax(v[q]) : warp_sum(v[q]);
    if (lane == 0) sh[q][warp] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      double r = sh[q][0];
      for (int w = 1; w < kBlock / 32; ++w)
        r = (is_max >> q) & 1u ? nan_max(r, sh[q][w]) : __dadd_rn(r, sh[q][w]);
      out[q] = r;
    }
  }
}

// one warp per slice: width = longest row of the slice; out[s] = 32 * width
__global__ void k_slice_extent(const int64_t *__restrict__ rowptr, int64_t nrows, int64_t nslices,
                               int64_t *__restrict__ extent) {
  int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= nslices) return;
  int64_t r = s * kSlice + lane;
  int64_t len = r < nrows ? rowptr[r + 1] - rowptr[r] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if (lane == 0) extent[s] = len * kSlice;
}

// position of `v` (compared by bit pattern) in the sorted dictionary, or -1
__device__ __forceinline__ int dict_find(const unsigned long long *__restrict__ dict, int ndict, double v) {
  const unsigned long long key = (unsigned long long)__double_as_longlong(v);
  int lo = 0, hi = ndict - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const unsigned long long d = dict[mid];
    if (d == key) return mid;
    if (d < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// one warp per slice: copy CSR entries into the column-major slice, pad with idx = kPad.
// With a dictionary the value is folded into the index word as a code and `val` is not written.
__global__ void k_fill_sell(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ indices,
                            const double *__restrict__ values, int64_t nrows, int64_t nslices,
                            const int64_t *__restrict__ slice_ptr, int32_t *__restrict__ idx,
                            double *__restrict__ val, const unsigned long long *__restrict__ dict, int ndict,
                            int idx_bits) {
  int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= nslices) return;
  int64_t r = s * kSlice + lane;
  int64_t p0 = slice_ptr[s], p1 = slice_ptr[s + 1];
  int64_t e0 = 0, len = 0;
  if (r < nrows) {
    e0 = rowptr[r];
    len = rowptr[r + 1] - e0;
  }
  int64_t width = (p1 - p0) / kSlice;
  for (int64_t k = 0; k < width; ++k) {
    int64_t p = p0 + k * kSlice + lane;
    if (k < len) {
      int32_t w = indices[e0 + k];
      if (dict) {
        const int code = dict_find(dict, ndict, values[e0 + k]);
        w = (w & kEqBit) | (w & ((1 << idx_bits) - 1)) | (code << idx_bits);
      } else {
        val[p] = values[e0 + k];
      }
      idx[p] = w;
    } else {
      idx[p] = kPad;
      if (!dict) val[p] = 0.0;
    }
  }
}

// flag[0] = 1 when some value is not in the dictionary
__global__ void k_dict_check(const double *__restrict__ values, int64_t nnz, const unsigned long long *__restrict__ dict,
                             int ndict, int *__restrict__ flag) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
    if (dict_find(dict, ndict, values[e]) < 0) *flag = 1;
}

// flag[0] = 1 when some element differs (bitwise) from the first one
__global__ void k_not_constant(const double *__restrict__ v, int64_t count, int *__restrict__ flag) {
  const long long first = __double_as_longlong(v[0]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    if (__double_as_longlong(v[i]) != first) *flag = 1;
}

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

// sort key of a row / column: (owner, [is_ineq,] bucket); also counts per owner
__global__ void k_sort_keys(const int32_t *__restrict__ key, int64_t count, int32_t granule,
                            const int32_t *__restrict__ owner_of_bucket, int64_t m_eq, int is_rows,
                            const int64_t *__restrict__ rowptr, const int32_t *__restrict__ len32,
                            uint64_t *__restrict__ out_key, uint32_t *__restrict__ out_id,
                            int32_t *__restrict__ count_per_owner, int32_t *__restrict__ eq_per_owner) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  int32_t q = key[i] / granule;
  int32_t o = owner_of_bucket[q];
  // inside a bucket, rows / columns of equal length sit together (SELL sigma-sorting: slices of
  // 32 neighbours then have nearly equal widths and little padding)
  int64_t len = rowptr ? rowptr[i + 1] - rowptr[i] : (int64_t)len32[i];
  uint64_t len12 = (uint64_t)(len > 4095 ? 4095 : len);
  uint64_t major = is_rows ? (uint64_t)o * 2 + (i >= m_eq ? 1 : 0) : (uint64_t)o;
  out_key[i] = (major << 44) | ((uint64_t)(uint32_t)q << 12) | len12;
  out_id[i] = (uint32_t)i;
  atomicAdd(count_per_owner + o, 1);
  if (is_rows && i < m_eq) atomicAdd(eq_per_owner + o, 1);
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

// what this rank must send to peer t: its columns hit by t's rows, its rows hitting t's columns
__global__ void k_mark_sends(const int32_t *__restrict__ indices, const uint32_t *__restrict__ row_of, int64_t nnz,
                             const int32_t *__restrict__ row_pos, const int32_t *__restrict__ col_pos, int32_t rs,
                             int32_t re, int32_t cs, int32_t ce, int32_t trs, int32_t tre, int32_t tcs, int32_t tce,
                             int32_t *__restrict__ sendx_flag, int32_t *__restrict__ sendy_flag) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int32_t rp = row_pos[row_of[e]], cp = col_pos[indices[e]];
  if (cp >= cs && cp < ce && rp >= trs && rp < tre) sendx_flag[cp - cs] = 1;
  if (rp >= rs && rp < re && cp >= tcs && cp < tce) sendy_flag[rp - rs] = 1;
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

// ------------------------------------------------------------------------------------------
// the two hot kernels
// ------------------------------------------------------------------------------------------
// Primal half-iteration (:198-228).  Thread j owns column j of A (row j of A^T).
// Loads that do not depend on the matrix (c, T, x) are issued first so that they are in flight
// together with the slice entries; matrix entries are read once (ld.global.cs).
// kDict: entries are single 32-bit words [pad][eq][code][index]; values come from a <= 256 entry
// dictionary staged in shared memory.
template <bool kWriteD, bool kDict>
__global__ void __launch_bounds__(kBlock, 8)
k_primal(SellView AT, const double *__restrict__ y, Vec c, Vec T, Vec lb, Vec ub, double *__restrict__ x,
         double *__restrict__ xbar, double *__restrict__ d_out, int64_t n, int has_eq, int has_ineq,
         double theta, double one_plus_theta) {
  __shared__ double sdict[kDict ? 256 : 1];
  if (kDict) {
    if ((int)threadIdx.x < AT.ndict) sdict[threadIdx.x] = AT.dict[threadIdx.x];
    __syncthreads();
  }
  const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t s = j >> 5;
  if (s >= AT.nslices) return;
  const int lane = threadIdx.x & 31;
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
#pragma unroll 4
    for (int k = 0; k < width; ++k) {
      const int32_t r = __ldcs(ip + k * kSlice);
      double a;
      if (kDict) a = sdict[(r >> AT.idx_bits) & AT.code_mask]; else a = __ldcs(vp + k * kSlice);
      if (r >= 0) {
        const double t = __dmul_rn(a, __ldg(y + (r & mask)));
        if (r & kEqBit) s_eq = __dadd_rn(s_eq, t); else s_in = __dadd_rn(s_in, t);
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
  xbar[j] = __dsub_rn(__dmul_rn(one_plus_theta, x2), __dmul_rn(theta, xo));
  x[j] = x2;
  if (kWriteD) d_out[j] = d;
}

// Dual half-iteration (:231-240, :333-341).  Thread i owns row i of A.
template <bool kDict>
__global__ void __launch_bounds__(kBlock, 8)
k_dual(SellView A, const double *__restrict__ xbar, Vec b, Vec sigma, double *__restrict__ y, int64_t m,
       int64_t m_eq) {
  __shared__ double sdict[kDict ? 256 : 1];
  if (kDict) {
    if ((int)threadIdx.x < A.ndict) sdict[threadIdx.x] = A.dict[threadIdx.x];
    __syncthreads();
  }
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t s = i >> 5;
  if (s >= A.nslices) return;
  const int lane = threadIdx.x & 31;
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
#pragma unroll 4
    for (int k = 0; k < width; ++k) {
      const int32_t jc = __ldcs(ip + k * kSlice);
      double a;
      if (kDict) a = sdict[(jc >> A.idx_bits) & A.code_mask]; else a = __ldcs(vp + k * kSlice);
      if (jc >= 0) acc = __dadd_rn(acc, __dmul_rn(a, __ldg(xbar + (jc & mask))));
    }
  }
  if (!live) return;
  const double r = __dsub_rn(acc, bi);
  double yn = __dadd_rn(yi, __dmul_rn(si, r));
  if (i >= m_eq) yn = (yn < 0.0) ? 0.0 : yn;  // np.maximum(y_ineq, 0): NaN stays NaN, -0.0 stays
  y[i] = yn;
}

// ------------------------------------------------------------------------------------------
// stats block (:248-291)
// ------------------------------------------------------------------------------------------
// Column pass: c.x, c.x4, c.xr, #(xbar == 0); turns the d buffer into x4 in place and
// (force_integer) stores xr into xr_out.
__global__ void __launch_bounds__(kBlock)
k_stats_cols(Vec c, const double *__restrict__ x, const double *__restrict__ xbar, Vec lb, Vec ub,
             double *__restrict__ d_x4, double *__restrict__ xr_out, int64_t n, int force_integer,
             double *__restrict__ part) {
  double v[kColQ] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x; j < n; j += (int64_t)gridDim.x * kBlock) {
    const double cj = c.at(j), xj = x[j];
    const double x4 = d_x4[j] < 0.0 ? ub.at(j) : lb.at(j);  // x4 = lb; x4[d < 0] = ub[d < 0]  (:260-261)
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
  }
  block_reduce_write<kColQ>(v, 0u, part + (int64_t)blockIdx.x * kColQ);
}

// Row pass: A x, A x4, A xbar, A xr per row -> energy terms and violation maxima.
__global__ void __launch_bounds__(kBlock)
k_stats_rows(SellView A, const double *__restrict__ x, const double *__restrict__ x4,
             const double *__restrict__ xbar, const double *__restrict__ xr, Vec b,
             const double *__restrict__ y, int64_t m, int64_t m_eq, int force_integer,
             double *__restrict__ part) {
  const double ninf = -INFINITY;
  double v[kRowQ] = {0.0, 0.0, 0.0, 0.0, ninf, ninf, ninf};
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
    }
  }
  block_reduce_write<kRowQ>(v, 0x70u, part + (int64_t)blockIdx.x * kRowQ);
}

// One CTA: fold this rank's per-CTA partials in a fixed order into kStatQ numbers.
__global__ void __launch_bounds__(kBlock)
k_stats_local(const double *__restrict__ colpart, int nbc, const double *__restrict__ rowpart, int nbr,
              double *__restrict__ out) {
  double cv[kColQ] = {0.0, 0.0, 0.0, 0.0};
  const double ninf = -INFINITY;
  double rv[kRowQ] = {0.0, 0.0, 0.0, 0.0, ninf, ninf, ninf};
  for (int bi = threadIdx.x; bi < nbc; bi += kBlock)
#pragma unroll
    for (int q = 0; q < kColQ; ++q) cv[q] = __dadd_rn(cv[q], colpart[(int64_t)bi * kColQ + q]);
  for (int bi = threadIdx.x; bi < nbr; bi += kBlock) {
#pragma unroll
    for (int q = 0; q < 4; ++q) rv[q] = __dadd_rn(rv[q], rowpart[(int64_t)bi * kRowQ + q]);
#pragma unroll
    for (int q = 4; q < kRowQ; ++q) rv[q] = nan_max(rv[q], rowpart[(int64_t)bi * kRowQ + q]);
  }
  __shared__ double fin[kStatQ];
  block_reduce_write<kColQ>(cv, 0u, fin);
  __syncthreads();
  block_reduce_write<kRowQ>(rv, 0x70u, fin + kColQ);
  __syncthreads();
  if (threadIdx.x < kStatQ) out[threadIdx.x] = fin[threadIdx.x];
}

// One thread: fold the ranks' numbers in rank order, then apply :248-291's scalar logic.
__global__ void k_stats_final(const double *__restrict__ all, int world, int64_t n_glob, int has_eq, int has_ineq,
                              int64_t niter, StatsDev *out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double fin[kStatQ];
  for (int q = 0; q < kStatQ; ++q) fin[q] = all[q];
  for (int r = 1; r < world; ++r)
    for (int q = 0; q < kStatQ; ++q) {
      const double v = all[r * kStatQ + q];
      fin[q] = (q >= kColQ + 4) ? nan_max(fin[q], v) : __dadd_rn(fin[q], v);
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
  s.energy_rounded = fin[2];
  s.frac_zero_xbar = n_glob > 0 ? fin[3] / (double)n_glob : 0.0;
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

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Halo push over peer memory: entry k of the send list goes to peer push_peer[k], element
// push_dst[k] of that peer's vector (its ghost slot).  The last CTA to finish raises, on every
// neighbour it sent to, the flag [kind * world + me] to the new stamp (release at system scope after
// every CTA fenced its stores).
__global__ void __launch_bounds__(kBlock)
k_push(const double *__restrict__ vec, const int32_t *__restrict__ src, const int64_t *__restrict__ dst,
       const int32_t *__restrict__ peer, int64_t count, PeerPtrs P, int kind, int world, int me,
       unsigned long long send_mask, SyncState *st) {
  const int64_t k = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (k < count) P.vec[peer[k]][dst[k]] = vec[src[k]];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x != 0) return;
  const unsigned int ticket = atomicAdd(&st->ticket[kind], 1u);
  if (ticket != gridDim.x - 1) return;
  __threadfence_system();
  st->ticket[kind] = 0;
  const unsigned long long stamp = st->push_stamp[kind] + 1;
  st->push_stamp[kind] = stamp;
  for (int t = 0; t < world; ++t)
    if ((send_mask >> t) & 1ull) st_release_sys(P.flags[t] + kind * world + me, stamp);
}

// Wait until every neighbour this rank receives from has pushed its halo for this exchange.
__global__ void k_wait(const unsigned long long *__restrict__ flags, int kind, int world,
                       unsigned long long recv_mask, SyncState *st) {
  const int t = threadIdx.x;
  const unsigned long long want = st->wait_stamp[kind] + 1;
  if (t < world && ((recv_mask >> t) & 1ull)) {
    while (ld_acquire_sys(flags + kind * world + t) < want) __nanosleep(200);
  }
  __syncthreads();
  if (t == 0) st->wait_stamp[kind] = want;
}

// halo staging: buf[k] = vec[idx[k]]
__global__ void k_pack(const double *__restrict__ vec, const int32_t *__restrict__ idx, int64_t count,
                       double *__restrict__ buf) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) buf[k] = vec[idx[k]];
}

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
  if (int rc = tmp.get(&full, full_count)) return rc;
  if (int rc = upload_f64(h, full, host_full, full_count)) return rc;
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
  if (num_h >= 1 && num_h <= 256) {
    if (int rc = alloc_array(h, &h->dict, 256)) return rc;
    CK(cudaMemsetAsync(h->dict, 0, sizeof(unsigned long long) * 256, st));
    CK(cudaMemcpyAsync(h->dict, uniq, sizeof(unsigned long long) * num_h, cudaMemcpyDeviceToDevice, st));
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
  CK(cudaMalloc(&pp.flags, sizeof(unsigned long long) * 2 * N));
  pp.own.push_back(pp.flags);
  CK(cudaMalloc(&pp.state, sizeof(SyncState)));
  pp.own.push_back(pp.state);
  CK(cudaMemsetAsync(pp.flags, 0, sizeof(unsigned long long) * 2 * N, st));
  CK(cudaMemsetAsync(pp.state, 0, sizeof(SyncState), st));
  PeerRecord mine;
  memset(&mine, 0, sizeof mine);
  CK(cudaIpcGetMemHandle(&mine.xbar, h->xbar));
  CK(cudaIpcGetMemHandle(&mine.y, h->y));
  CK(cudaIpcGetMemHandle(&mine.flags, pp.flags));
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
  memset(pp.ptrs, 0, sizeof pp.ptrs);
  for (int t = 0; t < N; ++t) {
    if (t == me) continue;
    const bool nb = h->hx.send_count[t] || h->hx.recv_count[t] || h->hy.send_count[t] || h->hy.recv_count[t];
    if (!nb) continue;
    void *px = nullptr, *py = nullptr, *pf = nullptr;
    cudaError_t e1 = cudaIpcOpenMemHandle(&px, all[t].xbar, cudaIpcMemLazyEnablePeerAccess);
    cudaError_t e2 = e1 == cudaSuccess ? cudaIpcOpenMemHandle(&py, all[t].y, cudaIpcMemLazyEnablePeerAccess) : e1;
    cudaError_t e3 = e2 == cudaSuccess ? cudaIpcOpenMemHandle(&pf, all[t].flags, cudaIpcMemLazyEnablePeerAccess) : e2;
    if (e3 != cudaSuccess) {
      cudaGetLastError();
      return fail(h, CPPPD_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s (use CPPPD_FLAG_NO_P2P for the NCCL path)",
                  t, cudaGetErrorString(e3));
    }
    pp.opened.insert(pp.opened.end(), {px, py, pf});
    pp.ptrs[0].vec[t] = (double *)px;
    pp.ptrs[1].vec[t] = (double *)py;
    pp.ptrs[0].flags[t] = pp.ptrs[1].flags[t] = (unsigned long long *)pf;
  }
  // per-entry destinations of the two send lists
  for (int kind = 0; kind < 2; ++kind) {
    Halo &H = kind ? h->hy : h->hx;
    std::vector<int32_t> peer(H.send_total);
    std::vector<int64_t> dst(H.send_total);
    for (int t = 0; t < N; ++t) {
      if (H.send_count[t]) pp.send_mask[kind] |= 1ull << t;
      if (H.recv_count[t]) pp.recv_mask[kind] |= 1ull << t;
      const int64_t base = (kind ? all[t].owned_y : all[t].owned_x) + (kind ? all[t].recv_off_y[me] : all[t].recv_off_x[me]);
      for (int64_t k = 0; k < H.send_count[t]; ++k) {
        peer[H.send_off[t] + k] = t;
        dst[H.send_off[t] + k] = base + k;
      }
    }
    if (int rc = alloc_array(h, &pp.push_peer[kind], H.send_total)) return rc;
    if (int rc = alloc_array(h, &pp.push_dst[kind], H.send_total)) return rc;
    if (H.send_total) {
      CK(cudaMemcpy(pp.push_peer[kind], peer.data(), sizeof(int32_t) * H.send_total, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(pp.push_dst[kind], dst.data(), sizeof(int64_t) * H.send_total, cudaMemcpyHostToDevice));
    }
  }
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
  if (H.send_total)
    k_push<<<grid_for(H.send_total), kBlock, 0, h->stream>>>(vec, H.send_idx, pp.push_dst[kind], pp.push_peer[kind],
                                                            H.send_total, pp.ptrs[kind], kind, h->world, h->rank,
                                                            pp.send_mask[kind], pp.state);
  if (pp.recv_mask[kind]) k_wait<<<1, kMaxWorld, 0, h->stream>>>(pp.flags, kind, h->world, pp.recv_mask[kind], pp.state);
  return 0;
}

int setup(cpppd_solver *h, const cpppd_problem *P) {
  const int64_t n = h->n_glob, m = h->m_glob, nnz = h->nnz_glob, m_eq = h->m_eq_glob;
  const int N = h->world, me = h->rank;
  cudaStream_t st = h->stream;
  Scratch tmp(h);
  // ---- CSR of the whole A on the device (temporary; every rank analyses the same pattern)
  int64_t *rowptr = nullptr;
  int32_t *indices = nullptr;
  double *values = nullptr;
  if (int rc = tmp.get(&rowptr, m + 1)) return rc;
  if (int rc = tmp.get(&indices, nnz)) return rc;
  if (int rc = tmp.get(&values, nnz)) return rc;
  if (P->indptr_bits == 64) {
    CK(cudaMemcpyAsync(rowptr, P->indptr, sizeof(int64_t) * (m + 1), cudaMemcpyHostToDevice, st));
  } else {
    int32_t *tmp32 = nullptr;
    if (int rc = tmp.get(&tmp32, m + 1)) return rc;
    CK(cudaMemcpyAsync(tmp32, P->indptr, sizeof(int32_t) * (m + 1), cudaMemcpyHostToDevice, st));
    k_widen_indptr<<<grid_for(m + 1), kBlock, 0, st>>>(tmp32, rowptr, m + 1);
    CK(cudaStreamSynchronize(st));
    tmp.release(tmp32);
  }
  if (nnz) {
    CK(cudaMemcpyAsync(indices, P->indices, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(values, P->values, sizeof(double) * nnz, cudaMemcpyHostToDevice, st));
  }
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
  if ((h->flags & CPPPD_FLAG_VALUE_DICT) && nnz)
    if (int rc = detect_dictionary(h, tmp, values, nnz)) return rc;
  uint32_t *row_of = nullptr, *entry_id = nullptr;
  if (int rc = tmp.get(&row_of, nnz)) return rc;
  if (int rc = tmp.get(&entry_id, nnz)) return rc;
  if (nnz) k_row_of_entry<<<grid_for(nnz), kBlock, 0, st>>>(rowptr, m, nnz, row_of, entry_id);

  bool reorder = N > 1 || (h->flags & CPPPD_FLAG_REORDER);
  if (!reorder && nnz && !(h->flags & CPPPD_FLAG_NO_REORDER)) {
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
    // ---- locality keys -> buckets -> owners (oracle/partition_oracle.py restates this block)
    const int64_t G = h->granule > 0 ? h->granule : default_granule(n);
    h->granule = G;
    const int64_t nb = n / G + 2;
    int32_t *row_key = nullptr, *col_key = nullptr, *col_len = nullptr, *owner_dev = nullptr, *counts = nullptr;
    unsigned long long *work = nullptr;
    if (int rc = tmp.get(&row_key, m)) return rc;
    if (int rc = tmp.get(&col_key, n)) return rc;
    if (int rc = tmp.get(&col_len, n)) return rc;
    if (int rc = tmp.get(&work, nb)) return rc;
    if (int rc = tmp.get(&owner_dev, nb)) return rc;
    if (int rc = tmp.get(&counts, 3 * (int64_t)N)) return rc;
    if (m) k_row_key<<<grid_for(m), kBlock, 0, st>>>(rowptr, indices, m, (int32_t)n, row_key);
    if (n) k_fill_i32<<<grid_for(n), kBlock, 0, st>>>(col_key, n, (int32_t)n);
    CK(cudaMemsetAsync(col_len, 0, sizeof(int32_t) * std::max<int64_t>(n, 1), st));
    if (nnz) k_col_key<<<grid_for(nnz), kBlock, 0, st>>>(indices, row_of, nnz, row_key, col_key, col_len);
    CK(cudaMemsetAsync(work, 0, sizeof(unsigned long long) * nb, st));
    if (m) k_bucket_work<<<grid_for(m), kBlock, 0, st>>>(row_key, rowptr, nullptr, m, (int32_t)G, work);
    if (n) k_bucket_work<<<grid_for(n), kBlock, 0, st>>>(col_key, nullptr, col_len, n, (int32_t)G, work);
    std::vector<unsigned long long> work_h(nb);
    CK(cudaMemcpyAsync(work_h.data(), work, sizeof(unsigned long long) * nb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<int32_t> owner_h(nb, 0);
    unsigned long long total = 0, before = 0;
    for (auto w : work_h) total += w;
    for (int64_t q = 0; q < nb; ++q) {
      owner_h[q] = total ? (int32_t)std::min<unsigned long long>(N - 1, (unsigned __int128)before * N / total) : 0;
      before += work_h[q];
    }
    CK(cudaMemcpyAsync(owner_dev, owner_h.data(), sizeof(int32_t) * nb, cudaMemcpyHostToDevice, st));
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
    if (m) k_sort_keys<<<grid_for(m), kBlock, 0, st>>>(row_key, m, (int32_t)G, owner_dev, m_eq, 1, rowptr, nullptr, rk_a, ro_a, counts, counts + N);
    if (n) k_sort_keys<<<grid_for(n), kBlock, 0, st>>>(col_key, n, (int32_t)G, owner_dev, 0, 0, nullptr, col_len, ck_a, co_a, counts + 2 * N, nullptr);
    const int end_bit = 44 + bits_for((uint64_t)2 * N + 1);
    cub::DoubleBuffer<uint64_t> rk(rk_a, rk_b), ck(ck_a, ck_b);
    cub::DoubleBuffer<uint32_t> rov(ro_a, ro_b), cov(co_a, co_b);
    if (int rc = sort_pairs(h, rk, rov, m, end_bit)) return rc;
    if (int rc = sort_pairs(h, ck, cov, n, end_bit)) return rc;
    row_order = rov.Current();
    col_order = cov.Current();
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
    tmp.release(row_key); tmp.release(col_key); tmp.release(col_len); tmp.release(work);
    // ---- ghosts of this rank
    int32_t *gcol_flag = nullptr, *grow_flag = nullptr;
    if (int rc = tmp.get(&gcol_flag, n + 1)) return rc;
    if (int rc = tmp.get(&grow_flag, m + 1)) return rc;
    if (int rc = tmp.get(&gcol_scan, n + 1)) return rc;
    if (int rc = tmp.get(&grow_scan, m + 1)) return rc;
    CK(cudaMemsetAsync(gcol_flag, 0, sizeof(int32_t) * (n + 1), st));
    CK(cudaMemsetAsync(grow_flag, 0, sizeof(int32_t) * (m + 1), st));
    if (nnz && N > 1) k_mark_ghosts<<<grid_for(nnz), kBlock, 0, st>>>(indices, row_of, nnz, row_pos, col_pos, rs, re, cs, ce, gcol_flag, grow_flag);
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
      for (int t = 0; t < N; ++t) {
        if (t == me) continue;
        CK(cudaMemsetAsync(sx_flag, 0, sizeof(int32_t) * (nloc + 1), st));
        CK(cudaMemsetAsync(sy_flag, 0, sizeof(int32_t) * (mloc + 1), st));
        if (nnz) k_mark_sends<<<grid_for(nnz), kBlock, 0, st>>>(indices, row_of, nnz, row_pos, col_pos, rs, re, cs, ce,
                                                              row_start[t], row_start[t + 1], col_start[t], col_start[t + 1],
                                                              sx_flag, sy_flag);
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
    const int idx_bits = bits_for((uint64_t)std::max<int64_t>(std::max(nloc + n_ghost, mloc + m_ghost), 2) - 1);
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
  // ---- this rank's rows of A -> SELL-32
  if (!reorder) {
    h->nnz_rows = nnz;
    if (int rc = build_sell(h, rowptr, indices, values, m, &h->A)) return rc;
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
    if (int rc = build_sell(h, lrowptr, lidx, lval, mloc, &h->A)) return rc;
    tmp.release(len); tmp.release(lrowptr); tmp.release(lidx); tmp.release(lval);
  }
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
    if (int rc = build_sell(h, lcolptr, t_idx, t_val, nloc, &h->AT)) return rc;
    tmp.release(lcolptr); tmp.release(t_idx); tmp.release(t_val);
  }
  // ---- vectors in local layout
  for (double **v : {&h->c, &h->T, &h->lb, &h->ub, &h->best})
    if (int rc = alloc_array(h, v, nloc)) return rc;
  const bool want_p2p = N > 1 && !(h->flags & CPPPD_FLAG_NO_P2P);
  for (double **v : {&h->x, &h->dbuf})
    if (int rc = alloc_array(h, v, nloc + n_ghost)) return rc;
  for (double **v : {&h->b, &h->sigma})
    if (int rc = alloc_array(h, v, mloc)) return rc;
  if (want_p2p) {  // the two vectors with peer-written ghost tails: plain cudaMalloc, exportable by IPC
    CK(cudaMalloc(&h->xbar, sizeof(double) * std::max<int64_t>(nloc + n_ghost, 1)));
    h->p2p.own.push_back(h->xbar);
    CK(cudaMalloc(&h->y, sizeof(double) * std::max<int64_t>(mloc + m_ghost, 1)));
    h->p2p.own.push_back(h->y);
    h->device_bytes += 8 * (nloc + n_ghost + mloc + m_ghost);
  } else {
    if (int rc = alloc_array(h, &h->xbar, nloc + n_ghost)) return rc;
    if (int rc = alloc_array(h, &h->y, mloc + m_ghost)) return rc;
  }
  if (int rc = upload_local(h, tmp, P->c, n, h->col_old, nloc, h->c)) return rc;
  if (int rc = upload_local(h, tmp, P->lb, n, h->col_old, nloc, h->lb)) return rc;
  if (int rc = upload_local(h, tmp, P->ub, n, h->col_old, nloc, h->ub)) return rc;
  if (int rc = upload_local(h, tmp, P->b, m, h->row_old, mloc, h->b)) return rc;
  if (P->x0) {
    if (int rc = upload_local(h, tmp, P->x0, n, h->col_old, nloc + n_ghost, h->x)) return rc;
  } else {
    CK(cudaMemsetAsync(h->x, 0, sizeof(double) * std::max<int64_t>(nloc + n_ghost, 1), st));
  }
  CK(cudaMemcpyAsync(h->xbar, h->x, sizeof(double) * (nloc + n_ghost), cudaMemcpyDeviceToDevice, st));  // x3 = x (:190)
  CK(cudaMemsetAsync(h->dbuf, 0, sizeof(double) * std::max<int64_t>(nloc + n_ghost, 1), st));
  CK(cudaMemsetAsync(h->y, 0, sizeof(double) * std::max<int64_t>(mloc + m_ghost, 1), st));              // :166,:177
  // ---- preconditioners (:122-179): complete columns / rows are local, so no exchange is needed
  const int has_eq = h->m_eq_glob > 0, has_ineq = h->m_ineq_glob > 0;
  if (h->AT.nslices)
    k_precond_cols<<<grid_for(h->AT.nslices * 32), kBlock, 0, st>>>(view(h->AT), nloc, has_eq, has_ineq, 2.0 - h->alpha, h->T);
  if (h->A.nslices)
    k_precond_rows<<<grid_for(h->A.nslices * 32), kBlock, 0, st>>>(view(h->A), mloc, h->alpha, h->sigma);
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
  if (want_p2p)
    if (int rc = setup_p2p(h)) return rc;
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

template <bool kWriteD, bool kDict>
void launch_primal_t(cpppd_solver *h) {
  const int has_eq = h->m_eq_glob > 0, has_ineq = h->m_ineq_glob > 0;
  k_primal<kWriteD, kDict><<<grid_for(h->AT.nslices * 32), kBlock, 0, h->stream>>>(
      view(h->AT), h->y, h->vc, h->vT, h->vlb, h->vub, h->x, h->xbar, h->dbuf, h->n, has_eq, has_ineq, h->theta,
      h->one_plus_theta);
}

int launch_primal(cpppd_solver *h, bool write_d) {
  if (h->AT.nslices) {
    const bool dict = h->AT.dict != nullptr;
    if (write_d) dict ? launch_primal_t<true, true>(h) : launch_primal_t<true, false>(h);
    else dict ? launch_primal_t<false, true>(h) : launch_primal_t<false, false>(h);
  }
  return h->p2p.active ? exchange_p2p(h, 0) : exchange(h, h->xbar, h->hx);
}

int launch_dual(cpppd_solver *h) {
  if (h->A.nslices) {
    const int grid = grid_for(h->A.nslices * 32);
    if (h->A.dict)
      k_dual<true><<<grid, kBlock, 0, h->stream>>>(view(h->A), h->xbar, h->vb, h->vsigma, h->y, h->m, h->m_eq);
    else
      k_dual<false><<<grid, kBlock, 0, h->stream>>>(view(h->A), h->xbar, h->vb, h->vsigma, h->y, h->m, h->m_eq);
  }
  return h->p2p.active ? exchange_p2p(h, 1) : exchange(h, h->y, h->hy);
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

int run_iterations(cpppd_solver *h, int64_t k) {
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
  // every entry is owned by exactly one rank, the others contribute +0.0: the sum is exact
  if (h->world > 1 && full) NK(g_nccl.AllReduce(buf, buf, (size_t)full, ncclFloat64, ncclSum, h->comm, h->stream));
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
  if (world > 1 && !P->comm_id) return fail(h, CPPPD_ERR_INVALID, "world_size > 1 needs comm_id (cpppd_comm_unique_id)");
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
    if (P->stream) {
      h->stream = (cudaStream_t)P->stream;
    } else {
      if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        rc = fail(h, CPPPD_ERR_CUDA, "cudaStreamCreate failed");
        break;
      }
      h->own_stream = true;
    }
    if (world > 1) {
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
  if (h->comm) g_nccl.CommDestroy(h->comm);
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
      if (int rc = alloc_array(h, &h->xr_scratch, h->n + h->hx.ghost)) return rc;
    xr = h->xr_scratch;
  }
  k_stats_cols<<<h->stat_blocks_c, kBlock, 0, h->stream>>>(h->vc, h->x, h->xbar, h->vlb, h->vub, h->dbuf, xr, h->n,
                                                          force_integer, h->colpart);
  // the row pass reads x, x4 and xr at ghost columns too
  if (int rc = exchange(h, h->x, h->hx)) return rc;
  if (int rc = exchange(h, h->dbuf, h->hx)) return rc;
  if (force_integer)
    if (int rc = exchange(h, xr, h->hx)) return rc;
  k_stats_rows<<<h->stat_blocks_r, kBlock, 0, h->stream>>>(view(h->A), h->x, h->dbuf, h->xbar, xr, h->vb, h->y, h->m,
                                                          h->m_eq, force_integer, h->rowpart);
  k_stats_local<<<1, kBlock, 0, h->stream>>>(h->colpart, h->stat_blocks_c, h->rowpart, h->stat_blocks_r, h->stat_local);
  const double *all = h->stat_local;
  if (h->world > 1) {
    NK(g_nccl.AllGather(h->stat_local, h->stat_all, kStatQ, ncclFloat64, h->comm, h->stream));
    all = h->stat_all;
  }
  k_stats_final<<<1, 32, 0, h->stream>>>(all, h->world, h->n_glob, has_eq, has_ineq, h->niter, h->stats_dev);
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
  return 0;
}

int cpppd_read_stats(cpppd_handle h, cpppd_stats *out) {
  CHECK_HANDLE(h);
  if (!out) return fail(h, CPPPD_ERR_INVALID, "null output");
  if (!h->stats_pending) return fail(h, CPPPD_ERR_STATE, "no stats step has been issued");
  CK(cudaStreamSynchronize(h->stream));
  *out = *h->stats_host;
  return 0;
}

int cpppd_time_iterations(cpppd_handle h, int64_t k, float *elapsed_ms) {
  CHECK_HANDLE(h);
  if (!elapsed_ms || k < 0) return fail(h, CPPPD_ERR_INVALID, "bad argument");
  if (h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "cpppd_dual_step must close the open iteration first");
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, h->stream));
  int rc = run_iterations(h, k);
  if (!rc) {
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(elapsed_ms, e0, e1));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

int cpppd_time_kernels(cpppd_handle h, int64_t k, float *primal_ms, float *dual_ms) {
  CHECK_HANDLE(h);
  if (!primal_ms || !dual_ms || k < 0 || k > 64) return fail(h, CPPPD_ERR_INVALID, "bad argument (k must be in [0, 64])");
  if (h->mid_iteration) return fail(h, CPPPD_ERR_STATE, "cpppd_dual_step must close the open iteration first");
  // the halo exchanges (world > 1) fall inside the brackets of the kernel that produces the data
  std::vector<cudaEvent_t> ev(2 * k + 1);
  for (auto &e : ev) CK(cudaEventCreate(&e));
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
  for (auto &e : ev) cudaEventDestroy(e);
  return 0;
}

int cpppd_get_vector(cpppd_handle h, int32_t which, double *host_dst) {
  CHECK_HANDLE(h);
  double *p = nullptr;
  bool is_col = true;
  if (int rc = vector_ptr(h, which, &p, &is_col)) return rc;
  if (!host_dst) return fail(h, CPPPD_ERR_INVALID, "null destination");
  return fetch_vector(h, p, is_col, host_dst);
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
  out->bytes_per_iteration_actual = (h->A.padded + h->AT.padded) * entry_bytes +
                                    (h->A.uniform_width >= 0 ? 0 : 8 * (h->A.nslices + 1)) +
                                    (h->AT.uniform_width >= 0 ? 0 : 8 * (h->AT.nslices + 1)) + vec_bytes;
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
  return 0;
}

int64_t cpppd_iteration_count(cpppd_handle h) { return h ? h->niter : -1; }

}  // extern "C"
