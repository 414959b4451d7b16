/*
 * cpppd.h — C ABI of the B200-native Chambolle-Pock PPD LP solver core (libcpppd.so).
 *
 * This is the drop-in boundary for the hot path of martinResearch/PySparseLP:
 * everything the reference does inside
 *     pysparselp/ChambollePockPPD.py:122-343   (preconditioners, main loop, stats block)
 * happens behind these entry points, on the GPU, in IEEE fp64.  The host side
 * (pysparselp_b200/ChambollePockPPD.py) keeps the reference's Python signature and
 * calls this library through ctypes; see INTEGRATION.md for the stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - plain C types only; every pointer in cpppd_problem is a HOST pointer that stays
 *     owned by the caller and is only read during cpppd_create();
 *   - every function returns 0 on success and a negative cpppd_status on failure; a
 *     human readable message is available from cpppd_last_error(); CUDA errors are
 *     sticky on the handle; nothing throws across the boundary;
 *   - one host thread drives one handle; all work of a handle is issued on one CUDA
 *     stream (the one given at creation, or an internal one) and is asynchronous
 *     unless the function says it synchronises;
 *   - there is NO CPU fallback: without a CUDA device cpppd_create() fails.
 *
 * Environment variables read by the library (all optional)
 *   CPPPD_NCCL_LIB          path of the libnccl.so.2 to dlopen (world_size > 1 only)
 *   CPPPD_HALO_TIMEOUT_S    seconds a peer-memory halo wait may spin before CPPPD_ERR_COMM (default 60, 0: for ever)
 *   CPPPD_KERNEL_VARIANT    like cpppd_problem.kernel_variant when that field is 0
 *   CPPPD_AUTOTUNE_MIN_NNZ  smallest operand (entries) whose kernel variants are timed at creation (default 2^22)
 *   CPPPD_AUTOTUNE_CACHE    0: time the variants at every creation instead of once per operand shape and process
 *   CPPPD_BAND_WINDOW_MB    megabytes of the gathered vector per window of a banded operand (default 56)
 *   CPPPD_BAND_SHAPE        0..7: compiled shape of the banded window kernels (cpppd_info.band_shape_ms) instead of the timed choice
 *   CPPPD_BAND_WINDOW       the same in elements (tests: windows of a few dozen elements on small LPs); both are
 *                           only read when cpppd_problem.band_window is 0
 *
 * Problem statement (reference ChambollePockPPD.py:55-65 after the one-sided
 * conversion of :74-88, which stays in Python):
 *     min c.x   s.t.  A[0:m_eq] x = b[0:m_eq],   A[m_eq:m] x <= b[m_eq:m],   lb <= x <= ub
 * with A = [A_eq ; A_ineq] stacked row-wise in CSR (entry order inside a row is
 * preserved: it defines the floating-point summation order).
 */
#ifndef CPPPD_H
#define CPPPD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPPPD_ABI_VERSION 7

#define CPPPD_KERNEL_VARIANTS 9

typedef struct cpppd_solver *cpppd_handle;

typedef enum {
  CPPPD_OK = 0,
  CPPPD_ERR_INVALID = -1,  /* bad argument / malformed CSR */
  CPPPD_ERR_CUDA = -2,     /* CUDA runtime error (sticky) */
  CPPPD_ERR_NOMEM = -3,    /* allocation failed */
  CPPPD_ERR_NODEVICE = -4, /* no CUDA device: there is no CPU path */
  CPPPD_ERR_COMM = -5,     /* NCCL / multi-GPU exchange error */
  CPPPD_ERR_STATE = -6     /* call sequence violated (e.g. stats before a primal step) */
} cpppd_status;

/* Device buffer provider.  The Python front passes callbacks that hand out
 * torch.uint8 CUDA tensors (PyTorch is only the buffer allocator); NULL means
 * cudaMalloc/cudaFree.  Returned pointers must be 256-byte aligned device memory
 * on the problem's device. */
typedef void *(*cpppd_alloc_fn)(size_t bytes, void *user);
typedef void (*cpppd_free_fn)(void *ptr, void *user);

enum {
  CPPPD_FLAG_NONE = 0,
  /* compress matrix values through a dictionary when they take few distinct values
   * (bit-exact: the dictionary holds the original doubles) */
  CPPPD_FLAG_VALUE_DICT = 1u << 0,
  /* replace constant vectors (b, sigma, lb, ub, c, T) by scalars inside the kernels */
  CPPPD_FLAG_CONST_VECTORS = 1u << 1,
  /* do not capture the inner iterations into CUDA graphs */
  CPPPD_FLAG_NO_GRAPH = 1u << 2,
  /* on one GPU: renumber rows / columns by locality and length (always done when world_size > 1;
   * done automatically on one GPU when SELL-32 would otherwise pad the matrix by more than 15 %) */
  CPPPD_FLAG_REORDER = 1u << 3,
  /* world_size > 1: capture the iterations including the NCCL halo exchanges into CUDA graphs */
  CPPPD_FLAG_GRAPH_COMM = 1u << 4,
  /* world_size > 1: exchange the halos with NCCL send/recv instead of the peer-memory push kernels */
  CPPPD_FLAG_NO_P2P = 1u << 5,
  /* one GPU: never renumber, whatever the padding */
  CPPPD_FLAG_NO_REORDER = 1u << 6,
  /* world_size > 1, peer-memory halos: let k_primal / k_dual wait for, store and signal the halos
   * themselves instead of the separate push / wait kernels (experimental: validated on the CPU
   * emulation of the library only) */
  CPPPD_FLAG_FUSED_HALO = 1u << 7,
  /* never time the kernel variants at creation: use variant 1 (see cpppd_problem.kernel_variant) */
  CPPPD_FLAG_NO_AUTOTUNE = 1u << 8,
  /* small LPs: prefer the one-CTA persistent kernel (k_tiny_iterate) over the thread-block-cluster kernel when the LP
   * fits one SM (see CPPPD_FLAG_NO_TINY_PERSISTENT for what runs by default) */
  CPPPD_FLAG_TINY_PERSISTENT = 1u << 9,
  /* one GPU, caller's numbering: store A and A^T window-major ("banded") and run one launch per window of the
   * gathered vector, so that the gathers of a launch stay inside a window that fits the L2 (cpppd_banded.cuh).
   * Without this flag the banded copies are built only for patterns without locality over vectors of more than
   * two windows (random sparse LPs) and used only if they time faster than the SELL kernels at creation; with
   * it they are built and used whenever the entry order allows (ascending window index along every row —
   * otherwise the operand silently stays with the SELL kernels; see cpppd_info.band_in_use).  Same bits. */
  CPPPD_FLAG_BANDED = 1u << 10,
  /* never build the banded copies */
  CPPPD_FLAG_NO_BANDED = 1u << 11,
  /* One GPU, small LPs, no forced kernel variant: cpppd_iterate(k) runs all k iterations in ONE launch instead of 2k
   * graph nodes — such LPs are bound by launch latency, not bandwidth.  LPs whose operands and vectors fit the shared
   * memory of 16 SMs (<= 131072 stored entries per operand) run in one thread-block cluster of 16 CTAs, gathers
   * through distributed shared memory, the hardware cluster barrier between the two halves of an iteration
   * (k_cluster_iterate, csrc/cpppd_cluster.cuh: Potts 50x50 793 000 iterations/s against 144 000 through CUDA graphs,
   * netlib SC105 1 130 000 against 211 000); such LPs are renumbered by locality like CPPPD_FLAG_REORDER.  Where the
   * cluster launch is refused, LPs that fit one SM (n, m <= 4096, <= 16384 stored entries) run in one CTA
   * (k_tiny_iterate).  Same per-row arithmetic, same bits.  This flag keeps the graph path. */
  CPPPD_FLAG_NO_TINY_PERSISTENT = 1u << 12,
  /* world_size > 1, patterns without locality (balanced split + banded operands): keep only the ghosts the pattern
   * really touches and push them through index lists, instead of the dense halo (every foreign column / row is a
   * ghost; one contiguous copy per peer).  For comparison runs. */
  CPPPD_FLAG_NO_DENSE_HALO = 1u << 13
};

typedef struct {
  int32_t abi_version;  /* must be CPPPD_ABI_VERSION */
  int32_t device;       /* CUDA device ordinal */
  int64_t n;            /* variables (columns) */
  int64_t m_eq;         /* equality rows: rows [0, m_eq) of the stacked matrix */
  int64_t m_ineq;       /* inequality rows: rows [m_eq, m_eq + m_ineq) */
  int64_t nnz;          /* stored entries (explicit zeros count, as in scipy) */
  const void *indptr;   /* m+1 row pointers, int32 or int64 */
  const void *indices;  /* nnz column indices, int32 (wider indices are narrowed by the host front) */
  const double *values; /* nnz values */
  int32_t indptr_bits;  /* 32 or 64 */
  int32_t index_bits;   /* must be 32 */
  const double *c;      /* n   costs                           (ChambollePockPPD.py:198)   */
  const double *b;      /* m   right-hand sides [b_eq; b_ineq] (:235,:240)                 */
  const double *lb;     /* n   lower bounds, -inf allowed      (:221)                      */
  const double *ub;     /* n   upper bounds, +inf allowed      (:222)                      */
  const double *x0;     /* n   initial point or NULL for zeros (:91-94)                    */
  double alpha;         /* preconditioner exponent             (:133,:143,:160,:171)       */
  double theta;         /* extrapolation                       (:226)                      */
  double one_plus_theta;/* (1 + theta) as evaluated by the host language (:226)            */
  void *stream;         /* cudaStream_t to issue on, or NULL for an internal stream        */
  uint32_t flags;       /* CPPPD_FLAG_* */
  int32_t kernel_variant; /* 0: automatic — LPs with >= 2^22 entries time every variant of k_primal / k_dual on
                             their own operands during cpppd_create and keep the fastest, smaller ones use
                             variant 1.  Otherwise bits 0-7 force the variant of k_primal and bits 8-15 that of
                             k_dual (1 .. CPPPD_KERNEL_VARIANTS; bits 8-15 zero: same as k_primal).  All
                             variants produce bit-identical iterates; they differ in loads kept in flight (variants 6 and 7 walk two rows per lane). */
  cpppd_alloc_fn alloc; /* may be NULL */
  cpppd_free_fn free;   /* may be NULL */
  void *alloc_user;
  /* multi-GPU (one process per GPU).  Every rank passes the SAME full LP; the library keeps
   * only its share on the device.  comm_id: the 128 bytes produced by cpppd_comm_unique_id()
   * on one rank and broadcast by the host program (torch.distributed in the Python front). */
  int32_t rank;         /* 0 .. world_size-1 */
  int32_t world_size;   /* <= 1: single GPU */
  const void *comm_id;  /* 128 bytes, required when world_size > 1 */
  int64_t partition_granule; /* locality bucket width in columns; <= 0: default */
  void *comm;           /* optional cpppd_comm created by cpppd_comm_create(): reused (and not destroyed) by
                           this solver instead of building a new communicator from comm_id */
  int64_t long_row_threshold; /* rows of A / columns of A with more entries than this are summed by a CTA per
                           16384-entry segment instead of by one thread (skewed patterns: L1-SVM weight columns, dense
                           budget rows).  0: default (2048); < 0: never.  LPs with such rows agree with the reference
                           to rounding (fixed summation tree) instead of bit for bit; others are unaffected. */
  int64_t band_window;  /* banded operands: elements of the gathered vector per window.  0: CPPPD_BAND_WINDOW_MB
                           megabytes (default 56 MB = 7 340 032 elements: the L2-resident plateau measured by
                           tools/probe/gather_probe.cu) */
} cpppd_problem;

/* The numbers the reference's stats block produces (ChambollePockPPD.py:242-291). */
typedef struct {
  int64_t niter;                       /* iteration index the block belongs to                */
  double energy1;                      /* c.x + y.(Ax - b)          (:248,:267,:271)          */
  double energy2;                      /* c.x4 + y.(A x4 - b)       (:260-263,:268,:272)      */
  double max_violated_equality;        /* max |A_eq xbar - b_eq|    (:269), 0 without eq rows */
  double max_violated_inequality;      /* max (A_ineq xr - b_ineq)  (:283), -inf without rows */
  double energy_rounded;               /* c.xr                      (:278)                    */
  double max_violated_equality_rounded;/* max |A_eq xr - b_eq|      (:280-282)                */
  double best_integer_energy;          /* running minimum over feasible xr (:289-291)         */
  double frac_zero_xbar;               /* mean(xbar == 0)           (:308)                    */
  /* what SparseLP.solve()'s per-callback curves need, so that x can stay on the device
   * (reference SparseLP.py:186-204, :1074-1089) */
  double max_bound_violation;          /* max(max(lb - x), max(x - ub))                       */
  double distance_to_ground_truth;     /* mean |gt - x[idx]|        (cpppd_set_ground_truth)  */
  double distance_to_ground_truth_rounded; /* mean |gt - round(x[idx])|                       */
  /* the two row maxima of max_constraint_violation for the caller's FULL LP when variables were eliminated
   * before the solve (cpppd_set_row_offsets); equal to max_violated_equality_rounded / max_violated_inequality
   * without offsets */
  double max_violated_equality_full;   /* max |A_eq xr - b_eq - off|                          */
  double max_violated_inequality_full; /* max (A_ineq xr - b_ineq - off)                      */
  int32_t feasible;                    /* exact test of :284                                   */
  int32_t improved;                    /* xr became the best integer solution at this block   */
  int32_t have_best_integer;           /* a best integer solution exists                       */
  int32_t reserved;
} cpppd_stats;

typedef struct {
  int64_t n, m_eq, m_ineq, nnz;
  int64_t a_padded_entries;  /* entries stored for A   in SELL-32 (>= nnz) */
  int64_t at_padded_entries; /* entries stored for A^T in SELL-32 (>= nnz) */
  int64_t device_bytes;      /* resident device memory of the solver state */
  int64_t bytes_per_iteration_algorithmic; /* 2*nnz*12 + P(m+1) + P(n+1) + 8(8n+5m) */
  int64_t bytes_per_iteration_actual;      /* what the kernels of this handle really stream    */
  int32_t value_bytes;       /* bytes per stored matrix value: 8, or 0 with a dictionary (the code lives in the index word) */
  int32_t const_vector_mask; /* bit0 b, bit1 sigma, bit2 lb, bit3 ub, bit4 c, bit5 T folded to scalars */
  int32_t sm_count;
  int32_t world_size;
  int32_t rank;
  int32_t primal_variant;    /* kernel variant in use for k_primal (1-based) */
  int64_t n_local, m_local, m_eq_local;  /* columns / rows / equality rows owned by this rank      */
  int64_t n_ghost, m_ghost;              /* ghost columns (xbar) / ghost rows (y) kept by this rank */
  int64_t nnz_local_rows, nnz_local_cols;/* entries of the owned rows of A / owned columns of A     */
  int64_t halo_send_bytes_per_iteration; /* bytes this rank sends per iteration (xbar + y halos)    */
  int64_t partition_granule;
  int32_t dual_variant;      /* kernel variant in use for k_dual (1-based) */
  int32_t autotuned;         /* 1 when the variants were chosen by timing (at this creation, or at an earlier one
                                of operands of the same shape in this process) */
  /* milliseconds per launch measured at creation for k_primal ([0][v-1]) and k_dual ([1][v-1]); 0 = not timed */
  float variant_ms[2][CPPPD_KERNEL_VARIANTS];
  int64_t long_rows, long_cols; /* rows / columns handled by the long-row path on this rank */
  int64_t long_entries;         /* their entries */
  int32_t balanced_split;       /* world_size > 1: the locality buckets left some rank with more than 1.5x its share of
                                   the row or column entries, so rows / columns were dealt out by prefix sums instead */
  int32_t tiny_persistent;      /* iterations run in a persistent kernel: 1 one CTA (k_tiny_iterate), 2 one thread-block
                                   cluster (k_cluster_iterate); 0: CUDA graphs of k_primal / k_dual */
  /* banded operands (CPPPD_FLAG_BANDED; [0]: A, used by the dual half, [1]: A^T, used by the primal half) */
  int32_t band_windows[2];      /* windows of the gathered vector (= launches per half-iteration); 0: not built */
  int32_t band_in_use[2];       /* the half-iteration runs the banded kernels */
  float band_ms[2];             /* milliseconds per half-iteration measured at creation; 0 = not timed */
  float band_sectors_per_gather[2]; /* sampled locality of the operand: distinct 32-byte sectors per gather of a warp */
  int64_t band_window_bytes;    /* bytes of the gathered vector per window (largest window) */
  int32_t dense_halo;           /* world_size > 1: ghosts = all foreign columns / rows, contiguous per-peer pushes */
  int32_t reserved2;
  int32_t band_shape[2];        /* compiled shape of the window kernels in use (0-based, see band_shape_ms) */
  float band_shape_ms[2][8];    /* ms per half-iteration measured at creation for the shapes flat4/4cta, flat3/6cta,
                                   flat2/6cta, flat2/8cta (entries through registers: flat entries per lane and trip /
                                   CTAs per SM) and bulk-g8/5cta, bulk-g10/5cta, bulk-g12/4cta, bulk-g6/5cta (entries
                                   staged in shared memory by cp.async.bulk: gathers in flight per lane); 0 = not timed */
} cpppd_info;

typedef enum {
  CPPPD_VEC_X = 0,       /* n  primal iterate x                         */
  CPPPD_VEC_XBAR = 1,    /* n  extrapolated point (x3 in the reference) */
  CPPPD_VEC_Y = 2,       /* m  duals [y_eq; y_ineq]                     */
  CPPPD_VEC_T = 3,       /* n  diag_t   (:153)                          */
  CPPPD_VEC_SIGMA = 4,   /* m  [diag_sigma_eq; diag_sigma_ineq] (:164,:175) */
  CPPPD_VEC_BEST_INTEGER = 5, /* n, valid when stats.have_best_integer  */
  CPPPD_VEC_D = 6        /* n  d = c + A^T y of the last stats step (:198-217) */
} cpppd_vector;

/* -- lifetime ------------------------------------------------------------------ */
/* Upload the LP, build A and A^T in SELL-32 on the device, compute diag_t and the
 * sigmas (ChambollePockPPD.py:122-179), set x = x0, y = 0.  Synchronises. */
int cpppd_create(const cpppd_problem *problem, cpppd_handle *out);
int cpppd_destroy(cpppd_handle h);
/* Message of the last failure on this handle (or of the last failed cpppd_create
 * on this thread when h is NULL). */
const char *cpppd_last_error(cpppd_handle h);
int cpppd_abi_version(void);
/* 128-byte NCCL unique id for a new multi-GPU solve (call on one rank, broadcast, pass as comm_id). */
int cpppd_comm_unique_id(void *out128);
/* A communicator that outlives one solver: building one costs far more than a short solve, so the
 * Python front keeps one per process group and passes it through cpppd_problem.comm. */
typedef struct cpppd_comm_s *cpppd_comm;
int cpppd_comm_create(const void *id128, int32_t rank, int32_t world_size, int32_t device, cpppd_comm *out);
int cpppd_comm_destroy(cpppd_comm comm);

/* -- the iteration (ChambollePockPPD.py:195-343) ----------------------------------- */
/* k full iterations: [x,xbar <- primal(y)] ; [y <- dual(xbar)], asynchronous. */
int cpppd_iterate(cpppd_handle h, int64_t k);
/* First half of an iteration (:198-240): x, xbar from the current y.  With keep_d != 0
 * the vector d = c + A^T y is also stored (needed by a following cpppd_stats_step).
 * Asynchronous. */
int cpppd_primal_step(cpppd_handle h, int32_t keep_d);
/* The stats block (:248-291), evaluated between the two halves of an iteration, i.e.
 * with the new x / xbar and the still unmodified y.  Requires the preceding
 * cpppd_primal_step(keep_d = 1).  Updates the best integer solution on the device and
 * starts the copy of the result to the host.  Asynchronous. */
int cpppd_stats_step(cpppd_handle h, int32_t force_integer);
/* Second half (:333-343): y from xbar; increments the iteration counter. */
int cpppd_dual_step(cpppd_handle h);
/* Wait for everything issued so far. */
int cpppd_sync(cpppd_handle h);
/* Wait, then return the stats of the last cpppd_stats_step. */
int cpppd_read_stats(cpppd_handle h, cpppd_stats *out);
/* Run k iterations bracketed by CUDA events on the solver's stream; returns the
 * elapsed device time in milliseconds.  Synchronises. */
int cpppd_time_iterations(cpppd_handle h, int64_t k, float *elapsed_ms);
/* Run k (<= 64) iterations with a CUDA event between every kernel; returns the summed
 * device time of the primal kernels and of the dual kernels (ms).  Synchronises. */
int cpppd_time_kernels(cpppd_handle h, int64_t k, float *primal_ms, float *dual_ms);

/* -- state access (synchronising copies to/from HOST memory) ------------------------- */
/* x and the best integer solution are what the reference returns (ChambollePockPPD.py:344-346) and hands to its
 * callback (:319-329); y, xbar (x3), diag_t, diag_sigma and d are the locals of :153-217 that the parity tests compare. */
int cpppd_get_vector(cpppd_handle h, int32_t which, double *host_dst);
/* Warm start (the reference can only start from x0, :91-94): X, XBAR, Y only. */
int cpppd_set_vector(cpppd_handle h, int32_t which, const double *host_src);
int cpppd_get_info(cpppd_handle h, cpppd_info *out);
/* Ground truth for the distance curves of the stats block: values[k] is compared with x[indices[k]]
 * (original column ids).  count = 0 removes it.  With world_size > 1 every rank passes the full list. */
int cpppd_set_ground_truth(cpppd_handle h, const int32_t *indices, const double *values, int64_t count);
/* Per-row constants subtracted from the residuals A xr - b in the two *_full maxima of the stats block (m values,
 * original row order; NULL removes them).  SparseLP.solve() eliminates fixed variables before the solve (reference
 * SparseLP.py:632-674) and evaluates max_constraint_violation on the FULL LP at the mapped-back point (:1064-1093,
 * :186-204): the full residual of a row is the reduced one minus a constant, so x can stay on the device. */
int cpppd_set_row_offsets(cpppd_handle h, const double *offsets);
/* Partition of this rank: number of owned and ghost columns (columns != 0) or rows (columns == 0)
 * and, when ids is not NULL, their original indices in local order (owned first, then ghosts in
 * exchange order).  A pure function of (indptr, indices, m_eq, world_size, granule):
 * oracle/partition_oracle.py restates it. */
int cpppd_get_layout(cpppd_handle h, int32_t columns, int64_t *owned, int64_t *ghost, int32_t *ids);
int64_t cpppd_iteration_count(cpppd_handle h);

#ifdef __cplusplus
}
#endif
#endif /* CPPPD_H */
