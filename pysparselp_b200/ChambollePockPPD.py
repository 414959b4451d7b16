"""Chambolle-Pock diagonally preconditioned primal-dual LP solver — B200 front end.

Drop-in for the reference's ``pysparselp/ChambollePockPPD.py:36-346``: same function name,
positional arguments, defaults, callback protocol and ``(x, best_integer_solution)``
return.  The Python here only prepares the operator ([A_eq; A_ineq] in CSR, one-sided
inequality rows as in ``:70-88``) and drives the schedule; every floating point operation
of the solver (preconditioners ``:122-179``, the loop ``:195-343``, the stats block
``:242-291``) runs in hand-written CUDA behind the C ABI of ``include/cpppd.h``.
PyTorch is used for one thing: handing out device buffers (``torch.uint8`` tensors) and
the current CUDA stream.  There is no CPU fallback.

Extra keyword-only arguments (defaults keep the reference behaviour):
``device`` (CUDA ordinal or torch.device), ``verbose`` (print the reference's per-block
line), ``flags`` (CPPPD_FLAG_* bit mask), ``return_solver`` (also return the live
``CpPpdSolver`` for inspection), ``kernel_variant`` (0: pick the fastest variant of the two hot kernels
automatically; see ``cpppd_problem.kernel_variant`` — the iterates do not depend on it),
``long_row_threshold`` (rows / columns with more entries are summed by many threads, see
``cpppd_problem.long_row_threshold``), ``distributed`` (None: use the initialised torch.distributed
world when it has more than one rank — one process per GPU, every rank passes the same LP and
gets the same result; False: this GPU only; or an explicit ProcessGroup), ``n_gpus`` (N > 1: this ONE process
drives N GPUs of the node — helper processes are spawned for the other devices for the duration of the call, see
``pysparselp_b200/multi_gpu.py``; ``device`` may then list the N ordinals), ``y0`` (dual warm start
``[y_eq; y_ineq]`` in the row order of the one-sided system; the reference always starts from zero).
"""
import atexit
import ctypes as C
import time

import numpy as np
import scipy.sparse as sp

from . import _cabi

# (process group, device) -> cpppd_comm, kept for the life of the process
_COMM_CACHE = {}
# device ordinal -> the torch stream solvers use when the caller is on the legacy default stream
_SOLVER_STREAMS = {}


def _destroy_cached_comms():
    lib = _cabi._lib
    while _COMM_CACHE and lib is not None:
        _, comm = _COMM_CACHE.popitem()
        try:
            lib.cpppd_comm_destroy(comm)
        except Exception:
            pass


atexit.register(_destroy_cached_comms)


def _as_f64(v, size, name):
    a = np.ascontiguousarray(v, dtype=np.float64).ravel()
    if a.size != size:
        raise ValueError("%s has %d entries, expected %d" % (name, a.size, size))
    return a


def one_sided_rows(a_ineq, b_lower, b_upper):
    """``b_lower <= A x <= b_upper``  ->  ``A' x <= b'``  (reference ``:74-88``).

    Rows of ``A'``: those with a finite upper bound (original order), then the negation of
    those with a finite lower bound (original order).  Deviation from the reference: when
    only one side has finite entries the reference keeps / negates the whole matrix without
    dropping the unbounded rows and then fails its own shape assertion (``:141``); here the
    unbounded rows are dropped, which is what the two-sided branch does and is a no-op
    whenever the reference does not crash.
    """
    if a_ineq is None or b_lower is None:
        return a_ineq, b_upper
    b_lower = np.asarray(b_lower, dtype=np.float64)
    b_upper = np.asarray(b_upper, dtype=np.float64)
    up = np.flatnonzero(b_upper != np.inf)
    lo = np.flatnonzero(b_lower != -np.inf)
    m = a_ineq.shape[0]
    if up.size and lo.size:
        a = sp.vstack((a_ineq[up, :], -a_ineq[lo, :])).tocsr()
    elif lo.size:
        a = -a_ineq if lo.size == m else -a_ineq[lo, :]
    else:
        a = a_ineq if up.size == m else a_ineq[up, :]
    return a, np.concatenate((b_upper[up], -b_lower[lo]))


class _TorchBuffers:
    """Device buffer provider for libcpppd: every buffer is a ``torch.uint8`` CUDA tensor."""

    def __init__(self, device, stream=None):
        import torch

        self.torch = torch
        self.device = device
        self.stream = stream  # torch.cuda.Stream the library issues on: blocks are handed out (and reused) in its order
        self.live = {}
        self.alloc_cb = _cabi.ALLOC_FN(self._alloc)
        self.free_cb = _cabi.FREE_FN(self._free)

    def _alloc(self, nbytes, _user):
        try:
            if self.stream is not None:
                with self.torch.cuda.stream(self.stream):
                    t = self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.device)
            else:
                t = self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.device)
        except Exception:  # out of memory -> NULL -> CPPPD_ERR_NOMEM
            return None
        self.live[t.data_ptr()] = t
        return t.data_ptr()

    def _free(self, ptr, _user):
        self.live.pop(ptr, None)


def prepare_problem(c, a, m_eq, b, lb, ub, x0, alpha, theta, flags, partition_granule, kernel_variant=0,
                    long_row_threshold=0):
    """``cpppd_problem`` with the host-side fields filled from numpy / scipy operands.

    Returns ``(problem, keepalive)``; the arrays in ``keepalive`` must outlive ``cpppd_create``.
    Device, stream, allocator and multi-GPU fields are left to the caller.
    """
    a = sp.csr_matrix(a) if not sp.isspmatrix_csr(a) else a
    m, n = a.shape
    c = _as_f64(c, n, "c")
    lb = _as_f64(lb, n, "lb")
    ub = _as_f64(ub, n, "ub")
    b = _as_f64(b, m, "b")
    x0 = None if x0 is None else _as_f64(x0, n, "x0")
    data = np.ascontiguousarray(a.data, dtype=np.float64)
    indices = np.ascontiguousarray(a.indices)
    if indices.dtype != np.int32:
        indices = indices.astype(np.int32)
    indptr = np.ascontiguousarray(a.indptr)
    if indptr.dtype not in (np.int32, np.int64):
        indptr = indptr.astype(np.int64)
    p = _cabi.Problem()
    p.abi_version = _cabi.ABI_VERSION
    p.n, p.m_eq, p.m_ineq, p.nnz = n, int(m_eq), m - int(m_eq), int(indptr[-1]) if m else 0
    p.indptr = indptr.ctypes.data
    p.indices = indices.ctypes.data
    p.values = data.ctypes.data
    p.indptr_bits = 32 if indptr.dtype == np.int32 else 64
    p.index_bits = 32
    p.c, p.b, p.lb, p.ub = c.ctypes.data, b.ctypes.data, lb.ctypes.data, ub.ctypes.data
    p.x0 = None if x0 is None else x0.ctypes.data
    p.alpha, p.theta = float(alpha), float(theta)
    p.one_plus_theta = float(1 + theta)
    p.flags = int(flags)
    p.kernel_variant = int(kernel_variant)
    p.long_row_threshold = int(long_row_threshold)
    p.partition_granule = int(partition_granule)
    p.rank, p.world_size = 0, 1
    return p, (c, lb, ub, b, x0, data, indices, indptr)


class SolverHandle:
    """Methods of a live ``cpppd_handle`` (``self.lib``, ``self.handle``, ``self.n``, ``self.m``)."""

    def _result_buffer(self, size):
        return np.empty(size, dtype=np.float64)

    def layout(self, columns=True):
        """(owned original ids, ghost original ids) of this rank, local order."""
        owned, ghost = C.c_int64(), C.c_int64()
        self._call(self.lib.cpppd_get_layout, int(columns), C.byref(owned), C.byref(ghost), None)
        ids = np.empty(owned.value + ghost.value, dtype=np.int32)
        self._call(self.lib.cpppd_get_layout, int(columns), C.byref(owned), C.byref(ghost), ids.ctypes.data)
        return ids[: owned.value], ids[owned.value:]

    # -- schedule ---------------------------------------------------------------------------
    def _call(self, fn, *args):
        return _cabi.check(self.lib, self.handle, fn(self.handle, *args))

    def iterate(self, k):
        self._call(self.lib.cpppd_iterate, int(k))

    def primal_step(self, keep_d=False):
        self._call(self.lib.cpppd_primal_step, int(bool(keep_d)))

    def stats_step(self, force_integer=False):
        self._call(self.lib.cpppd_stats_step, int(bool(force_integer)))

    def dual_step(self):
        self._call(self.lib.cpppd_dual_step)

    def sync(self):
        self._call(self.lib.cpppd_sync)

    def read_stats(self):
        s = _cabi.Stats()
        self._call(self.lib.cpppd_read_stats, C.byref(s))
        return s.as_dict()

    def read_stats_or_none(self):
        """Stats of the last stats step, or None when none was issued yet."""
        try:
            return self.read_stats()
        except _cabi.CpppdError as e:
            if e.code == -6:
                return None
            raise

    def time_iterations(self, k):
        ms = C.c_float()
        self._call(self.lib.cpppd_time_iterations, int(k), C.byref(ms))
        return ms.value

    def time_kernels(self, k):
        """(primal_ms, dual_ms) summed over k iterations, one CUDA event between every kernel."""
        a, b = C.c_float(), C.c_float()
        self._call(self.lib.cpppd_time_kernels, int(k), C.byref(a), C.byref(b))
        return a.value, b.value

    # -- state ------------------------------------------------------------------------------
    def _get(self, which, size):
        out = self._result_buffer(size)
        self._call(self.lib.cpppd_get_vector, which, out.ctypes.data)
        return out

    def _set(self, which, v, size):
        v = _as_f64(v, size, "vector")
        self._call(self.lib.cpppd_set_vector, which, v.ctypes.data)

    def get_x(self):
        return self._get(_cabi.VEC_X, self.n)

    def get_xbar(self):
        return self._get(_cabi.VEC_XBAR, self.n)

    def get_y(self):
        return self._get(_cabi.VEC_Y, self.m)

    def get_d(self):
        return self._get(_cabi.VEC_D, self.n)

    def get_preconditioners(self):
        return self._get(_cabi.VEC_T, self.n), self._get(_cabi.VEC_SIGMA, self.m)

    def get_best_integer(self):
        return self._get(_cabi.VEC_BEST_INTEGER, self.n)

    def set_ground_truth(self, indices, values):
        """Ground truth for the distance curves of the stats block: values[k] vs x[indices[k]]."""
        idx = np.ascontiguousarray(np.ravel(indices), dtype=np.int32)
        val = np.ascontiguousarray(np.ravel(values), dtype=np.float64)
        if idx.size != val.size:
            raise ValueError("ground truth indices and values differ in size")
        self._call(self.lib.cpppd_set_ground_truth, idx.ctypes.data, val.ctypes.data, idx.size)

    def set_row_offsets(self, offsets):
        """Per-row constants of the full-LP residuals (see ``cpppd_set_row_offsets``); None removes them."""
        if offsets is None:
            self._call(self.lib.cpppd_set_row_offsets, None)
            return
        v = _as_f64(offsets, self.m, "row offsets")
        self._call(self.lib.cpppd_set_row_offsets, v.ctypes.data)

    def set_x(self, v):
        self._set(_cabi.VEC_X, v, self.n)

    def set_xbar(self, v):
        self._set(_cabi.VEC_XBAR, v, self.n)

    def set_y(self, v):
        self._set(_cabi.VEC_Y, v, self.m)

    def info(self):
        i = _cabi.Info()
        self._call(self.lib.cpppd_get_info, C.byref(i))
        return i.as_dict()

    @property
    def niter(self):
        return int(self.lib.cpppd_iteration_count(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cpppd_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class CpPpdSolver(SolverHandle):
    """Live solver state on one GPU: the only way the product creates a ``cpppd_handle``."""

    def __init__(self, c, a, m_eq, b, lb, ub, x0=None, alpha=1, theta=1, device=None, flags=0,
                 process_group=None, partition_granule=0, kernel_variant=0, long_row_threshold=0):
        import torch

        self.lib = _cabi.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("pysparselp_b200 needs a CUDA device (B200); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise ValueError("device must be a CUDA device")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.rank, self.world = 0, 1
        comm = None
        if process_group is not None:
            comm = self._shared_comm(process_group, dev)
        p, keep = prepare_problem(c, a, m_eq, b, lb, ub, x0, alpha, theta, flags, partition_granule, kernel_variant,
                                  long_row_threshold)
        self.n, self.m, self.m_eq = int(p.n), int(p.m_eq + p.m_ineq), int(p.m_eq)
        # Stream contract: the solver issues everything on ONE stream that torch knows about, and
        # its buffers come from torch's caching allocator *on that stream*, so a block torch recycles is never
        # handed out while work that still uses it is pending.  The caller's current stream is used as is unless
        # it is the legacy default stream (handle 0): then the solver gets a stream of its own that first waits
        # for the work already queued on the default stream.
        current = torch.cuda.current_stream(dev)
        if current.cuda_stream == 0:
            # one such stream per device and process: torch's caching allocator pools blocks per stream, so a fresh
            # stream per solve could not reuse what the previous solve freed (gigabytes of cudaMalloc per call)
            if dev.index not in _SOLVER_STREAMS:
                _SOLVER_STREAMS[dev.index] = torch.cuda.Stream(dev)
            self._stream = _SOLVER_STREAMS[dev.index]
            self._stream.wait_stream(current)
        else:
            self._stream = current
        self._buffers = _TorchBuffers(dev, self._stream)
        p.device = dev.index
        p.stream = self._stream.cuda_stream
        p.alloc = self._buffers.alloc_cb
        p.free = self._buffers.free_cb
        p.alloc_user = None
        p.rank, p.world_size = self.rank, self.world
        p.comm_id = None
        p.comm = comm
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            _cabi.check(self.lib, None, self.lib.cpppd_create(C.byref(p), C.byref(handle)))
        del keep  # host arrays are only read during create
        self.handle = handle

    def _result_buffer(self, size):
        # large results land in pinned host memory (torch's caching host allocator): the D2H copy
        # then runs at PCIe speed instead of going through a pageable staging buffer
        if size >= (1 << 16):
            import torch

            return torch.empty(size, dtype=torch.float64, pin_memory=True).numpy()
        return np.empty(size, dtype=np.float64)

    def close(self):
        if getattr(self, "handle", None):
            super().close()
            self._buffers.live.clear()

    def agree(self, flag, any_rank=False):
        """One decision for all ranks of a distributed solve: rank 0's `flag` (default) or the OR over the ranks.
        The schedule branches on it (time-out, "does anybody have a callback"): ranks that decided differently
        would issue different collectives and hang."""
        if self.world <= 1:
            return bool(flag)
        import torch
        import torch.distributed as dist

        on_gpu = dist.get_backend(self._group) == "nccl"
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=self.device if on_gpu else "cpu")
        if any_rank:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self._group)
        else:
            dist.broadcast(t, src=dist.get_global_rank(self._group, 0), group=self._group)
        return bool(int(t.item()))

    def _shared_comm(self, group, dev):
        """NCCL communicator of (process group, device), built once and kept for later solves: rank 0
        draws the id, torch.distributed broadcasts it, cpppd_comm_create joins."""
        import os

        import torch
        import torch.distributed as dist

        if group is True:
            group = dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world == 1:
            return None
        self._group = group
        # keyed by the member ranks, not by id(group): ids are reused after garbage collection
        key = (tuple(dist.get_process_group_ranks(group)), dev.index)
        if key in _COMM_CACHE:
            return _COMM_CACHE[key]
        if "CPPPD_NCCL_LIB" not in os.environ:  # use the NCCL that torch itself loaded
            try:
                import nvidia.nccl

                # (a namespace package: no __file__, the directories are in __path__)
                for base in list(getattr(nvidia.nccl, "__path__", [])):
                    cand = os.path.join(base, "lib", "libnccl.so.2")
                    if os.path.isfile(cand):
                        os.environ["CPPPD_NCCL_LIB"] = cand
                        break
            except Exception:
                pass
        ident = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            _cabi.check(self.lib, None, self.lib.cpppd_comm_unique_id(ident.ctypes.data))
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.from_numpy(ident).to(dev) if on_gpu else torch.from_numpy(ident)
        dist.broadcast(t, src=dist.get_global_rank(group, 0), group=group)
        ident = t.cpu().numpy().copy()
        comm = C.c_void_p()
        with torch.cuda.device(dev):
            _cabi.check(self.lib, None, self.lib.cpppd_comm_create(ident.ctypes.data, self.rank, self.world, dev.index,
                                                                   C.byref(comm)))
        _COMM_CACHE[key] = comm
        return comm


def stack_operator(a_eq, beq, a_ineq, b_ineq, n):
    """[A_eq; A_ineq] as one CSR (entry order inside rows untouched) and [b_eq; b_ineq]."""
    has_eq = a_eq is not None and a_eq.shape[0] > 0
    has_ineq = a_ineq is not None and a_ineq.shape[0] > 0
    for a in (a_eq, a_ineq):
        if a is not None:
            if not sp.issparse(a):
                raise ValueError("constraint matrices must be scipy sparse matrices")
            if a.shape[1] != n:
                raise ValueError("constraint matrix has %d columns, expected %d" % (a.shape[1], n))
    if has_eq and has_ineq:
        a_eq, a_ineq = sp.csr_matrix(a_eq), sp.csr_matrix(a_ineq)
        wide = np.int64 if (a_eq.nnz + a_ineq.nnz) >= 2**31 - 1 else np.int32
        indptr = np.concatenate((a_eq.indptr.astype(wide), a_ineq.indptr[1:].astype(wide) + wide(a_eq.nnz)))
        a = sp.csr_matrix((np.concatenate((a_eq.data, a_ineq.data)).astype(np.float64, copy=False),
                           np.concatenate((a_eq.indices, a_ineq.indices)), indptr),
                          shape=(a_eq.shape[0] + a_ineq.shape[0], n))
        return a, np.concatenate((np.ravel(beq), np.ravel(b_ineq))).astype(np.float64), a_eq.shape[0]
    if has_eq:
        return sp.csr_matrix(a_eq), np.asarray(beq, dtype=np.float64), a_eq.shape[0]
    if has_ineq:
        return sp.csr_matrix(a_ineq), np.asarray(b_ineq, dtype=np.float64), 0
    return None, None, 0


def make_solver(c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, x0=None, alpha=1, theta=1, device=None, flags=0,
                distributed=None, partition_granule=0, kernel_variant=0, long_row_threshold=0):
    """Upload an LP given with the arguments of ``chambolle_pock_ppd`` and return the live solver.

    Returns None when the LP has no constraint row at all (closed-form case, ``:147-151``).
    """
    n = np.size(c)
    if a_eq is not None and a_eq.shape[0] == 0:  # :70-72
        a_eq, beq = None, None
    a_ineq, b_ineq = one_sided_rows(a_ineq, b_lower, b_upper)
    if a_eq is not None and a_eq.shape[0] != np.size(beq):
        raise ValueError("a_eq has %d rows but beq has %d entries" % (a_eq.shape[0], np.size(beq)))
    if a_ineq is not None and a_ineq.shape[0] != np.size(b_ineq):
        raise ValueError("a_ineq has %d rows but its bounds have %d entries" % (a_ineq.shape[0], np.size(b_ineq)))
    a, b, m_eq = stack_operator(a_eq, beq, a_ineq, b_ineq, n)
    if a is None:
        return None
    return CpPpdSolver(c, a, m_eq, b, lb, ub, x0=x0, alpha=alpha, theta=theta, device=device, flags=flags,
                       process_group=_resolve_group(distributed), partition_granule=partition_granule,
                       kernel_variant=kernel_variant, long_row_threshold=long_row_threshold)


def _resolve_group(distributed):
    """None -> the default torch.distributed group when one with more than one rank is initialised
    (SPMD use: every rank calls the solver with the same LP); False -> single GPU; a ProcessGroup -> it."""
    if distributed is False:
        return None
    try:
        import torch.distributed as dist
    except Exception:
        return None
    if distributed is None or distributed is True:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.group.WORLD
        return None
    return distributed


def run_schedule(solver, nb_max_iter, callback_func=None, max_time=None, force_integer=False, nb_iter_plot=10,
                 verbose=False, start=None, stats_func=None):
    """The reference's ``while`` loop (``:195-343``) driven over a live solver handle.

    Iterations are issued asynchronously in blocks of ``nb_iter_plot``; the host only waits at the
    stats iterations (``niter % nb_iter_plot == 0``, including 0).  Returns ``(x, best_integer)``.
    ``stats_func(niter, stats, elapsed)`` receives the whole stats block without x leaving the device.
    """
    start = time.perf_counter() if start is None else start
    nb_iter_plot = int(nb_iter_plot)
    if nb_iter_plot < 1:
        raise ValueError("nb_iter_plot must be >= 1")
    niter = 0
    agree = getattr(solver, "agree", None) if getattr(solver, "world", 1) > 1 else None
    # get_x() is a collective when the solve is distributed: every rank fetches x if any rank has a callback
    fetch_x = callback_func is not None
    if agree is not None:
        fetch_x = agree(fetch_x, any_rank=True)
    while niter < nb_max_iter:
        # iteration `niter` is a stats iteration (niter % nb_iter_plot == 0 by construction)
        solver.primal_step(keep_d=True)
        # the reference reads its clock after the (synchronous) primal step (:243): wait for the device first, so
        # that `elapsed` — the time-out test and the time axis of the caller's curves — includes the block of
        # iterations that was queued before this one
        solver.sync()
        elapsed = time.perf_counter() - start
        if max_time is not None:
            timed_out = elapsed > max_time
            if agree is not None:  # ranks read different clocks: rank 0 decides for all
                timed_out = agree(timed_out)
            if timed_out:
                break
        solver.stats_step(force_integer)
        st = solver.read_stats()
        if verbose:
            print("iter%d: energy1= %r energy2=%r elapsed %r second max violated inequality:%r "
                  "max violated equality:%r x3 has %r %% of zeros" % (
                      niter, st["energy1"], st["energy2"], elapsed, st["max_violated_inequality"],
                      st["max_violated_equality"], 100 * st["frac_zero_xbar"]))
        if stats_func is not None:
            stats_func(niter, st, elapsed)
        if fetch_x:
            x_now = solver.get_x()
            if callback_func is not None:
                callback_func(niter, x_now, st["energy1"], st["energy2"], elapsed,
                              st["max_violated_equality"], st["max_violated_inequality"])
        solver.dual_step()
        niter += 1
        k = min(nb_iter_plot - 1, nb_max_iter - niter)
        if k > 0:
            solver.iterate(k)
            niter += k
    x = solver.get_x()
    last = solver.read_stats_or_none()
    best = solver.get_best_integer() if last is not None and last["have_best_integer"] else None
    return x, best


def chambolle_pock_ppd(
    c,
    a_eq,
    beq,
    a_ineq,
    b_lower,
    b_upper,
    lb,
    ub,
    x0=None,
    alpha=1,
    theta=1,
    nb_max_iter=100,
    callback_func=None,
    max_time=None,
    save_problem=False,
    force_integer=False,
    nb_iter_plot=10,
    *,
    device=None,
    verbose=False,
    flags=0,
    return_solver=False,
    distributed=None,
    partition_granule=0,
    kernel_variant=0,
    long_row_threshold=0,
    n_gpus=None,
    y0=None,
    timings=None,
):
    """minimise ``c.x``  s.t.  ``a_eq x = beq``, ``b_lower <= a_ineq x <= b_upper``, ``lb <= x <= ub``.

    Semantics kept from the reference (file:line in ``pysparselp/ChambollePockPPD.py``):

    * the stats block runs when ``niter % nb_iter_plot == 0`` — including iteration 0 and
      when ``callback_func`` is None (``:242``); it sees the x / xbar of this iteration and
      the y of the previous one;
    * ``max_time`` is only tested there, after the primal half of the iteration and before
      the stats, the callback and the dual half (``:243-247``) — on time-out the returned x
      already contains that primal step and the callback of that iteration is not made;
    * ``callback_func(niter, x, energy1, energy2, elapsed, max_violated_equality,
      max_violated_inequality)`` (``:319-329``) receives a fresh host array;
      ``max_violated_inequality`` is evaluated at x (``:283``), ``max_violated_equality`` at
      the extrapolated point (``:269``);
    * returns ``(x, best_integer_solution)`` (``:344-346``); with neither equality nor
      inequality rows the reference returns the bare closed-form x (``:147-151``), so does this.

    ``timings`` (keyword-only extension): a dict that receives the wall-clock seconds of the phases of this call
    (``setup``: upload + operator build, ``iterate``: the schedule incl. the final read-back of x, ``close``).
    """
    start = time.perf_counter()
    c = np.ascontiguousarray(c, dtype=np.float64)
    n = c.size
    lb = _as_f64(lb, n, "lb")  # the reference asserts these sizes (:95-96)
    ub = _as_f64(ub, n, "ub")
    if n_gpus is not None and int(n_gpus) > 1:
        # one call, one process, several GPUs: helper processes for the other devices (pysparselp_b200/multi_gpu.py)
        from .multi_gpu import solve_on_gpus

        if return_solver:
            raise ValueError("return_solver is not available with n_gpus > 1 (the solver state lives in several processes)")
        kw = dict(alpha=alpha, theta=theta, nb_max_iter=nb_max_iter, max_time=max_time, save_problem=save_problem,
                  force_integer=force_integer, nb_iter_plot=nb_iter_plot, verbose=verbose, flags=flags,
                  partition_granule=partition_granule, kernel_variant=kernel_variant, long_row_threshold=long_row_threshold)
        return solve_on_gpus(n_gpus, dict(c=c, a_eq=a_eq, beq=beq, a_ineq=a_ineq, b_lower=b_lower, b_upper=b_upper, lb=lb,
                                          ub=ub, x0=x0), kw, callback_func=callback_func,
                             devices=device if isinstance(device, (list, tuple)) else None)
    if save_problem:  # :99-112
        import pickle

        a_in_1s, b_in_1s = one_sided_rows(a_ineq, b_lower, b_upper)
        with open("LP_problem2.pkl", "wb") as f:
            pickle.dump({"c": c, "a_eq": a_eq, "beq": beq, "a_ineq": a_in_1s, "b_ineq": b_in_1s,
                         "lb": lb, "ub": ub}, f)
    solver = make_solver(c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, x0=x0, alpha=alpha, theta=theta,
                         device=device, flags=flags, distributed=distributed, partition_granule=partition_granule,
                         kernel_variant=kernel_variant, long_row_threshold=long_row_threshold)
    if timings is not None:
        if solver is not None:
            solver.sync()
        timings["setup"] = time.perf_counter() - start
    if solver is not None and y0 is not None:
        # dual warm start (extension: the reference always starts from y = 0, :166,:177).  y0 = [y_eq; y_ineq] in the
        # row order of the ONE-SIDED system (finite uppers, then negated finite lowers, :74-88)
        try:
            solver.set_y(y0)
        except BaseException:
            solver.close()
            raise
    if solver is None:  # no constraint row: closed form, bare vector (:147-151)
        x = np.zeros_like(lb)
        x[c > 0] = lb[c > 0]
        x[c < 0] = ub[c < 0]
        return x
    try:
        x, best = run_schedule(solver, nb_max_iter, callback_func, max_time, force_integer, nb_iter_plot, verbose, start)
    except BaseException:
        solver.close()
        raise
    if timings is not None:
        timings["iterate"] = time.perf_counter() - start - timings["setup"]
    if return_solver:
        return x[:n], best, solver
    solver.close()
    if timings is not None:
        timings["close"] = time.perf_counter() - start - timings["setup"] - timings["iterate"]
    return x[:n], best

