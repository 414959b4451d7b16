"""TEST INFRASTRUCTURE — drive the CPU-emulated build of libcpppd (tests/emul/make_emul.py) through the
same C ABI and the same Python schedule as the product, without CUDA and without torch.

`EmulatedSolver` binds tests/emul/_build/libcpppd_emul.so directly with ctypes (it never touches
`pysparselp_b200._cabi.load_library`, so the product keeps having exactly one library and no CPU path) and
reuses the product's `prepare_problem`, `SolverHandle` methods and `run_schedule`.
"""
import ctypes as C
import time

import numpy as np

from pysparselp_b200 import _cabi
from pysparselp_b200.ChambollePockPPD import (SolverHandle, one_sided_rows, prepare_problem, run_schedule,
                                              stack_operator)

from . import make_emul

_lib = None


def emulated_library():
    global _lib
    if _lib is None:
        lib = C.CDLL(make_emul.build())
        for name, (restype, argtypes) in _cabi.SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        assert lib.cpppd_abi_version() == _cabi.ABI_VERSION
        _lib = lib
    return _lib


class EmulatedSolver(SolverHandle):
    def __init__(self, c, a, m_eq, b, lb, ub, x0=None, alpha=1, theta=1, flags=0, partition_granule=0,
                 rank=0, world=1, comm_id=None, kernel_variant=0, long_row_threshold=0, comm=None):
        self.lib = emulated_library()
        p, keep = prepare_problem(c, a, m_eq, b, lb, ub, x0, alpha, theta, flags, partition_granule, kernel_variant,
                                  long_row_threshold)
        self.n, self.m, self.m_eq = int(p.n), int(p.m_eq + p.m_ineq), int(p.m_eq)
        p.device = 0
        p.stream = None
        p.alloc = C.cast(None, _cabi.ALLOC_FN)
        p.free = C.cast(None, _cabi.FREE_FN)
        p.rank, p.world_size = rank, world
        p.comm_id = None if comm_id is None else comm_id.ctypes.data
        p.comm = comm  # a cpppd_comm kept between solves (emulated_comm): owner of the peer-memory pool
        handle = C.c_void_p()
        _cabi.check(self.lib, None, self.lib.cpppd_create(C.byref(p), C.byref(handle)))
        del keep
        self.handle = handle


def make_emulated_solver(c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, x0=None, alpha=1, theta=1, flags=0,
                         partition_granule=0, rank=0, world=1, comm_id=None, kernel_variant=0, long_row_threshold=0,
                         comm=None):
    if a_eq is not None and a_eq.shape[0] == 0:
        a_eq, beq = None, None
    a_ineq, b_ineq = one_sided_rows(a_ineq, b_lower, b_upper)
    a, b, m_eq = stack_operator(a_eq, beq, a_ineq, b_ineq, np.size(c))
    return EmulatedSolver(c, a, m_eq, b, lb, ub, x0=x0, alpha=alpha, theta=theta, flags=flags,
                          partition_granule=partition_granule, rank=rank, world=world, comm_id=comm_id,
                          kernel_variant=kernel_variant, long_row_threshold=long_row_threshold, comm=comm)


def emulated_comm(comm_id, rank, world):
    """cpppd_comm_create on the emulated library: a communicator that outlives the solves made on it."""
    comm = C.c_void_p()
    _cabi.check(emulated_library(), None, emulated_library().cpppd_comm_create(comm_id.ctypes.data, rank, world, 0,
                                                                                C.byref(comm)))
    return comm


def emulated_chambolle_pock_ppd(c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, x0=None, alpha=1, theta=1,
                                nb_max_iter=100, callback_func=None, max_time=None, force_integer=False,
                                nb_iter_plot=10, flags=0, partition_granule=0, kernel_variant=0,
                                long_row_threshold=0):
    """(x, best_integer, solver) — the product schedule over the emulated library."""
    start = time.perf_counter()
    solver = make_emulated_solver(c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, x0=x0, alpha=alpha, theta=theta,
                                  flags=flags, partition_granule=partition_granule, kernel_variant=kernel_variant,
                                  long_row_threshold=long_row_threshold)
    x, best = run_schedule(solver, nb_max_iter, callback_func, max_time, force_integer, nb_iter_plot, False, start)
    return x, best, solver


def new_comm_id():
    """128-byte id of a fresh emulated communicator (the in-process NCCL of tests/emul/emul_nccl.cpp)."""
    import os

    os.environ["CPPPD_NCCL_LIB"] = make_emul.LIB  # libcpppd resolves "NCCL" inside the emulated library itself
    ident = np.zeros(128, dtype=np.uint8)
    _cabi.check(emulated_library(), None, emulated_library().cpppd_comm_unique_id(ident.ctypes.data))
    return ident


def run_ranks(world, body, timeout_s=600):
    """Run ``body(rank, world, comm_id)`` on one Python thread per emulated rank (ctypes releases the GIL inside
    the library, so the ranks really run concurrently); returns the list of results, re-raises the first error."""
    import threading

    comm_id = new_comm_id()
    results, errors = [None] * world, []

    def runner(r):
        try:
            results[r] = body(r, world, comm_id)
        except BaseException as e:  # noqa: BLE001 - reported to the test
            errors.append((r, e))

    threads = [threading.Thread(target=runner, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    deadline = time.time() + timeout_s
    for t in threads:
        t.join(max(0.0, deadline - time.time()))
    if any(t.is_alive() for t in threads):
        raise TimeoutError("emulated ranks did not finish (dead-lock?); errors so far: %r" % (errors,))
    if errors:
        raise errors[0][1]
    return results
