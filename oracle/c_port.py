"""TEST INFRASTRUCTURE — ctypes wrapper of the plain-C oracle (oracle/cpppd_oracle.c).

Used as an independent second restatement (tests) and as the multi-threaded CPU baseline in
bench.py.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "cpppd_oracle.c")
LIB = os.path.join(_HERE, "_build", "libcpppd_oracle.so")
_lib = None


def _cpu_signature():
    """The library is built with -march=native and travels with the repo snapshot (built files are not
    git-tracked but are shipped to the GPU box): rebuild when the host CPU is not the one it was built on."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " ".join(sorted(line.split(":", 1)[1].split()))
    except OSError:
        pass
    return "unknown"


def build(force=False):
    stamp = LIB + ".cpu"
    sig = _cpu_signature()
    fresh = os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC)
    same_cpu = os.path.isfile(stamp) and open(stamp).read() == sig
    if not force and fresh and same_cpu:
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True, capture_output=True)
    with open(stamp, "w") as f:
        f.write(sig)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def set_threads(threads=None):
    """Use `threads` OpenMP threads (default: every core this process may run on) whatever OMP_NUM_THREADS
    says — torchrun exports OMP_NUM_THREADS=1 to its ranks.  Returns the count the loops really use."""
    if threads is None:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    fn = lib().cpppd_c_set_threads
    fn.restype = C.c_int
    return int(fn(C.c_int(int(threads))))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class COracle:
    """Bare loop (no stats block) of the reference solver in C; same inputs as the numpy oracle."""

    def __init__(self, c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, x0=None, alpha=1, theta=1):
        from .cpppd_oracle import one_sided_system

        if a_eq is not None and a_eq.shape[0] == 0:
            a_eq, beq = None, None
        a_ineq, b_ineq = one_sided_system(a_ineq, b_lower, b_upper)
        blocks = [a for a in (a_eq, a_ineq) if a is not None]
        a = sp.vstack(blocks).tocsr() if len(blocks) > 1 else sp.csr_matrix(blocks[0])
        self.m_eq = a_eq.shape[0] if a_eq is not None else 0
        self.has_eq, self.has_ineq = int(a_eq is not None), int(a_ineq is not None)
        self.m, self.n = a.shape
        self.b = np.ascontiguousarray(np.concatenate([np.ravel(v) for v in (beq, b_ineq) if v is not None]), dtype=np.float64)
        self.rowptr = a.indptr.astype(np.int64)
        self.colidx = a.indices.astype(np.int32)
        self.val = np.ascontiguousarray(a.data, dtype=np.float64)
        # CSC with row-sorted columns (scipy's csr->csc walks the rows in order, so entries of a
        # column come out sorted by row — the accumulation order of csc_matvec)
        csc = a.tocsc()
        self.rowidx = np.ascontiguousarray(csc.indices, dtype=np.int32)
        self.cval = np.ascontiguousarray(csc.data, dtype=np.float64)
        self.colptr = csc.indptr.astype(np.int64)
        del csc
        self.c = np.ascontiguousarray(c, dtype=np.float64)
        self.lb = np.ascontiguousarray(lb, dtype=np.float64)
        self.ub = np.ascontiguousarray(ub, dtype=np.float64)
        self.theta, self.opt = float(theta), float(1 + theta)
        self.x = np.zeros(self.n) if x0 is None else np.array(x0, dtype=np.float64)
        self.xbar = self.x.copy()
        self.y = np.zeros(self.m)
        self.T = np.empty(self.n)
        self.sigma = np.empty(self.m)
        lib().cpppd_c_precond(C.c_int64(self.n), C.c_int64(self.m), C.c_int64(self.m_eq), self.has_eq, self.has_ineq,
                              _p(self.rowptr), _p(self.val), _p(self.colptr), _p(self.rowidx), _p(self.cval),
                              C.c_double(float(alpha)), _p(self.T), _p(self.sigma))

    def iterate(self, k):
        lib().cpppd_c_iterate(C.c_int64(k), C.c_int64(self.n), C.c_int64(self.m), C.c_int64(self.m_eq), self.has_eq,
                              self.has_ineq, _p(self.rowptr), _p(self.colidx), _p(self.val), _p(self.colptr),
                              _p(self.rowidx), _p(self.cval), _p(self.c), _p(self.b), _p(self.lb), _p(self.ub),
                              _p(self.T), _p(self.sigma), C.c_double(self.theta), C.c_double(self.opt),
                              _p(self.x), _p(self.xbar), _p(self.y))

    def primal_step(self):
        """Primal half of one iteration (reference :198-228): what a callback of that iteration sees in x."""
        lib().cpppd_c_primal(C.c_int64(self.n), C.c_int64(self.m_eq), self.has_eq, self.has_ineq, _p(self.colptr),
                             _p(self.rowidx), _p(self.cval), _p(self.y), _p(self.c), _p(self.T), _p(self.lb),
                             _p(self.ub), C.c_double(self.theta), C.c_double(self.opt), _p(self.x), _p(self.xbar))

    def dual_step(self):
        """Dual half (reference :231-240, :333-341)."""
        lib().cpppd_c_dual(C.c_int64(self.m), C.c_int64(self.m_eq), _p(self.rowptr), _p(self.colidx), _p(self.val),
                           _p(self.xbar), _p(self.b), _p(self.sigma), _p(self.y))
