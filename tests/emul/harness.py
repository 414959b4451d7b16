"""TEST INFRASTRUCTURE — run the real k_primal / k_dual CUDA sources on the CPU.

`tests/emul/emul_kernels.cpp` compiles `pysparselp_b200/csrc/cpppd_device_types.cuh` and
`cpppd_hot_kernels.cuh` with g++ through `cuda_shim.h`.  This module builds the operands the kernels
expect — A and A^T in SELL-32 exactly as `k_fill_sell` lays them out (entry order kept, A^T entries in
original row order with the equality bit, INT32_MIN padding, optional dictionary packing) — and drives
the iteration.  What it checks is the kernel *logic* (indexing, predicates, accumulation order, epilogue);
what only hardware can check (memory model, scheduling) stays with the `-m gpu` tests.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pysparselp_b200", "csrc")
LIB = os.path.join(HERE, "_build", "libemul.so")
_lib = None


class EmulVec(C.Structure):
    _fields_ = [("p", C.c_void_p), ("c", C.c_double)]


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("emul_kernels.cpp", "cuda_shim.h")] + [
        os.path.join(CSRC, f) for f in ("cpppd_device_types.cuh", "cpppd_hot_kernels.cuh")]
    if not force and os.path.isfile(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
           "-o", LIB, os.path.join(HERE, "emul_kernels.cpp")]
    subprocess.run(cmd, check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def to_sell(indptr, indices, values, dictionary=None, idx_bits=30):
    """CSR (entry order kept) -> SELL-32 arrays as k_fill_sell writes them."""
    kpad = np.int32(-2**31)
    indptr = np.asarray(indptr, dtype=np.int64)
    nrows = indptr.size - 1
    nslices = (nrows + 31) // 32
    lens = np.diff(indptr)
    widths = np.array([lens[32 * s: 32 * s + 32].max() if lens[32 * s: 32 * s + 32].size else 0 for s in range(nslices)],
                      dtype=np.int64)
    slice_ptr = np.concatenate(([0], np.cumsum(widths * 32))).astype(np.int64)
    idx = np.full(int(slice_ptr[-1]), kpad, dtype=np.int32)
    val = np.zeros(int(slice_ptr[-1]), dtype=np.float64)
    codes = None
    if dictionary is not None:
        bits = values.view(np.uint64)
        codes = np.searchsorted(dictionary.view(np.uint64), bits)
        assert np.array_equal(dictionary.view(np.uint64)[codes], bits)
    for r in range(nrows):
        s, lane = divmod(r, 32)
        e0, n = indptr[r], lens[r]
        pos = slice_ptr[s] + np.arange(n) * 32 + lane
        w = indices[e0: e0 + n].astype(np.int64)
        if dictionary is not None:
            w = (w & 0x40000000) | (w & ((1 << idx_bits) - 1)) | (codes[e0: e0 + n] << idx_bits)
        idx[pos] = w.astype(np.int32)
        val[pos] = values[e0: e0 + n]
    uniform = int(widths[0]) if nslices and np.all(widths == widths[0]) else -1
    return dict(slice_ptr=slice_ptr, idx=idx, val=val, nrows=nrows, nslices=nslices, uniform_width=uniform)


class EmulSolver:
    """The solver loop with the real kernel sources on the CPU (single rank, original numbering)."""

    def __init__(self, c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, x0=None, alpha=1, theta=1,
                 value_dict=False, const_vectors=False, variant=0):
        from oracle.cpppd_oracle import CpPpdOracle, one_sided_system

        if a_eq is not None and a_eq.shape[0] == 0:
            a_eq, beq = None, None
        a_ineq1, b_ineq = one_sided_system(a_ineq, b_lower, b_upper)
        blocks = [sp.csr_matrix(a) for a in (a_eq, a_ineq1) if a is not None]
        data = np.concatenate([b_.data for b_ in blocks]).astype(np.float64)
        indices = np.concatenate([b_.indices for b_ in blocks]).astype(np.int64)
        indptr = np.concatenate(([0], np.cumsum(np.concatenate([np.diff(b_.indptr) for b_ in blocks])))).astype(np.int64)
        m, n = indptr.size - 1, c.size
        self.m_eq = a_eq.shape[0] if a_eq is not None else 0
        self.has_eq, self.has_ineq = int(a_eq is not None), int(a_ineq1 is not None)
        self.n, self.m = n, m
        self.b = np.concatenate([np.ravel(v) for v in (beq, b_ineq) if v is not None]).astype(np.float64)
        ref = CpPpdOracle(c, a_eq, beq, a_ineq, b_lower, b_upper, lb, ub, alpha=alpha, theta=theta)
        self.T = ref.diag_t.copy()  # preconditioners come from the oracle: only the two hot kernels are emulated
        self.sigma = np.concatenate([v for v in (ref.sig_eq, ref.sig_ineq) if v is not None])
        self.c, self.lb, self.ub = (np.ascontiguousarray(v, dtype=np.float64) for v in (c, lb, ub))
        dictionary, idx_bits = None, 30
        if value_dict:
            uniq = np.unique(data.view(np.uint64)).view(np.float64)
            bits_idx = max(int(max(n, m) - 1).bit_length(), 1)
            bits_code = max(int(uniq.size - 1).bit_length(), 1)
            if uniq.size <= 256 and bits_idx + bits_code <= 30:
                dictionary, idx_bits = np.ascontiguousarray(uniq), bits_idx
        self.dict = dictionary
        self.dict256 = None if dictionary is None else np.concatenate((dictionary, np.zeros(256 - dictionary.size)))
        self.A = to_sell(indptr, indices, data, dictionary, idx_bits)
        csr = sp.csr_matrix((data, indices, indptr), shape=(m, n))
        csc = csr.tocsc()  # entries of a column sorted by original row
        t_idx = csc.indices.astype(np.int64)
        t_idx = t_idx | np.where(t_idx < self.m_eq, 0x40000000, 0)
        self.AT = to_sell(csc.indptr, t_idx, csc.data, dictionary, idx_bits)
        self.idx_bits, self.ndict = idx_bits, 0 if dictionary is None else dictionary.size
        self.theta, self.opt = float(theta), float(1 + theta)
        self.x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
        self.xbar = self.x.copy()
        self.y = np.zeros(m)
        self.d = np.zeros(n)
        self.const_vectors = const_vectors
        self.variant = int(variant)  # 0-based index into kVariants

    def _vec(self, v):
        if self.const_vectors and v.size and np.all(v.view(np.uint64) == v.view(np.uint64)[0]):
            return EmulVec(None, float(v[0]))
        return EmulVec(v.ctypes.data, 0.0)

    def _sell_args(self, S):
        d = self.dict256
        return (S["slice_ptr"].ctypes.data_as(C.c_void_p), S["idx"].ctypes.data_as(C.c_void_p),
                S["val"].ctypes.data_as(C.c_void_p), C.c_int64(S["nrows"]), C.c_int64(S["nslices"]),
                C.c_int64(S["uniform_width"]), None if d is None else d.ctypes.data_as(C.c_void_p),
                C.c_int(self.idx_bits), C.c_int(self.ndict))

    def primal(self, write_d=False):
        lib().emul_primal(C.c_int(self.variant), C.c_int(int(write_d)), *self._sell_args(self.AT), self.y.ctypes.data_as(C.c_void_p),
                          self._vec(self.c), self._vec(self.T), self._vec(self.lb), self._vec(self.ub),
                          self.x.ctypes.data_as(C.c_void_p), self.xbar.ctypes.data_as(C.c_void_p),
                          self.d.ctypes.data_as(C.c_void_p), C.c_int(self.has_eq), C.c_int(self.has_ineq),
                          C.c_double(self.theta), C.c_double(self.opt))

    def dual(self):
        lib().emul_dual(C.c_int(self.variant), *self._sell_args(self.A), self.xbar.ctypes.data_as(C.c_void_p), self._vec(self.b),
                        self._vec(self.sigma), self.y.ctypes.data_as(C.c_void_p), C.c_int64(self.m_eq))

    def iterate(self, k):
        for _ in range(k):
            self.primal()
            self.dual()
