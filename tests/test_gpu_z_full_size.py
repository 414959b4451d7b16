"""GPU, BASELINE.json's full sizes: the bench workload itself (Potts 4096^2, configs[4]), a random sparse LP of the
configs[3] family and an L1-SVM LP of the configs[2] family with the bench's entry count, checked through properties
that do not need a golden file of that size:

* bit-identity with the plain-C OpenMP oracle port (oracle/cpppd_oracle.c) after a few iterations — the port
  itself is pinned to the reference goldens by tests/test_oracle_golden.py;
* iterate invariants (box, sign of y_ineq);
* a checksum of checksums: the renumbered + compressed storage (other kernels, other slice shapes, other index
  words) must give the same bits as the generic one.

(Sorted last on purpose: with `-x` the small, golden-pinned cases run first.)
"""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# the sizes can be scaled down to dry-run this file on the CPU emulation of the library (tests/README.md)
POTTS_SIDE = int(os.environ.get("CPPPD_FULL_SIZE_POTTS", "4096"))
RANDOM_N = int(os.environ.get("CPPPD_FULL_SIZE_RANDOM_N", "2000000"))
SVM_SAMPLES = int(os.environ.get("CPPPD_FULL_SIZE_SVM_SAMPLES", "50000"))
SVM_FEATURES = int(os.environ.get("CPPPD_FULL_SIZE_SVM_FEATURES", "1000"))


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


def solve_on_gpu(args, iters, **kw):
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    trace = []
    x, _, solver = chambolle_pock_ppd(*args, nb_max_iter=iters, nb_iter_plot=iters, return_solver=True,
                                      callback_func=lambda k, xx, e1, e2, el, a, b: trace.append((k, e1, e2, a, b)),
                                      **kw)
    try:
        return x, solver.get_y(), solver.info(), trace
    finally:
        solver.close()


def c_port_iterates(args, iters):
    from oracle.c_port import COracle

    co = COracle(*args)
    co.iterate(iters)
    return co.x.copy(), co.y.copy()


def test_potts_4096_bench_workload_bit_identical_to_the_c_port():
    """configs[4] — n = 50 323 456, m = 67 092 480, 201 277 440 entries: x and y after 6 iterations."""
    from pysparselp_b200 import generators

    lp = generators.potts_lp(POTTS_SIDE)
    args = generators.lp_args(lp)
    iters = 6
    xo, yo = c_port_iterates(args, iters)
    x, y, info, trace = solve_on_gpu(args, iters)
    edges = 2 * POTTS_SIDE * (POTTS_SIDE - 1)
    assert (info["n"], info["m_ineq"], info["nnz"]) == (POTTS_SIDE**2 + edges, 2 * edges, 6 * edges)
    if POTTS_SIDE == 4096:
        assert (info["n"], info["m_ineq"], info["nnz"]) == (50323456, 67092480, 201277440)
    # the hot kernels are chosen by measurement on these operands (>= 2^22 entries)
    assert info["autotuned"] == (1 if info["nnz"] >= 1 << 22 else 0)
    assert np.array_equal(x, xo) and np.array_equal(y, yo)
    assert np.all(x >= lp.lb) and np.all(x <= lp.ub) and np.all(y >= 0)
    # iteration 0 stats: x = clip(-T c) in {0, 1}, y = 0 -> energy1 = c.x exactly representable terms
    k, e1, e2, mv_eq, mv_ineq = trace[0]
    assert k == 0 and mv_eq == 0.0 and np.isfinite(e1) and np.isfinite(e2) and np.isfinite(mv_ineq)
    want = digest(xo, yo)
    del xo, yo
    # renumbering + value dictionary + constant vectors: different operands, same bits
    x2, y2, info2, _ = solve_on_gpu(args, iters, flags=11)
    assert info2["value_bytes"] == 0 and info2["const_vector_mask"] != 0
    assert digest(x2, y2) == want == digest(x, y)


def test_random_lp_with_equalities_bit_identical_to_the_c_port():
    """configs[3] family at 1/10 size (2 M variables, 3.6 M inequalities + 0.4 M equalities, 32 M entries): ragged
    columns (automatic renumbering), equality rows, two-digit values."""
    from pysparselp_b200 import generators

    n, m_eq = RANDOM_N, RANDOM_N // 5
    lp, _ = generators.random_sparse_lp(n, 2 * n - m_eq, n_eq=m_eq, seed=0)
    args = generators.lp_args(lp)
    iters = 10
    xo, yo = c_port_iterates(args, iters)
    x, y, info, _ = solve_on_gpu(args, iters)
    assert info["nnz"] == 16 * n and info["m_eq"] == m_eq
    assert np.array_equal(x, xo) and np.array_equal(y, yo)
    assert np.all(x >= lp.lb) and np.all(x <= lp.ub) and np.all(y[m_eq:] >= 0)
    x2, y2, _, _ = solve_on_gpu(args, iters, flags=64)  # CPPPD_FLAG_NO_REORDER: the caller's numbering
    assert digest(x2, y2) == digest(xo, yo)


def test_l1svm_with_1000_features_agrees_with_the_c_port():
    """configs[2] family — 50 000 samples x 1 000 features, K = 3 (106 006 rows, 56 006 columns, 200 M entries: 1/20
    of the samples of configs[2], the entry count of the bench workload).  Rows of 2 003 entries stay on the
    thread-per-row path; the 3 003 weight / bias columns (66 667 entries each) go through the long-row kernels, whose fixed summation tree agrees with scipy's sequential sums to rounding: 1e-9 relative
    (BASELINE.json's bound for the iterates), not bit for bit."""
    from pysparselp_b200 import generators

    lp, _ = generators.l1svm_lp(SVM_SAMPLES, SVM_FEATURES)
    args = generators.lp_args(lp)
    iters = 5
    xo, yo = c_port_iterates(args, iters)
    x, y, info, trace = solve_on_gpu(args, iters)
    assert info["nnz"] == lp.a_ineq.nnz and info["long_rows"] == 0
    if SVM_SAMPLES * 4 // 3 > 2048:
        assert info["long_cols"] >= 3 * SVM_FEATURES and info["long_entries"] > info["nnz"] // 2
    for got, want in ((x, xo), (y, yo)):
        assert np.max(np.abs(got - want)) <= 1e-9 * max(np.max(np.abs(want)), 1e-300)
    assert np.all(x >= lp.lb) and np.all(x <= lp.ub)
    assert all(np.isfinite(row[1]) for row in trace)  # energy1 (energy2 is NaN by construction: free weights, :260-263)
    # the same LP without the long-row path: bit for bit (one thread per weight column: slow, but exact)
    if SVM_SAMPLES <= 5000:
        x2, y2, info2, _ = solve_on_gpu(args, iters, long_row_threshold=-1)
        assert info2["long_cols"] == 0 and np.array_equal(x2, xo) and np.array_equal(y2, yo)
