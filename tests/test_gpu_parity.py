"""GPU: the CUDA path (through the C ABI) against the oracle and the golden vectors.

Tolerances are BASELINE.json's: iterates after 100 iterations within 1e-9 relative
(max-norm), energy / violation curves within 1e-6 relative at every callback, integer
decisions (preconditioner "replaced by 1" masks, rounding, feasibility flags) exact.
On one GPU the kernels accumulate in the reference's order without FMA, so the iterates are
expected — and asserted — to be bit-identical.
"""
import json
import os

import numpy as np
import pytest

from conftest import CASE_PARAMS, GOLDEN, GOLDEN_CASES, case_args

pytestmark = pytest.mark.gpu

REL_ITERATE = 1e-9
REL_CURVE = 1e-6


def rel_inf(a, b):
    denom = max(np.max(np.abs(b)), 1e-300)
    return np.max(np.abs(a - b)) / denom


def traced(args, nb_max_iter, nb_iter_plot, **kw):
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    trace, xs = [], []

    def cb(niter, x, e1, e2, elapsed, mv_eq, mv_ineq):
        trace.append((niter, e1, e2, mv_eq, mv_ineq))
        xs.append(x.copy())

    x, best, solver = chambolle_pock_ppd(*args, nb_max_iter=nb_max_iter, nb_iter_plot=nb_iter_plot,
                                         callback_func=cb, return_solver=True, **kw)
    return x, best, np.array(trace), xs, solver


def assert_curves_close(got, want, scale_cols=(1, 2)):
    """NaN/inf-aware comparison of (niter, e1, e2, mv_eq, mv_ineq) rows."""
    assert got.shape == want.shape
    assert np.array_equal(got[:, 0], want[:, 0])
    for col in range(1, want.shape[1]):
        g, w = got[:, col], want[:, col]
        assert np.array_equal(np.isnan(g), np.isnan(w)), "NaN pattern differs in column %d" % col
        fin = np.isfinite(w)
        assert np.array_equal(g[~fin & ~np.isnan(w)], w[~fin & ~np.isnan(w)]), "inf pattern differs"
        # energies are sums with cancellation: tie the absolute floor to the curve's own magnitude
        floor = REL_CURVE * max(np.max(np.abs(w[fin])) if fin.any() else 0.0, 1e-30)
        assert np.all(np.abs(g[fin] - w[fin]) <= REL_CURVE * np.abs(w[fin]) + (floor if col in scale_cols else 0.0)), \
            "column %d: %r vs %r" % (col, g[fin], w[fin])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_iterates_and_curves_vs_golden(name):
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    x, best, trace, xs, solver = traced(args, 100, 10, **kw)
    try:
        y = solver.get_y()
        T, sigma = solver.get_preconditioners()
        m_eq = g["y_eq"].size if "y_eq" in g else 0
        y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
        s_gold = np.concatenate([g[k] for k in ("diag_sigma_eq", "diag_sigma_ineq") if k in g])
        assert solver.niter == 100
        # preconditioners: exact sparsity of the "replaced by 1" entries + values
        assert np.array_equal(T == 1.0, g["diag_t"] == 1.0)
        if "alpha" not in kw:
            assert np.array_equal(T, g["diag_t"]) and np.array_equal(sigma, s_gold)
        else:  # general exponent goes through pow(): last-ulp differences allowed
            assert rel_inf(T, g["diag_t"]) < 1e-14 and rel_inf(sigma, s_gold) < 1e-14
        assert rel_inf(x, g["x_100"]) <= REL_ITERATE
        assert rel_inf(y, y_gold) <= REL_ITERATE
        assert rel_inf(xs[5], g["x_at_50"]) <= REL_ITERATE
        if "alpha" not in kw:
            assert np.array_equal(x, g["x_100"]), "single-GPU iterates should be bit-identical"
            assert np.array_equal(y, y_gold)
        assert_curves_close(trace, g["trace_10"])
        assert (best is None) == (g["best_100"].size == 0)
        assert m_eq == solver.m_eq
    finally:
        solver.close()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_force_integer_bookkeeping_vs_golden(name):
    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    x, best, trace, xs, solver = traced(args, 300, 20, force_integer=True, **kw)
    solver.close()
    assert rel_inf(x, g["x_300_fi"]) <= REL_ITERATE
    assert_curves_close(trace, g["trace_20_fi"])
    if g["best_300_fi"].size:
        assert best is not None and np.array_equal(best, g["best_300_fi"])  # rounded values: exact
    else:
        assert best is None


@pytest.mark.parametrize("name", ["potts50", "sc105", "random_small"])
def test_against_oracle_live(name):
    """Same seeded inputs through the oracle and the CUDA path, different schedule than the goldens."""
    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle

    args, _ = case_args(name)
    tr_o, st = [], {}
    with np.errstate(invalid="ignore"):
        xo, bo = chambolle_pock_ppd_oracle(
            *args, nb_max_iter=237, nb_iter_plot=7, state_out=st,
            callback_func=lambda k, x, e1, e2, el, a, b: tr_o.append((k, e1, e2, a, b)))
    x, best, trace, xs, solver = traced(args, 237, 7)
    y = solver.get_y()
    xbar = solver.get_xbar()
    solver.close()
    yo = np.concatenate([v for v in (st["y_eq"], st["y_ineq"]) if v is not None])
    assert np.array_equal(x, xo) and np.array_equal(y, yo) and np.array_equal(xbar, st["xbar"])
    assert_curves_close(trace, np.array(tr_o))


def test_x0_warm_start_and_no_callback():
    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    args, _ = case_args("random_small")
    rng = np.random.default_rng(3)
    x0 = rng.standard_normal(args[0].size)
    x0_copy = x0.copy()
    with np.errstate(invalid="ignore"):
        xo, _ = chambolle_pock_ppd_oracle(*args, x0=x0, nb_max_iter=64, nb_iter_plot=1000)
    x, best = chambolle_pock_ppd(*args, x0=x0, nb_max_iter=64, nb_iter_plot=1000)
    assert np.array_equal(x0, x0_copy), "inputs must not be mutated"
    assert np.array_equal(x, xo)


def test_reference_golden_curves_through_solve():
    """The reference's own regression pins, through SparseLP.solve on the GPU
    (reference tests/test_netlib.py:90-117, tests/test_pott_segmentation.py:20-37)."""
    from pysparselp_b200.examples.example_pott_segmentation import build_linear_program
    from pysparselp_b200.netlib import get_problem
    from pysparselp_b200.SparseLP import SparseLP

    with open(os.path.join(GOLDEN, "reference_curves.json")) as f:
        ref = json.load(f)
    lp, gt, gti, _ = build_linear_program(50, 0.5, 500)
    lp.solve(method="chambolle_pock_ppd", get_timing=True, nb_iter=27500, max_time=150, ground_truth=gt,
             ground_truth_indices=gti, nb_iter_plot=500)
    curve = lp.distance_to_ground_truth
    assert len(curve) == 55
    np.testing.assert_almost_equal(curve, ref["potts50"][: len(curve)])
    assert curve[-1] == 0.0

    d = get_problem("SC105")
    gt = d["solution"]
    lp = SparseLP()
    lp.add_variables_array(len(d["cost_vector"]), lower_bounds=d["lower_bounds"],
                           upper_bounds=np.minimum(d["upper_bounds"], np.max(gt) * 2), costs=d["cost_vector"])
    lp.add_equality_constraints_sparse(d["a_eq"], d["b_eq"])
    lp.add_inequality_constraints_sparse(d["a_ineq"], d["b_lower"], d["b_upper"])
    lp.convert_to_one_sided_inequality_system()
    assert lp.check_solution(gt)
    lp.solve(method="chambolle_pock_ppd", get_timing=True, nb_iter=41500, max_time=100, ground_truth=gt,
             ground_truth_indices=np.arange(len(gt)), nb_iter_plot=500)
    curve = lp.distance_to_ground_truth
    assert len(curve) == 83
    np.testing.assert_almost_equal(curve, ref["SC105"][: len(curve)])


def test_reference_l1svm_accuracy():
    """reference tests/test_l1_svm.py: 99.4 % after 2000 iterations."""
    from pysparselp_b200.examples.example_l1_svm import run

    with open(os.path.join(GOLDEN, "reference_curves.json")) as f:
        ref = json.load(f)
    assert run()["chambolle_pock_ppd"] == ref["l1svm"]


def test_max_time_semantics():
    """:243-247 — time-out is only seen at a stats iteration; x then already holds that primal step
    and the callback of that iteration is not made."""
    from oracle.cpppd_oracle import CpPpdOracle
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    args, _ = case_args("sc105")
    calls = []
    x, best = chambolle_pock_ppd(*args, nb_max_iter=10**9, nb_iter_plot=50, max_time=0.0,
                                 callback_func=lambda *a: calls.append(a[0]))
    assert calls == []
    o = CpPpdOracle(*args)
    o.primal_step()
    assert np.array_equal(x, o.x)


def test_edge_cases():
    import scipy.sparse as sp

    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    rng = np.random.default_rng(5)
    n = 37  # ragged: n and m not multiples of the slice height, empty rows and empty columns
    a = sp.random(45, n, density=0.1, random_state=2, format="csr")
    a.data = np.round(a.data * 6 - 3, 1)
    a = a.tolil()
    a[3, :] = 0
    a[:, 5] = 0
    a = a.tocsr()
    c = rng.standard_normal(n)
    lb, ub = -np.ones(n), np.ones(n)
    b_up = rng.random(45)
    empty_eq = sp.csr_matrix((0, n))
    with np.errstate(invalid="ignore"):
        xo, _ = chambolle_pock_ppd_oracle(c, empty_eq, np.empty(0), a, None, b_up, lb, ub, nb_max_iter=50, nb_iter_plot=10)
    x, _ = chambolle_pock_ppd(c, empty_eq, np.empty(0), a, None, b_up, lb, ub, nb_max_iter=50, nb_iter_plot=10)
    assert np.array_equal(x, xo)
    # only equalities: the reference crashes at :283 (max over a missing inequality block);
    # here the maximum over an empty set is -inf and the solve goes through
    a_eq = a[:20, :]
    beq = a_eq @ (rng.random(n) * 0.5)
    seen = []
    x, _ = chambolle_pock_ppd(c, a_eq, beq, None, None, None, lb, ub, nb_max_iter=30, nb_iter_plot=10,
                              callback_func=lambda k, x, e1, e2, el, mv_eq, mv_in: seen.append(mv_in))
    assert seen and all(v == -np.inf for v in seen)
    # no constraint at all: closed form, bare vector (reference :147-151)
    x = chambolle_pock_ppd(c, empty_eq, np.empty(0), None, None, None, lb, ub)
    assert np.array_equal(x, np.where(c > 0, lb, np.where(c < 0, ub, 0.0)))
    # nb_max_iter = 0: x0 comes back untouched
    x, best = chambolle_pock_ppd(c, empty_eq, np.empty(0), a, None, b_up, lb, ub, nb_max_iter=0)
    assert np.array_equal(x, np.zeros(n)) and best is None


def test_large_properties_potts256():
    """Size-independent properties on a mid-size Potts LP (196k variables):
    agreement with the oracle after 30 iterations, and iterate invariants."""
    from oracle.cpppd_oracle import chambolle_pock_ppd_oracle
    from pysparselp_b200 import generators
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    lp = generators.potts_lp(256)
    args = generators.lp_args(lp)
    xo, _ = chambolle_pock_ppd_oracle(*args, nb_max_iter=30, nb_iter_plot=1000)
    x, _, solver = chambolle_pock_ppd(*args, nb_max_iter=30, nb_iter_plot=1000, return_solver=True)
    y = solver.get_y()
    info = solver.info()
    solver.close()
    assert np.array_equal(x, xo)
    assert np.all(x >= lp.lb) and np.all(x <= lp.ub) and np.all(y >= 0)
    assert info["nnz"] == lp.a_ineq.nnz and info["a_padded_entries"] >= info["nnz"]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_locality_reordering_keeps_iterates_bit_identical(name):
    """CPPPD_FLAG_REORDER renumbers rows and columns (the multi-GPU layout on one GPU); entry order
    inside rows / columns is untouched, so every iterate must stay bit-identical."""
    from pysparselp_b200 import _cabi

    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    x, best, trace, xs, solver = traced(args, 100, 10, flags=_cabi.FLAG_REORDER, **kw)
    try:
        y = solver.get_y()
        T, sigma = solver.get_preconditioners()
        y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
        if "alpha" not in kw:
            assert np.array_equal(x, g["x_100"]) and np.array_equal(y, y_gold) and np.array_equal(T, g["diag_t"])
        else:
            assert rel_inf(x, g["x_100"]) <= REL_ITERATE and rel_inf(y, y_gold) <= REL_ITERATE
        assert_curves_close(trace, g["trace_10"])
        owned, ghost = solver.layout(columns=True)
        assert ghost.size == 0 and np.array_equal(np.sort(owned), np.arange(args[0].size))
    finally:
        solver.close()


def test_partition_matches_python_restatement():
    """The row / column order computed on the device equals oracle/partition_oracle.py (world = 1)."""
    from oracle import partition_oracle as po
    from pysparselp_b200 import _cabi, generators
    from pysparselp_b200.ChambollePockPPD import make_solver

    for lp, m_eq in ((generators.potts_lp(40), 0), (generators.random_sparse_lp(700, 900, n_eq=150, seed=4)[0], 150)):
        solver = make_solver(*generators.lp_args(lp), flags=_cabi.FLAG_REORDER, partition_granule=32)
        cols, _ = solver.layout(columns=True)
        rows, _ = solver.layout(columns=False)
        solver.close()
        blocks = [a for a in (lp.a_eq, lp.a_ineq) if a is not None]
        import scipy.sparse as sp

        a = sp.vstack(blocks).tocsr() if len(blocks) > 1 else blocks[0]
        part = po.partition(a.indptr, a.indices, a.shape[1], m_eq, 1, granule=32)
        assert np.array_equal(cols, part["col_order"]) and np.array_equal(rows, part["row_order"])


@pytest.mark.parametrize("world", [2])
def test_multi_gpu_parity(world):
    """Row/column-partitioned solve on `world` GPUs (one process per GPU, NCCL halo exchange)."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29613", os.path.join(here, "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "DIST_WORKER_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_lossless_compression_keeps_iterates_bit_identical(name):
    """CPPPD_FLAG_VALUE_DICT | CPPPD_FLAG_CONST_VECTORS change the storage, not a single bit of the
    arithmetic: dictionary entries are the original doubles, folded vectors were constant."""
    from pysparselp_b200 import _cabi

    args, g = case_args(name)
    kw = CASE_PARAMS.get(name, {})
    flags = _cabi.FLAG_VALUE_DICT | _cabi.FLAG_CONST_VECTORS
    x, best, trace, xs, solver = traced(args, 100, 10, flags=flags, **kw)
    try:
        info = solver.info()
        y = solver.get_y()
        T, sigma = solver.get_preconditioners()
        y_gold = np.concatenate([g[k] for k in ("y_eq", "y_ineq") if k in g])
        if "alpha" not in kw:
            assert np.array_equal(x, g["x_100"]) and np.array_equal(y, y_gold) and np.array_equal(T, g["diag_t"])
        else:
            assert rel_inf(x, g["x_100"]) <= REL_ITERATE and rel_inf(y, y_gold) <= REL_ITERATE
        assert_curves_close(trace, g["trace_10"])
        if name == "potts50":  # values are +-1, b / sigma / lb / ub are constant vectors
            assert info["value_bytes"] == 0 and (info["const_vector_mask"] & 0xF) == 0xF
        if name == "random_small":  # hundreds of distinct values: the generic format must be kept
            assert info["value_bytes"] == 8
    finally:
        solver.close()
    x, best, trace, xs, solver = traced(args, 300, 20, force_integer=True, flags=flags | _cabi.FLAG_REORDER, **kw)
    solver.close()
    assert_curves_close(trace, g["trace_20_fi"])
    if g["best_300_fi"].size:
        assert best is not None and np.array_equal(best, g["best_300_fi"])


def test_solve_curves_on_device_equal_host_curves():
    """SparseLP.solve(): curves evaluated inside the stats block (x stays on the GPU) vs the reference-style
    host evaluation (x downloaded at every callback)."""
    from pysparselp_b200.examples.example_pott_segmentation import build_linear_program

    out = {}
    for device_curves in (True, False):
        lp, gt, gti, _ = build_linear_program(50, 0.5, 500)
        x, _ = lp.solve(method="chambolle_pock_ppd", nb_iter=3001, nb_iter_plot=500, ground_truth=gt,
                        ground_truth_indices=gti, device_curves=device_curves)
        out[device_curves] = (x, {k: np.array(getattr(lp, k), dtype=float) for k in (
            "distance_to_ground_truth", "distanceToGroundTruthAfterRounding", "pobj_curve", "dobj_curve",
            "max_violated_constraint", "max_violated_equality", "max_violated_inequality")})
    (xd, cd), (xh, ch) = out[True], out[False]
    assert np.array_equal(xd, xh)
    for k in ch:
        if k.startswith("distance"):
            assert np.allclose(cd[k], ch[k], rtol=1e-13, atol=1e-15), k
        else:
            assert np.array_equal(cd[k], ch[k], equal_nan=True), k


def test_dual_warm_start_y0():
    """``y0=`` (extension; the reference always starts from y = 0, :166,:177): the solve continues from the given duals.
    Checked against the numpy oracle stepped from the same state: x0, y0 given, xbar = x0 as the reference sets it."""
    from oracle.cpppd_oracle import CpPpdOracle
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd

    args, g = case_args("random_small")
    rng = np.random.default_rng(11)
    m_eq, m_in = g["y_eq"].size, g["y_ineq"].size
    x0 = rng.standard_normal(args[0].size)
    y0 = np.concatenate((rng.standard_normal(m_eq), np.abs(rng.standard_normal(m_in))))
    o = CpPpdOracle(*args, x0=x0)
    o.y_eq, o.y_ineq = y0[:m_eq].copy(), y0[m_eq:].copy()
    for _ in range(40):
        o.primal_step()
        o.dual_step()
    x, best, solver = chambolle_pock_ppd(*args, x0=x0, y0=y0, nb_max_iter=40, nb_iter_plot=1000, return_solver=True)
    try:
        y = solver.get_y()
    finally:
        solver.close()
    assert np.array_equal(x, o.x) and np.array_equal(y, np.concatenate((o.y_eq, o.y_ineq)))
    with pytest.raises(ValueError):
        chambolle_pock_ppd(*args, y0=y0[:-1], nb_max_iter=1)
